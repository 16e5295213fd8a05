set -x
mkdir -p gpurun_out/r02
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 300 --warmup 10 > gpurun_out/r02/bench_8gpu.json 2> gpurun_out/r02/bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 tests/multi_gpu/run_sharded_nccl.py > gpurun_out/r02/sharded_nccl_8gpu.json 2> gpurun_out/r02/sharded_nccl_8gpu.err
cut -c1-300 gpurun_out/r02/bench_8gpu.json; cat gpurun_out/r02/sharded_nccl_8gpu.json; tail -3 gpurun_out/r02/bench_8gpu.err
