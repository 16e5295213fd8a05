set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q -k "sharded or keys_merge" > gpurun_out/r02/pytest_gpu11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu11.log
tail -15 gpurun_out/r02/pytest_gpu11.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu/run_sharded_nccl.py > gpurun_out/r02/sharded_nccl_2gpu.json 2> gpurun_out/r02/sharded_nccl_2gpu.err
cat gpurun_out/r02/sharded_nccl_2gpu.json; tail -5 gpurun_out/r02/sharded_nccl_2gpu.err
