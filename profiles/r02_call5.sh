set -x
mkdir -p gpurun_out/r02
timeout 600 python profiles/tune_chamfer.py 0 14 15 > gpurun_out/r02/tune_chamfer5.json 2> gpurun_out/r02/tune_chamfer5.err
timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02/bench_torch.json 2> gpurun_out/r02/bench_torch.err
timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-ref-gpu --sched priority > gpurun_out/r02/bench_prio.json 2> gpurun_out/r02/bench_prio.err
cat gpurun_out/r02/tune_chamfer5.json; cut -c1-330 gpurun_out/r02/bench_torch.json; cut -c1-330 gpurun_out/r02/bench_prio.json; tail -5 gpurun_out/r02/bench_prio.err
