set -x
mkdir -p gpurun_out/r02
timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-ref-gpu --sched priority > gpurun_out/r02/bench_prio.json 2> gpurun_out/r02/bench_prio.err
cut -c1-330 gpurun_out/r02/bench_prio.json; tail -5 gpurun_out/r02/bench_prio.err
