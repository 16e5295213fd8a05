set -x
mkdir -p gpurun_out/r03
timeout 300 python profiles/probe_split.py > gpurun_out/r03/probe_split.json 2> gpurun_out/r03/probe_split.err; cat gpurun_out/r03/probe_split.json; tail -3 gpurun_out/r03/probe_split.err
timeout 300 python -m pytest tests -m gpu -q -k "chamfer or Chamfer or sweep or fullsize or losses" > gpurun_out/r03/pytest_split.log 2>&1; tail -4 gpurun_out/r03/pytest_split.log
timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu > gpurun_out/r03/bench_split.json 2> gpurun_out/r03/bench_split.err; cut -c1-300 gpurun_out/r03/bench_split.json; grep -o '"roofline.*ms_per_launch[^,]*' gpurun_out/r03/bench_split.json | tail -c 200
