mkdir -p gpurun_out/r03
time timeout 600 python bench.py > gpurun_out/r03/bench_v7.json 2> gpurun_out/r03/bench_v7.err; tail -n 4 gpurun_out/r03/bench_v7.err; python - <<PY
import json
d=json.load(open("gpurun_out/r03/bench_v7.json"))
print(round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["ms_per_launch"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["ref_gpu"]["ms_per_step"])
print(json.dumps(d["other_kernels"], indent=0))
PY
time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03/bench_ref_v7.json 2> gpurun_out/r03/bench_ref_v7.err; tail -n 4 gpurun_out/r03/bench_ref_v7.err; cut -c1-400 gpurun_out/r03/bench_ref_v7.json
