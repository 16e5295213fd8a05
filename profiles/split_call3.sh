mkdir -p gpurun_out/r03
timeout 300 python -m pytest tests -m gpu -q -k "chamfer or Chamfer or sweep or fullsize or losses or split" > gpurun_out/r03/pytest_split2.log 2>&1; tail -4 gpurun_out/r03/pytest_split2.log
timeout 300 python profiles/probe_split.py > gpurun_out/r03/probe_split_v2.json 2> gpurun_out/r03/probe_split.err; cat gpurun_out/r03/probe_split_v2.json; tail -3 gpurun_out/r03/probe_split.err
timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu > gpurun_out/r03/bench_split_v2.json 2> gpurun_out/r03/bench_split.err; tail -3 gpurun_out/r03/bench_split.err; python - <<PY
import json
d=json.load(open("gpurun_out/r03/bench_split_v2.json"))
print(round(d["value"]), round(d["ms_per_step"]*1e3,1), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"]*1e3,1), d["roofline"]["ms_per_launch"], d["roofline"]["ms_per_launch_unsplit"], d["roofline"]["frac"])
PY
