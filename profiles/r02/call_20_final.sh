#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02/pytest_gpu_full_final.txt
tail -3 gpurun_out/r02/pytest_gpu_full_final.txt
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; echo bench rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_ref_final.json 2>> gpurun_out/r02/bench_final.err; cut -c1-300 gpurun_out/r02/bench_ref_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_bench_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-configs > gpurun_out/r02/bench_under_ncu.log 2>&1; echo ncu rc=$?
timeout 200 python profiles/time_patchify.py > /dev/null 2>&1
timeout 120 python profiles/trace_patchify.py > /dev/null 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
r = d["roofline"]
print(r["ms_per_launch"], r["frac"], r["share_of_step"])
print(json.dumps(d["ref_gpu"]["kernels_only_speedup"]))
for k, v in d["other_kernels"].items(): print(k, round(v["us"], 1))
for k, v in d["configs"].items():
    if k[:2] in ("C2", "C4", "C5"):
        print(k, json.dumps(v)[:600])
PY
