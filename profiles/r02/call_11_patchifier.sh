#!/bin/bash
mkdir -p gpurun_out/r02
for m in tail first overlap; do
  timeout 280 python bench.py --steps 20 --warmup 5 --patchifier $m --no-configs --no-cpu-baseline --no-ref-gpu > gpurun_out/r02/bench_p_$m.json 2> gpurun_out/r02/bench_p_$m.err; echo $m rc=$?
  tail -2 gpurun_out/r02/bench_p_$m.err
  python - "$m" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02/bench_p_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
