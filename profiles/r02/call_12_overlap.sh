#!/bin/bash
mkdir -p gpurun_out/r02
timeout 100 python profiles/trace_chamfer_tc.py 2>&1 | grep "forward us"
for cfg in "3 tail2" "3 tail"; do
  set -- $cfg
  PDAE_CHAMFER_TC=$1 timeout 280 python bench.py --steps 20 --warmup 5 --patchifier $2 --no-configs --no-cpu-baseline --no-ref-gpu > gpurun_out/r02/bench_q_$1_$2.json 2> gpurun_out/r02/bench_q_$1_$2.err; echo "mode $1 patchifier $2 rc=$?"
  tail -2 gpurun_out/r02/bench_q_$1_$2.err
  python - "$1_$2" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r02/bench_q_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"])
PY
done
