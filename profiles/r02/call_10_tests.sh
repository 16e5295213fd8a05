#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02/pytest_gpu_full_v2.txt
tail -8 gpurun_out/r02/pytest_gpu_full_v2.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_04.json 2> gpurun_out/r02/bench_04.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_04.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
