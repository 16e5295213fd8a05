#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r02/pytest_gpu_full.txt
tail -6 gpurun_out/r02/pytest_gpu_full.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_03.json 2> gpurun_out/r02/bench_03.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_03.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'])
print(json.dumps(d['ref_gpu'].get('ball_query_scene_scale')))
print(json.dumps(d['other_kernels'])[:900])
PY
