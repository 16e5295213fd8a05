#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_02.json 2> gpurun_out/r02/bench_02.err; echo rc=$?
tail -c 300 gpurun_out/r02/bench_02.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_02.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','timed_steps','timed_region_ms')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['alone_with_column_split']['frac'])
print(json.dumps(d.get('configs'))[:3000])
print(json.dumps(d.get('ref_gpu'))[:1200])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_ref_02.json 2>> gpurun_out/r02/bench_02.err; cut -c1-700 gpurun_out/r02/bench_ref_02.json
