#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02/bench_4gpu_v10.json 2> gpurun_out/r02/bench_4gpu_v10.err; echo rc=$?
tail -c 200 gpurun_out/r02/bench_4gpu_v10.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_4gpu_v10.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','timed_steps')}, d['e2e']['value'], d['e2e']['ms_per_step'])
s=d.get('sharded_c5') or {}
print(json.dumps(s.get('chamfer_forward'))[:400])
PY
