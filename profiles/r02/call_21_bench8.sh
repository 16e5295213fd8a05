#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/bench_8gpu_v10.json 2> gpurun_out/r02/bench_8gpu_v10.err; echo rc=$?
tail -c 300 gpurun_out/r02/bench_8gpu_v10.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02/bench_8gpu_v10.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','timed_steps')}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(json.dumps(d.get('sharded_c5'))[:1800])
PY
