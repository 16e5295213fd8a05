mkdir -p gpurun_out/r03
timeout 300 python -m pytest tests -m gpu -q -k "split or shard or Shard or chamfer or Chamfer or sweep or fullsize" > gpurun_out/r03/pytest_split3.log 2>&1; tail -n 4 gpurun_out/r03/pytest_split3.log
python - <<'PY'
import sys, json; sys.path.insert(0, ".")
import torch
from pointdae_b200 import ops, synth, _native
L = _native.lib()
a = torch.from_numpy(synth.prediction(synth.clouds(1, 100000, seed=5), seed=5)).cuda(); c = torch.from_numpy(synth.clouds(1, 100000, seed=5)).cuda()
out = {}
for w in (2, 8):
    sl = c[:, : 100000 // w].contiguous()
    for nc in (1, 0):
        L.pdae_tune_chamfer_split(nc)
        for _ in range(3): ops.chamfer_sharded_local(a, sl, 0)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.chamfer_sharded_local(a, sl, 0)
        e1.record(); torch.cuda.synchronize()
        out["per-rank sharded forward, world %d, split %s" % (w, "auto" if nc == 0 else "off")] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
L.pdae_tune_chamfer_split(0)
print(json.dumps(out))
open("gpurun_out/r03/sharded_split_local.json", "w").write(json.dumps(out))
PY
