mkdir -p gpurun_out/r03
timeout 300 compute-sanitizer --tool memcheck python profiles/sanitize_smoke.py > gpurun_out/r03/sanitize_memcheck.log 2>&1; tail -2 gpurun_out/r03/sanitize_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python profiles/sanitize_smoke.py > gpurun_out/r03/sanitize_racecheck.log 2>&1; tail -2 gpurun_out/r03/sanitize_racecheck.log
cat > /tmp/fwd_alone.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch
from pointdae_b200 import ops, synth, group
c = torch.from_numpy(synth.clouds(128, 2048, seed=1)).cuda(); p = torch.from_numpy(synth.prediction(synth.clouds(128, 2048, seed=1), seed=1)).cuda()
for _ in range(3): ops.chamfer_forward(p, c)
g = group.Group(64, 32)
for _ in range(2): g.forward_corrupted(c, mats=torch.randn(128, 3, 3, 3))
torch.cuda.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"chamfer_min_kernel|chamfer_col_recover_list|knn3_kernel|fill_keys" -s 6 -c 8 -o gpurun_out/r03/ncu_split python /tmp/fwd_alone.py > gpurun_out/r03/ncu_split.log 2>&1; tail -3 gpurun_out/r03/ncu_split.log
ls -la gpurun_out/r03/*.ncu-rep
