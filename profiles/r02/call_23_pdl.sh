timeout 300 python -m pytest tests/test_gpu_patchify.py tests/test_gpu_chamfer_tc.py -x -q -m gpu 2>&1 | tail -2
for cfg in "--patchifier tail" "--patchifier pdl" "--patchifier pdl --no-graphs"; do
  timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu --no-configs $cfg 2>gpurun_out/pdl_err.txt | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', round(d['value']), round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']), d['gpu_launches'])" || tail -5 gpurun_out/pdl_err.txt
done
