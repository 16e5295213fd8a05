timeout 600 python -m pytest tests/test_gpu_chamfer_tc.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
for bc in 0 12 20 28; do
  PDAE_TCC_BUILD_COST=$bc timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu --no-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('build_cost=$bc', round(d['value']), round(d['ms_per_step']*1e3,1),'us  fwd alone', round(d['roofline']['ms_per_launch']*1e3,1), 'e2e', round(d['e2e']['value']))"
done
