timeout 240 python -m pytest tests/test_gpu_patchify.py -x -q 2>&1 | tail -3
timeout 200 python profiles/time_patchify.py 2>&1 | head -2 | cut -c1-800
for ncw in 4 6 8; do
  for qw in 1 2; do
    PDAE_PATCHIFY_NCW=$ncw PDAE_PATCHIFY_QW=$qw timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu --no-configs 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('fused ncw=$ncw qw=$qw', round(d['value']), round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']), d['gpu_launches'])"
  done
done
