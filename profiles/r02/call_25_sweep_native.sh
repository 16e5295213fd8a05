for env in "A=0" "PDAE_PATCHIFY_NCW=12" "PDAE_PATCHIFY_QW=2" "PDAE_TCC_BUILD_COST=16" "PDAE_TCC_BUILD_COST=24"; do
  env $env timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu --no-configs --steps 200 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$env', round(d['value']), round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']))"
done
