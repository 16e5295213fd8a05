for cfg in "--patchifier tail" "--patchifier overlap --sched priority" "--patchifier overlap" "--patchifier tail --sched priority"; do
  timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu --no-configs $cfg 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', round(d['value']), round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']), d['gpu_launches'])"
done
