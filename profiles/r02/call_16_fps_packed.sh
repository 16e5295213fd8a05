timeout 400 python -m pytest tests/test_gpu_patchify.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_random_sweep.py -x -q -k "fps or patchif or group or Group" 2>&1 | tail -3
timeout 200 python profiles/time_patchify.py 2>&1 | head -2 | cut -c1-500
timeout 200 python bench.py --no-cpu-baseline --no-ref-gpu 2>/dev/null > gpurun_out/bench_z.json; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step']*1e3,1),'us  e2e', round(d['e2e']['value']), d['gpu_launches'])
for k,v in d['other_kernels'].items(): print(k, round(v['us'],1))
c=d['configs']
for k in c:
    if isinstance(c[k],dict):
        for kk,vv in c[k].items():
            if isinstance(vv,dict) and 'fps' in kk: print(k,kk,vv)
P
