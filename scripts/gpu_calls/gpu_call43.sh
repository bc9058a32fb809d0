#!/bin/bash
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_ref43.json 2> gpurun_out/r2_bench_ref43.err; echo ref rc=$?; cut -c1-700 gpurun_out/r2_bench_ref43.json
timeout 900 python bench.py > gpurun_out/r2_bench43.json 2> gpurun_out/r2_bench43.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench43.json'))
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'plain',d['config']['ms_per_step_plain'],'r1',d['config']['ms_per_step_r1'],'e2e',round(d['e2e']['value'],2),'launches',d['gpu_launches'],d['clocks'])
for k,v in d.items():
    if k.startswith('roofline'):
        print(k, v['kernel'][:40], round(v['achieved'],1), v['unit'], 'frac', round(v['frac'],3))
print(d.get('cpu_baseline'))
lb=d.get('library_baseline',{})
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in list(v.items())[:8]}) for k,v in lb.items() if k!='cfg2'})
PY
