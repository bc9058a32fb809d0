#!/bin/bash
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench50.json 2> gpurun_out/r2_bench50.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench50.json'))
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],2),'plain',d['config']['ms_per_step_plain'],'r1',d['config']['ms_per_step_r1'],'e2e',round(d['e2e']['value'],2),'launches',d['gpu_launches'],d['clocks'])
for k,v in d.items():
    if k.startswith('roofline'):
        print(k, v['kernel'][:40], round(v['achieved'],1), v['unit'], 'frac', round(v['frac'],3))
print(d.get('tf32_gemm_peak')); print(d.get('cpu_baseline'))
PY
