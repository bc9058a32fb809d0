#!/bin/bash
mkdir -p gpurun_out
export TORCH_EXTENSIONS_DIR=/root/repo/baseline/_ref/_ext TORCH_CUDA_ARCH_LIST=10.0a
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final.json'))
print(d['value'], d['ms_per_step'], d['config']['ms_per_step_plain'], d['config']['ms_per_step_r1'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
for k,v in d.items():
    if k.startswith('roofline'): print(k, round(v['achieved'],1), v['unit'], round(v['frac'],3), v.get('traffic'))
print(d['tf32_gemm_peak']); print(d['library_baseline']['cfg4']); print(d['cpu_baseline'])
PY
bash scripts/ncu_cases.sh r2 conv_lowch_e256 conv_fwd_cfg3 2>&1 | tail -2
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline --no-library-baseline --no-graphs > gpurun_out/launches_r2.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r2.csv
