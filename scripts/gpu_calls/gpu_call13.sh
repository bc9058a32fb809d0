#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "nosplit::--no-split-dreal" "earlyg::--early-g"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab13_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'])"
done
export TORCH_EXTENSIONS_DIR=/root/repo/baseline/_ref/_ext TORCH_CUDA_ARCH_LIST=10.0a
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err
echo "bench rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_bench13.json'))
print(d['value'], d['ms_per_step'], d['config']['ms_per_step_plain'], d['config']['ms_per_step_r1'], d['e2e']['value'], d['gpu_launches'])
for k,v in d.items():
    if k.startswith('roofline'): print(k, round(v['achieved'],1), v['unit'], round(v['frac'],3))
print(d['tf32_gemm_peak']); print(d['library_baseline']['cfg3']); print(d['library_baseline']['cfg4']); print(d['cpu_baseline'])"
