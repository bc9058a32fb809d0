#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "concg::--concurrent-g" "concg_split::--concurrent-g --split-dreal"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab11_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'])"
done
