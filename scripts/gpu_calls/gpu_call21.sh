#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_nets.py tests/test_gpu_ops.py tests/test_checkpoint.py -m gpu -q 2>&1 | tail -5
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "default2::"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab21_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'], d['clocks']['sm_mhz'])"
done
timeout 600 python scripts/profile_step.py 32 > gpurun_out/r2_profile_step21.txt 2>&1
head -50 gpurun_out/r2_profile_step21.txt | cut -c1-150
