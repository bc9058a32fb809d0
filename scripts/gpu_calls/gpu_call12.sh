#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "split::--split-dreal" "noconc::--no-concurrent-g"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab12_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'])"
done
timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q -x 2>&1 | tail -3
