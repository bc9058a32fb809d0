#!/bin/bash
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "split1::--split-dreal" "default::" "split2::--split-dreal"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'], d['clocks'])"
done
