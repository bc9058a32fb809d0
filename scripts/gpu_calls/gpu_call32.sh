#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "pair1::" "pair0:IDEAS_OPTS=pair=0:" "pair1b::"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab32_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['clocks']['sm_mhz'])"
done
