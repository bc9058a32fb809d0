#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/bench_halo_epi.py 2>&1 | tail -12
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "epi0:IDEAS_OPTS=halo_epi=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['clocks']['sm_mhz'])"
done
