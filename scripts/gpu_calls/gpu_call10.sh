#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/bench_pmh.py 2>&1 | sed -n 1,8p
IDEAS_OPTS=pmh_resident=0 timeout 300 python scripts/bench_pmh.py 2>&1 | sed -n 2,5p
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "noresident:IDEAS_OPTS=pmh_resident=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab10_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches')"
done
