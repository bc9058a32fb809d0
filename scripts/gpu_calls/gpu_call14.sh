#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -k "multi_stream" --tb=short 2>&1 | tail -40 > gpurun_out/r2_pytest14.log
grep -n "assert\|Error\|passed\|failed" gpurun_out/r2_pytest14.log | head -20
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "default2::" "nohoist:IDEAS_AB_NO_DCO_HOIST=1:" "nosplit::--no-split-dreal" "nosplit_nohoist:IDEAS_AB_NO_DCO_HOIST=1:--no-split-dreal"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab14_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'], d['clocks']['sm_mhz'])"
done
