#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2_pytest6.log
tail -12 gpurun_out/r2_pytest6.log
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "notail:IDEAS_OPTS=tail_split=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab6_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'],'GiB')"
  tail -2 gpurun_out/r2_ab6_$name.err
done
for c in conv_fwd_cfg3 conv_dgrad_cfg3; do python bench.py --case $c --iters 10; IDEAS_OPTS=tail_split=0 python bench.py --case $c --iters 10; done
timeout 600 python scripts/profile_step.py 32 > gpurun_out/r2_profile_step6.txt 2>&1
head -40 gpurun_out/r2_profile_step6.txt
