#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for b in 32 16 8; do
  timeout 600 $B --batch $b 2> gpurun_out/r2_ab9_b$b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch $b', round(d['ms_per_step'],2), 'ms', round(d['value'],2), 'img/s', d['config']['peak_mem_gib'],'GiB')"
done
bash scripts/ncu_cases.sh r2 conv_fwd_cfg3 conv_dgrad_cfg3 conv_wgrad_cfg3 conv_halo_g256 conv_lowch_e256 blur_cfg2 blur_act_bwd_cfg2 lrelu_fwd_cfg2 lrelu_bwd_cfg2 2>&1 | tail -10
ls -la gpurun_out/*.ncu-rep | head -20
