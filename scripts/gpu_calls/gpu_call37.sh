#!/bin/bash
bash scripts/ncu_cases.sh r2 conv_fwd_cfg3 conv_dgrad_cfg3
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-roofline --no-library-baseline --no-graphs > gpurun_out/launches_r2.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r2.csv
