#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "pmh0:IDEAS_OPTS=pmh=0:" "nobatchg::--no-batch-g" "cublaslin:IDEAS_AB_CUBLAS_LINEAR=1:" "all3:IDEAS_OPTS=pmh=0 IDEAS_AB_CUBLAS_LINEAR=1:--no-batch-g"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'],'GiB')"
done
timeout 600 python scripts/profile_step.py 32 > gpurun_out/r2_profile_step4.txt 2>&1
head -45 gpurun_out/r2_profile_step4.txt
# pmh variants on the lowch shapes
for o in "pmh_xstages=2" "pmh_xstages=3" "pmh_minbw=100" "pmh_minbw=100,pmh_xstages=3"; do
  echo "== $o"; IDEAS_OPTS=$o ITERS=5 timeout 300 python scripts/bench_pmh.py 2>&1 | sed -n 2,5p
done
bash scripts/ncu_cases.sh r2 conv_lowch_e256 2>&1 | tail -2
ncu -i gpurun_out/r2_conv_lowch_e256.ncu-rep --page details --csv 2>/dev/null > gpurun_out/r2_conv_lowch_e256.details.csv
