#!/bin/bash
# One `ncu --set full` capture per roofline kernel of bench.py (run on the GPU box through gpurun):
#   scripts/ncu_cases.sh r2 conv_fwd_cfg3 conv_wgrad_cfg3 ...        -> gpurun_out/<tag>_<case>.ncu-rep (+ .raw.csv)
# then, back in the build container:  python scripts/ncu_summary.py --traffic profiles/traffic.json gpurun_out/<tag>_*.ncu-rep
TAG=$1; shift
mkdir -p gpurun_out
for CASE in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -c 1 --launch-skip 2 \
      -k regex:'conv_umma|blur4|bias_act|upfirdn' -f -o gpurun_out/${TAG}_${CASE} \
      python bench.py --case $CASE --iters 1 > gpurun_out/${TAG}_${CASE}.log 2>&1
  ncu -i gpurun_out/${TAG}_${CASE}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${CASE}.raw.csv 2>/dev/null
  echo "$CASE rc=$?"
done
