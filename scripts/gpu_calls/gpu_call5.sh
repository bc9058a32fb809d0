#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest5.log
tail -12 gpurun_out/r2_pytest5.log
timeout 300 python scripts/bench_pmh.py 2>&1 | sed -n 1,12p
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "splitdreal::--split-dreal" "pmh0:IDEAS_OPTS=pmh=0:"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab5_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'],'GiB')"
done
# racecheck / synccheck on small shapes of the three tcgen05 kernel families + pmh
K='(test_umma_forward_matches_simt and (case0 or case5)) or (test_halo_forward_matches_simt and (case0 or case5)) or (test_pmh_forward_matches_simt and (case0 or case6)) or (test_umma_wgrad_matches_simt and (case0 or case5)) or (test_pmh_dgrad_matches_simt and case2)'
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_umma.py -q -x -k "$K" > gpurun_out/sanitizer_${tool}_r2.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_r2.log | tail -3
done
