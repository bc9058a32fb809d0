#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_nets.py -m gpu -q -x -k "shared_reconstruction or pruned or train_step_matches" 2>&1 | tail -4
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "share::--share-recon" "share_prune::--share-recon --prune-dead-backward" "default2::"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab46_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'], 'GiB', d['clocks']['sm_mhz'])"
done
