#!/bin/bash
timeout 600 python scripts/debug_graph_streams.py 2>&1 | grep graphs=
ONLY=1 CH=32 timeout 600 python scripts/debug_graph_streams.py 2>&1 | grep graphs=
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline --no-e2e"
for cfg in "default::" "noconc::--no-concurrent-g" "earlyg::--early-g"; do
  name=${cfg%%:*}; rest=${cfg#*:}; envs=${rest%%:*}; flags=${rest#*:}
  env $envs timeout 600 $B $flags 2> gpurun_out/r2_ab18_$name.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name', round(d['ms_per_step'],2), 'ms', d['gpu_launches'], 'launches', d['config']['peak_mem_gib'], d['clocks']['sm_mhz'])"
done
