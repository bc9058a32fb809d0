#!/bin/bash
timeout 300 python scripts/bench_resample.py 2>&1 | tail -14
for v in 0 4 5 6 7 3; do
  echo "== blur_variant=$v"; IDEAS_OPTS=blur_variant=$v timeout 120 python bench.py --case blur_cfg2 --iters 20 2>&1 | tail -1 | cut -c90-200
done
