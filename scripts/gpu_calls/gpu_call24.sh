#!/bin/bash
for o in "pmh_align=0" "pmh_align=8" "pmh_align=16"; do
  echo "== $o"; PMH_PROBE=1 IDEAS_OPTS=$o ITERS=20 timeout 300 python scripts/bench_pmh.py 2>&1 | tail -6
done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -3
