#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -8
for o in "pair_epi=0" "pair_epi=1"; do
  echo "== $o"; IDEAS_OPTS=$o timeout 120 python scripts/bench_pair.py 2>&1 | tail -12
done
