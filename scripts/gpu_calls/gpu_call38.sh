#!/bin/bash
for o in "pair_stages=4" "pair_stages=5" "pair_stages=6" "pair_stages=7"; do
  echo "== $o"; IDEAS_OPTS=$o timeout 120 python scripts/bench_pair.py 2>&1 | sed -n 2,5p
done
