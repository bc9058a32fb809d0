#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q -x 2>&1 | tail -8
for o in "halo_epi=3" "halo_epi=1"; do
  echo "== $o"; IDEAS_OPTS=$o timeout 120 python bench.py --case conv_halo_g256 --iters 20 2>&1 | tail -1 | cut -c1-200
  IDEAS_OPTS=$o timeout 120 python scripts/bench_dgrad_phases.py 2>&1 | tail -8 | cut -c1-75
done
