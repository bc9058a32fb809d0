#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "pair" 2>&1 | tail -8
timeout 300 python scripts/bench_pair.py 2>&1 | tail -13
timeout 300 python scripts/bench_sustained.py pair 2>&1 | tail -7
