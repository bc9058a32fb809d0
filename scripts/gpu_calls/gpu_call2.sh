#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest2a.log
tail -15 gpurun_out/r2_pytest2a.log
timeout 600 python scripts/bench_pmh.py > gpurun_out/r2_bench_pmh.txt 2>&1
cat gpurun_out/r2_bench_pmh.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_umma.py --deselect tests/test_gpu_ops.py 2>&1 | tail -40 > gpurun_out/r2_pytest2b.log
tail -15 gpurun_out/r2_pytest2b.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-library-baseline --no-roofline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
tail -c 1200 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
