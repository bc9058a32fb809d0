#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x -k "pair" 2>&1 | tail -3
timeout 600 python scripts/profile_step.py 32 > gpurun_out/profile_step_r2b.txt 2> gpurun_out/profile_step_r2b.err; tail -3 gpurun_out/profile_step_r2b.err; head -75 gpurun_out/profile_step_r2b.txt
