#!/bin/bash
# round-2 GPU call 1: health of the tests, baseline bench with comparators, reference on the same GPU, launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl gpurun_out/train_py_compat.log
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
export TORCH_EXTENSIONS_DIR=/root/repo/baseline/_ref/_ext TORCH_CUDA_ARCH_LIST=10.0a
timeout 900 python scripts/reference_step.py --ref baseline/_ref/IDEAS --device cuda --batch 32 --steps 32 --warmup 8 --budget 300 \
   > gpurun_out/reference_gpu_b32.json 2> gpurun_out/reference_gpu_b32.err
if ! grep -q images_per_s gpurun_out/reference_gpu_b32.json; then
  timeout 900 python scripts/reference_step.py --ref baseline/_ref/IDEAS --device cuda --batch 16 --steps 32 --warmup 8 --budget 300 \
     > gpurun_out/reference_gpu_b16.json 2> gpurun_out/reference_gpu_b16.err
fi
timeout 600 python scripts/reference_step.py --ref baseline/_ref/IDEAS --device cuda --batch 8 --steps 20 --warmup 4 --budget 120 --cudnn-tf32 0 \
   > gpurun_out/reference_gpu_b8_fp32.json 2> gpurun_out/reference_gpu_b8_fp32.err
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?"
timeout 900 python scripts/profile_convs.py 32 > gpurun_out/r2_profile_convs1.txt 2>&1
tail -5 gpurun_out/r2_pytest1.log; cat gpurun_out/reference_gpu_b32.json | cut -c1-600; tail -c 1500 gpurun_out/r2_bench1.json
