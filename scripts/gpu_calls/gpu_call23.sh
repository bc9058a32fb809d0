#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 4 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n2b.json')); print(d['value'], d['ms_per_step'], d['replicas_identical'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/r2_bench_n2b.err
