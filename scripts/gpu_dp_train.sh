#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_train_check.py 2>&1 | grep -v Warning | tail -5 | tee gpurun_out/dp_train_check.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/train_bench.py --steps 10 --warmup 3 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/train_bench_2gpu.json
