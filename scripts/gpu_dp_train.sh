#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_train_check.py 2>&1 | grep -v Warning | tail -2 | tee gpurun_out/dp_train_check.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/train_bench.py --steps 20 --warmup 4 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/train_bench_2gpu.json
