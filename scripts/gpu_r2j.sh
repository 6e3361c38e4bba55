#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_train_check.py 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/dp_train_check_r2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_overlap_probe.py 2>&1 | grep -v Warning | tail -1 | tee gpurun_out/dp_overlap_probe.json
