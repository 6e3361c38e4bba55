#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -m gpu 2>&1 | tail -5
timeout 600 python scripts/train_bench.py --steps 5 --warmup 2 2>&1 | tail -3 | tee gpurun_out/train_bench_bf16.json
