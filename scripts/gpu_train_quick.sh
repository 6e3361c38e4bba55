#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_golden.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/train_bench.py --steps 20 --warmup 4 2>&1 | tail -1
