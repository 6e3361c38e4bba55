#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_golden.py -q -m gpu 2>&1 | tail -3
for h in 0 1; do echo "LDP_TRAIN_PDL=$h"; LDP_TRAIN_PDL=$h timeout 300 python scripts/train_bench.py --steps 20 --warmup 4 2>&1 | tail -1; done
