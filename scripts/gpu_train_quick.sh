#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_golden.py -q -m gpu 2>&1 | tail -3
for st in 1 0; do echo "LDP_TRAIN_STREAMS=$st (1 = single stream)"; LDP_TRAIN_STREAMS=$st timeout 300 python scripts/train_bench.py --steps 10 --warmup 3 2>&1 | tail -1; done
