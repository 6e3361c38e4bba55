#!/bin/bash
# N1 bring-up: compute-sanitizer on one small case, then the training tests
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_train_gpu.py -x -q -m gpu -k "planner_loss_and_grads and 6-dims0" > gpurun_out/train_sanitizer.log 2>&1
tail -5 gpurun_out/train_sanitizer.log
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -40
