#!/bin/bash
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 300 python scripts/ops_profile.py ${TAG}_pf > gpurun_out/ops_${TAG}_pf.log 2>&1
LDP_L2PF=0 timeout 300 python scripts/ops_profile.py ${TAG}_nopf > gpurun_out/ops_${TAG}_nopf.log 2>&1
head -1 gpurun_out/ops_${TAG}_pf.log gpurun_out/ops_${TAG}_nopf.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "planner or unet or reverse" 2>&1 | tail -3
