#!/bin/bash
# Loop-kernel bring-up: parity tests (loop path on), then timings with the loop on and off.
TAG=${1:-loop}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/ops_profile.py ${TAG}_on > gpurun_out/ops_${TAG}_on.log 2>&1
LDP_LOOP=0 timeout 300 python scripts/ops_profile.py ${TAG}_off > gpurun_out/ops_${TAG}_off.log 2>&1
head -1 gpurun_out/ops_${TAG}_on.log gpurun_out/ops_${TAG}_off.log
LDP_LOOP_DBG=50 LDP_REPS=1 timeout 300 python scripts/loop_dbg.py > gpurun_out/loopdbg_$TAG.log 2>&1; tail -40 gpurun_out/loopdbg_$TAG.log
