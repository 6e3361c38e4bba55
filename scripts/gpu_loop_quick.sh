#!/bin/bash
# quick loop-kernel timing + per-layer stamps
TAG=${1:-q}
mkdir -p gpurun_out
LDP_LOOP=1 timeout 300 python scripts/ops_profile.py ${TAG}_on 2>&1 | head -1
LDP_LOOP=1 LDP_LOOP_DBG=50 timeout 300 python scripts/loop_dbg.py > gpurun_out/loopdbg_$TAG.log 2>&1; python scripts/loop_dbg_table.py gpurun_out/loopdbg_$TAG.log
if [ -n "$2" ]; then LDP_DBG_STAGES=1 LDP_PAIR=0 timeout 300 python scripts/ops_profile.py ${TAG}_stages > gpurun_out/stages_$TAG.log 2>&1; grep "epilogue deltas" gpurun_out/stages_$TAG.log | sed 's/stage arrivals.*|//' ; fi
