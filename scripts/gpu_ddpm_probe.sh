#!/bin/bash
for f in 0 512 1024 2048 3584 3616; do
  echo "== LDP_LOOP_FLAGS=$f"
  LDP_LOOP=1 LDP_LOOP_FLAGS=$f LDP_LOOP_DBG=50 timeout 300 python scripts/loop_dbg.py > gpurun_out/loopdbg_f$f.log 2>&1; python scripts/loop_dbg_table.py gpurun_out/loopdbg_f$f.log | tail -2 | head -1
done
