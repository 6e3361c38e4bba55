#!/bin/bash
# per-kernel device time, DRAM bytes and tensor-pipe activity of one LDPAgent.update (the last of four under ncu)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 4000 --csv --log-file gpurun_out/train_kernels.csv python scripts/train_bench.py --steps 1 --warmup 3 > gpurun_out/train_ncu2.log 2>&1
tail -1 gpurun_out/train_ncu2.log
wc -l gpurun_out/train_kernels.csv
