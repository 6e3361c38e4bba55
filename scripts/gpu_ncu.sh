#!/bin/bash
# ncu passes for profiles/: launch list (device time, DRAM bytes, tensor-pipe activity per launch) and one --set full capture
TAG=${1:-r1f}
mkdir -p gpurun_out
LDP_STEPS=2 LDP_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log
LDP_STEPS=2 LDP_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 42 -c 8 -o gpurun_out/prof_tc_$TAG python scripts/profile_step.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | grep $TAG
