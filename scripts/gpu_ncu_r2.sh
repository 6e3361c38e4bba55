#!/bin/bash
# r2 ncu evidence: planner launch list (2 denoising steps) + --set full of the step's tc_gemm launches; VAE --set full of the
# dominant level-0 convolution.  Summaries are produced here (no GPU) by scripts/ncu_summary.py / vae_launch_summary.py.
TAG=${1:-r2b}
WHAT=${2:-all}     # planner | vae | all  (the two .ncu-rep files together exceed gpurun's 64 MiB return limit: one visit each)
mkdir -p gpurun_out
if [ "$WHAT" != "vae" ]; then
LDP_STEPS=2 LDP_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py > gpurun_out/ncu_list_$TAG.log 2>&1
echo "list rc=$?"; tail -1 gpurun_out/ncu_list_$TAG.log
LDP_STEPS=2 LDP_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 32 -c 30 -o gpurun_out/prof_tc_$TAG -f python scripts/profile_step.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"; tail -1 gpurun_out/ncu_full_$TAG.log
fi
if [ "$WHAT" != "planner" ]; then
LDP_REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 30 -c 6 -o gpurun_out/prof_vae_$TAG -f python scripts/profile_vae.py > gpurun_out/ncu_vae_full_$TAG.log 2>&1
echo "vae full rc=$?"; tail -1 gpurun_out/ncu_vae_full_$TAG.log
fi
ls -la gpurun_out | grep $TAG
