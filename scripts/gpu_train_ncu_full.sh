#!/bin/bash
# ncu --set full on a few launches of the two kernels that dominate the training step
mkdir -p gpurun_out
LDP_TRAIN_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"prep_kernel|tc_gemm_kernel" -s 700 -c 12 -o gpurun_out/prof_train python scripts/train_bench.py --steps 1 --warmup 2 > gpurun_out/ncu_train_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_train_full.log; ls -la gpurun_out/prof_train.ncu-rep
