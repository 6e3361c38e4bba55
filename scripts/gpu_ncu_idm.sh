#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:idm_loop_kernel -s 2 -c 1 -o gpurun_out/prof_idm_loop_r2 -f python scripts/idm_bench.py > gpurun_out/ncu_idm_loop_r2.log 2>&1
tail -3 gpurun_out/ncu_idm_loop_r2.log
ls -la gpurun_out/prof_idm_loop_r2.ncu-rep
