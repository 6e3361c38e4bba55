#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel timings under the tuning switches, ncu launch list.
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/ops_profile.py ${TAG}_default > gpurun_out/ops_${TAG}_default.log 2>&1
LDP_BN64=0 timeout 300 python scripts/ops_profile.py ${TAG}_bn128 > gpurun_out/ops_${TAG}_bn128.log 2>&1
LDP_NO_PDL=1 timeout 300 python scripts/ops_profile.py ${TAG}_nopdl > gpurun_out/ops_${TAG}_nopdl.log 2>&1
LDP_B=128 timeout 300 python scripts/ops_profile.py ${TAG}_b128 > gpurun_out/ops_${TAG}_b128.log 2>&1
head -1 gpurun_out/ops_${TAG}_*.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json | cut -c1-600
LDP_STEPS=2 LDP_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python scripts/profile_step.py > gpurun_out/ncu_list_$TAG.log 2>&1
LDP_STEPS=2 LDP_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 40 -c 4 -o gpurun_out/prof_tc_$TAG python scripts/profile_step.py > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -20
