#!/bin/bash
# ncu launch list of one 256-image VAE chunk (second encode call: warm), device time + DRAM bytes + tensor-pipe activity per launch
TAG=${1:-r2}
mkdir -p gpurun_out
LDP_REPS=2 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_vae_$TAG.csv python scripts/profile_vae.py > gpurun_out/ncu_vae_$TAG.log 2>&1
echo "list rc=$?"; tail -2 gpurun_out/ncu_vae_$TAG.log
