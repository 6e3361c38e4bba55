#!/bin/bash
# ncu --set full of the VAE's GroupNorm-apply pass (level-0 launches of the second encode call)
TAG=${1:-gn}
mkdir -p gpurun_out
LDP_VAE_CHUNK=256 LDP_REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vae_gn_apply -s 22 -c 3 -o gpurun_out/prof_gn_$TAG -f python scripts/profile_vae.py > gpurun_out/ncu_gn_$TAG.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_gn_$TAG.log
