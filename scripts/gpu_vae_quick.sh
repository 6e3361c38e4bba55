#!/bin/bash
# VAE tests + throughput under a few environment settings: bash scripts/gpu_vae_quick.sh TAG "VAR=val VAR2=val" "..."
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k vae > gpurun_out/vae_tests_$TAG.log 2>&1; echo "vae tests rc=$?"; tail -2 gpurun_out/vae_tests_$TAG.log
timeout 300 python -m pytest tests/test_bench_config_parity_gpu.py tests/test_agent_gpu.py -x -q -m gpu -k "vae or act" > gpurun_out/vae_tests2_$TAG.log 2>&1; echo "vae chunk/act test rc=$?"; tail -2 gpurun_out/vae_tests2_$TAG.log
i=0
for cfg in "$@"; do
  env $cfg timeout 300 python scripts/vae_bench.py > gpurun_out/vae_${TAG}_$i.log 2>&1; echo "[$cfg] rc=$?"; tail -1 gpurun_out/vae_${TAG}_$i.log | cut -c1-170
  i=$((i+1))
done
