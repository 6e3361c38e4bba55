#!/bin/bash
# VAE encoder: TMA epilogue / paired persistent convolutions vs the earlier paths.  Tests first (a hang must not eat the visit).
TAG=${1:-vp}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k vae > gpurun_out/vae_tests_$TAG.log 2>&1; echo "vae tests rc=$?"; tail -3 gpurun_out/vae_tests_$TAG.log
timeout 300 python -m pytest tests/test_bench_config_parity_gpu.py -x -q -m gpu -k vae > gpurun_out/vae_tests2_$TAG.log 2>&1; echo "vae chunk test rc=$?"; tail -3 gpurun_out/vae_tests2_$TAG.log
for cfg in "1 1" "1 0" "0 1" "0 0"; do
  set -- $cfg
  LDP_VAE_EPI_TMA=$1 LDP_VAE_PAIR=$2 timeout 300 python scripts/vae_bench.py > gpurun_out/vae_${TAG}_tma$1_pair$2.log 2>&1; echo "tma=$1 pair=$2 rc=$?"; tail -1 gpurun_out/vae_${TAG}_tma$1_pair$2.log | cut -c1-200
done
LDP_VAE_DBG=1 LDP_REPS=4 timeout 200 python scripts/profile_vae.py 2>&1 | tail -32 > gpurun_out/vae_dbg_$TAG.log
