#!/bin/bash
# One GPU-box visit: parity tests, per-kernel timings, VAE throughput, bench line.
TAG=${1:-r1e}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 300 python scripts/ops_profile.py ${TAG}_default > gpurun_out/ops_${TAG}_default.log 2>&1
LDP_PAIR=0 timeout 300 python scripts/ops_profile.py ${TAG}_nopair > gpurun_out/ops_${TAG}_nopair.log 2>&1
head -1 gpurun_out/ops_${TAG}_*.log
timeout 300 python scripts/vae_bench.py > gpurun_out/vae_${TAG}.log 2>&1; tail -3 gpurun_out/vae_${TAG}.log
LDP_PERSIST=0 timeout 300 python scripts/vae_bench.py > gpurun_out/vae_${TAG}_nopersist.log 2>&1; tail -3 gpurun_out/vae_${TAG}_nopersist.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json | cut -c1-1500
