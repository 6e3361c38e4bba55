#!/bin/bash
# compute-sanitizer memcheck over the VAE's paired persistent convolutions with the TMA epilogue (benchmark topology, 64x64 images: level 0 is
# 160 tiles -> persistent CTA pairs) and over the small configurations (non-persistent launches, TMA epilogue, narrow channels).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool tag pytest-args...
  local tool=$1 tag=$2; shift 2
  timeout 1200 $CS --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitizer_${tool}_${tag}.log 2>&1
  echo "== $tool $tag rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_${tag}.log | tail -3
}
run memcheck vae_bench_topology "tests/test_gpu_parity.py::test_vae_benchmark_topology"
run memcheck vae_small "tests/test_gpu_parity.py::test_vae_small_configs"
run memcheck vae_decoder "tests/test_gpu_parity.py::test_vae_decoder_small_configs"
