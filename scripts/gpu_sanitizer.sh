#!/bin/bash
# compute-sanitizer over the mbarrier / TMEM kernels: memcheck + racecheck on a dense tcgen05 GEMM, one bf16 UNet forward,
# the persistent IDM loop kernel, the persistent planner loop kernel (opt-in path) and a VAE encode.  Summaries -> gpurun_out/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool tag pytest-args...
  local tool=$1 tag=$2; shift 2
  timeout 1500 $CS --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitizer_${tool}_${tag}.log 2>&1
  echo "== $tool $tag rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_${tool}_${tag}.log | tail -3
}
run memcheck tc_dense "tests/test_gpu_parity.py::test_tc_dense"
run racecheck tc_dense "tests/test_gpu_parity.py::test_tc_dense"
run memcheck unet_bf16 "tests/test_gpu_parity.py::test_unet_forward_bf16" -k "4-8-50"
run racecheck unet_bf16 "tests/test_gpu_parity.py::test_unet_forward_bf16" -k "4-8-50"
run memcheck idm_loop "tests/test_idm_loop_gpu.py::test_loop_kernel_matches_oracle_and_per_layer_path" -k "265-7-128-4"
run racecheck idm_loop "tests/test_idm_loop_gpu.py::test_loop_kernel_matches_oracle_and_per_layer_path" -k "265-7-128-4"
LDP_LOOP=1 run memcheck planner_loop "tests/test_loop_kernel_gpu.py::test_loop_equals_graph_path_philox_and_ddim"
run memcheck vae "tests/test_gpu_parity.py::test_vae_small_configs"
