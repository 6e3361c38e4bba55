#!/bin/bash
# Round-2 visit A: the full GPU test suite (incl. the bench-configuration parity tests) and the new bench line.
TAG=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_bench_configs.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
grep -E "parity\]|passed|failed|Error|error" gpurun_out/pytest_$TAG.log | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$TAG.err
cut -c1-6000 gpurun_out/bench_$TAG.json
