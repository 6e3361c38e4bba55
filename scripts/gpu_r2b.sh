#!/bin/bash
# quick check of planner changes: planner parity tests + bare bench line
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_config_parity_gpu.py tests/test_golden.py tests/test_loop_kernel_gpu.py tests/test_agent_gpu.py -m gpu -q -x -s -p no:cacheprovider -k "not train and not vae" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
grep -E "parity\]|passed|failed|Error|error|rc=" gpurun_out/pytest_$TAG.log | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
python - <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_%s.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/bench_r2b.json"))
print(d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["isolated"])
PY
