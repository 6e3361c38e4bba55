#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_idm_loop_gpu.py tests/test_gpu_parity.py tests/test_bench_config_parity_gpu.py tests/test_agent_gpu.py -m gpu -q -x -s -p no:cacheprovider -k "idm or act or agent" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "parity\] idm|passed|failed|Error|error|timeout|rc=" gpurun_out/pytest_$TAG.log | tail -20
timeout 300 python scripts/idm_bench.py 2>&1 | tail -5
LDP_IDM_LOOP=0 timeout 300 python scripts/idm_bench.py 2>&1 | tail -3
