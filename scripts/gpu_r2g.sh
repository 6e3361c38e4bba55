#!/bin/bash
LDP_IDM_LOOP_DBG=1 timeout 300 python scripts/idm_bench.py 2>&1 | tail -2
timeout 300 python scripts/idm_bench.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_idm_loop_gpu.py tests/test_gpu_parity.py tests/test_bench_config_parity_gpu.py -m gpu -q -x -s -p no:cacheprovider -k "idm" 2>&1 | grep -E "parity\] idm|passed|failed|Error|error|timeout" | tail
