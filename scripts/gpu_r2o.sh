#!/bin/bash
timeout 300 python scripts/ops_profile.py wholegraph 2>&1 | head -1
LDP_LOOP_GRAPH=0 timeout 300 python scripts/ops_profile.py stepgraph 2>&1 | head -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_config_parity_gpu.py tests/test_agent_gpu.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider -k "planner or loop or act or agent" 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
