#!/bin/bash
timeout 600 python scripts/update_host_profile.py 2>&1 | grep -E "host enqueue|as_tensor|_update_step"
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_agent_gpu.py tests/test_golden.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
