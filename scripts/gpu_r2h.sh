#!/bin/bash
timeout 300 python scripts/vae_bench.py 2>&1 | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bench_config_parity_gpu.py tests/test_agent_gpu.py tests/test_golden.py -m gpu -q -s -p no:cacheprovider -k "vae or agent or act or sdvae" 2>&1 | grep -E "parity\] vae|passed|failed|Error|error|assert" | tail
