#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('plans/s', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
for bn in 128 0; do echo "LDP_TRAIN_BN=$bn"; LDP_TRAIN_BN=$bn timeout 300 python scripts/train_bench.py --steps 10 --warmup 3 2>&1 | tail -1; done
