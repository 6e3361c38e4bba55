#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json'))
print('plans/s', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
print(d['extras'])"
python __graft_entry__.py smoke 2>&1 | tail -5
