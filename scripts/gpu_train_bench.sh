#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_bench.py --steps 5 --warmup 2 2>&1 | tail -3 | tee gpurun_out/train_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/train_launches.csv python scripts/train_bench.py --steps 1 --warmup 3 > gpurun_out/train_ncu.log 2>&1
tail -2 gpurun_out/train_ncu.log
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/train_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
n = len(rows) - 1
half = rows[1 + 3 * n // 4:]        # last of four updates
agg = collections.defaultdict(lambda: [0, 0.0])
for r in half:
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] == 'ns' else v
    k = r[ki].split('(')[0][:60]; agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("last-quarter launches", len(half), 'total us', tot)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'{t:10.1f} us {100*t/tot:5.1f}% n={c:4d} {k}')
PY
