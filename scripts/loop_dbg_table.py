"""Turn the LDP_LOOP_DBG stamps into per-layer durations."""
import re
import sys

rows = []
for line in open(sys.argv[1]):
    m = re.match(r"loop layer\s+(\d+): enter\s+(-?\d+) polled\s+(-?\d+) first-operands\s+(-?\d+) mma-issued\s+(-?\d+) epilogue-done\s+(-?\d+) tile0-done\s+(-?\d+)", line)
    if m:
        rows.append([int(x) for x in m.groups()])
print("layer  sync(epi-done(l-1)->polled)  poll->operands  main(operands->issued)  epilogue(issued->done)  layer total")
tot = [0, 0, 0, 0]
for i, (l, ent, pol, fo, mi, ed, t0d) in enumerate(rows):
    prev_ed = rows[i - 1][5] if i else 0
    a, b, c, d = (pol - prev_ed if i else 0), fo - max(pol, 0), mi - fo, ed - mi
    for k, v in enumerate((a, b, c, d)):
        tot[k] += v
    print(f"{l:5d}  {a:10d}  {b:10d}  {c:10d}  {d:10d}  {ed - prev_ed:10d}   tile0 epilogue {t0d - mi:8d}")
print("total ", *("%10d" % v for v in tot), "%10d" % (rows[-1][5] - rows[0][1]) if rows else "")
