"""Summarise gpurun_out/launches_vae_<tag>.csv (scripts/gpu_ncu_vae.sh) into profiles/<tag>_vae_launches.md: the LAST encode
call's launches of one 256-image chunk, per kernel and per launch."""
import csv
import collections
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
rows = list(csv.reader(open(ROOT / "gpurun_out" / f"launches_vae_{tag}.csv")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ci = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    d = launches.setdefault(int(r[ci["ID"]]), {"name": r[ci["Kernel Name"]], "grid": r[ci["Grid Size"]]})
    d[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
    d[r[ci["Metric Name"]] + ":unit"] = r[ci["Metric Unit"]]
ls = list(launches.values())
# the last call = from the last im2col / conv_in launch on
first = max(i for i, l in enumerate(ls) if "im2col" in l["name"] or "conv_in" in l["name"])
ls = ls[first:]


def us(l):
    v = l.get("gpu__time_duration.sum", 0.0)
    u = l.get("gpu__time_duration.sum:unit", "ns")
    return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)


def mb(l, k):
    v = l.get(k, 0.0)
    u = l.get(k + ":unit", "byte")
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)


tot = sum(us(l) for l in ls)
agg = collections.OrderedDict()
for l in ls:
    k = l["name"].split("(")[0].replace("void ldp::", "").replace("ldp::", "")[:48]
    a = agg.setdefault(k, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += us(l)
    a[2] += mb(l, "dram__bytes_read.sum") + mb(l, "dram__bytes_write.sum")
    a[3] += l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us(l)
    a[4] += mb(l, "lts__t_bytes.sum")
out = [f"# ncu launch list `{tag}`: one 256-image chunk of the VAE encoder (4-block SD-VAE, 64x64x3 uint8 -> 8x8x4), bf16 path", "",
       "`scripts/gpu_ncu_vae.sh` (`ncu --metrics gpu__time_duration.sum,dram__bytes_*,sm__pipe_tensor_cycles_active,lts__t_bytes --clock-control none`);",
       "per-launch times are cold-cache and serialised - compare SHARES, not absolutes.", "",
       f"{len(ls)} launches, {tot / 1e3:.2f} ms summed device time ({256 / tot * 1e6:.0f} img/s if they ran back to back).", "",
       "| kernel | launches | total us | share | DRAM GB | DRAM GB/s | L2 GB | tensor-pipe active % (time-weighted) |", "|---|---|---|---|---|---|---|---|"]
for k, (n, t, d, tp, l2) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {d / 1e3:.2f} | {d / t * 1e3 / 1e3:.0f} | {l2 / 1e3:.2f} | {tp / t if t else 0:.1f} |")
out += ["", "## the ten longest launches", "", "| kernel | grid | us | DRAM MB | L2 MB | tensor % |", "|---|---|---|---|---|---|"]
for l in sorted(ls, key=lambda l: -us(l))[:10]:
    out.append(f"| `{l['name'].split('(')[0].replace('void ldp::', '')[:40]}` | {l['grid']} | {us(l):.0f} | {mb(l, 'dram__bytes_read.sum') + mb(l, 'dram__bytes_write.sum'):.0f} | "
               f"{mb(l, 'lts__t_bytes.sum'):.0f} | {l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0.0):.1f} |")
(ROOT / "profiles" / f"{tag}_vae_launches.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
