"""Summarise ncu output brought back in gpurun_out/ into profiles/ (tracked).

    python scripts/ncu_summary.py <tag>      # reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_tc_<tag>.ncu-rep
Writes profiles/<tag>_launches.md (per-kernel share of a denoising step) and profiles/<tag>_top_kernel.md
(`ncu --set full` metrics of the captured tc_gemm launches + the most-sampled SASS instructions)."""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
out = ROOT / "profiles"
out.mkdir(exist_ok=True)


def read_csv(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            return rows[i], rows[i + 1:]
    return None, []


lc = ROOT / "gpurun_out" / f"launches_{tag}.csv"
if lc.exists():
    h, data = read_csv(lc)
    ix = {n: h.index(n) for n in h}
    per = collections.OrderedDict()
    for r in data:
        if len(r) < len(h):
            continue
        d = per.setdefault(r[0], {"k": r[ix["Kernel Name"]], "g": r[ix["Grid Size"]], "b": r[ix["Block Size"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    agg = collections.OrderedDict()
    for d in per.values():
        name = d["k"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list `{tag}` (scripts/profile_step.py: 2 bf16 denoising steps at B=1024, T=8, D=265, incl. handle creation)",
             "", "`ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum,sm__pipe_tensor_cycles_active... --clock-control none`;",
             "per-launch times are cold-cache and serialised - compare SHARES, not absolutes.", "",
             "| kernel | launches | total us | share | avg us | DRAM MB/launch | tensor-pipe active % (avg) |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.2f} | {a[2] / a[0] / 1e6:.2f} | {a[3] / a[0]:.1f} |")
    tc = [d for d in per.values() if "tc_gemm" in d["k"]]
    if tc:
        n_step = 30
        last = tc[-n_step:]
        lines += ["", f"## the last denoising step's {len(last)} tcgen05 launches", "",
                  "| # | kernel | grid | block | us | DRAM read MB | tensor % |", "|---|---|---|---|---|---|---|"]
        for i, d in enumerate(last):
            lines.append(f"| {i} | `{d['k'].split('(')[0].replace('void ', '')}` | {d['g']} | {d['b']} | {d.get('gpu__time_duration.sum', 0) / 1e3:.2f} | "
                         f"{d.get('dram__bytes_read.sum', 0) / 1e6:.2f} | {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):.1f} |")
        lines.append(f"\nsum = {sum(d.get('gpu__time_duration.sum', 0) for d in last) / 1e3:.1f} us")
    (out / f"{tag}_launches.md").write_text("\n".join(lines) + "\n")
    print("wrote", out / f"{tag}_launches.md")

rep = ROOT / "gpurun_out" / f"prof_tc_{tag}.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_active.avg"]
    lines = [f"# ncu --set full capture `{tag}`: tcgen05 implicit-GEMM launches of one denoising step (B=1024, T=8, D=265)", "",
             "`ncu --set full --clock-control none --import-source on -k regex:tc_gemm`; read with `ncu -i ... --page raw --csv`.", ""]
    for r in rows[2:]:
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in keys:
            if k in h:
                lines.append(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |")
        lines.append("")
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    secs, cur = [], None
    hdr = None
    for r in srows:
        if r and r[0] == "Kernel Name":
            cur = [r[1]]
            secs.append(cur)
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if cur is not None and len(r) > 5:
            cur.append(r)
    if secs and hdr:
        st = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        sec = secs[min(1, len(secs) - 1)]
        ins = sec[1:]
        tot = sum(int(r[4] or 0) for r in ins)
        c = collections.Counter()
        for r in ins:
            for i in st:
                c[hdr[i]] += int(r[i] or 0)
        lines += [f"## warp-state sampling of `{sec[0]}` ({len(ins)} SASS instructions, {tot} samples)", "",
                  "stall reasons (all samples): " + ", ".join(f"{k[6:]} {v}" for k, v in c.most_common(8)), "",
                  "| SASS index | samples | instruction |", "|---|---|---|"]
        top = sorted(range(len(ins)), key=lambda i: -int(ins[i][4] or 0))[:25]
        for i in sorted(top):
            lines.append(f"| {i} | {ins[i][4]} | `{ins[i][1].strip()[:90]}` |")
        mn = [r[1].strip().split()[0] for r in ins if r[1].strip()]
        proof = {m: sum(1 for x in mn if x.startswith(m)) for m in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA")}
        lines += ["", "SASS evidence (instruction counts in this kernel): " + ", ".join(f"{k} {v}" for k, v in proof.items())]
    (out / f"{tag}_top_kernel.md").write_text("\n".join(lines) + "\n")
    print("wrote", out / f"{tag}_top_kernel.md")
