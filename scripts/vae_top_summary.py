"""profiles/<tag>_vae_top_kernel.md from gpurun_out/prof_vae_<tag>.ncu-rep (`scripts/gpu_ncu_r2.sh <tag> vae`): python scripts/vae_top_summary.py <tag>"""
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1]
raw = subprocess.run(["ncu", "-i", str(ROOT / "gpurun_out" / f"prof_vae_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, units = r[0], r[1]
keys = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = [f"# ncu `--set full` of the VAE encoder's convolutions, build `{tag}` (`scripts/gpu_ncu_r2.sh {tag} vae`: 6 launches of the second encode call of one 256-image chunk)\n",
       "Kernel: `tc_gemm_kernel<BN, PLAIN, pair, persistent>`; level-0 launches run the shared-tap-row layout (6 stages of 56 KB per tile).\n"]
for row in r[2:]:
    d, u = dict(zip(h, row)), dict(zip(h, units))
    out.append("| metric | value | unit |\n|---|---|---|")
    out += [f"| {k} | {d[k]} | {u.get(k, '')} |" for k in keys if k in d]
    out.append("")
(ROOT / "profiles" / f"{tag}_vae_top_kernel.md").write_text("\n".join(out))
print("wrote", ROOT / "profiles" / f"{tag}_vae_top_kernel.md")
