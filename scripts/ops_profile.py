"""Per-kernel timing of one bf16 denoising step + whole-loop timing at the benchmark shape (diagnostics).

    python scripts/ops_profile.py [tag]      # honours LDP_BN64 / LDP_NO_PDL / LDP_NO_GRAPH / LDP_B
Writes gpurun_out/ops_<tag>.json and prints a table.
"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "default"
B = int(os.environ.get("LDP_B", "1024"))
T, D = 8, 265
p = P.init_params(P.unet_spec(D, D), seed=0)
pl = H.Planner(p, D, D)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, D, generator=g).cuda()
c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
for _ in range(2):
    pl.sample(x, c, seed=1, n_steps=100, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for i in range(reps):
    pl.sample(x, c, seed=i, n_steps=100, precision="bf16")
e1.record()
torch.cuda.synchronize()
loop_ms = e0.elapsed_time(e1) / reps
ops = pl.profile_step(B, T, reps=30)
tot = sum(o["us"] for o in ops)
print(f"[{tag}] B={B} loop {loop_ms:.2f} ms / 100 steps = {loop_ms * 10:.1f} us/step; sum of isolated kernels {tot:.1f} us; "
      f"{B / loop_ms * 1e3:.0f} plans/s")
for i, o in enumerate(ops):
    ctas = -(-o["M"] // 128) * -(-o["N"] // o["block_n"])
    fl = 2.0 * o["M"] * o["N"] * o["K"]
    print(f"  {i:2d} {o['epilogue']:5s} M={o['M']:5d} N={o['N']:4d} K={o['K']:5d} bn={o['block_n']:3d} acc={o['n_acc']}+{o['aux']} ctas={ctas:3d} "
          f"{o['us']:7.2f} us  {fl / o['us'] / 1e6:7.1f} TF/s(padded)  cyc: " + " ".join(f"{v:6.0f}" for v in o["phases"][1:]))
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
(out / f"ops_{tag}.json").write_text(json.dumps({"tag": tag, "B": B, "loop_ms": loop_ms, "sum_isolated_us": tot, "ops": ops}))
