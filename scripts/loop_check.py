"""Bring-up check: persistent loop kernel vs the per-layer CUDA-graph path on identical inputs."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

D = 265
p = P.init_params(P.unet_spec(D, D), seed=0)


def run(loop, B, T, n, flags=0):
    os.environ["LDP_LOOP"] = "1" if loop else "0"
    os.environ["LDP_LOOP_FLAGS"] = str(flags)
    pl = H.Planner(p, D, D)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, D, generator=g).cuda()
    c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
    z = torch.randn(n, B, T, D, generator=g).cuda()
    out = pl.sample(x, c, noise=z, n_steps=n, precision="bf16")
    torch.cuda.synchronize()
    return out


CASES = [(8, 8, 1), (16, 8, 1), (32, 8, 1), (48, 8, 1), (63, 8, 1), (64, 8, 1), (65, 8, 1), (100, 8, 1), (5, 16, 1), (40, 16, 2)]
for B, T, n in CASES:
    ref = run(False, B, T, n)
    for flags in (0,):
        out = run(True, B, T, n, flags)
        d = (out - ref).abs()
        bad = (d > 1e-3)
        rows = bad.any(dim=2).nonzero()
        dt = d.amax(dim=(0, 2)).tolist()
        print("   max err by t:", ["%.1e" % v for v in dt], " by channel block of 32:", ["%.0e" % float(d[:, :, i:i + 32].max()) for i in range(0, D, 32)])
        print(f"B={B} T={T} n={n} flags={flags}: max err {float(d.max()):.3e}, bad elems {int(bad.sum())} of {d.numel()}, "
              f"bad samples {sorted(set(rows[:, 0].tolist()))[:12]} bad cols {sorted(set(bad.any(dim=0).any(dim=0).nonzero().flatten().tolist()))[:12]}")
