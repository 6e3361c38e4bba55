"""Per-layer clock stamps of the persistent planner loop kernel (CTA 0, one iteration): LDP_LOOP_DBG=<iteration>."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

import os
B, T, D = int(os.environ.get("LDP_B", "1024")), 8, 265
p = P.init_params(P.unet_spec(D, D), seed=0)
pl = H.Planner(p, D, D)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, D, generator=g).cuda()
c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
pl.sample(x, c, seed=1, n_steps=100, precision="bf16")
torch.cuda.synchronize()
