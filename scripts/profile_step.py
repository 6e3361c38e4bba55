"""Short planner run for ncu: a few bf16 denoising steps at the benchmark shape (B=1024, T=8, D=265)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

B = int(os.environ.get("LDP_B", "1024"))
STEPS = int(os.environ.get("LDP_STEPS", "3"))
D = 265
p = P.init_params(P.unet_spec(D, D), seed=0)
pl = H.Planner(p, D, D)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 8, D, generator=g).cuda()
c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
for _ in range(int(os.environ.get("LDP_REPS", "2"))):
    pl.sample(x, c, seed=1, n_steps=STEPS, precision="bf16")
torch.cuda.synchronize()
print("done")
