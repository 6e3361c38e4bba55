"""BASELINE config #5 shapes: planner loop at B=512, T=16, D=270 (aloha), 100 DDIM steps, bf16; per-op table."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

B, T, D = 512, 16, 270
p = P.init_params(P.unet_spec(D, D), seed=0)
pl = H.Planner(p, D, D)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, D, generator=g).cuda()
c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
for _ in range(2):
    pl.sample(x, c, seed=1, n_steps=100, sampler="ddim", precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3):
    pl.sample(x, c, seed=i, n_steps=100, sampler="ddim", precision="bf16")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"aloha planner loop B={B} T={T} D={D}: {ms:.2f} ms / 100 DDIM steps, {B / ms * 1e3:.0f} plans/s")
ops = pl.profile_step(B, T, reps=20)
for i, o in enumerate(ops):
    print(f"  {i:2d} {o['epilogue']:5s} M={o['M']:5d} N={o['N']:4d} K={o['K']:5d} bn={o['block_n']:3d} acc={o['n_acc']}+{o['aux']} {o['us']:7.2f} us")
print("sum", sum(o["us"] for o in ops))
