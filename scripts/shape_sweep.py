"""Odd batch sizes through every entry point (finite outputs, no launch errors): diagnostics for the chunk / tile edge cases."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

ok = True


def check(name, fn):
    global ok
    try:
        out = fn()
        fin = bool(torch.isfinite(out).all())
        print(f"{name}: shape {tuple(out.shape)} finite={fin}")
        ok &= fin
    except Exception as e:  # noqa: BLE001
        ok = False
        print(f"{name}: FAILED {type(e).__name__}: {str(e)[:200]}")


D, T = 265, 8
pl = H.Planner(P.init_params(P.unet_spec(D, D), seed=0), D, D)
g = torch.Generator().manual_seed(0)
for B in (1, 3, 17, 100, 129, 1000, 1025):
    x = torch.randn(B, T, D, generator=g).cuda()
    c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
    check(f"planner B={B}", lambda: pl.sample(x, c, seed=1, n_steps=3, precision="bf16"))
blocks = (128, 256, 512, 512)
enc = H.VaeEncoder(P.init_params(P.vae_encoder_spec(blocks), seed=0), blocks)
for B in (1, 2, 5, 37, 591, 593, 1185):
    img = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.int32).to(torch.uint8).cuda()
    check(f"vae encode B={B}", lambda: enc.encode(img, precision="bf16"))
dec = H.VaeDecoder(P.init_params(P.vae_decoder_spec(blocks), seed=0), blocks)
for B in (1, 3, 37, 295, 297, 300, 593):
    z = torch.randn(B, 8, 8, 4, generator=g).cuda()
    check(f"vae decode B={B}", lambda: dec.decode(z, precision="bf16"))
A, Ha = 7, 4
idm = H.Idm(P.init_params(P.idm_spec(D, A), seed=0), D, A)      # s rows are [z_t, z_{t+1}]: 2 x obs_dim columns
for N in (1, 5, 127, 129, 1000, 4097):
    s = torch.randn(N, 2 * D, generator=g).cuda()
    a = torch.randn(N, A, generator=g).cuda()
    check(f"idm N={N}", lambda: idm.sample(s, a, seed=1, n_steps=3, precision="bf16"))
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
