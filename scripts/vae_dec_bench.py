"""VAE decoder throughput (plan_viz path): SD-VAE [128,256,512,512], 8x8x4 -> 64x64x3, bf16."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
blocks = (128, 256, 512, 512)
dec = H.VaeDecoder(P.init_params(P.vae_decoder_spec(blocks), seed=0), blocks)
z = torch.randn(B, 8, 8, 4, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(2):
    out = dec.decode(z, precision="bf16")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = dec.decode(z, precision="bf16")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"vae decode B={B}: {ms:.1f} ms, {B / ms * 1e3:.0f} frames/s, {B * 38.8 / ms:.0f} TF/s useful (38.8 GF/frame), finite={bool(torch.isfinite(out).all())}")
