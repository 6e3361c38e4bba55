"""Short VAE-encode run for ncu: one 256-image chunk of the benchmark topology (bf16 path)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import handles as H, params as P  # noqa: E402

B = int(os.environ.get("LDP_B", "256"))
blocks = (128, 256, 512, 512)
p = P.init_params(P.vae_encoder_spec(blocks), seed=0)
vae = H.VaeEncoder(p, blocks)
img = torch.randint(0, 256, (B, 64, 64, 3), dtype=torch.int32).to(torch.uint8).cuda()
for _ in range(int(os.environ.get("LDP_REPS", "2"))):
    out = vae.encode(img, lat_min=-10.0, lat_max=10.0, precision="bf16")
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
