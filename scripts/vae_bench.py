"""VAE-encode throughput at BASELINE config #3: stable_vae_model.encode on 64x64x3 agentview images, B=4096
(process_sdvae_data.py path).  Prints one JSON line; diagnostics, the driver's bench is bench.py."""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import _native, handles as H, params as P  # noqa: E402

B = int(os.environ.get("LDP_B", "4096"))
blocks = (128, 256, 512, 512) if os.environ.get("LDP_VAE", "bench") == "bench" else (128, 256, 256, 256, 256, 256)
gf_per_img = 16.92 if len(blocks) == 4 else 11.48
p = P.init_params(P.vae_encoder_spec(blocks), seed=0)
vae = H.VaeEncoder(p, blocks)
g = torch.Generator().manual_seed(4)
img = torch.randint(0, 256, (B, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8).cuda()
lib = _native.load()
for _ in range(2):
    out = vae.encode(img, lat_min=-10.0, lat_max=10.0, precision="bf16")
torch.cuda.synchronize()
lib.ldp_launch_count_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    out = vae.encode(img, lat_min=-10.0, lat_max=10.0, precision="bf16")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
tf = B * gf_per_img / ms / 1e3 * 1e0
print(json.dumps({"metric": "vae_encode_imgs_per_sec", "value": B / ms * 1e3, "unit": "img/s", "B": B, "blocks": list(blocks),
                  "ms_per_batch": ms, "tflops_useful": B * gf_per_img / ms, "frac_of_sustained_peak":
                  B * gf_per_img / ms / float(peaks.get("bf16_tflops_sustained", 1397.0)),
                  "gpu_launches": int(lib.ldp_launch_count()) // reps, "finite": bool(torch.isfinite(out).all())}))
