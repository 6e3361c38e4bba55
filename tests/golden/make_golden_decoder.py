"""Generates tests/golden/ldp_golden_decoder.npz - golden vectors for the VAE decoder (scope row N2, plan_viz).

PARITY UNPINNED, like ldp_golden.npz: produced by oracle/ldp_oracle.py:vae_decode (float64), not by the reference
(diffusers 0.27.2 FlaxAutoencoderKL.decode cannot be imported here).  Weights and inputs are stored, not re-seeded.

    python tests/golden/make_golden_decoder.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import params as P  # noqa: E402
from oracle import ldp_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "ldp_golden_decoder.npz"
DEC = dict(blocks=(32, 64), layers=1, groups=8, size=16)


def main():
    g = torch.Generator().manual_seed(4321)
    d = DEC
    out = {}
    dp = P.init_params(P.vae_decoder_spec(d["blocks"], 3, 4, d["layers"]), seed=11, perturb=0.1)
    for k, v in dp.items():
        out[f"dec/p/{k}"] = v.astype(np.float32)
    hw = d["size"] >> (len(d["blocks"]) - 1)
    z = torch.randn(2, hw, hw, 4, generator=g, dtype=torch.float64)
    out["dec/z"] = z.numpy()
    out["dec/sample"] = O.vae_decode(dp, z, d["blocks"], d["layers"], d["groups"]).numpy()
    np.savez_compressed(OUT, **out)
    print(OUT, f"{OUT.stat().st_size / 1e3:.0f} kB, {len(out)} arrays")


if __name__ == "__main__":
    main()
