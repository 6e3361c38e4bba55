"""Generates tests/golden/ldp_golden.npz - golden input/output vectors for the LDP hot path.

PARITY UNPINNED: the reference (JAX 0.4.26 / Flax 0.8.4 / diffusers 0.27.2) cannot be imported in this image and
ships no tests or fixtures, so these vectors are produced by oracle/ldp_oracle.py (float64), not by the reference.
What they pin: (1) the oracle against silent regressions, (2) the CUDA path against a committed artefact that does
not depend on any RNG or library version at test time (weights and inputs are stored, not re-seeded).

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import params as P  # noqa: E402
from oracle import ldp_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "ldp_golden.npz"

UNET = dict(D=6, Dc=6, dims=(8, 16, 32), n_groups=8, step_embed=16)
IDM = dict(D=6, A=3, hidden=32, blocks=2, time_dim=16, cond=(16, 16))
VAE = dict(blocks=(32, 64), layers=1, groups=8, size=16)


def main():
    g = torch.Generator().manual_seed(1234)
    out = {}
    b, a, c = O.ddpm_schedule(100)
    out["sched/betas"], out["sched/alphas"], out["sched/acp"] = b, a, c
    sched = (b, a, c)

    # scheduler step / add_noise on fixed arrays
    x = torch.randn(3, 8, 6, generator=g, dtype=torch.float64)
    eps = torch.randn(3, 8, 6, generator=g, dtype=torch.float64)
    z = torch.randn(3, 8, 6, generator=g, dtype=torch.float64)
    out["step/x"], out["step/eps"], out["step/z"] = x.numpy(), eps.numpy(), z.numpy()
    for t in (0, 1, 50, 99):
        out[f"step/ddpm_t{t}"] = O.ddpm_step(sched, eps, t, x, z).numpy()
        out[f"step/ddim_t{t}"] = O.ddim_step(sched, eps, t, x).numpy()
    tt = np.array([0, 50, 99], dtype=np.int32)
    out["step/add_noise_t"] = tt
    out["step/add_noise"] = O.add_noise(sched, x, z, tt).numpy()

    # planner UNet (tiny widths so the fixture stays small); weights stored
    u = UNET
    up = P.init_params(P.unet_spec(u["D"], u["Dc"], u["dims"], 5, u["step_embed"]), seed=7, perturb=0.1)
    for k, v in up.items():
        out[f"unet/p/{k}"] = v
    ux = torch.randn(2, 8, u["D"], generator=g, dtype=torch.float64)
    uc = torch.rand(2, u["Dc"], generator=g, dtype=torch.float64) * 2 - 1
    out["unet/x"], out["unet/c"] = ux.numpy(), uc.numpy()
    kw = dict(down_dims=u["dims"], n_groups=u["n_groups"], step_embed_dim=u["step_embed"])
    for k in (0, 50, 99):
        out[f"unet/eps_k{k}"] = O.unet_forward(up, ux, k, uc, **kw).numpy()
    out["unet/eps_rows"] = O.unet_forward(up, ux, np.array([3, 77]), uc, **kw).numpy()
    ux16 = torch.randn(2, 16, u["D"], generator=g, dtype=torch.float64)
    out["unet/x16"] = ux16.numpy()
    out["unet/eps16_k50"] = O.unet_forward(up, ux16, 50, uc, **kw).numpy()
    zz = torch.randn(10, 2, 8, u["D"], generator=g, dtype=torch.float64)
    out["unet/loop_noise"] = zz.numpy()
    out["unet/loop_ddpm10"] = O.planner_sample(up, sched, ux, uc, zz, 10, **kw).numpy()
    out["unet/loop_ddim10"] = O.planner_sample(up, sched, ux, uc, None, 10, sampler="ddim", **kw).numpy()

    # IDM
    m = IDM
    ip = P.init_params(P.idm_spec(m["D"], m["A"], m["hidden"], m["blocks"], m["time_dim"], m["cond"]), seed=8, perturb=0.1)
    for k, v in ip.items():
        out[f"idm/p/{k}"] = v
    s = torch.randn(5, 2 * m["D"], generator=g, dtype=torch.float64)
    a_ = torch.randn(5, m["A"], generator=g, dtype=torch.float64)
    out["idm/s"], out["idm/a"] = s.numpy(), a_.numpy()
    for k in (0, 50, 99):
        out[f"idm/eps_k{k}"] = O.idm_forward(ip, s, a_, k, time_dim=m["time_dim"]).numpy()
    za = torch.randn(10, 5, m["A"], generator=g, dtype=torch.float64)
    out["idm/loop_noise"] = za.numpy()
    out["idm/loop_ddpm10"] = _idm_loop(ip, sched, s, a_, za, 10, m["time_dim"]).numpy()

    # VAE encoder
    v = VAE
    vp = P.init_params(P.vae_encoder_spec(v["blocks"], 3, 4, v["layers"]), seed=9, perturb=0.1)
    for k, val in vp.items():
        out[f"vae/p/{k}"] = val
    img = torch.randint(0, 256, (2, v["size"], v["size"], 3), generator=g, dtype=torch.int32).to(torch.uint8)
    out["vae/img_u8"] = img.numpy()
    xin = img.to(torch.float64) / 255.0 * 2 - 1
    out["vae/mean"] = O.vae_encode_mean(vp, xin, v["blocks"], v["layers"], v["groups"], 4).numpy()

    np.savez_compressed(OUT, **{k: (np.asarray(val, dtype=np.float32) if np.asarray(val).dtype == np.float64 and "/p/" in k
                                    else np.asarray(val)) for k, val in out.items()})
    print(OUT, f"{OUT.stat().st_size / 1e3:.0f} kB, {len(out)} arrays")


def _idm_loop(p, sched, s, a_T, noise, n, time_dim):
    a = a_T
    for i in range(n):
        k = n - 1 - i
        eps = O.idm_forward(p, s, a, k, time_dim=time_dim)
        a = O.ddpm_step(sched, eps, k, a, noise[i])
    return a


if __name__ == "__main__":
    main()
