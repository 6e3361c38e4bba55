"""Generates tests/golden/ldp_golden_train.npz - golden vectors for the training row (N1): losses, parameter gradients
and two optax.adam steps of the planner and IDM objectives on small configurations.

PARITY UNPINNED, like the other fixtures: produced by oracle/ldp_oracle.py (float64 autograd on the restated forward),
not by the reference (jax / optax cannot be imported here).  Weights and inputs are stored, not re-seeded.

    python tests/golden/make_golden_train.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from latent_diffusion_planning_b200 import params as P  # noqa: E402
from oracle import ldp_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent / "ldp_golden_train.npz"
UNET = dict(D=6, dims=(8, 16, 32), n_groups=8, step_embed=16, B=3, T=8)
IDM = dict(D=5, A=3, hidden=32, blocks=2, time_dim=16, cond=(16, 16), B=3, H=5)
OPT = dict(init=1e-5, peak=1e-3, warmup=2, decay=10, end=1e-5)


def main():
    g = torch.Generator().manual_seed(2024)
    sched = O.ddpm_schedule(100)
    out = {}
    u = UNET
    up = P.init_params(P.unet_spec(u["D"], u["D"], u["dims"], 5, u["step_embed"]), seed=21, perturb=0.1)
    obs = torch.randn(u["B"], u["T"] + 1, u["D"], generator=g, dtype=torch.float64) * 0.7
    tp = torch.randint(0, 100, (u["B"],), generator=g).numpy().astype(np.int32)
    zp = torch.randn(u["B"], u["T"], u["D"], generator=g, dtype=torch.float64)
    kw = dict(down_dims=u["dims"], n_groups=u["n_groups"], step_embed_dim=u["step_embed"])
    f = lambda q: O.planner_loss(q, sched, obs, 1, tp, zp, **kw)
    loss, grads = O.loss_and_grads(f, up)
    out.update({f"unet/p/{k}": v.astype(np.float32) for k, v in up.items()})
    out.update({f"unet/g/{k}": v.numpy() for k, v in grads.items()})
    out.update({"unet/obs_emb": obs.numpy(), "unet/t": tp, "unet/noise": zp.numpy(), "unet/loss": np.asarray(float(loss))})

    i = IDM
    ip = P.init_params(P.idm_spec(i["D"], i["A"], i["hidden"], i["blocks"], i["time_dim"], i["cond"]), seed=22, perturb=0.1)
    emb = torch.randn(i["B"], i["H"], i["D"], generator=g, dtype=torch.float64)
    act = torch.randn(i["B"], i["H"], i["A"], generator=g, dtype=torch.float64)
    n = i["B"] * (i["H"] - 1)
    ti = torch.randint(0, 100, (n, 1), generator=g).numpy().astype(np.int32)
    zi = torch.randn(n, i["A"], generator=g, dtype=torch.float64)
    fi = lambda q: O.idm_loss(q, sched, emb, act, 1, ti, zi, time_dim=i["time_dim"])
    loss_i, grads_i = O.loss_and_grads(fi, ip)
    out.update({f"idm/p/{k}": v.astype(np.float32) for k, v in ip.items()})
    out.update({f"idm/g/{k}": v.numpy() for k, v in grads_i.items()})
    out.update({"idm/obs_emb": emb.numpy(), "idm/actions": act.numpy(), "idm/t": ti, "idm/noise": zi.numpy(),
                "idm/loss": np.asarray(float(loss_i))})

    # two Adam steps of the IDM objective on the same batch with the warm-up schedule (lr at the pre-increment count)
    sc = O.warmup_cosine_decay_schedule(OPT["init"], OPT["peak"], OPT["warmup"], OPT["decay"], OPT["end"])
    q = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in ip.items()}
    mu = {k: torch.zeros_like(v) for k, v in q.items()}
    nu = {k: torch.zeros_like(v) for k, v in q.items()}
    lrs = []
    for step in range(2):
        _, gq = O.loss_and_grads(fi, q)
        lrs.append(sc(step))
        for k in q:
            q[k], mu[k], nu[k] = O.adam_update(q[k], gq[k], mu[k], nu[k], step + 1, sc(step))
    out.update({f"idm/p2/{k}": v.numpy() for k, v in q.items()})
    out["idm/lrs"] = np.asarray(lrs)
    np.savez_compressed(OUT, **out)
    print(OUT, f"{OUT.stat().st_size / 1e3:.0f} kB, {len(out)} arrays, losses {float(loss):.6f} {float(loss_i):.6f}")


if __name__ == "__main__":
    main()
