"""Persistent IDM loop kernel (csrc/idm_loop.cu: the whole reverse-diffusion loop of reference agent/ldp_agent.py:492-503 in one
launch) against (a) the float64 oracle with identical injected noise and (b) the per-layer CUDA path (LDP_IDM_LOOP=0)."""
import os

import numpy as np
import pytest
import torch

from oracle import ldp_oracle as O
from latent_diffusion_planning_b200 import params as P

pytestmark = pytest.mark.gpu


def _case(D, A, n, seed):
    g = torch.Generator().manual_seed(seed)
    p = P.init_params(P.idm_spec(D, A), seed=seed, perturb=0.1)
    s = torch.rand(n, 2 * D, generator=g) * 2 - 1
    a = torch.randn(n, A, generator=g)
    return p, s, a, g


def _per_layer(fn):
    os.environ["LDP_IDM_LOOP"] = "0"
    try:
        return fn()
    finally:
        os.environ.pop("LDP_IDM_LOOP", None)


@pytest.mark.parametrize("D,A,n,steps,sampler", [(265, 7, 5, 3, "ddpm"), (265, 7, 128, 4, "ddpm"), (270, 14, 300, 3, "ddim"),
                                                 (25, 7, 4096, 2, "ddpm"), (265, 7, 129, 100, "ddpm")])
def test_loop_kernel_matches_oracle_and_per_layer_path(cuda, D, A, n, steps, sampler):
    from latent_diffusion_planning_b200 import handles as H
    p, s, a, g = _case(D, A, n, seed=3)
    idm = H.Idm(p, D, A)
    z = torch.randn(steps, n, A, generator=g)
    got = idm.sample(s.cuda(), a.cuda(), noise=z.cuda(), n_steps=steps, sampler=sampler, precision="bf16")
    ref_layers = _per_layer(lambda: idm.sample(s.cuda(), a.cuda(), noise=z.cuda(), n_steps=steps, sampler=sampler, precision="bf16"))
    assert torch.isfinite(got).all()
    if steps <= 4:                                       # late steps (k <= 3) are contractive: per-step tolerance applies end to end
        with torch.no_grad():
            ref = O.idm_sample(p, O.ddpm_schedule(100), s, a, z, steps, sampler=sampler)
        err = float((got.cpu().double() - ref).abs().max())
        assert err < 1e-2 * max(1.0, float(ref.abs().max())), err
        assert float((got - ref_layers).abs().max()) < 5e-3
    else:                                                # 100 chained bf16 steps: same distribution, both paths bounded by the clip
        assert float(got.abs().max()) <= 1.0 + 1e-4 and float(ref_layers.abs().max()) <= 1.0 + 1e-4
        assert abs(float(got.mean()) - float(ref_layers.mean())) < 0.05 and abs(float(got.std()) - float(ref_layers.std())) < 0.05
    idm.close()


def test_loop_kernel_philox_noise_is_the_per_layer_path_noise_and_shard_invariant(cuda):
    from latent_diffusion_planning_b200 import handles as H
    D, A, n = 265, 7, 200
    p, s, a, _ = _case(D, A, n, seed=4)
    idm = H.Idm(p, D, A)
    run = lambda lo, hi: idm.sample(s[lo:hi].cuda(), a[lo:hi].cuda(), seed=11, row_offset=lo, n_steps=3, precision="bf16")
    full = run(0, n)
    assert torch.equal(torch.cat([run(0, 77), run(77, n)]), full)                     # rows are independent, noise keyed by global row
    layers = _per_layer(lambda: run(0, n))
    assert float((full - layers).abs().max()) < 5e-3                                  # same Philox draws in both paths
    idm.close()
