"""The persistent planner loop kernel (csrc/planner_loop.cu, opt-in with LDP_LOOP=1) against the per-layer CUDA-graph
path on identical inputs.  Both run the same tile arithmetic in the same order, so the results must be bit-identical;
the batch sizes cover full groups, partial groups (CTAs that own no tile in some layers still have to take part in the
group barriers), several tiles per CTA, and the T = 16 level lengths."""
import os

import pytest
import torch

from latent_diffusion_planning_b200 import params as P

pytestmark = pytest.mark.gpu


def _run(H, p, D, loop, B, T, n, sampler="ddpm", seed=None):
    old = os.environ.get("LDP_LOOP")
    os.environ["LDP_LOOP"] = "1" if loop else "0"       # read when the (B, T) workspace prepares its loop table
    try:
        planner = H.Planner(p, D, D)
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, T, D, generator=g).cuda()
        c = (torch.rand(B, D, generator=g) * 2 - 1).cuda()
        if seed is None:
            z = torch.randn(n, B, T, D, generator=g).cuda()
            out = planner.sample(x, c, noise=z, n_steps=n, sampler=sampler, precision="bf16")
        else:
            out = planner.sample(x, c, seed=seed, row_offset=5, n_steps=n, sampler=sampler, precision="bf16")
        torch.cuda.synchronize()
        return out
    finally:
        if old is None:
            os.environ.pop("LDP_LOOP", None)
        else:
            os.environ["LDP_LOOP"] = old


@pytest.fixture(scope="module")
def H(cuda):
    from latent_diffusion_planning_b200 import handles
    return handles


@pytest.fixture(scope="module")
def params265():
    return P.init_params(P.unet_spec(265, 265), seed=0)


@pytest.mark.parametrize("B,T,n", [(8, 8, 2), (48, 8, 1), (65, 8, 2), (100, 8, 1), (256, 8, 3)])
def test_loop_equals_graph_path(H, params265, B, T, n):
    ref = _run(H, params265, 265, False, B, T, n)
    out = _run(H, params265, 265, True, B, T, n)
    assert torch.equal(out, ref)


def test_loop_equals_graph_path_philox_and_ddim(H, params265):
    for sampler in ("ddpm", "ddim"):
        ref = _run(H, params265, 265, False, 70, 8, 3, sampler=sampler, seed=11)
        out = _run(H, params265, 265, True, 70, 8, 3, sampler=sampler, seed=11)
        assert torch.equal(out, ref)


def test_loop_equals_graph_path_t16(H):
    D = 270
    p = P.init_params(P.unet_spec(D, D), seed=3)
    ref = _run(H, p, D, False, 40, 16, 2)
    out = _run(H, p, D, True, 40, 16, 2)
    assert torch.equal(out, ref)


def test_loop_benchmark_shape_finite(H, params265):
    out = _run(H, params265, 265, True, 1024, 8, 4, seed=1)
    ref = _run(H, params265, 265, False, 1024, 8, 4, seed=1)
    assert torch.isfinite(out).all() and torch.equal(out, ref)
