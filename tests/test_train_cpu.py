"""Training row (N1) on CPU: the oracle's optimiser pieces against independent statements, the host schedule against
the oracle, and the data-parallel contract over a world_size-2 gloo group: the all-reduced sum of per-shard gradients
times 1/world equals the gradient of the global-batch mean (reference semantics, train_bc.py:73 + jnp.mean)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from latent_diffusion_planning_b200 import params as P
from latent_diffusion_planning_b200 import train as TR
from oracle import ldp_oracle as O


def test_schedule_known_answers():
    """warmup_cosine_decay_schedule(init=end_lr, peak=lr, 1000, 500k, end=end_lr) (agent/ldp_agent.py:580-587)."""
    for mk in (O.warmup_cosine_decay_schedule, TR.warmup_cosine_decay_schedule):
        s = mk(1e-6, 1e-4, 1000, 500000, 1e-6)
        assert s(0) == pytest.approx(1e-6, rel=1e-12)
        assert s(500) == pytest.approx(1e-6 + 0.5 * (1e-4 - 1e-6), rel=1e-12)          # linear warm-up
        assert s(1000) == pytest.approx(1e-4, rel=1e-12)                               # joined at the boundary
        mid = 1000 + (500000 - 1000) // 2
        assert s(mid) == pytest.approx(1e-6 + 0.5 * (1e-4 - 1e-6), rel=1e-6)           # cosine midpoint
        assert s(500000) == pytest.approx(1e-6, rel=1e-9) and s(10 ** 7) == pytest.approx(1e-6, rel=1e-9)
        assert all(s(i) >= s(i + 1) for i in range(1000, 500000, 9973))
    a, b = O.warmup_cosine_decay_schedule(1e-6, 1e-4, 1000, 500000, 1e-6), TR.warmup_cosine_decay_schedule(1e-6, 1e-4, 1000, 500000, 1e-6)
    assert all(a(i) == pytest.approx(b(i), rel=1e-14) for i in (0, 1, 999, 1000, 1001, 77777, 499999, 500001))


def test_adam_matches_torch_optim():
    """optax.adam's update (eps outside the root, bias-corrected moments) is the one torch.optim.Adam implements."""
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, generator=g, dtype=torch.float64)
    tw = torch.nn.Parameter(w.clone())
    opt = torch.optim.Adam([tw], lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    mu, nu = torch.zeros_like(w), torch.zeros_like(w)
    for i in range(5):
        gr = torch.randn(64, generator=g, dtype=torch.float64)
        tw.grad = gr.clone()
        opt.step()
        w, mu, nu = O.adam_update(w, gr, mu, nu, i + 1, 3e-3)
        assert float((w - tw.detach()).abs().max()) < 1e-14


def test_oracle_grads_match_finite_differences():
    """loss_and_grads (autograd on the restated forward) against central differences on a few IDM parameters."""
    spec = P.idm_spec(5, 3, 32, 1, 16, (16, 16))
    p = P.init_params(spec, seed=1, perturb=0.1)
    sched = O.ddpm_schedule(100)
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(3, 5, 5, generator=g, dtype=torch.float64)
    act = torch.randn(3, 5, 3, generator=g, dtype=torch.float64)
    t = torch.randint(0, 100, (12, 1), generator=g).numpy()
    noise = torch.randn(12, 3, generator=g, dtype=torch.float64)
    f = lambda q: O.idm_loss(q, sched, obs, act, 1, t, noise, time_dim=16)
    _, grads = O.loss_and_grads(f, p)
    for key, idx in (("MLPResNet_0/Dense_0/kernel", (2, 3)), ("MLP_0/Dense_0/bias", (4,)),
                     ("MLPResNet_0/MLPResNetBlock_0/LayerNorm_0/scale", (7,))):
        q = {k: torch.as_tensor(v, dtype=torch.float64).clone() for k, v in p.items()}
        h = 1e-6
        q[key][idx] += h
        up = float(f(q))
        q[key][idx] -= 2 * h
        dn = float(f(q))
        assert (up - dn) / (2 * h) == pytest.approx(float(grads[key][idx]), rel=1e-5, abs=1e-9)


def test_idm_pairs_layout():
    obs = torch.arange(2 * 4 * 3, dtype=torch.float64).reshape(2, 4, 3)
    act = torch.arange(2 * 4 * 2, dtype=torch.float64).reshape(2, 4, 2)
    ssp, a = O.idm_pairs(obs, act, 1)
    assert ssp.shape == (6, 6) and a.shape == (6, 2)
    assert torch.equal(ssp[0], torch.cat([obs[0, 0], obs[0, 1]])) and torch.equal(ssp[5], torch.cat([obs[1, 2], obs[1, 3]]))
    assert torch.equal(a[4], act[1, 1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    spec = P.idm_spec(4, 2, 32, 1, 16, (16, 16))
    p = P.init_params(spec, seed=3, perturb=0.1)
    g = torch.Generator().manual_seed(3)
    obs = torch.randn(4, 3, 4, generator=g, dtype=torch.float64)
    act = torch.randn(4, 3, 2, generator=g, dtype=torch.float64)
    t = torch.randint(0, 100, (8, 1), generator=g).numpy()
    noise = torch.randn(8, 2, generator=g, dtype=torch.float64)
    return spec, p, obs, act, t, noise


def _flat(spec, grads):
    return torch.cat([grads[k].reshape(-1) for k in spec])


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec, p, obs, act, t, noise = _problem()
        sched = O.ddpm_schedule(100)
        B = obs.shape[0] // world
        rows = slice(rank * B, (rank + 1) * B)
        n = B * 2                                        # transitions per rank (H = 2)
        f = lambda qq: O.idm_loss(qq, sched, obs[rows], act[rows], 1, t[rank * n:(rank + 1) * n], noise[rank * n:(rank + 1) * n], time_dim=16)
        _, grads = O.loss_and_grads(f, p)
        flat = _flat(spec, grads).clone()
        scale = TR.allreduce_grads(flat)                 # the product's exchange step, on the gloo group
        q.put((rank, (flat * scale).numpy(), scale))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_equals_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    spec, p, obs, act, t, noise = _problem()
    sched = O.ddpm_schedule(100)
    _, grads = O.loss_and_grads(lambda qq: O.idm_loss(qq, sched, obs, act, 1, t, noise, time_dim=16), p)
    full = _flat(spec, grads).numpy()
    for _, flat, scale in res:
        assert scale == 0.5
        assert np.abs(flat - full).max() < 1e-12 * max(1.0, np.abs(full).max())


def test_allreduce_is_identity_without_a_group():
    g = torch.ones(5)
    assert TR.allreduce_grads(g) == 1.0 and torch.equal(g, torch.ones(5))


def test_unflatten_roundtrip():
    spec = P.idm_spec(4, 2, 32, 1, 16, (16, 16))
    p = P.init_params(spec, seed=0)
    back = TR.unflatten_params(spec, P.flatten_params(spec, p))
    assert all(np.array_equal(back[k], p[k]) for k in spec)


def test_update_gating_follows_the_reference_rules():
    """LDPAgent._gates == the boolean algebra of `update` (agent/ldp_agent.py:229-236), checked without a GPU handle."""
    from latent_diffusion_planning_b200.agent import LDPAgent
    ag = LDPAgent.__new__(LDPAgent)
    ag.use_planner, ag.use_idm = True, True
    ag.config = dict(update_planner_every=2, update_idm_every=3, update_idm_after=4, update_planner_until=9, update_planner_after=2)

    def ref(step, c):
        use_planner = step % c["update_planner_every"] == 0
        use_idm = step % c["update_idm_every"] == 0 and step >= c["update_idm_after"]
        upd = (c["update_planner_until"] < 0 or step < c["update_planner_until"]) and step >= c["update_planner_after"]
        return use_planner and upd, use_idm
    for step in range(14):
        assert ag._gates(step) == ref(step, ag.config)
    assert ag._gates(0) == (False, False) and ag._gates(2) == (True, False) and ag._gates(6) == (True, True) and ag._gates(10) == (False, False)
    ag.config.update(update_planner_until=-1, update_planner_after=-1, update_idm_after=-1, update_planner_every=1, update_idm_every=1)
    assert all(ag._gates(s) == (True, True) for s in range(5))          # the yaml defaults (agent/ldp_agent.yaml:56-60)
    ag.use_idm = False
    assert ag._gates(3) == (True, False)
