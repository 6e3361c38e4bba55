"""Training row (N1) on the GPU: gradients of the planner / IDM denoising losses from the CUDA path (through the C ABI)
against torch autograd on the float64 oracle restatement with identical timesteps and noise; the Adam kernel against
the oracle's optax.adam; LDPAgent.update end to end.  Tolerance: fp32 path, 5e-4 of each tensor's largest gradient
(+2e-6 of the largest gradient overall) - accumulation order differs (atomics), arithmetic does not."""
import numpy as np
import pytest
import torch

from oracle import ldp_oracle as O

pytestmark = pytest.mark.gpu


def _check_grads(got, ref, rtol=5e-4):
    """Per tensor: err <= rtol * max|g_tensor| + 2e-6 * max|g_any|.  The absolute floor covers gradients that are
    analytically zero (a conv bias feeding GroupNorm is cancelled by the mean) and come out as fp32 rounding noise."""
    gmax = max(float(r.abs().max()) for r in ref.values())
    for k, r in ref.items():
        r = r.numpy()
        scale = np.abs(r).max()
        err = np.abs(got[k].astype(np.float64) - r).max()
        assert err <= rtol * scale + 2e-6 * gmax, f"{k}: err {err:.3e} vs max|g| {scale:.3e} (global {gmax:.3e})"


def _planner_case(seed, D, dims, B, T, ds=16, groups=8):
    from latent_diffusion_planning_b200 import params as P
    spec = P.unet_spec(D, D, dims, 5, ds)
    p = P.init_params(spec, seed=seed, perturb=0.1)
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(B, T + 1, D, generator=g, dtype=torch.float64) * 0.7
    t = torch.randint(0, 100, (B,), generator=g)
    noise = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    return spec, p, obs, t, noise


# (GroupNorm groups of only 4 elements - e.g. dims (8,16) at T=4 - are ill-conditioned in float32: the oracle run in
#  float32 deviates from its float64 self by 1.6e-3 there, against 8e-5 on these cases, so they are not used as a gate)
@pytest.mark.parametrize("D,dims,B,T", [(6, (8, 16, 32), 5, 8), (6, (16, 32), 3, 4), (25, (64, 128, 256), 9, 8), (10, (16,), 2, 8),
                                        (12, (32, 64, 128), 3, 16)])
def test_planner_loss_and_grads_match_oracle(cuda, D, dims, B, T):
    from latent_diffusion_planning_b200 import _native as N, train as TR
    spec, p, obs, t, noise = _planner_case(0, D, dims, B, T)
    sched = O.ddpm_schedule(100)
    kw = dict(down_dims=dims, n_groups=8, step_embed_dim=16)
    loss, grads = O.loss_and_grads(lambda q: O.planner_loss(q, sched, obs, 1, t.numpy(), noise, **kw), p)
    ts = TR.TrainState("planner", spec, N.unet_config(D, D, dims, 16, 5, 8, 100), p, lambda c: 1e-3)
    for rep in range(2):                                   # second pass reuses the workspace: same answer
        ts.zero_grad()
        got_loss = ts.planner_loss_grad(obs[:, 1:].float().cuda(), noise.float().cuda(), t.cuda(), obs[:, 0].float().cuda())
        assert float(got_loss) == pytest.approx(float(loss), rel=2e-5)
        _check_grads(ts.grads_dict(), grads)


def _check_grads_bf16(got, ref, tol=2e-2):
    """bf16 operands, fp32 accumulation: relative L2 error per tensor (north_star's 1e-2 bf16 bar is per forward value;
    a gradient passes through ~3x as many bf16 contractions) with the same absolute floor as the fp32 check."""
    gmax = max(float(r.abs().max()) for r in ref.values())
    for k, r in ref.items():
        r = r.numpy().ravel()
        d = got[k].astype(np.float64).ravel() - r
        assert np.sqrt((d * d).mean()) <= tol * np.sqrt((r * r).mean()) + 2e-5 * gmax, f"{k}: rel L2 {np.linalg.norm(d) / max(np.linalg.norm(r), 1e-30):.3e}"


@pytest.mark.parametrize("D,dims,B,T", [(25, (64, 128, 256), 9, 8), (12, (32, 64, 128), 20, 16), (265, (64, 128), 33, 8)])
def test_planner_grads_bf16_tensor_core_path(cuda, D, dims, B, T):
    from latent_diffusion_planning_b200 import _native as N, train as TR
    spec, p, obs, t, noise = _planner_case(3, D, dims, B, T)
    sched = O.ddpm_schedule(100)
    kw = dict(down_dims=dims, n_groups=8, step_embed_dim=16)
    loss, grads = O.loss_and_grads(lambda q: O.planner_loss(q, sched, obs, 1, t.numpy(), noise, **kw), p)
    ts = TR.TrainState("planner", spec, N.unet_config(D, D, dims, 16, 5, 8, 100), p, lambda c: 1e-3, precision="bf16")
    for rep in range(2):
        ts.zero_grad()
        got_loss = ts.planner_loss_grad(obs[:, 1:].float().cuda(), noise.float().cuda(), t.cuda(), obs[:, 0].float().cuda())
        assert float(got_loss) == pytest.approx(float(loss), rel=1e-2)
        _check_grads_bf16(ts.grads_dict(), grads)


def test_idm_grads_bf16_tensor_core_path(cuda):
    from latent_diffusion_planning_b200 import _native as NV, params as P, train as TR
    D, A, H, blocks, N = 25, 7, 256, 3, 300
    spec = P.idm_spec(D, A, H, blocks, 16, (32, 32))
    p = P.init_params(spec, seed=1, perturb=0.1)
    g = torch.Generator().manual_seed(1)
    s = torch.randn(N, 2 * D, generator=g, dtype=torch.float64)
    a0 = torch.randn(N, A, generator=g, dtype=torch.float64)
    t = torch.randint(0, 100, (N, 1), generator=g)
    noise = torch.randn(N, A, generator=g, dtype=torch.float64)
    sched = O.ddpm_schedule(100)

    def f(q):
        noisy = O.add_noise(sched, a0, noise, t.numpy())
        return ((O.idm_forward(q, s, noisy, t.numpy().reshape(-1), 16) - noise) ** 2).mean()
    loss, grads = O.loss_and_grads(f, p)
    ts = TR.TrainState("idm", spec, NV.idm_config(D, A, H, blocks, 16, (32, 32), 100), p, lambda c: 1e-3, precision="bf16")
    ts.zero_grad()
    got = ts.idm_loss_grad(s.float().cuda(), a0.float().cuda(), noise.float().cuda(), t.cuda())
    assert float(got) == pytest.approx(float(loss), rel=1e-2)
    # the first Dense sees the longest bf16 backward chain (3 blocks x {Dense, Dense, LayerNorm}): 3.2e-2 measured
    _check_grads_bf16(ts.grads_dict(), grads, tol=5e-2)


def test_planner_loss_weight_and_shard_sum(cuda):
    """grads ADD into the buffer and scale with loss_weight: two half-batches at weight 1/2 == the full batch."""
    from latent_diffusion_planning_b200 import _native as N, train as TR
    D, dims, B, T = 6, (8, 16, 32), 6, 8
    spec, p, obs, t, noise = _planner_case(2, D, dims, B, T)
    ts = TR.TrainState("planner", spec, N.unet_config(D, D, dims, 16, 5, 8, 100), p, lambda c: 1e-3)
    args = lambda s: (obs[s, 1:].float().cuda(), noise[s].float().cuda(), t[s].cuda(), obs[s, 0].float().cuda())
    ts.zero_grad()
    ts.planner_loss_grad(*args(slice(0, B)))
    full = ts.grads.clone()
    ts.zero_grad()
    l0 = ts.planner_loss_grad(*args(slice(0, 3)), weight=0.5)
    l1 = ts.planner_loss_grad(*args(slice(3, 6)), weight=0.5)
    assert float((ts.grads - full).abs().max()) <= 1e-5 * float(full.abs().max())
    assert float(l0) > 0 and float(l1) > 0


@pytest.mark.parametrize("D,A,H,blocks,N", [(5, 3, 32, 2, 12), (25, 7, 256, 3, 40), (30, 14, 64, 1, 7)])
def test_idm_loss_and_grads_match_oracle(cuda, D, A, H, blocks, N):
    from latent_diffusion_planning_b200 import _native as NV, params as P, train as TR
    spec = P.idm_spec(D, A, H, blocks, 16, (32, 32))
    p = P.init_params(spec, seed=1, perturb=0.1)
    g = torch.Generator().manual_seed(1)
    s = torch.randn(N, 2 * D, generator=g, dtype=torch.float64)
    a0 = torch.randn(N, A, generator=g, dtype=torch.float64)
    t = torch.randint(0, 100, (N, 1), generator=g)
    noise = torch.randn(N, A, generator=g, dtype=torch.float64)
    sched = O.ddpm_schedule(100)

    def f(q):
        noisy = O.add_noise(sched, a0, noise, t.numpy())
        return ((O.idm_forward(q, s, noisy, t.numpy().reshape(-1), 16) - noise) ** 2).mean()
    loss, grads = O.loss_and_grads(f, p)
    ts = TR.TrainState("idm", spec, NV.idm_config(D, A, H, blocks, 16, (32, 32), 100), p, lambda c: 1e-3)
    ts.zero_grad()
    got = ts.idm_loss_grad(s.float().cuda(), a0.float().cuda(), noise.float().cuda(), t.cuda())
    assert float(got) == pytest.approx(float(loss), rel=2e-5)
    _check_grads(ts.grads_dict(), grads)


def test_adam_kernel_matches_oracle(cuda):
    from latent_diffusion_planning_b200 import _native as N
    lib = N.load()
    g = torch.Generator().manual_seed(0)
    n = 100003
    w = torch.randn(n, generator=g)
    ref_w, mu, nu = w.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    dw, dmu, dnu = w.cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    sched = O.warmup_cosine_decay_schedule(1e-6, 1e-4, 10, 100, 1e-6)
    for count in range(1, 6):
        gr = torch.randn(n, generator=g) * (10.0 ** torch.randint(-6, 2, (n,), generator=g))
        lr = sched(count - 1)
        ref_w, mu, nu = O.adam_update(ref_w, gr.double() * 0.5, mu, nu, count, lr)
        N.check(lib.ldp_adam_update(dw.data_ptr(), gr.cuda().data_ptr(), dmu.data_ptr(), dnu.data_ptr(), n, lr, 0.9, 0.999,
                                    1e-8, count, 0.5, torch.cuda.current_stream().cuda_stream))
        assert float((dw.cpu().double() - ref_w).abs().max()) < 1e-6
        assert float((dmu.cpu().double() - mu).abs().max()) <= 2e-6 * float(mu.abs().max())
    assert float((dw.cpu() - w).abs().max()) > 1e-5         # it moved


LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [16]}
DIMS = (32, 64)


def _norm():
    rng = np.random.default_rng(0)
    obs = {"latent_agentview_image": {"min": np.full(16, -10.0, np.float32), "max": np.full(16, 10.0, np.float32)}}
    for k in LOWDIM:
        n = SHAPES[k][0]
        lo = rng.uniform(-1.5, -0.5, n).astype(np.float32)
        obs[k] = {"min": lo, "max": lo + rng.uniform(1.0, 3.0, n).astype(np.float32)}
    return {"obs": obs, "actions": {"clip_min": np.full(7, -1.0, np.float32), "clip_max": np.full(7, 1.0, np.float32)}}


def _train_batch(B, Hh=9, seed=0):
    """What the latent dataloader yields (data/robomimic_latent_data.py:116-147): precomputed latents + low-dim + actions."""
    g = torch.Generator().manual_seed(seed)
    obs = {"latent_agentview_image": torch.randn(B, Hh, 16, generator=g) * 3}
    for k in LOWDIM:
        obs[k] = torch.rand(B, Hh, SHAPES[k][0], generator=g) * 2 - 1
    return {"obs": obs, "actions": torch.randn(B, Hh, 7, generator=g)}


def test_agent_update_matches_oracle_training_step(cuda):
    """Two LDPAgent.update steps == two oracle steps (losses, optax.adam with the warm-up schedule) fed the same
    timesteps and noise; afterwards the sampling handles run on the trained weights."""
    from latent_diffusion_planning_b200 import handles as H, params as P
    from latent_diffusion_planning_b200.agent import LDPAgent, STREAM_TRAIN_IDM, STREAM_TRAIN_PLANNER, normalize_unnormalize
    norm = _norm()
    ag = LDPAgent.create(5, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=DIMS, diffusion_step_embed_dim=32),
                         rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=norm, vae_feature_dim=16,
                         vae_block_out_channels=(32,) * 6, obs_horizon=1, pred_horizon=8, action_horizon=4,
                         planner_n_diffusion_steps=100, idm_n_diffusion_steps=100, precision="fp32", lr=1e-3, end_lr=1e-5,
                         idm_lr=2e-3, idm_end_lr=1e-5, warmup_steps=2, decay_steps=10)
    pp = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in P.unnest(ag.get_params()["planner_params"]).items()}
    ip = {k: torch.as_tensor(v, dtype=torch.float64) for k, v in P.unnest(ag.get_params()["idm_params"]).items()}
    st = {n: ({k: torch.zeros_like(v) for k, v in q.items()}, {k: torch.zeros_like(v) for k, v in q.items()})
          for n, q in (("p", pp), ("i", ip))}
    psched = O.warmup_cosine_decay_schedule(1e-5, 1e-3, 2, 10, 1e-5)
    isched = O.warmup_cosine_decay_schedule(1e-5, 2e-3, 2, 10, 1e-5)
    sched = O.ddpm_schedule(100)
    B, D = 6, 25
    for step in range(2):
        batch = _train_batch(B, seed=step)
        seed = 100 + step
        _, m = ag.update(batch, seed, step)
        # oracle side, same draws
        emb = torch.cat([normalize_unnormalize(batch["obs"][k], norm["obs"][k], True) for k in ["latent_agentview_image"] + LOWDIM],
                        dim=-1).double()
        act = batch["actions"].clamp(-1, 1).double()
        tp = torch.randint(0, 100, (B,), generator=torch.Generator().manual_seed(seed * 2)).numpy()
        ti = torch.randint(0, 100, (B * 8,), generator=torch.Generator().manual_seed(seed * 2 + 1)).numpy()
        zp = torch.tensor(O.philox_normal_rows(seed, STREAM_TRAIN_PLANNER, step, 0, B * 8, D)).reshape(B, 8, D).double()
        zi = torch.tensor(O.philox_normal_rows(seed, STREAM_TRAIN_IDM, step, 0, B * 8, 7)).double()
        lp, gp = O.loss_and_grads(lambda q: O.planner_loss(q, sched, emb, 1, tp, zp, down_dims=DIMS, n_groups=8, step_embed_dim=32), pp)
        li, gi = O.loss_and_grads(lambda q: O.idm_loss(q, sched, emb, act, 1, ti, zi), ip)
        assert float(m["plan_loss"]) == pytest.approx(float(lp), rel=5e-5)
        assert float(m["idm_loss"]) == pytest.approx(float(li), rel=5e-5)
        assert float(m["loss"]) == pytest.approx(float(lp + li), rel=5e-5)
        gn = torch.sqrt(sum((v ** 2).sum() for v in gp.values()) + sum((v ** 2).sum() for v in gi.values()))
        assert float(m["g_norm"]) == pytest.approx(float(gn), rel=1e-4)
        assert m["planner_step"] == step and m["idm_step"] == step
        assert m["planner_lr"] == pytest.approx(isched(step)) and m["idm_lr"] == pytest.approx(isched(step))   # reference quirk
        for q, gq, (mu, nu), sc in ((pp, gp, st["p"], psched), (ip, gi, st["i"], isched)):
            for k in q:
                q[k], mu[k], nu[k] = O.adam_update(q[k], gq[k], mu[k], nu[k], step + 1, sc(step))
    got_p = P.unnest(ag.get_params()["planner_params"])
    got_i = P.unnest(ag.get_params()["idm_params"])
    # Adam's first steps move every weight by ~lr regardless of gradient size, so tiny-gradient weights are
    # ill-conditioned: compare against the step size
    for got, ref in ((got_p, pp), (got_i, ip)):
        for k in ref:
            assert float(np.abs(got[k] - ref[k].numpy()).max()) < 2e-4, k
    # the sampling handles pick the trained weights up
    x = torch.randn(2, 8, D, generator=torch.Generator().manual_seed(9))
    c = torch.randn(2, D, generator=torch.Generator().manual_seed(10))
    eps = ag.planner.forward(x.cuda(), 7, c.cuda(), precision="fp32")
    ref = O.unet_forward({k: v for k, v in pp.items()}, x.double(), 7, c.double(), down_dims=DIMS, n_groups=8, step_embed_dim=32)
    assert float((eps.cpu().double() - ref).abs().max()) < 2e-3
    action, info = ag.sample_viz({"obs": {k: v[:, :1] for k, v in _train_batch(2)["obs"].items()}}, 3)
    assert tuple(action.shape) == (2, 4, 7) and torch.isfinite(action).all()


def test_update_gating(cuda):
    from latent_diffusion_planning_b200.agent import LDPAgent
    ag = LDPAgent.create(1, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=(16,), diffusion_step_embed_dim=16),
                         rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=_norm(), vae_feature_dim=16,
                         vae_block_out_channels=(32,) * 6, precision="fp32", update_planner_every=2, update_idm_after=1)
    _, m0 = ag.update(_train_batch(2), 0, 0)          # planner only
    assert m0["idm_lr"] == 0 and m0["idm_step"] == 0 and float(m0["idm_loss"]) == 0 and float(m0["plan_loss"]) > 0
    _, m1 = ag.update(_train_batch(2), 1, 1)          # IDM only
    assert m1["planner_lr"] == 0 and m1["noise_diff"] == 0 and float(m1["plan_loss"]) == 0 and float(m1["idm_loss"]) > 0
    _, m2 = ag.update_mixed(_train_batch(2), _train_batch(2, seed=5), 2, 2)
    assert m2["planner_step"] == 1 and m2["idm_step"] == 1


def test_workspace_run_logs_saves_and_restores(cuda, tmp_path):
    """train_bc Workspace mirror: 6 updates on window-sampled latent episodes, the reference's cadence (log / dump /
    save / eval), then a snapshot restored into a fresh agent reproduces the trained network."""
    import csv
    from latent_diffusion_planning_b200 import params as P, train_bc as TB
    from latent_diffusion_planning_b200.agent import LDPAgent
    g = np.random.default_rng(0)
    eps = {}
    for d, n in enumerate([12, 9, 15]):
        eps[f"demo_{d}"] = {"obs": {"latent_agentview_image": g.normal(0, 3, (n, 16)).astype(np.float32),
                                    **{k: g.uniform(-0.4, 0.4, (n, SHAPES[k][0])).astype(np.float32) for k in LOWDIM}},
                            "actions": g.normal(0, 0.5, (n, 7)).astype(np.float32)}
    keys = ["latent_agentview_image"] + LOWDIM
    ds = TB.LatentSequenceDataset(eps, keys, seq_length=9, n_frame_stack=1)
    mk = lambda: LDPAgent.create(2, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=(256,), diffusion_step_embed_dim=32),
                                 rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=_norm(), vae_feature_dim=16,
                                 vae_block_out_channels=(32,) * 6, planner_n_diffusion_steps=4, idm_n_diffusion_steps=4,
                                 precision="bf16", lr=1e-3, warmup_steps=2, decay_steps=20)
    ds_gpu = TB.LatentSequenceDataset(eps, keys, seq_length=9, n_frame_stack=1).to("cuda")      # one gather per key on the device
    b0, b1 = ds.sample_batch(8, np.random.default_rng(4)), ds_gpu.sample_batch_fast(8, np.random.default_rng(4))
    assert torch.equal(b0["obs"]["latent_agentview_image"], b1["obs"]["latent_agentview_image"].cpu()) and torch.equal(b0["actions"], b1["actions"].cpu())
    ws = TB.Workspace(mk(), ds_gpu, tmp_path, eval_dataset=ds, batch_size=8, n_grad_steps=6, log_every_step=2, dump_every_step=3,
                      save_every_step=6, eval_every_step=6, n_eval_batches=1)
    last = ws.run()
    assert ws.step == 6 and np.isfinite(last["loss"])
    train_rows = list(csv.DictReader(open(tmp_path / "train.csv")))
    assert [r["step"] for r in train_rows] == ["3", "6"] and float(train_rows[-1]["planner_step"]) >= 3
    eval_rows = list(csv.DictReader(open(tmp_path / "eval.csv")))
    assert len(eval_rows) == 1 and {"evaldata/action_mse", "evaldata/plan_loss", "evaldata/full_action_mse", "evaldata/plan_mse"} <= set(eval_rows[0])
    ck = tmp_path / "ckpt" / "6.ckpt.npz"
    assert ck.exists()
    fresh = TB.Workspace(mk(), ds, tmp_path / "b", batch_size=8)
    fresh.load_snapshot(ck)
    assert fresh.step == 6
    a, b = P.unnest(ws.agent.get_params()["planner_params"]), P.unnest(fresh.agent.get_params()["planner_params"])
    assert all(np.array_equal(a[k], b[k]) for k in a)
    batch = ds.sample_batch(4, np.random.default_rng(1))
    x = ws.agent.sample_action(batch, 3)
    y = fresh.agent.sample_action(batch, 3)
    assert torch.equal(x, y)
    with pytest.raises(TypeError):
        TB.Workspace(mk(), ds, tmp_path / "c", bogus=1)
