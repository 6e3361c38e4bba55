"""Device jax.random (ldp_jax_random) against the oracle restatement: bits bit-exact (integer work), normal to float32
rounding of erf_inv; and LDPAgent sampling driven by a raw JAX key equals the oracle pipeline fed the oracle's draws."""
import numpy as np
import pytest
import torch

from oracle import ldp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 65537])
def test_bits_bit_exact_and_normal_close(cuda, n):
    from latent_diffusion_planning_b200 import jax_random as JR
    keys = np.stack([O.jax_prng_key(s) for s in (0, 42, 2 ** 33 + 5)])
    got_b = JR.bits(keys, n).cpu().numpy().view(np.uint32)
    got_n = JR.normal(keys, n).cpu().numpy()
    for i, k in enumerate(keys):
        with np.errstate(over="ignore"):
            assert np.array_equal(got_b[i], O.jax_threefry_bits(k, n))
            ref = O.jax_normal(k, (n,))
        assert np.abs(got_n[i] - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    assert float(JR.normal(O.jax_prng_key(42), 1)[0, 0]) == pytest.approx(-0.18471177, abs=2e-7)   # the documented value


LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [16]}


def test_sample_viz_with_a_jax_key_follows_the_reference_key_threading(cuda):
    """agent/ldp_agent.py:461-503: split sequence, start noise, per-step scheduler keys (planner, then IDM)."""
    from latent_diffusion_planning_b200 import params as P
    from latent_diffusion_planning_b200.agent import LDPAgent
    dims, n_steps, B, T, D, Ha, A = (64, 128), 5, 3, 8, 25, 4, 7
    norm = {"obs": {"latent_agentview_image": {"min": np.full(16, -10.0, np.float32), "max": np.full(16, 10.0, np.float32)},
                    **{k: {"min": -np.ones(SHAPES[k][0], np.float32), "max": np.ones(SHAPES[k][0], np.float32)} for k in LOWDIM}},
            "actions": {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}}
    ag = LDPAgent.create(4, None, {"ac_dim": A, "all_shapes": SHAPES}, planner=dict(down_dims=dims), rgb_obs=["latent_agentview_image"],
                         lowdim_obs=LOWDIM, obs_normalization=norm, vae_feature_dim=16, vae_block_out_channels=(32,) * 6,
                         pred_horizon=T, action_horizon=Ha, planner_n_diffusion_steps=n_steps, idm_n_diffusion_steps=n_steps,
                         precision="fp32")
    g = torch.Generator().manual_seed(0)
    obs = {"latent_agentview_image": torch.randn(B, 1, 16, generator=g) * 3, **{k: torch.rand(B, 1, SHAPES[k][0], generator=g) * 2 - 1 for k in LOWDIM}}
    key = O.jax_prng_key(123)
    action, info = ag.sample_viz({"obs": obs}, key)
    # oracle side
    pp, ip = P.unnest(ag.get_params()["planner_params"]), P.unnest(ag.get_params()["idm_params"])
    emb = torch.cat([(obs["latent_agentview_image"] + 10) / 20 * 2 - 1] + [obs[k] for k in LOWDIM], dim=-1).double()
    sched = O.ddpm_schedule(n_steps)
    with np.errstate(over="ignore"):
        k0, ks, rest = O.jax_sampling_keys(key, n_steps)
        xT = torch.tensor(O.jax_normal(k0, (B, T, D))).double()
        z = torch.stack([torch.tensor(O.jax_normal(k, (B, T, D))) for k in ks]).double()
        x0 = O.planner_sample(pp, sched, xT, emb[:, 0], z, n_steps, down_dims=dims)
        plan, ssp = O.make_transitions(emb[:, 0:1], x0, Ha)
        k0, ks, _ = O.jax_sampling_keys(rest, n_steps)
        aT = torch.tensor(O.jax_normal(k0, (B * Ha, A))).double()
        za = torch.stack([torch.tensor(O.jax_normal(k, (B * Ha, A))) for k in ks]).double()
    a = O.idm_sample(ip, sched, ssp, aT, za, n_steps).reshape(B, Ha, A).clamp(-1, 1)
    assert float((info["plan"].cpu().double() - plan).abs().max()) < 1e-3
    assert float((action.cpu().double() - a).abs().max()) < 1e-3
    with pytest.raises(ValueError):
        ag.sample_viz({"obs": obs}, key, row_offset=1)
    a2 = ag.sample_action({"obs": {k: v.repeat(1, 3, 1) for k, v in obs.items()}}, key)
    assert tuple(a2.shape) == (B, 2, A) and torch.isfinite(a2).all()


def test_update_with_a_jax_key_draws_the_reference_timesteps_and_noise(cuda):
    from latent_diffusion_planning_b200 import jax_random as JR, params as P
    from latent_diffusion_planning_b200.agent import LDPAgent
    key = O.jax_prng_key(77)
    with np.errstate(over="ignore"):
        assert np.array_equal(JR.randint(key, 1001, 0, 100).cpu().numpy(), O.jax_randint(key, 1001, 0, 100))
        assert np.array_equal(JR.randint(key, 64, 3, 10).cpu().numpy(), O.jax_randint(key, 64, 3, 10))
    dims, B, D, A = (32, 64), 4, 25, 7
    norm = {"obs": {"latent_agentview_image": {"min": np.full(16, -10.0, np.float32), "max": np.full(16, 10.0, np.float32)},
                    **{k: {"min": -np.ones(SHAPES[k][0], np.float32), "max": np.ones(SHAPES[k][0], np.float32)} for k in LOWDIM}},
            "actions": {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}}
    ag = LDPAgent.create(4, None, {"ac_dim": A, "all_shapes": SHAPES}, planner=dict(down_dims=dims, diffusion_step_embed_dim=32),
                         rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=norm, vae_feature_dim=16,
                         vae_block_out_channels=(32,) * 6, precision="fp32")
    g = torch.Generator().manual_seed(0)
    batch = {"obs": {"latent_agentview_image": torch.randn(B, 9, 16, generator=g) * 3,
                     **{k: torch.rand(B, 9, SHAPES[k][0], generator=g) * 2 - 1 for k in LOWDIM}},
             "actions": torch.rand(B, 9, A, generator=g) * 2 - 1}
    pp, ip = P.unnest(ag.get_params()["planner_params"]), P.unnest(ag.get_params()["idm_params"])
    _, m = ag.update(batch, key, 0)
    emb = torch.cat([(batch["obs"]["latent_agentview_image"] + 10) / 20 * 2 - 1] + [batch["obs"][k] for k in LOWDIM], dim=-1).double()
    sched = O.ddpm_schedule(100)
    with np.errstate(over="ignore"):
        keys = O.jax_update_keys(key)
        tp, zp = O.jax_randint(keys["planner"][0], B, 0, 100), O.jax_normal(keys["planner"][1], (B, 8, D))
        ti, zi = O.jax_randint(keys["idm"][0], B * 8, 0, 100), O.jax_normal(keys["idm"][1], (B * 8, A))
    lp = O.planner_loss(pp, sched, emb, 1, tp, torch.tensor(zp).double(), down_dims=dims, n_groups=8, step_embed_dim=32)
    li = O.idm_loss(ip, sched, emb, batch["actions"].double(), 1, ti, torch.tensor(zi).double())
    assert float(m["plan_loss"]) == pytest.approx(float(lp), rel=5e-5)
    assert float(m["idm_loss"]) == pytest.approx(float(li), rel=5e-5)
