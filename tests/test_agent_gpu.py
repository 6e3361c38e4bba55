"""LDPAgent.act() end to end on the GPU (VAE encode -> planner loop -> IDM loop -> un-normalised actions) against the
oracle pipeline restated from reference agent/ldp_agent.py:435-506 with the same counter-based noise."""
import numpy as np
import pytest
import torch

from oracle import ldp_oracle as O

pytestmark = pytest.mark.gpu

LOWDIM = ["robot0_eef_pos", "robot0_eef_quat", "robot0_gripper_qpos"]
SHAPES = {"robot0_eef_pos": [3], "robot0_eef_quat": [4], "robot0_gripper_qpos": [2], "latent_agentview_image": [16]}
VAE_BLOCKS = (32,) * 6          # 64x64 -> 2x2x4 = 16 features: the reference's default latent size (agent/ldp_agent.yaml:5)
DIMS = (64, 128, 256)
N_STEPS = 6


def _norm():
    rng = np.random.default_rng(0)
    obs = {"agentview_image": {"min": 0, "max": 255},
           "latent_agentview_image": {"min": np.full(16, -10.0, np.float32), "max": np.full(16, 10.0, np.float32)}}
    for k in LOWDIM:
        n = SHAPES[k][0]
        lo = rng.uniform(-1.5, -0.5, n).astype(np.float32)
        obs[k] = {"min": lo, "max": lo + rng.uniform(1.0, 3.0, n).astype(np.float32)}
    return {"obs": obs, "actions": {"clip_min": np.full(7, -1.0, np.float32), "clip_max": np.full(7, 1.0, np.float32)}}


def _batch(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    obs = {"agentview_image": torch.randint(0, 256, (B, 1, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.float32)}
    for k in LOWDIM:
        obs[k] = torch.rand(B, 1, SHAPES[k][0], generator=g) * 2 - 1
    return {"obs": obs}


@pytest.fixture(scope="module")
def agent(cuda):
    from latent_diffusion_planning_b200.agent import LDPAgent
    return LDPAgent.create(3, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=DIMS),
                           rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=_norm(),
                           vae_feature_dim=16, vae_block_out_channels=VAE_BLOCKS, obs_horizon=1, pred_horizon=8,
                           action_horizon=4, planner_n_diffusion_steps=N_STEPS, idm_n_diffusion_steps=N_STEPS,
                           precision="fp32")


def _oracle_act(agent, batch, seed):
    from latent_diffusion_planning_b200 import params as P
    norm = agent.obs_normalization["obs"]
    pp, ip = P.unnest(agent.get_params()["planner_params"]), P.unnest(agent.get_params()["idm_params"])
    vp = P.init_params(P.vae_encoder_spec(VAE_BLOCKS), seed=3 + 2)
    img = batch["obs"]["agentview_image"]
    B = img.shape[0]
    x = img.reshape(B, 64, 64, 3).to(torch.float64) / 255.0 * 2 - 1
    z = O.vae_encode_mean(vp, x, VAE_BLOCKS).reshape(B, 1, 16)
    feats = (z + 10.0) / 20.0 * 2 - 1
    low = torch.cat([(batch["obs"][k].double() - torch.tensor(norm[k]["min"]).double())
                     / (torch.tensor(norm[k]["max"]).double() - torch.tensor(norm[k]["min"]).double()) * 2 - 1 for k in LOWDIM], dim=-1)
    obs_emb = torch.cat([feats, low], dim=-1)
    D, T, Ha, A = 25, 8, 4, 7
    sched = O.ddpm_schedule(N_STEPS)
    xT = torch.tensor(O.philox_normal_rows(seed, 2, 0, 0, B * T, D)).reshape(B, T, D)
    noise = torch.stack([torch.tensor(O.philox_normal_rows(seed, 0, N_STEPS - 1 - i, 0, B * T, D)).reshape(B, T, D)
                         for i in range(N_STEPS)])
    x0 = O.planner_sample(pp, sched, xT, obs_emb[:, 0], noise, N_STEPS, down_dims=DIMS)
    plan, ssp = O.make_transitions(obs_emb[:, 0:1], x0, Ha)
    aT = torch.tensor(O.philox_normal_rows(seed, 3, 0, 0, B * Ha, A))
    an = torch.stack([torch.tensor(O.philox_normal_rows(seed, 1, N_STEPS - 1 - i, 0, B * Ha, A)) for i in range(N_STEPS)])
    a = O.idm_sample(ip, sched, ssp, aT, an, N_STEPS).reshape(B, Ha, A).clamp(-1, 1)
    return a, plan


def test_act_matches_oracle_pipeline(agent):
    batch = _batch(3)
    action, metrics = agent.act(batch, 42)
    ref_a, ref_plan = _oracle_act(agent, batch, 42)
    assert tuple(action.shape) == (3, 4, 7) and tuple(metrics["plan"].shape) == (3, 5, 25)
    assert float((metrics["plan"].cpu().double() - ref_plan).abs().max()) < 1e-3
    assert float((action.cpu().double() - ref_a).abs().max()) < 1e-3
    assert float(action.abs().max()) <= 1.0


def test_act_is_shard_invariant(agent):
    batch = _batch(5, seed=1)
    full, _ = agent.sample_viz(batch, 7)
    part = {"obs": {k: v[2:] for k, v in batch["obs"].items()}}
    tail, _ = agent.sample_viz(part, 7, row_offset=2)
    assert torch.equal(tail, full[2:])


def test_sample_action_and_from_plan(agent):
    g = torch.Generator().manual_seed(5)
    B, Hh = 2, 3
    obs = {"agentview_image": torch.randint(0, 256, (B, Hh, 64, 64, 3), generator=g, dtype=torch.int32).to(torch.uint8)}
    for k in LOWDIM:
        obs[k] = torch.rand(B, Hh, SHAPES[k][0], generator=g) * 2 - 1
    a = agent.sample_action({"obs": obs}, 9)
    assert tuple(a.shape) == (B, Hh - 1, 7) and torch.isfinite(a).all()
    nxt = torch.rand(B, Hh, 25, generator=g).cuda() * 2 - 1
    a2 = agent.sample_action_from_plan({"obs": obs}, nxt, 9)
    assert tuple(a2.shape) == (B, Hh, 7) and torch.isfinite(a2).all()


def test_process_sdvae_data_matches_direct_encode(cuda, tmp_path):
    """process_sdvae_data mirror (reference process_sdvae_data.py:52-118): the latent file holds exactly what
    VaeEncoder.encode returns for the assembled episode frames, in shards, with the reference's min/max attributes."""
    import numpy as np
    from latent_diffusion_planning_b200 import handles as H, params as P, process_sdvae_data as PS
    blocks = (32, 64)
    vp = P.init_params(P.vae_encoder_spec(blocks, 3, 4, 1), seed=2, perturb=0.1)
    vae = H.VaeEncoder(vp, blocks, 3, 4, 1, 8, 16)
    g = np.random.default_rng(0)
    eps = {f"demo_{i}": {"obs": {"agentview_image": g.integers(0, 256, (n, 16, 16, 3), dtype=np.uint8)},
                         "next_obs": {"agentview_image": g.integers(0, 256, (n, 16, 16, 3), dtype=np.uint8)}}
           for i, n in enumerate([5, 2])}
    path = PS.process_sdvae_data(eps, ["agentview_image"], vae, tmp_path, data_name="rm_lift", shard=4)
    with np.load(path) as z:
        got = z["data/demo_0/latent/agentview_image"]
        frames = PS.episode_frames(eps["demo_0"], "agentview_image", "rm_lift")
        ref = vae.encode(torch.from_numpy(frames).cuda(), precision="bf16").cpu().numpy()
        assert got.shape == (6, 8, 8, 4) and np.array_equal(got, ref)           # images are independent: shards == one batch
        lo, hi = float(z["data.attrs/min_z"]), float(z["data.attrs/max_z"])
        allz = np.concatenate([z[k].ravel() for k in z.files if k.startswith("data/")])
        assert lo == min(0.0, float(allz.min())) and hi == max(0.0, float(allz.max())) and int(z["data.attrs/total"]) == 2


def test_sample_viz_decodes_the_plan(cuda):
    """viz=True: plan_viz = vae_decode(plan) (reference agent/ldp_agent.py:66-85, :483): first 16 features of every plan
    frame -> (2,2,4) latents -> unnormalize_obs -> decoder -> (B, Ha+1, 3, 64, 64)."""
    from latent_diffusion_planning_b200 import params as P
    from latent_diffusion_planning_b200.agent import LDPAgent
    ag = LDPAgent.create(3, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=DIMS),
                         rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=_norm(),
                         vae_feature_dim=16, vae_block_out_channels=VAE_BLOCKS, obs_horizon=1, pred_horizon=8,
                         action_horizon=4, planner_n_diffusion_steps=2, idm_n_diffusion_steps=2, precision="fp32", viz=True)
    action, info = ag.sample_viz(_batch(2, seed=4), 7)
    viz, plan = info["plan_viz"], info["plan"]
    assert viz.shape == (2, 5, 3, 64, 64) and torch.isfinite(viz).all()
    dp = P.init_params(P.vae_decoder_spec(VAE_BLOCKS), seed=3 + 3)
    z = plan[..., :16].reshape(10, 2, 2, 4).cpu().double()
    z = torch.clamp((z + 1) / 2 * 20.0 - 10.0, -10.0, 10.0)                       # unnormalize_obs with min -10 / max 10
    ref = O.vae_decode(dp, z, VAE_BLOCKS, 2, 32).permute(0, 3, 1, 2).reshape(2, 5, 3, 64, 64)
    assert float((viz.cpu().double() - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))
    # without viz the reference-shaped dict still carries the key
    assert ag.viz and LDPAgent.create(3, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=DIMS),
                                      rgb_obs=["latent_agentview_image"], lowdim_obs=LOWDIM, obs_normalization=_norm(),
                                      vae_feature_dim=16, vae_block_out_channels=VAE_BLOCKS, planner_n_diffusion_steps=2,
                                      idm_n_diffusion_steps=2, precision="fp32").sample_viz(_batch(1), 1)[1]["plan_viz"] is None
