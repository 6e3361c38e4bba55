"""Host-side contract checks that need no GPU: the reference's config keys are honoured or rejected (never silently
ignored), parameter trees must match the network spec exactly, normalisation glue (reference utils/data_utils.py:9-68)."""
import numpy as np
import pytest
import torch

from latent_diffusion_planning_b200 import agent as A, params as P

SHAPES = {"robot0_eef_pos": [3], "latent_agentview_image": [16]}


def _create(**kw):
    return A.LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=(32, 64)),
                             lowdim_obs=["robot0_eef_pos"], rgb_obs=[], vae_feature_dim=16, **kw)


def test_unsupported_reference_options_raise_before_any_device_work():
    # reference agent/ldp_agent.yaml:17-23: idm_net has n_blocks / use_layer_norm / dropout_rate
    with pytest.raises(NotImplementedError, match="use_layer_norm"):
        _create(idm_net=dict(n_blocks=3, use_layer_norm=False))
    with pytest.raises(NotImplementedError, match="dropout"):
        _create(idm_net=dict(dropout_rate=0.1))
    with pytest.raises(NotImplementedError, match="downsample"):
        A.LDPAgent.create(0, None, {"ac_dim": 7, "all_shapes": SHAPES}, planner=dict(down_dims=(32, 64), downsample=False),
                          lowdim_obs=["robot0_eef_pos"], rgb_obs=[], vae_feature_dim=16)


def test_idm_spec_follows_n_blocks_and_extra_tensors_are_an_error():
    s3, s4 = P.idm_spec(25, 7, n_blocks=3), P.idm_spec(25, 7, n_blocks=4)
    assert len(s4) == len(s3) + 6 and "MLPResNet_0/MLPResNetBlock_3/Dense_1/kernel" in s4
    p4 = P.init_params(s4, seed=0)
    assert P.flatten_params(s4, p4).size == P.spec_size(s4)
    with pytest.raises(ValueError, match="does not contain"):      # a 4-block tree handed to a 3-block network
        P.flatten_params(s3, p4)
    p3 = dict(P.init_params(s3, seed=0))
    p3.pop("MLPResNet_0/Dense_1/bias")
    with pytest.raises(KeyError):
        P.flatten_params(s3, p3)


def test_action_unnormalisation_matches_reference_formulas():
    """utils/data_utils.py:13-15 (min/max: (a+1)/2*(max-min)+min then clip) and :61-65 (clip spec), aloha / robomimic."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn(5, 4, 14, generator=g) * 1.5
    lo = np.linspace(-2.0, -0.5, 14).astype(np.float32)
    hi = np.linspace(0.5, 3.0, 14).astype(np.float32)
    got = A.normalize_unnormalize(a, {"min": lo, "max": hi}, False).numpy()
    ref = np.clip((a.numpy() + 1) / 2 * (hi - lo) + lo, lo, hi)
    assert np.allclose(got, ref, atol=1e-6)
    back = A.normalize_unnormalize(torch.from_numpy(ref), {"min": lo, "max": hi}, True).numpy()
    inside = np.abs(a.numpy()) < 1
    assert np.allclose(back[inside], a.numpy()[inside], atol=1e-5)
    got = A.normalize_unnormalize(a[..., :7], {"clip_min": -np.ones(7, np.float32), "clip_max": np.ones(7, np.float32)}, False)
    assert torch.equal(got, a[..., :7].clamp(-1, 1))
    with pytest.raises(NotImplementedError):
        A.normalize_unnormalize(a, {"mean": 0, "std": 1}, True)


def test_resize_bilinear_is_half_pixel_triangle():
    """process_sdvae_data.py:66-69 `jax.image.resize(..., 'bilinear')`: output pixel i samples input coordinate
    (i + 0.5) * in/out - 0.5 with edge clamping; checked against an explicit loop on a 2x upsampling."""
    from latent_diffusion_planning_b200 import process_sdvae_data as PS
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 4, 4, 3, generator=g)
    y = PS.resize_bilinear(x, 8).numpy()
    xn = x.numpy()
    ref = np.zeros((2, 8, 8, 3), np.float32)
    for i in range(8):
        for j in range(8):
            fi, fj = (i + 0.5) / 2 - 0.5, (j + 0.5) / 2 - 0.5
            i0, j0 = int(np.floor(fi)), int(np.floor(fj))
            wi, wj = fi - i0, fj - j0
            c = lambda v: min(max(v, 0), 3)
            ref[:, i, j] = ((1 - wi) * (1 - wj) * xn[:, c(i0), c(j0)] + (1 - wi) * wj * xn[:, c(i0), c(j0 + 1)]
                            + wi * (1 - wj) * xn[:, c(i0 + 1), c(j0)] + wi * wj * xn[:, c(i0 + 1), c(j0 + 1)])
    assert np.allclose(y, ref, atol=1e-6)
