"""Consumes tests/golden/ref_jax_goldens.npz - outputs of the REAL reference stack (jax 0.4.26 / flax 0.8.4 / diffusers 0.27.2),
written by scripts/make_reference_goldens.py on a box that has it.  That stack cannot be installed in this image (no
network, not in the wheelhouse), so the file is absent here and these tests skip; the day it exists they pin the oracle's
from-memory halves (diffusers scheduler / VAE, Flax layer semantics, jax.random) to the reference itself."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ldp_oracle as O
from latent_diffusion_planning_b200 import params as P

GOLD = Path(__file__).resolve().parent / "golden" / "ref_jax_goldens.npz"
pytestmark = pytest.mark.skipif(not GOLD.exists(), reason="tests/golden/ref_jax_goldens.npz not generated (needs the real JAX reference stack)")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def test_scheduler_tables_and_steps(gold):
    s = O.ddpm_schedule(100)
    for i, k in enumerate(("betas", "alphas", "alphas_cumprod")):
        assert np.array_equal(np.asarray(s[i], np.float32), gold[f"ddpm/{k}"]), k         # float32 tables, bit for bit
    x, eps = gold["ddpm/x"], gold["ddpm/eps"]
    for t in (99, 98, 50, 1, 0):
        ref = gold[f"ddpm/prev_t{t}"]
        cands = [O.ddpm_step(s, eps, t, x, gold[f"ddpm/{n}_t{t}"].reshape(x.shape), dtype=torch.float32).numpy()
                 for n in ("noise_split1_0", "noise_split1") if gold[f"ddpm/{n}_t{t}"].size == x.size]
        assert min(np.abs(c - ref).max() for c in cands) < 2e-6, t
    got = O.add_noise(s, gold["ddpm/x"], gold["ddpm/eps"], gold["ddpm/add_noise_t"], dtype=torch.float32).numpy()
    assert np.abs(got - gold["ddpm/add_noise"]).max() < 1e-6


@pytest.mark.parametrize("name,D,dims", [("unet_small", 25, (64, 128, 256)), ("unet_full", 265, (256, 512, 1024)), ("unet_t16", 12, (32, 64, 128))])
def test_unet(gold, name, D, dims):
    spec = P.unet_spec(D, D, dims)
    names = set(gold[f"{name}/param_names"].tolist())
    assert set(P.canonicalize_flax_names({k: 0 for k in names})) == set(spec), "Flax parameter-tree names differ from params.unet_spec"
    p = P.init_params(spec, seed=0, perturb=0.1)
    got = O.unet_forward(p, gold[f"{name}/x"], int(gold[f"{name}/k"]), gold[f"{name}/cond"], down_dims=dims).numpy()
    assert np.abs(got - gold[f"{name}/out"]).max() < 2e-5 * max(1.0, np.abs(got).max())     # reference computes in float32


@pytest.mark.parametrize("name,D,A", [("idm_rm", 265, 7), ("idm_aloha", 270, 14)])
def test_idm(gold, name, D, A):
    p = P.init_params(P.idm_spec(D, A), seed=1, perturb=0.1)
    got = O.idm_forward(p, gold[f"{name}/s"], gold[f"{name}/a"], int(gold[f"{name}/k"])).numpy()
    assert np.abs(got - gold[f"{name}/out"]).max() < 2e-5 * max(1.0, np.abs(got).max())


@pytest.mark.parametrize("name,blocks", [("vae_ref6", (32, 64, 64, 64, 64, 64)), ("vae_sd4", (32, 64, 128, 128))])
def test_vae(gold, name, blocks):
    enc = P.init_params(P.vae_encoder_spec(blocks), seed=2, perturb=0.1)
    dec = P.init_params(P.vae_decoder_spec(blocks), seed=3, perturb=0.1)
    img = np.transpose(gold[f"{name}/img_nchw"], (0, 2, 3, 1))
    mean = O.vae_encode_mean(enc, img, blocks).numpy()
    assert np.abs(mean - gold[f"{name}/mean"]).max() < 5e-5 * max(1.0, np.abs(mean).max())
    rec = O.vae_decode(dec, gold[f"{name}/mean"], blocks).numpy()
    ref = gold[f"{name}/decoded"]
    ref = np.transpose(ref, (0, 2, 3, 1)) if ref.shape[1] == 3 else ref
    assert np.abs(rec - ref).max() < 5e-5 * max(1.0, np.abs(rec).max())


def test_jax_random_and_key_threading(gold):
    start_key, step_keys, _ = O.jax_sampling_keys(O.jax_prng_key(5), 4)
    assert np.array_equal(np.asarray(step_keys, np.uint32), gold["rng/step_keys"].astype(np.uint32))
    assert np.abs(O.jax_normal(start_key, (2, 8, 25)) - gold["rng/x_T"]).max() < 1e-6
    assert np.array_equal(O.jax_randint(O.jax_prng_key(9), 64, 0, 100), gold["rng/randint"])
    assert np.array_equal(np.asarray(O.jax_split(O.jax_prng_key(9), 3), np.uint32), gold["rng/split3"].astype(np.uint32))


def test_optax(gold):
    sched = O.warmup_cosine_decay_schedule(1e-6, 1e-4, 1000, 500000, 1e-6)
    got = np.array([sched(int(i)) for i in gold["optax/lr_steps"]])
    assert np.allclose(got, gold["optax/lr"], rtol=1e-6, atol=0)
    w = torch.from_numpy(gold["optax/adam_w"]).double()
    g = torch.from_numpy(gold["optax/adam_g"]).double()
    w1, _, _ = O.adam_update(w, g, torch.zeros_like(w), torch.zeros_like(w), 1, 1e-3)
    assert np.abs(w1.numpy() - gold["optax/adam_w1"]).max() < 1e-7
