#!/usr/bin/env python
"""Pin-the-oracle recipe for a box that HAS the reference stack (this image does not: no jax / flax / diffusers, no network):

    pip install "jax==0.4.26" "flax==0.8.4" "diffusers==0.27.2" "optax==0.2.2"      # reference env.yml / README.md:44
    python scripts/make_reference_goldens.py --reference /path/to/latent_diffusion_planning [--out tests/golden/ref_jax_goldens.npz]

Imports the REAL reference (`networks/*`, its `FlaxDDPMScheduler` / `FlaxAutoencoderKL` exactly as agent/ldp_agent.py:571-650
and model/stable_vae_model.yaml construct them) and dumps, for seeded inputs and `params.init_params` weights under the Flax
names: scheduler tables and single steps (t = 99, 98, 50, 1, 0) with the noise `step` really draws from its key, add_noise,
UNet / IDM outputs (float32, as the reference computes), VAE encoder means (6-block reference yaml topology and the 4-block
benchmark topology, small widths), jax.random draws behind `sample_viz_step`'s key threading (agent/ldp_agent.py:461-476) and
the optax schedule / one Adam step.  tests/test_reference_jax_goldens.py consumes the file when it exists (skips otherwise)
and checks the oracle against every entry - that test going green is what turns "parity unpinned" into "pinned".

Every ⚠ item of DESIGN.md section 2 has an entry here: `split(key, 1)[0]` inside `step` (ddpm/noise_*), XLA cumprod
(ddpm/alphas_cumprod), Flax 'SAME' stride-2 and ConvTranspose padding (unet/*), GroupNorm fast variance (unet/*, vae/*),
FlaxDownsample2D's (0,1) pad and the attention scale (vae/*), parameter-tree names (a wrong name raises in `apply`).
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "ref_jax_goldens.npz"))
    a = ap.parse_args()
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, a.reference)
    import jax
    import jax.numpy as jnp
    from functools import partial
    from diffusers import FlaxAutoencoderKL, FlaxDDPMScheduler
    from latent_diffusion_planning_b200 import params as P
    from networks.diffusion import FourierFeatures
    from networks.diffusion_nets_v2 import ConditionalUnet1D
    from networks.mlp_diffusion_nets import MLPDiffusion, MLPResNet
    from networks.mlp_nets import MLP
    out = {"meta/versions": np.array([jax.__version__, __import__("flax").__version__, __import__("diffusers").__version__])}
    tree = lambda flat: jax.tree_util.tree_map(jnp.asarray, P.nest(flat))
    g = np.random.default_rng(7)

    # ---- scheduler (agent/ldp_agent.py:637-650) ----
    sch = FlaxDDPMScheduler(num_train_timesteps=100, beta_schedule="squaredcos_cap_v2", clip_sample=True, prediction_type="epsilon")
    st = sch.create_state()
    out["ddpm/betas"], out["ddpm/alphas"], out["ddpm/alphas_cumprod"] = (np.asarray(v) for v in (st.common.betas, st.common.alphas, st.common.alphas_cumprod))
    x = g.standard_normal((3, 8, 25)).astype(np.float32)
    eps = g.standard_normal((3, 8, 25)).astype(np.float32)
    out["ddpm/x"], out["ddpm/eps"] = x, eps
    key = jax.random.PRNGKey(11)
    out["ddpm/key"] = np.asarray(key)
    for t in (99, 98, 50, 1, 0):
        out[f"ddpm/prev_t{t}"] = np.asarray(sch.step(st, jnp.asarray(eps), t, jnp.asarray(x), key).prev_sample)
        # the noise `step` draws: variance noise = normal(split(key, 1), shape) in diffusers 0.27.2
        out[f"ddpm/noise_split1_t{t}"] = np.asarray(jax.random.normal(jax.random.split(key, num=1), shape=eps.shape, dtype=jnp.float32))
        out[f"ddpm/noise_split1_0_t{t}"] = np.asarray(jax.random.normal(jax.random.split(key, num=1)[0], shape=eps.shape, dtype=jnp.float32))
    tt = np.array([0, 50, 99], np.int32)
    out["ddpm/add_noise_t"] = tt
    out["ddpm/add_noise"] = np.asarray(sch.add_noise(st, jnp.asarray(x), jnp.asarray(eps), jnp.asarray(tt)))

    # ---- score networks ----
    for name, D, dims, B, T, k in (("unet_small", 25, (64, 128, 256), 3, 8, 50), ("unet_full", 265, (256, 512, 1024), 2, 8, 50),
                                   ("unet_t16", 12, (32, 64, 128), 2, 16, 7)):
        p = P.init_params(P.unet_spec(D, D, dims), seed=0, perturb=0.1)
        xs = g.standard_normal((B, T, D)).astype(np.float32)
        c = g.uniform(-1, 1, (B, D)).astype(np.float32)
        net = ConditionalUnet1D(input_dim=D, global_cond_dim=D, diffusion_step_embed_dim=256, down_dims=dims, kernel_size=5, n_groups=8)
        out[f"{name}/x"], out[f"{name}/cond"], out[f"{name}/k"] = xs, c, np.int32(k)
        out[f"{name}/out"] = np.asarray(net.apply({"params": tree(p)}, jnp.asarray(xs), k, jnp.asarray(c)))
        init = net.init(jax.random.PRNGKey(0), jnp.asarray(xs), k, jnp.asarray(c))["params"]          # the names Flax really registers
        out[f"{name}/param_names"] = np.array(sorted("/".join(q.key for q in path) for path, _ in jax.tree_util.tree_flatten_with_path(init)[0]))
    for name, D, A, N, k in (("idm_rm", 265, 7, 6, 10), ("idm_aloha", 270, 14, 5, 42)):
        p = P.init_params(P.idm_spec(D, A), seed=1, perturb=0.1)
        s = g.uniform(-1, 1, (N, 2 * D)).astype(np.float32)
        act = g.standard_normal((N, A)).astype(np.float32)
        idm = MLPDiffusion(partial(MLP, hidden_dims=(256, 256), activations="mish", activate_final=False),
                           partial(MLPResNet, n_blocks=3, out_dim=A, dropout_rate=None, use_layer_norm=True, hidden_dim=256),
                           partial(FourierFeatures, output_size=256, learnable=False))
        out[f"{name}/s"], out[f"{name}/a"], out[f"{name}/k"] = s, act, np.int32(k)
        out[f"{name}/out"] = np.asarray(idm.apply({"params": tree(p)}, jnp.asarray(s), jnp.asarray(act), k))

    # ---- VAE encoder (model/stable_vae_model.yaml; agent/ldp_agent.py:58-59) ----
    for name, blocks, size in (("vae_ref6", (32, 64, 64, 64, 64, 64), 64), ("vae_sd4", (32, 64, 128, 128), 32)):
        vae = FlaxAutoencoderKL(in_channels=3, out_channels=3, down_block_types=("DownEncoderBlock2D",) * len(blocks),
                                up_block_types=("UpDecoderBlock2D",) * len(blocks), block_out_channels=blocks, layers_per_block=2,
                                act_fn="silu", latent_channels=4, norm_num_groups=32, sample_size=size)
        p = {**P.init_params(P.vae_encoder_spec(blocks), seed=2, perturb=0.1), **P.init_params(P.vae_decoder_spec(blocks), seed=3, perturb=0.1)}
        img = g.uniform(-1, 1, (2, 3, size, size)).astype(np.float32)                     # NCHW, as the agent passes it
        out[f"{name}/img_nchw"] = img
        dist = vae.apply({"params": tree(p)}, jnp.asarray(img), method=vae.encode).latent_dist
        out[f"{name}/mean"] = np.asarray(dist.mean)
        z = np.asarray(dist.mean)
        out[f"{name}/decoded"] = np.asarray(vae.apply({"params": tree(p)}, jnp.asarray(z).transpose(0, 3, 1, 2), method=vae.decode).sample)

    # ---- jax.random behind sample_viz_step's key threading (agent/ldp_agent.py:461-476) ----
    eval_rng = jax.random.PRNGKey(5)
    eval_rng, noise_rng = jax.random.split(eval_rng)
    out["rng/x_T"] = np.asarray(jax.random.normal(noise_rng, (2, 8, 25), dtype=jnp.float32))
    s_rng, eval_rng = jax.random.split(eval_rng)
    carry, keys = s_rng, []
    for i in range(4):
        k_i, carry = jax.random.split(carry)
        keys.append(np.asarray(k_i))
    out["rng/step_keys"] = np.stack(keys)
    out["rng/randint"] = np.asarray(jax.random.randint(jax.random.PRNGKey(9), (64,), 0, 100))
    out["rng/split3"] = np.asarray(jax.random.split(jax.random.PRNGKey(9), 3))

    # ---- optimiser (agent/ldp_agent.py:580-587) ----
    import optax
    sched = optax.warmup_cosine_decay_schedule(init_value=1e-6, peak_value=1e-4, warmup_steps=1000, decay_steps=500000, end_value=1e-6)
    out["optax/lr_steps"] = np.array([0, 1, 500, 1000, 1001, 250000, 499999, 500000, 600000])
    out["optax/lr"] = np.array([float(sched(int(i))) for i in out["optax/lr_steps"]])
    w, gr = jnp.asarray(g.standard_normal(16).astype(np.float32)), jnp.asarray(g.standard_normal(16).astype(np.float32))
    tx = optax.adam(1e-3)
    os_ = tx.init(w)
    up, os_ = tx.update(gr, os_, w)
    out["optax/adam_w"], out["optax/adam_g"], out["optax/adam_w1"] = np.asarray(w), np.asarray(gr), np.asarray(optax.apply_updates(w, up))
    np.savez_compressed(a.out, **out)
    print("wrote", a.out, "with", len(out), "entries")


if __name__ == "__main__":
    main()
