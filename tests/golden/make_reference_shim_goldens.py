"""Goldens produced by EXECUTING THE REFERENCE'S OWN network source (not a restatement of it):

    python tests/golden/make_reference_shim_goldens.py        # needs /root/reference (this container only)

`/root/reference/networks/{diffusion_nets_v2,mlp_diffusion_nets,diffusion,mlp_nets}.py` are imported unmodified with
`tests/flax_shim` standing in for jax / flax (see tests/flax_shim/README.md for exactly what that shim restates), and
`ConditionalUnet1D` / `MLPDiffusion` are applied - through the reference's `module.apply({"params": tree}, ...)` call
(agent/ldp_agent.py:123, :137, :470, :497) - to seeded inputs with parameter trees drawn by `params.init_params` under the
Flax parameter names.  Outputs go to tests/golden/ref_shim_goldens.npz; tests/test_reference_shim.py checks the oracle (CPU)
and the CUDA fp32 / bf16 paths (GPU) against them.  Weights are regenerated from their seeds, only inputs/outputs are stored.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
CASES = dict(
    unet_small=dict(kind="unet", D=25, dims=(64, 128, 256), B=3, T=8, seed=0, steps=[0, 50, 99]),
    unet_t16=dict(kind="unet", D=12, dims=(32, 64, 128), B=2, T=16, seed=3, steps=[7]),
    unet_full=dict(kind="unet", D=265, dims=(256, 512, 1024), B=2, T=8, seed=0, steps=[50]),
    unet_rowsteps=dict(kind="unet", D=25, dims=(64, 128, 256), B=4, T=8, seed=5, steps=[[0, 13, 50, 99]]),
    idm_rm=dict(kind="idm", D=265, A=7, N=6, seed=1, steps=[0, 10, 99]),
    idm_aloha=dict(kind="idm", D=270, A=14, N=5, seed=2, steps=[42]),
    idm_rowsteps=dict(kind="idm", D=25, A=7, N=4, seed=4, steps=[[3, 99, 0, 57]]),
)


def case_inputs(c):
    g = np.random.default_rng(1000 + c["seed"])
    if c["kind"] == "unet":
        return dict(x=g.standard_normal((c["B"], c["T"], c["D"])), cond=g.uniform(-1, 1, (c["B"], c["D"])))
    return dict(s=g.uniform(-1, 1, (c["N"], 2 * c["D"])), a=g.standard_normal((c["N"], c["A"])))


def case_params(c):
    sys.path.insert(0, str(ROOT))
    from latent_diffusion_planning_b200 import params as P
    spec = P.unet_spec(c["D"], c["D"], c["dims"]) if c["kind"] == "unet" else P.idm_spec(c["D"], c["A"])
    return P.init_params(spec, seed=c["seed"], perturb=0.1)


def main(out_path=None, only=None):
    if not REF.exists():
        raise SystemExit("/root/reference is not mounted here")
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT / "tests" / "flax_shim"))             # jax / flax resolve to the shim
    from functools import partial
    from latent_diffusion_planning_b200 import params as P
    from networks.diffusion import FourierFeatures
    from networks.diffusion_nets_v2 import ConditionalUnet1D
    from networks.mlp_diffusion_nets import MLPDiffusion, MLPResNet
    from networks.mlp_nets import MLP
    import flax
    assert "flax_shim" in flax.__file__
    out = {}
    for name, c in CASES.items():
        if only and name not in only:
            continue
        tree = P.nest(case_params(c))
        inp = case_inputs(c)
        for k, v in inp.items():
            out[f"{name}/{k}"] = v
        for i, step in enumerate(c["steps"]):
            if c["kind"] == "unet":
                # agent/ldp_agent.py:571-578 builds the planner from agent/ldp_agent.yaml:7-15
                net = ConditionalUnet1D(input_dim=c["D"], global_cond_dim=c["D"], diffusion_step_embed_dim=256,
                                        down_dims=tuple(c["dims"]), kernel_size=5, n_groups=8, downsample=True)
                t = np.asarray(step) if isinstance(step, list) else step
                y = net.apply({"params": tree}, inp["x"], t, inp["cond"])
            else:
                # agent/ldp_agent.py:601-606 with agent/ldp_agent.yaml:17-34
                idm = MLPDiffusion(partial(MLP, hidden_dims=(256, 256), activations="mish", activate_final=False),
                                   partial(MLPResNet, n_blocks=3, out_dim=c["A"], dropout_rate=None, use_layer_norm=True, hidden_dim=256),
                                   partial(FourierFeatures, output_size=256, learnable=False))
                t = np.asarray(step).reshape(-1, 1) if isinstance(step, list) else step
                y = idm.apply({"params": tree}, inp["s"], inp["a"], t)
            out[f"{name}/out_{i}"] = np.asarray(y, dtype=np.float64)
            print(f"{name} step {step}: out {y.shape} max|y| {np.abs(y).max():.4f}")
    out_path = Path(out_path) if out_path else ROOT / "tests" / "golden" / "ref_shim_goldens.npz"
    np.savez_compressed(out_path, **out)
    print(f"wrote {out_path}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, sys.argv[2].split(",") if len(sys.argv) > 2 else None)
