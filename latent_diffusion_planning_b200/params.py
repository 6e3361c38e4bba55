"""Parameter trees in the reference's Flax layout, their canonical flat order, and synthetic init.

The reference keeps weights in Flax pytrees (`planner_state.params`, `idm_state.params`, `vae_params`;
reference agent/ldp_agent.py:508-514, :543-554).  This module defines

* the tree *spec* (path -> shape) for the three networks on the hot path, following the module
  structure of networks/diffusion_nets_v2.py:104-169 (ConditionalUnet1D), networks/mlp_diffusion_nets.py:8-68
  (MLPDiffusion) and diffusers' FlaxAutoencoderKL encoder (model/stable_vae_model.yaml:4-16);
* the canonical flat order that the C ABI (`include/ldp_b200.h`) expects - a plain float32 blob, the
  tensors concatenated in spec order, each in its Flax layout (Dense kernel (in,out); Conv kernel
  (k..., in, out));
* synthetic initialisers matching the reference's init distributions (xavier-uniform where the
  reference passes `default_init()`, LeCun-normal otherwise, GN/LN scale 1 / bias 0) plus an optional
  perturbation so that tests exercise biases and norm affine terms.

Everything here is host-side numpy; nothing touches the GPU.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

Spec = "OrderedDict[str, Tuple[int, ...]]"


# ----------------------------------------------------------------------------------------------
# Planner: ConditionalUnet1D (reference networks/diffusion_nets_v2.py:104-169)
# ----------------------------------------------------------------------------------------------
def unet_block_plan(input_dim: int, down_dims: Sequence[int]) -> List[Tuple[int, int, bool]]:
    """(c_in, c_out, residual_proj) for ConditionalResidualBlock1D_0..N-1 in creation order.

    down: two blocks per level (first projects), mid: two blocks, up: for every level but the last,
    reversed; the first up block consumes concat([x, skip]) (diffusion_nets_v2.py:134-159).
    """
    blocks = []
    c = input_dim
    for d in down_dims:
        blocks.append((c, d, True))
        blocks.append((d, d, False))
        c = d
    mid = down_dims[-1]
    blocks.append((mid, mid, False))
    blocks.append((mid, mid, False))
    skips = list(down_dims)
    c = mid
    for d in reversed(down_dims[:-1]):
        skip = skips.pop()
        blocks.append((c + skip, d, True))
        blocks.append((d, d, False))
        c = d
    return blocks


def unet_spec(input_dim: int, global_cond_dim: int, down_dims: Sequence[int] = (256, 512, 1024),
              kernel_size: int = 5, step_embed_dim: int = 256) -> Spec:
    dsed = step_embed_dim
    cond_dim = dsed + global_cond_dim
    spec: Spec = OrderedDict()
    spec["Dense_0/kernel"] = (dsed, dsed * 4)
    spec["Dense_0/bias"] = (dsed * 4,)
    spec["Dense_1/kernel"] = (dsed * 4, dsed)
    spec["Dense_1/bias"] = (dsed,)
    for i, (cin, cout, proj) in enumerate(unet_block_plan(input_dim, down_dims)):
        p = f"ConditionalResidualBlock1D_{i}"
        spec[f"{p}/Conv1dBlock_0/Conv_0/kernel"] = (kernel_size, cin, cout)
        spec[f"{p}/Conv1dBlock_0/Conv_0/bias"] = (cout,)
        spec[f"{p}/Conv1dBlock_0/GroupNorm_0/scale"] = (cout,)
        spec[f"{p}/Conv1dBlock_0/GroupNorm_0/bias"] = (cout,)
        spec[f"{p}/Dense_0/kernel"] = (cond_dim, 2 * cout)
        spec[f"{p}/Dense_0/bias"] = (2 * cout,)
        spec[f"{p}/Conv1dBlock_1/Conv_0/kernel"] = (kernel_size, cout, cout)
        spec[f"{p}/Conv1dBlock_1/Conv_0/bias"] = (cout,)
        spec[f"{p}/Conv1dBlock_1/GroupNorm_0/scale"] = (cout,)
        spec[f"{p}/Conv1dBlock_1/GroupNorm_0/bias"] = (cout,)
        if proj:
            spec[f"{p}/Conv_0/kernel"] = (1, cin, cout)
            spec[f"{p}/Conv_0/bias"] = (cout,)
    for i, d in enumerate(down_dims[:-1]):
        spec[f"Downsample1d_{i}/Conv_0/kernel"] = (3, d, d)
        spec[f"Downsample1d_{i}/Conv_0/bias"] = (d,)
    for i, d in enumerate(reversed(down_dims[:-1])):
        spec[f"Upsample1d_{i}/ConvTranspose_0/kernel"] = (4, d, d)
        spec[f"Upsample1d_{i}/ConvTranspose_0/bias"] = (d,)
    d0 = down_dims[0]
    spec["Conv1dBlock_0/Conv_0/kernel"] = (kernel_size, d0, d0)
    spec["Conv1dBlock_0/Conv_0/bias"] = (d0,)
    spec["Conv1dBlock_0/GroupNorm_0/scale"] = (d0,)
    spec["Conv1dBlock_0/GroupNorm_0/bias"] = (d0,)
    spec["Conv_0/kernel"] = (1, d0, input_dim)
    spec["Conv_0/bias"] = (input_dim,)
    return spec


# ----------------------------------------------------------------------------------------------
# IDM: MLPDiffusion(FourierFeatures -> MLP cond encoder -> MLPResNet)
# (reference networks/mlp_diffusion_nets.py:8-68, agent/ldp_agent.yaml:17-34)
# ----------------------------------------------------------------------------------------------
def idm_spec(obs_dim: int, action_dim: int, hidden_dim: int = 256, n_blocks: int = 3,
             time_dim: int = 256, cond_hidden: Sequence[int] = (256, 256)) -> Spec:
    spec: Spec = OrderedDict()
    c = time_dim
    for i, h in enumerate(cond_hidden):
        spec[f"MLP_0/Dense_{i}/kernel"] = (c, h)
        spec[f"MLP_0/Dense_{i}/bias"] = (h,)
        c = h
    in_dim = action_dim + 2 * obs_dim + c          # concat([a, s, cond]) mlp_diffusion_nets.py:66
    spec["MLPResNet_0/Dense_0/kernel"] = (in_dim, hidden_dim)
    spec["MLPResNet_0/Dense_0/bias"] = (hidden_dim,)
    for b in range(n_blocks):
        p = f"MLPResNet_0/MLPResNetBlock_{b}"
        spec[f"{p}/LayerNorm_0/scale"] = (hidden_dim,)
        spec[f"{p}/LayerNorm_0/bias"] = (hidden_dim,)
        spec[f"{p}/Dense_0/kernel"] = (hidden_dim, hidden_dim * 4)
        spec[f"{p}/Dense_0/bias"] = (hidden_dim * 4,)
        spec[f"{p}/Dense_1/kernel"] = (hidden_dim * 4, hidden_dim)
        spec[f"{p}/Dense_1/bias"] = (hidden_dim,)
    spec["MLPResNet_0/Dense_1/kernel"] = (hidden_dim, action_dim)
    spec["MLPResNet_0/Dense_1/bias"] = (action_dim,)
    return spec


# ----------------------------------------------------------------------------------------------
# VAE encoder (diffusers 0.27.2 FlaxAutoencoderKL.encode; un-vendored, see oracle header)
# ----------------------------------------------------------------------------------------------
def _resnet_spec(spec: Spec, p: str, cin: int, cout: int) -> None:
    spec[f"{p}/norm1/scale"] = (cin,)
    spec[f"{p}/norm1/bias"] = (cin,)
    spec[f"{p}/conv1/kernel"] = (3, 3, cin, cout)
    spec[f"{p}/conv1/bias"] = (cout,)
    spec[f"{p}/norm2/scale"] = (cout,)
    spec[f"{p}/norm2/bias"] = (cout,)
    spec[f"{p}/conv2/kernel"] = (3, 3, cout, cout)
    spec[f"{p}/conv2/bias"] = (cout,)
    if cin != cout:
        spec[f"{p}/conv_shortcut/kernel"] = (1, 1, cin, cout)
        spec[f"{p}/conv_shortcut/bias"] = (cout,)


def vae_encoder_spec(block_out_channels: Sequence[int] = (128, 256, 512, 512), in_channels: int = 3,
                     latent_channels: int = 4, layers_per_block: int = 2) -> Spec:
    spec: Spec = OrderedDict()
    c0 = block_out_channels[0]
    spec["encoder/conv_in/kernel"] = (3, 3, in_channels, c0)
    spec["encoder/conv_in/bias"] = (c0,)
    c = c0
    n = len(block_out_channels)
    for i, co in enumerate(block_out_channels):
        for j in range(layers_per_block):
            _resnet_spec(spec, f"encoder/down_blocks_{i}/resnets_{j}", c, co)
            c = co
        if i != n - 1:
            spec[f"encoder/down_blocks_{i}/downsamplers_0/conv/kernel"] = (3, 3, co, co)
            spec[f"encoder/down_blocks_{i}/downsamplers_0/conv/bias"] = (co,)
    _resnet_spec(spec, "encoder/mid_block/resnets_0", c, c)
    a = "encoder/mid_block/attentions_0"
    spec[f"{a}/group_norm/scale"] = (c,)
    spec[f"{a}/group_norm/bias"] = (c,)
    for name in ("query", "key", "value", "proj_attn"):
        spec[f"{a}/{name}/kernel"] = (c, c)
        spec[f"{a}/{name}/bias"] = (c,)
    _resnet_spec(spec, "encoder/mid_block/resnets_1", c, c)
    spec["encoder/conv_norm_out/scale"] = (c,)
    spec["encoder/conv_norm_out/bias"] = (c,)
    spec["encoder/conv_out/kernel"] = (3, 3, c, 2 * latent_channels)
    spec["encoder/conv_out/bias"] = (2 * latent_channels,)
    spec["quant_conv/kernel"] = (1, 1, 2 * latent_channels, 2 * latent_channels)
    spec["quant_conv/bias"] = (2 * latent_channels,)
    return spec


def vae_decoder_spec(block_out_channels: Sequence[int] = (128, 256, 512, 512), out_channels: int = 3,
                     latent_channels: int = 4, layers_per_block: int = 2) -> Spec:
    """diffusers 0.27.2 FlaxAutoencoderKL.decode: post_quant_conv + FlaxDecoder (un-vendored, restated from its published
    structure): conv_in (L -> C_last), mid block, up blocks over reversed(block_out_channels) with layers_per_block + 1
    resnets each and a nearest x2 upsampler + 3x3 conv on all but the last, GroupNorm, swish, conv_out (C_0 -> 3)."""
    spec: Spec = OrderedDict()
    spec["post_quant_conv/kernel"] = (1, 1, latent_channels, latent_channels)
    spec["post_quant_conv/bias"] = (latent_channels,)
    rev = list(reversed(block_out_channels))
    c = rev[0]
    spec["decoder/conv_in/kernel"] = (3, 3, latent_channels, c)
    spec["decoder/conv_in/bias"] = (c,)
    _resnet_spec(spec, "decoder/mid_block/resnets_0", c, c)
    a = "decoder/mid_block/attentions_0"
    spec[f"{a}/group_norm/scale"] = (c,)
    spec[f"{a}/group_norm/bias"] = (c,)
    for name in ("query", "key", "value", "proj_attn"):
        spec[f"{a}/{name}/kernel"] = (c, c)
        spec[f"{a}/{name}/bias"] = (c,)
    _resnet_spec(spec, "decoder/mid_block/resnets_1", c, c)
    n = len(rev)
    for i, co in enumerate(rev):
        for j in range(layers_per_block + 1):
            _resnet_spec(spec, f"decoder/up_blocks_{i}/resnets_{j}", c, co)
            c = co
        if i != n - 1:
            spec[f"decoder/up_blocks_{i}/upsamplers_0/conv/kernel"] = (3, 3, co, co)
            spec[f"decoder/up_blocks_{i}/upsamplers_0/conv/bias"] = (co,)
    spec["decoder/conv_norm_out/scale"] = (c,)
    spec["decoder/conv_norm_out/bias"] = (c,)
    spec["decoder/conv_out/kernel"] = (3, 3, c, out_channels)
    spec["decoder/conv_out/bias"] = (out_channels,)
    return spec


# ----------------------------------------------------------------------------------------------
# init / flatten
# ----------------------------------------------------------------------------------------------
_XAVIER_SUFFIXES = ("Dense_0/kernel", "Dense_1/kernel")


def _fans(shape: Tuple[int, ...]) -> Tuple[int, int]:
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    return shape[-2] * receptive, shape[-1] * receptive


def init_params(spec: Spec, seed: int = 0, perturb: float = 0.02, xavier_dense: bool = True) -> Dict[str, np.ndarray]:
    """Synthetic float32 weights: Dense kernels xavier-uniform (the reference passes
    `kernel_init=default_init()` to every Dense on the hot path except the MLPResNet block Denses, which
    use Flax's LeCun-normal default - both are O(1/sqrt(fan)) so activations stay O(1)), conv kernels
    LeCun-normal, norm scale 1, biases 0; then `perturb`*N(0,1) is added to biases and norm affine
    terms so parity tests see them (SURVEY.md section 8d synthetic inputs)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = OrderedDict()
    for path, shape in spec.items():
        leaf = path.rsplit("/", 1)[-1]
        if leaf == "kernel":
            fan_in, fan_out = _fans(shape)
            is_dense = len(shape) == 2
            if is_dense and xavier_dense:
                lim = np.sqrt(6.0 / (fan_in + fan_out))
                w = rng.uniform(-lim, lim, size=shape)
            else:
                w = rng.standard_normal(shape) * np.sqrt(1.0 / fan_in)
        elif leaf == "scale":
            w = 1.0 + perturb * rng.standard_normal(shape)
        else:  # bias
            w = perturb * rng.standard_normal(shape)
        out[path] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def flatten_params(spec: Spec, params: Dict[str, np.ndarray], strict: bool = True) -> np.ndarray:
    """Concatenate tensors in canonical (spec) order into the float32 blob the C ABI takes.  `strict`: a tree with
    tensors the spec does not know (e.g. a 4-block IDM handed to a 3-block spec) is an error, not silently truncated."""
    missing = [k for k in spec if k not in params]
    if missing:
        raise KeyError(f"parameter tree lacks {len(missing)} tensors of the spec, first: {missing[:3]}")
    if strict:
        extra = [k for k in params if k not in spec]
        if extra:
            raise ValueError(f"parameter tree has {len(extra)} tensors that the network spec does not contain "
                             f"(wrong n_blocks / down_dims / block_out_channels?), first: {extra[:3]}")
    chunks = []
    for path, shape in spec.items():
        a = np.asarray(params[path], dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"param {path}: shape {a.shape} != spec {shape}")
        chunks.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(chunks))


def spec_size(spec: Spec) -> int:
    return int(sum(int(np.prod(s)) for s in spec.values()))


def nest(params: Dict[str, np.ndarray]) -> dict:
    """'a/b/c' flat keys -> nested dict (the Flax pytree shape `get_params()` returns)."""
    root: dict = {}
    for path, v in params.items():
        d = root
        parts = path.split("/")
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return root


def unnest(tree: dict, prefix: str = "") -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = OrderedDict()
    for k, v in tree.items():
        path = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(unnest(v, path))
        else:
            out[path] = np.asarray(v)
    return out


def canonicalize_flax_names(params: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Accept the alternative Flax registration `.../Sequential_k/layers_i/...` for modules created inside
    `nn.Sequential([...])` in a compact method (SURVEY.md section 8a, A2 detail): strip those two path
    components so both spellings map onto the canonical spec names."""
    out: Dict[str, np.ndarray] = OrderedDict()
    for path, v in params.items():
        parts = [p for p in path.split("/") if not (p.startswith("Sequential_") or p.startswith("layers_"))]
        out["/".join(parts)] = v
    return out
