"""Host mirror of the reference's `process_sdvae_data.py` (Workspace.run_rm / run_aloha, lines 52-118 / 121-175):
walk the demonstrations of a robomimic / ALOHA dataset, encode every RGB frame with the SD-VAE encoder
(`latent_dist.mean`, pixels 0..255 -> /255 -> (x - 0.5) / 0.5, lines 87-90), and write the latent dataset
`data/<demo>/latent/<rgb_key>` float32 (T_ep [+1], h, w, 4) with the attributes `total`, `min_z`, `max_z`
(lines 57-59, 114-118) that `data/robomimic_latent_data.py:95,127-147` reads back for min/max normalisation.

Differences, all forced by this image: the VAE runs through `handles.VaeEncoder` (hand-written sm_100a kernels, no
JAX); h5py is not installed, so the default container is an `.npz` whose keys are the HDF5 paths
(`data/<demo>/latent/<key>`, `data.attrs/total`, `data.attrs/min_z`, `data.attrs/max_z`) - when h5py is importable
the same function writes `latent.hdf5` in the reference's layout.  The reference pads the last shard of an episode
with zero images so that XLA sees one shape (lines 95-107); our kernels take any batch size, so no padding is done
(images are independent: same latents).  Hydra/wandb/orbax plumbing is out of scope.

Resize: with `pretrain_path` set the reference bilinearly resizes every frame to 3x256x256 before encoding
(`jax.image.resize(..., method="bilinear")`, lines 66-69 / 91-92).  `resize_to=S` does the same here (`resize_bilinear`:
half-pixel centres, no antialiasing when upsampling - jax.image.resize's semantics) on the normalised float frames.  The
sm_100a encoder kernels take square inputs up to 128 pixels, so `resize_to=256` raises instead of silently producing latents
of a different geometry; latents of a 64-pixel encode (8x8x4) are NOT interchangeable with the reference's 256-pixel
ones (32x32x4) and `LDPAgent.create(vae_pretrain_path=...)` says so.
"""
from __future__ import annotations

from pathlib import Path
from typing import Callable, Dict, Iterable, Mapping, Optional, Sequence

import numpy as np


def episode_frames(ep: Mapping, rgb_key: str, data_name: str) -> np.ndarray:
    """Frames of one episode for one camera, as the reference assembles them: robomimic ("rm") appends the last
    `next_obs` frame to `obs` (lines 74-79); ALOHA uses `obs` only (line 150)."""
    obs = np.asarray(ep["obs"][rgb_key])
    if "rm" in data_name:
        last = np.asarray(ep["next_obs"][rgb_key])[-1]
        obs = np.concatenate([obs, last[None]], axis=0)
    elif "aloha" not in data_name:
        raise ValueError(f"data_name must contain 'rm' or 'aloha' (reference run(), lines 47-51): {data_name!r}")
    if obs.ndim != 4 or obs.shape[-1] != 3:
        raise ValueError(f"{rgb_key}: expected (T, H, W, 3) frames, got {obs.shape}")
    return obs


def encode_dataset(episodes: Mapping[str, Mapping], rgb_keys: Sequence[str], encode: Callable[[np.ndarray], np.ndarray],
                   data_name: str = "rm_lift", shard: int = 256) -> Dict[str, np.ndarray]:
    """Pure host logic (CPU-testable): returns {hdf5 path: array} for the whole latent dataset.
    `encode(frames uint8 (n, H, W, 3)) -> (n, h, w, 4) float32` is the VAE call; frames go through it `shard` at a time."""
    if shard <= 0:
        raise ValueError("shard must be positive")
    out: Dict[str, np.ndarray] = {}
    min_z, max_z = 0.0, 0.0                                   # the reference starts both at 0 (line 71)
    for name, ep in episodes.items():
        for key in rgb_keys:
            frames = episode_frames(ep, key, data_name)
            if frames.dtype != np.uint8:
                frames = np.clip(np.rint(frames), 0, 255).astype(np.uint8)     # reference casts to float32 of 0..255 values
            zs = [np.asarray(encode(frames[i:i + shard]), dtype=np.float32) for i in range(0, len(frames), shard)]
            z = np.concatenate(zs, axis=0) if zs else np.zeros((0,), np.float32)
            if z.shape[0] != frames.shape[0]:
                raise RuntimeError(f"encoder returned {z.shape[0]} latents for {frames.shape[0]} frames")
            if z.size:
                min_z, max_z = min(min_z, float(z.min())), max(max_z, float(z.max()))
            out[f"data/{name}/latent/{key}"] = z
    out["data.attrs/total"] = np.asarray(len(episodes), dtype=np.int64)
    out["data.attrs/min_z"] = np.asarray(min_z, dtype=np.float32)
    out["data.attrs/max_z"] = np.asarray(max_z, dtype=np.float32)
    return out


def write_latents(tables: Mapping[str, np.ndarray], out_dir, prefer_hdf5: bool = True) -> Path:
    """`latent.hdf5` in the reference's layout when h5py exists, else `latent.npz` keyed by the HDF5 paths."""
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    h5py = None
    if prefer_hdf5:
        try:
            import h5py  # type: ignore
        except Exception:
            h5py = None
    if h5py is not None:
        path = out_dir / "latent.hdf5"
        with h5py.File(path, "w") as f:
            grp = f.create_group("data")
            for k, v in tables.items():
                if k.startswith("data.attrs/"):
                    grp.attrs[k.split("/", 1)[1]] = v
                else:
                    f.create_dataset(k, data=v)
        return path
    path = out_dir / "latent.npz"
    np.savez(path, **{k: v for k, v in tables.items()})
    return path


def read_latent_stats(path) -> Dict[str, float]:
    """total / min_z / max_z of a latent file written by write_latents (what robomimic_latent_data.py:95 reads)."""
    path = Path(path)
    if path.suffix == ".npz":
        with np.load(path) as z:
            return {k: z[f"data.attrs/{k}"].item() for k in ("total", "min_z", "max_z")}
    import h5py  # type: ignore
    with h5py.File(path, "r") as f:
        return {k: f["data"].attrs[k].item() for k in ("total", "min_z", "max_z")}


def resize_bilinear(frames, size: int):
    """`jax.vmap(jax.image.resize(obs, (3, S, S), "bilinear"))` of the reference (process_sdvae_data.py:66-69): separable
    triangle kernel on half-pixel centres; jax antialiases only when DOWN-sampling, which torch's `antialias=True` matches.
    frames: float tensor (n, H, W, 3) NHWC -> (n, S, S, 3)."""
    import torch
    import torch.nn.functional as F
    x = frames.permute(0, 3, 1, 2)
    down = size < x.shape[-1] or size < x.shape[-2]
    y = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=down)
    return y.permute(0, 2, 3, 1).contiguous()


def process_sdvae_data(episodes: Mapping[str, Mapping], rgb_keys: Sequence[str], vae, out_dir, data_name: str = "rm_lift",
                       shard: int = 256, precision: str = "bf16", device: str = "cuda", resize_to: Optional[int] = None) -> Path:
    """The reference's entry point on our encoder: `vae` is a `handles.VaeEncoder` (raises if the CUDA library or a GPU
    is missing - there is no CPU fallback).  `resize_to`: the reference's pretrain_path resize (see the module docstring)."""
    import torch
    if resize_to is not None and int(resize_to) != vae.image_size:
        raise ValueError(f"resize_to={resize_to} but the encoder handle was built for {vae.image_size}-pixel inputs "
                         f"(the kernels support square inputs up to 128 pixels; the reference's 256 is not available)")

    def encode(frames: np.ndarray) -> np.ndarray:
        img = torch.from_numpy(np.ascontiguousarray(frames)).to(device)
        if resize_to is not None and img.shape[1] != resize_to:
            img = resize_bilinear((img.to(torch.float32) / 255 - 0.5) / 0.5, int(resize_to))   # lines 89-92
        elif img.shape[1] != vae.image_size:
            raise ValueError(f"frames are {img.shape[1]} pixels, the encoder takes {vae.image_size}: pass resize_to=")
        return vae.encode(img, precision=precision).cpu().numpy()          # raw latent_dist.mean, no min/max normalisation

    return write_latents(encode_dataset(episodes, rgb_keys, encode, data_name, shard), out_dir)
