"""Weight import (scope row N4, the part that can be written without the reference's stack): Flax msgpack files.

`flax.serialization.to_bytes / msgpack_restore` (flax 0.8.4) is the container of the Hugging Face Flax weights the
reference points at (`pcuenq/sd-vae-ft-mse-flax`, agent/dp_repr_agent.yaml:4: `diffusion_flax_model.msgpack`) and of
`flax.training.checkpoints`.  Restated from its published definition - FORMAT UNPINNED, no such file exists in this
image: a msgpack map of nested string-keyed maps whose leaves are ExtType(1) = packb((shape, dtype name, raw bytes)),
ExtType(3) numpy scalars in the same form, and arrays above 2**30 bytes split into `{"__msgpack_chunked_array__": True,
"shape": ..., "chunks": {"0": ..., ...}}`.  The reference's own training snapshots are orbax `PyTreeCheckpointer`
directories (train_bc.py:197-208), whose tensorstore layout is not restated here; this package writes `.npz` snapshots.
"""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Union

import msgpack
import numpy as np

from . import params as P

_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3


def _decode_array(data: bytes) -> np.ndarray:
    shape, dtype_name, buf = msgpack.unpackb(data, raw=False, strict_map_key=False)
    if dtype_name == "bfloat16":                       # numpy has no bfloat16: widen through the bit pattern
        u16 = np.frombuffer(buf, dtype=np.uint16).astype(np.uint32) << 16
        return u16.view(np.float32).reshape(shape).copy()
    return np.frombuffer(buf, dtype=np.dtype(dtype_name)).reshape(shape).copy()


def _ext_hook(code: int, data: bytes):
    if code == _EXT_NDARRAY:
        return _decode_array(data)
    if code == _EXT_NPSCALAR:
        return _decode_array(data)[()]
    if code == _EXT_COMPLEX:
        re, im = msgpack.unpackb(data)
        return complex(re, im)
    return msgpack.ExtType(code, data)


def _unchunk(tree):
    if isinstance(tree, dict):
        if tree.get("__msgpack_chunked_array__"):
            chunks = tree["chunks"]
            flat = np.concatenate([np.asarray(chunks[str(i)]).reshape(-1) for i in range(len(chunks))])
            shape = tree["shape"]               # flax stores tuples as {'0': d0, '1': d1, ...} (_tuple_to_dict)
            if isinstance(shape, dict):
                shape = tuple(int(shape[str(i)]) for i in range(len(shape)))
            return flat.reshape(tuple(shape))
        return {k: _unchunk(v) for k, v in tree.items()}
    return tree


def load_flax_msgpack(src: Union[str, Path, bytes]) -> Dict[str, Any]:
    """`flax.serialization.msgpack_restore`: nested dict of numpy arrays."""
    raw = src if isinstance(src, (bytes, bytearray)) else Path(src).read_bytes()
    return _unchunk(msgpack.unpackb(raw, ext_hook=_ext_hook, raw=False, strict_map_key=False))


def _encode(tree):
    if isinstance(tree, dict):
        return {k: _encode(v) for k, v in tree.items()}
    a = np.asarray(tree)
    return msgpack.ExtType(_EXT_NDARRAY, msgpack.packb((list(a.shape), a.dtype.name, a.tobytes()), use_bin_type=True))


def save_flax_msgpack(tree: Dict[str, Any], path=None) -> bytes:
    """`flax.serialization.to_bytes` for trees of numpy arrays (no chunking: arrays here are far below 2**30 bytes)."""
    raw = msgpack.packb(_encode(tree), use_bin_type=True)
    if path is not None:
        Path(path).write_bytes(raw)
    return raw


def select_params(tree: Dict[str, Any], spec, prefix: str = "") -> Dict[str, np.ndarray]:
    """The tensors `spec` names, out of a (possibly larger) Flax tree, as float32 in canonical naming.  `prefix` strips a
    leading scope (e.g. 'params').  Raises with the first missing / mis-shaped name."""
    flat = P.canonicalize_flax_names(P.unnest(tree) if any(isinstance(v, dict) for v in tree.values()) else dict(tree))
    if prefix:
        flat = {k[len(prefix) + 1:]: v for k, v in flat.items() if k.startswith(prefix + "/")}
    out = {}
    for name, shape in spec.items():
        if name not in flat:
            raise KeyError(f"checkpoint has no tensor {name!r} (it has {len(flat)} tensors, e.g. {sorted(flat)[:3]})")
        a = np.asarray(flat[name], dtype=np.float32)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name}: checkpoint shape {a.shape} != config shape {tuple(shape)}")
        out[name] = a
    return out


def load_vae_flax(path, block_out_channels=(128, 256, 512, 512), in_channels: int = 3, latent_channels: int = 4,
                  layers_per_block: int = 2):
    """(encoder params, decoder params) of a diffusers `FlaxAutoencoderKL` msgpack file (or the directory holding
    `diffusion_flax_model.msgpack`), split along this package's two specs."""
    path = Path(path)
    if path.is_dir():
        path = path / "diffusion_flax_model.msgpack"
    tree = load_flax_msgpack(path)
    if set(tree) == {"params"}:
        tree = tree["params"]
    enc = select_params(tree, P.vae_encoder_spec(block_out_channels, in_channels, latent_channels, layers_per_block))
    dec = select_params(tree, P.vae_decoder_spec(block_out_channels, in_channels, latent_channels, layers_per_block))
    return enc, dec
