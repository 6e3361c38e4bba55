"""Host side of the jax.random compatibility layer (row N4, optional): key arithmetic of jax 0.4.26's default
`threefry2x32` generator in numpy, and the key threading of the reference's sampling loops, so that an agent handed a
raw JAX PRNG key `(k0, k1)` draws the noise the reference would.  Bulk draws run on the GPU (`ldp_jax_random`).

Keys are tiny (two words), so `split` stays on the host; only `normal` / `bits` over tensors go to the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from . import _native as N

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_M32 = 0xFFFFFFFF


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & _M32


def threefry2x32(k0: int, k1: int, x0: int, x1: int) -> Tuple[int, int]:
    """One Threefry-2x32-20 block on Python ints."""
    ks = (k0 & _M32, k1 & _M32, (k0 ^ k1 ^ 0x1BD11BDA) & _M32)
    x0, x1 = (x0 + ks[0]) & _M32, (x1 + ks[1]) & _M32
    for i in range(5):
        for r in _ROT[i % 2]:
            x0 = (x0 + x1) & _M32
            x1 = _rotl(x1, r) ^ x0
        x0 = (x0 + ks[(i + 1) % 3]) & _M32
        x1 = (x1 + ks[(i + 2) % 3] + i + 1) & _M32
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """jax.random.PRNGKey(seed)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & _M32], dtype=np.uint32)


def is_key(rng) -> bool:
    """True for a raw JAX key (two uint32 words) as opposed to the integer seeds of the Philox mode."""
    if isinstance(rng, (int, np.integer)):
        return False
    a = np.asarray(rng)
    return a.shape == (2,) and a.dtype.kind in "ui"


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split: threefry_2x32(key, iota(2 num)) - counters split into halves - reshaped to (num, 2)."""
    k0, k1 = int(key[0]), int(key[1])
    n = 2 * num
    half = n // 2
    out = [0] * n
    for p in range(half):
        out[p], out[half + p] = threefry2x32(k0, k1, p, half + p)
    return np.array(out, dtype=np.uint32).reshape(num, 2)


def sampling_keys(eval_rng, n_steps: int):
    """agent/ldp_agent.py:461-476 (planner) / :488-503 (IDM): returns (key of the start noise, the n_steps scheduler noise
    keys in iteration order, the eval_rng to carry on with).  FlaxDDPMScheduler.step draws from split(key, num=1)[0]."""
    eval_rng, noise_rng = split(eval_rng)
    carry, eval_rng = split(eval_rng)
    keys = []
    for _ in range(n_steps):
        s_rng, carry = split(carry)
        keys.append(split(s_rng, 1)[0])
    return noise_rng, np.stack(keys), eval_rng


def _draw(keys: np.ndarray, n: int, mode: int, device="cuda") -> torch.Tensor:
    keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint32).reshape(-1, 2))
    kd = torch.from_numpy(keys.view(np.int32).copy()).to(device)
    out = torch.empty((keys.shape[0], n), dtype=torch.float32 if mode == 1 else torch.int32, device=device)
    N.check(N.load().ldp_jax_random(kd.data_ptr(), keys.shape[0], n, mode, out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    return out


def normal(keys, n: int, device="cuda") -> torch.Tensor:
    """jax.random.normal(key, (n,)) for every key in `keys` ((2,) or (k,2)) -> (k, n) float32 on the GPU."""
    return _draw(keys, n, 1, device)


def bits(keys, n: int, device="cuda") -> torch.Tensor:
    """jax.random.bits(key, (n,), uint32) -> (k, n), returned as int32 bit patterns."""
    return _draw(keys, n, 0, device)


def randint(key, n: int, minval: int, maxval: int, device="cuda") -> torch.Tensor:
    """jax.random.randint(key, (n,), minval, maxval) int32: ((hi % span) * (2^32 % span) + lo % span) % span + minval with
    hi / lo drawn from the two halves of split(key)."""
    k1, k2 = split(key)
    hi = bits(k1, n, device)[0].to(torch.int64) & 0xFFFFFFFF
    lo = bits(k2, n, device)[0].to(torch.int64) & 0xFFFFFFFF
    span = max(int(maxval) - int(minval), 1)
    mult = ((1 << 16) % span) ** 2 % span
    return (((hi % span) * mult + lo % span) % span + int(minval)).to(torch.int32)


def update_keys(rng, use_planner: bool, use_idm: bool):
    """Key threading of `update_step` / `loss` / `plan_loss` / `idm_loss` (agent/ldp_agent.py:240, :143-155, :115, :133):
    returns {'planner': (t_key, noise_key), 'idm': (t_key, noise_key)} for the networks in use."""
    _, r = split(rng)                       # rng, g_rng = split(rng); the losses see g_rng
    out = {}
    if use_planner:
        r, net = split(r)
        _, t_key, z_key = split(net, 3)
        out["planner"] = (t_key, z_key)
    if use_idm:
        r, net = split(r)
        _, t_key, z_key = split(net, 3)
        out["idm"] = (t_key, z_key)
    return out
