"""`flax.linen` subset used by the reference networks, apply-only, numpy float64.

Module system: dataclass fields from annotations, `@compact` methods, submodules auto-named `<Class>_<i>` in construction
order under the module whose compact method is executing (Flax 0.8 `Module.__post_init__` / `_CallInfo` behaviour; also for
modules handed to `nn.Sequential([...])` inside a compact method - they already have a parent, so Sequential does not adopt
them).  `apply({"params": tree}, *args)` looks parameters up by that path in the nested `tree`.

Layer primitives are restated from Flax 0.8.4 / lax semantics as explicit index arithmetic (no torch, no scipy): they are the
independent second restatement of oracle/ldp_oracle.py's conv1d / downsample / upsample / GroupNorm / LayerNorm."""
import dataclasses
import functools
from typing import Any, Callable, Optional, Sequence, Tuple, Union

import numpy as np

from jax.nn import relu, softplus  # noqa: F401  (nn.relu / nn.softplus are re-exported by flax.linen)

_stack = []          # modules whose compact method is running (innermost last)
_params = [None]     # the tree given to the outermost apply()


def tanh(x):
    return np.tanh(x)


class _Initializers:
    @staticmethod
    def _noinit(*a, **k):
        def init(*aa, **kk):
            raise NotImplementedError("flax_shim: parameters are supplied, never initialised")
        return init
    xavier_uniform = he_uniform = normal = constant = zeros = ones = _noinit


initializers = _Initializers()


def compact(fn):
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        self._counters = {}
        _stack.append(self)
        try:
            return fn(self, *a, **k)
        finally:
            _stack.pop()
    return wrapped


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        dataclasses.dataclass(cls, eq=False, repr=False)

    def __post_init__(self):
        parent = _stack[-1] if _stack else None
        self._parent = parent
        self._counters = {}
        if parent is None:
            self._path = ()
        else:
            cname = type(self).__name__
            i = parent._counters.get(cname, 0)
            parent._counters[cname] = i + 1
            self._path = parent._path + (f"{cname}_{i}",)

    def param(self, name, init, *shape_and_dtype):
        node = _params[0]
        for key in self._path + (name,):
            if key not in node:
                raise KeyError("parameter tree has no '" + "/".join(self._path + (name,)) + "'")
            node = node[key]
        return np.asarray(node, dtype=np.float64)

    def apply(self, variables, *args, method=None, **kwargs):
        _params[0] = variables["params"]
        try:
            return (method or type(self).__call__)(self, *args, **kwargs)
        finally:
            _params[0] = None


class Sequential(Module):
    layers: Sequence[Any]

    def __call__(self, x, *a, **k):
        for layer in self.layers:
            x = layer(x)
        return x


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: Any = None
    bias_init: Any = None

    @compact
    def __call__(self, x):
        w = self.param("kernel", None)
        assert w.shape == (x.shape[-1], self.features), (w.shape, x.shape, self.features)
        y = x @ w
        return y + self.param("bias", None) if self.use_bias else y


def _same_pads(t_in, k, s):
    """lax padtype_to_pads('SAME'): out = ceil(in / stride); total = max((out-1)*stride + k - in, 0); lo = total // 2."""
    out = -(-t_in // s)
    total = max((out - 1) * s + k - t_in, 0)
    return total // 2, total - total // 2


class Conv(Module):
    """1-D nn.Conv over (B, T, C), kernel (k, C_in, C_out), cross-correlation: y[t] = sum_j x_pad[t*s + j] W[j] + b."""
    features: int
    kernel_size: Sequence[int] = (1,)
    strides: Union[None, int, Sequence[int]] = 1
    padding: Any = "SAME"
    use_bias: bool = True

    @compact
    def __call__(self, x):
        (k,) = tuple(self.kernel_size)
        s = self.strides if isinstance(self.strides, int) else (1 if self.strides is None else tuple(self.strides)[0])
        w = self.param("kernel", None)
        assert w.shape == (k, x.shape[-1], self.features), (w.shape, x.shape)
        B, T, _ = x.shape
        if isinstance(self.padding, str):
            lo, hi = _same_pads(T, k, s) if self.padding == "SAME" else (0, 0)
        elif isinstance(self.padding, int):
            lo = hi = self.padding
        else:
            lo, hi = tuple(self.padding)[0]
        xp = np.zeros((B, T + lo + hi, x.shape[-1]))
        xp[:, lo:lo + T] = x
        t_out = (T + lo + hi - k) // s + 1
        y = np.zeros((B, t_out, self.features))
        for t in range(t_out):
            for j in range(k):
                y[:, t] += xp[:, t * s + j] @ w[j]
        return y + self.param("bias", None) if self.use_bias else y


class ConvTranspose(Module):
    """nn.ConvTranspose(features, (k,), strides=(s,)), padding 'SAME', transpose_kernel=False (Flax default):
    lax.conv_transpose -> conv_general_dilated(lhs_dilation=s, window stride 1, padding from _conv_transpose_padding,
    kernel used as is - not flipped)."""
    features: int
    kernel_size: Sequence[int] = (1,)
    strides: Optional[Sequence[int]] = None
    padding: Any = "SAME"
    use_bias: bool = True

    @compact
    def __call__(self, x):
        (k,) = tuple(self.kernel_size)
        s = 1 if self.strides is None else tuple(self.strides)[0]
        assert self.padding == "SAME"
        w = self.param("kernel", None)
        assert w.shape == (k, x.shape[-1], self.features), (w.shape, x.shape)
        pad_len = k + s - 2                                    # jax.lax._conv_transpose_padding, 'SAME'
        pad_a = k - 1 if s > k - 1 else -(-pad_len // 2)
        pad_b = pad_len - pad_a
        B, T, C = x.shape
        xd = np.zeros((B, (T - 1) * s + 1 + pad_a + pad_b, C))
        xd[:, pad_a:pad_a + (T - 1) * s + 1:s] = x
        t_out = xd.shape[1] - k + 1
        y = np.zeros((B, t_out, self.features))
        for t in range(t_out):
            for j in range(k):
                y[:, t] += xd[:, t + j] @ w[j]
        return y + self.param("bias", None) if self.use_bias else y


class GroupNorm(Module):
    """nn.GroupNorm(num_groups): groups of C/G contiguous channels, statistics over every non-batch axis inside the group,
    epsilon 1e-6, use_fast_variance=True: var = max(E[x^2] - E[x]^2, 0)."""
    num_groups: int = 32
    epsilon: float = 1e-6

    @compact
    def __call__(self, x):
        C = x.shape[-1]
        G = self.num_groups
        out = np.empty_like(x)
        for b in range(x.shape[0]):
            for g in range(G):
                blk = x[b, ..., g * (C // G):(g + 1) * (C // G)]
                mu = blk.mean()
                var = max((blk * blk).mean() - mu * mu, 0.0)
                out[b, ..., g * (C // G):(g + 1) * (C // G)] = (blk - mu) / np.sqrt(var + self.epsilon)
        return out * self.param("scale", None) + self.param("bias", None)


class LayerNorm(Module):
    epsilon: float = 1e-6

    @compact
    def __call__(self, x):
        mu = x.mean(axis=-1, keepdims=True)
        var = np.maximum((x * x).mean(axis=-1, keepdims=True) - mu * mu, 0.0)
        return (x - mu) / np.sqrt(var + self.epsilon) * self.param("scale", None) + self.param("bias", None)


class Dropout(Module):
    rate: float = 0.0

    @compact
    def __call__(self, x, deterministic=True):
        assert deterministic or not self.rate
        return x
