"""`jax.numpy` as plain numpy: the reference's network code only uses calls that exist under the same names with the same
semantics (array, broadcast_to, concatenate, split, expand_dims, arange, sin, cos, tanh).  Arithmetic runs in float64 (the
"truth" the oracle is compared at), with one exception that mirrors JAX's float32 default where it is visible at 1e-6: `log`
and `exp` - used by the reference only to build the sinusoid / Fourier frequency tables from Python ints
(diffusion_nets_v2.py:27-28, diffusion.py:19-20) - evaluate in float32, as `jnp.log(10000)` and `jnp.exp(int32 * f32)` do."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import ndarray, pi, float32  # noqa: F401


def log(x):
    return _np.log(_np.float32(x)) if _np.isscalar(x) else _np.log(_np.asarray(x, dtype=_np.float32))


def exp(x):
    return _np.exp(_np.asarray(x, dtype=_np.float32))
