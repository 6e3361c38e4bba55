import numpy as np


def softplus(x):
    return np.logaddexp(x, 0.0)          # jax.nn.softplus = logaddexp(x, 0)


def relu(x):
    return np.maximum(x, 0.0)
