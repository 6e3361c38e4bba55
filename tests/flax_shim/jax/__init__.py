"""Stand-in for the parts of `jax` the reference's network files touch at import / apply time (see ../README.md)."""
from . import numpy  # noqa: F401
from . import nn, random  # noqa: F401
