"""Stand-in for the parts of `flax` the reference's network files touch (see ../README.md)."""
from . import linen  # noqa: F401
