def uniform(*a, **k):
    raise NotImplementedError("flax_shim: parameters are supplied, never initialised")


normal = split = PRNGKey = uniform
