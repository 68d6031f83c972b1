"""Minimal observation/action space descriptors (the reference takes gym.spaces.Box / Discrete; gym is a host-side
dependency of the env layer, which is out of scope here).  Any object with the same attributes works: `shape` and
`dtype` for both, `n` for discrete, `low`/`high` for boxes -- so real gym spaces can be passed unchanged."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shape = tuple(shape) if shape is not None else np.asarray(low).shape
        self.shape, self.dtype = shape, np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()

    def __repr__(self):
        return f"Box{self.shape}"


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.dtype(np.int64)

    def __repr__(self):
        return f"Discrete({self.n})"


def is_discrete(space) -> bool:
    return hasattr(space, "n") and not hasattr(space, "low")


def get_obs_shape(space):
    """stable_baselines3/common/preprocessing.py get_obs_shape for Box / Discrete."""
    return (1,) if is_discrete(space) else tuple(space.shape)


def get_action_dim(space) -> int:
    """stable_baselines3/common/preprocessing.py get_action_dim: Box -> prod(shape), Discrete -> 1."""
    return 1 if is_discrete(space) else int(np.prod(space.shape))
