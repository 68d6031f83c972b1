"""TEST INFRASTRUCTURE ONLY -- import shim that lets the *unmodified* reference
(/root/reference, shehryar-malik/icrl) be imported in the build container.

The reference imports `gym` and `matplotlib`, neither of which is installed
here (and there is no network).  This module registers just enough stub
modules in `sys.modules` for `stable_baselines3` (the vendored fork) and
`icrl.constraint_net` to import, then puts /root/reference on sys.path.

It is used by exactly one thing: `tests/golden/make_golden.py`, which runs the
reference on seeded inputs and commits the outputs as fixtures.  Nothing that
runs on the GPU box (tests -m gpu, smoke(), bench.py) may import this file:
/root/reference does not exist there.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("ICRL_REFERENCE_ROOT", "/root/reference")


class _Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)

    def contains(self, x):
        return True


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.asarray(low).shape
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def __repr__(self):
        return f"Box{self.shape}"


class Discrete(_Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return np.random.randint(self.n)

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(_Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)


class MultiBinary(_Space):
    def __init__(self, n):
        self.n = n
        super().__init__((n,), np.int8)


class Dict(_Space):
    def __init__(self, spaces=None):
        super().__init__(None, None)
        self.spaces = spaces or {}


class Tuple(_Space):
    def __init__(self, spaces=()):
        super().__init__(None, None)
        self.spaces = tuple(spaces)


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    observation_space = None
    action_space = None

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = env.observation_space
        self.action_space = env.action_space

    def __getattr__(self, name):
        return getattr(self.env, name)


class GoalEnv(Env):
    pass


def _flatdim(space):
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    raise NotImplementedError(space)


def install():
    """Register the stub modules and make the reference importable.  Idempotent."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_icrl_b200_shim", False):
        return
    gym = types.ModuleType("gym")
    gym._icrl_b200_shim = True
    gym.Env, gym.Wrapper, gym.GoalEnv, gym.Space = Env, Wrapper, GoalEnv, _Space
    gym.ObservationWrapper = gym.RewardWrapper = gym.ActionWrapper = Wrapper
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("gym.make is stubbed"))
    gym.logger = types.SimpleNamespace(warn=lambda *a, **k: None, set_level=lambda *a, **k: None)

    spaces = types.ModuleType("gym.spaces")
    for cls in (Box, Discrete, MultiDiscrete, MultiBinary, Dict, Tuple):
        setattr(spaces, cls.__name__, cls)
    spaces.Space = _Space
    spaces_utils = types.ModuleType("gym.spaces.utils")
    spaces_utils.flatdim = _flatdim
    spaces.utils = spaces_utils
    spaces.flatdim = _flatdim
    gym.spaces = spaces

    wrappers = types.ModuleType("gym.wrappers")
    monitoring = types.ModuleType("gym.wrappers.monitoring")
    video_recorder = types.ModuleType("gym.wrappers.monitoring.video_recorder")
    video_recorder.VideoRecorder = type("VideoRecorder", (), {})
    monitoring.video_recorder = video_recorder
    wrappers.monitoring = monitoring
    wrappers.TimeLimit = Wrapper
    gym.wrappers = wrappers

    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda *a, **k: None
    registration.EnvSpec = type("EnvSpec", (), {})
    envs.registration = registration
    gym.envs = envs

    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt

    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.spaces.utils": spaces_utils,
        "gym.wrappers": wrappers, "gym.wrappers.monitoring": monitoring,
        "gym.wrappers.monitoring.video_recorder": video_recorder,
        "gym.envs": envs, "gym.envs.registration": registration,
        "matplotlib": mpl, "matplotlib.pyplot": plt,
    })
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # torch >= 2.6 defaults torch.load(weights_only=True); the reference's
    # checkpoints hold numpy arrays (constraint_net.py:367 calls th.load(path)).
    import torch
    if not getattr(torch.load, "_icrl_b200_shim", False):
        _orig_load = torch.load

        def _load(*a, **k):
            k.setdefault("weights_only", False)
            return _orig_load(*a, **k)
        _load._icrl_b200_shim = True
        torch.load = _load


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "icrl"))
