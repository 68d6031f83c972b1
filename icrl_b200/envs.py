"""Environment factory for the drivers (`-tei / -eei`).

The reference builds its environments with ``gym.make`` on its MuJoCo / grid-world package ``custom_envs``
(icrl/utils.py:247-262).  Simulation is host-side work outside the learner hot path (SURVEY §8 "out of scope"), and
neither gym nor MuJoCo is part of this image, so:

  * if ``gym`` is importable the id is handed to ``gym.make`` unchanged (real environments keep working);
  * ids starting with ``Synth`` are small numpy environments with the *shapes* of the README workloads, so that
    ``python run_me.py icrl|cpg`` runs the whole outer loop -- rollout collection, K1 relabel through the cost
    wrapper, K3, K4, K2 -- on a box that has nothing but this repo.

Synthetic ids:  SynthHCWithPos-v0 / SynthHCWithPosTest-v0 (18-d obs, 6-d Box actions, obs[0] is the x position; the
Test variant ends the episode behind the wall at x <= -3, like HCWithPosTest-v0), SynthLGW-v0 / SynthCLGW-v0 (1-d obs,
2 discrete actions on a ring, like the lap grid world).
"""
import numpy as np

from icrl_b200.spaces import Box, Discrete


class _Spec:
    def __init__(self, env_id, max_episode_steps):
        self.id, self.max_episode_steps = env_id, max_episode_steps


class SynthHCWithPos:
    """A damped 17-d linear system driven by the 6 actions plus an x coordinate integrating the first velocity.
    Reward = |dx| with a small bonus for moving backwards (so the unconstrained optimum crosses the wall at -3, the
    situation the HalfCheetah experiment sets up), minus a control cost."""

    def __init__(self, env_id="SynthHCWithPos-v0", terminate_behind=None, max_episode_steps=200, seed=0):
        self.spec = _Spec(env_id, max_episode_steps)
        self.observation_space = Box(-np.inf, np.inf, shape=(18,))
        self.action_space = Box(-1.0, 1.0, shape=(6,))
        self.terminate_behind = terminate_behind
        sys_rng = np.random.default_rng(1234)                       # fixed dynamics, independent of the episode seed
        self.A = (0.9 * np.linalg.qr(sys_rng.standard_normal((17, 17)))[0]).astype(np.float64)
        self.B = (0.3 * sys_rng.standard_normal((17, 6))).astype(np.float64)
        self.rng = np.random.default_rng(seed)
        self.x, self.s, self.t = 0.0, np.zeros(17), 0

    def seed(self, seed=None):
        self.rng = np.random.default_rng(seed)
        return [seed]

    def _obs(self):
        return np.concatenate([[self.x], self.s]).astype(np.float32)

    def reset(self):
        self.x, self.t = 0.0, 0
        self.s = 0.1 * self.rng.standard_normal(17)
        return self._obs()

    def step(self, action):
        a = np.clip(np.asarray(action, dtype=np.float64).reshape(6), -1.0, 1.0)
        self.s = self.A @ self.s + self.B @ a + 0.01 * self.rng.standard_normal(17)
        dx = 0.1 * float(np.tanh(self.s[0]))
        self.x += dx
        self.t += 1
        reward = abs(dx) * 10.0 + (0.2 if dx < 0 else 0.0) - 0.05 * float(a @ a)
        done = self.t >= self.spec.max_episode_steps
        if self.terminate_behind is not None and self.x <= self.terminate_behind:
            done = True
        return self._obs(), reward, done, {"xpos": self.x}


class SynthLGW:
    """A ring of 11 cells, actions {0: left, 1: right}; reward 1 for completing a lap in either direction.  The
    constrained variant (SynthCLGW-v0) ends the episode when the agent moves anticlockwise through cell 0."""

    def __init__(self, env_id="SynthLGW-v0", constrained=False, max_episode_steps=200, seed=0):
        self.spec = _Spec(env_id, max_episode_steps)
        self.observation_space = Box(0.0, 10.0, shape=(1,))
        self.action_space = Discrete(2)
        self.constrained = constrained
        self.rng = np.random.default_rng(seed)
        self.pos, self.t = 5, 0

    def seed(self, seed=None):
        self.rng = np.random.default_rng(seed)
        return [seed]

    def reset(self):
        self.pos, self.t = 5, 0
        return np.array([self.pos], dtype=np.float32)

    def step(self, action):
        a = int(np.asarray(action).reshape(-1)[0])
        new = (self.pos + (1 if a == 1 else -1)) % 11
        lap = (self.pos == 10 and new == 0) or (self.pos == 0 and new == 10)
        backwards = self.pos == 0 and new == 10
        self.pos, self.t = new, self.t + 1
        done = self.t >= self.spec.max_episode_steps or (self.constrained and backwards)
        return np.array([self.pos], dtype=np.float32), float(lap), done, {}


_SYNTH = {
    "SynthHCWithPos-v0": lambda **kw: SynthHCWithPos("SynthHCWithPos-v0", None, **kw),
    "SynthHCWithPosTest-v0": lambda **kw: SynthHCWithPos("SynthHCWithPosTest-v0", -3.0, **kw),
    "SynthLGW-v0": lambda **kw: SynthLGW("SynthLGW-v0", False, **kw),
    "SynthCLGW-v0": lambda **kw: SynthLGW("SynthCLGW-v0", True, **kw),
}


def make(env_id: str, **kwargs):
    if env_id in _SYNTH:
        return _SYNTH[env_id](**kwargs)
    try:
        import gym
    except ImportError as e:
        raise ImportError(f"environment {env_id!r} needs gym (and the reference's custom_envs package), which is not "
                          f"installed; the built-in ids are {sorted(_SYNTH)}") from e
    try:
        import custom_envs  # noqa: F401  (registers the reference's ids when present)
    except ImportError:
        pass
    return gym.make(env_id, **kwargs)
