"""A deterministic scripted environment shared by make_golden.py (driving the reference's VecEnv wrappers) and
tests/test_vec_env.py (driving icrl_b200.vec_env): observations, rewards and episode ends come from a seeded RNG and
do not depend on the action, so both sides see the same stream."""
import numpy as np


class _Box:
    def __init__(self, shape):
        self.shape, self.dtype = shape, np.dtype(np.float32)
        self.low, self.high = -np.ones(shape, np.float32), np.ones(shape, np.float32)


class ScriptedEnv:
    metadata = {"render.modes": []}
    reward_range = (-np.inf, np.inf)
    spec = None

    def __init__(self, seed, obs_dim=5, act_dim=2, mean_len=17):
        self.observation_space, self.action_space = _Box((obs_dim,)), _Box((act_dim,))
        self.rng = np.random.RandomState(seed)
        self.scale = self.rng.uniform(0.2, 6.0, obs_dim)
        self.mean_len = mean_len

    def seed(self, seed=None):
        return [seed]

    def reset(self):
        return (self.rng.standard_normal(self.observation_space.shape) * self.scale + 1.5).astype(np.float32)

    def step(self, action):
        obs = (self.rng.standard_normal(self.observation_space.shape) * self.scale + 1.5).astype(np.float32)
        reward = float(self.rng.standard_normal() * 3 + 0.5)
        done = bool(self.rng.uniform() < 1.0 / self.mean_len)
        return obs, reward, done, {}

    def close(self):
        pass


def scripted_cost(obs, acs):
    """Stand-in for ConstraintNet.cost_function in the wrapper test (pure numpy, float32)."""
    return (1.0 / (1.0 + np.exp(-(obs[:, 0] * 0.3 + acs[:, 0])))).astype(np.float32)


def scripted_actions(seed, steps, n_envs, act_dim=2):
    return np.random.RandomState(seed).uniform(-1, 1, (steps, n_envs, act_dim)).astype(np.float32)
