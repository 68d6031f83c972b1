"""Host-side vectorised-environment plumbing either side of the learner hot path (SURVEY §8 row b callers).

These mirror the reference classes the ICRL / CPG drivers touch -- same names, argument meaning and statistics -- so
that ``VecCostWrapper`` can hand each step's (previous obs, action) batch to ``ConstraintNet.cost_function`` (K1) and
``VecNormalizeWithCost`` can feed ``RolloutBufferWithCost`` the normalised rewards/costs it expects:

  * RunningMeanStd            stable_baselines3/common/running_mean_std.py:6-43
  * DummyVecEnv               stable_baselines3/common/vec_env/dummy_vec_env.py (the reference drivers use the
                              subprocess flavour; simulation is out of scope here, the step/auto-reset contract is
                              the same)
  * VecCostWrapper            stable_baselines3/common/vec_env/vec_cost_wrapper.py:9-66
  * VecNormalize(WithCost)    stable_baselines3/common/vec_env/vec_normalize.py:11-282
  * sync_envs_normalization   stable_baselines3/common/vec_env/__init__.py

Environment simulation itself stays on the host: it is not part of the hot path.
"""
import pickle
import time
from copy import deepcopy
from typing import Callable, List, Optional, Sequence

import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon: float = 1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr: np.ndarray) -> None:
        self.update_from_moments(np.mean(arr, axis=0), np.var(arr, axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count) -> None:
        # Chan et al. parallel-variance merge, in the reference's operation order (bit-exact in f64).
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, m_2 / (self.count + batch_count), batch_count + self.count


class VecEnv:
    def __init__(self, num_envs: int, observation_space, action_space):
        self.num_envs, self.observation_space, self.action_space = num_envs, observation_space, action_space

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_async(self, actions): raise NotImplementedError
    def step_wait(self): raise NotImplementedError
    def reset(self): raise NotImplementedError
    def close(self): pass
    def seed(self, seed: Optional[int] = None): return [None] * self.num_envs
    def get_attr(self, name, indices=None): raise NotImplementedError

    @property
    def unwrapped(self):
        return self.venv.unwrapped if isinstance(self, VecEnvWrapper) else self


class DummyVecEnv(VecEnv):
    """In-process vector of environments with the reference's auto-reset contract: when an episode ends the returned
    observation is the first one of the next episode and the last one is kept in ``info['terminal_observation']``."""

    def __init__(self, env_fns: Sequence[Callable]):
        self.envs = [fn() for fn in env_fns]
        env = self.envs[0]
        super().__init__(len(self.envs), env.observation_space, env.action_space)
        self.actions = None
        # episode accounting of the reference's Monitor wrapper (common/monitor.py): raw reward sum / length / wall time
        self._t_start = time.time()
        self._ep_rew = [0.0] * self.num_envs
        self._ep_len = [0] * self.num_envs

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs, rews, dones, infos = [], [], [], []
        for i, (env, a) in enumerate(zip(self.envs, self.actions)):
            o, r, d, info = env.step(a)
            info = dict(info)
            self._ep_rew[i] += float(r)
            self._ep_len[i] += 1
            if d:
                info["episode"] = {"r": round(self._ep_rew[i], 6), "l": self._ep_len[i],
                                   "t": round(time.time() - self._t_start, 6)}
                self._ep_rew[i], self._ep_len[i] = 0.0, 0
                info["terminal_observation"] = o
                o = env.reset()
            obs.append(o), rews.append(r), dones.append(d), infos.append(info)
        return (np.stack(obs).astype(np.float32), np.asarray(rews, dtype=np.float32), np.asarray(dones, dtype=bool),
                infos)

    def reset(self):
        self._ep_rew, self._ep_len = [0.0] * self.num_envs, [0] * self.num_envs
        return np.stack([env.reset() for env in self.envs]).astype(np.float32)

    def seed(self, seed: Optional[int] = None):
        return [env.seed(None if seed is None else seed + i) if hasattr(env, "seed") else None
                for i, env in enumerate(self.envs)]

    def get_attr(self, name, indices=None):
        return [getattr(env, name) for env in self.envs]

    def close(self):
        for env in self.envs:
            if hasattr(env, "close"):
                env.close()


class VecEnvWrapper(VecEnv):
    def __init__(self, venv: VecEnv, observation_space=None, action_space=None):
        self.venv = venv
        super().__init__(venv.num_envs, observation_space or venv.observation_space, action_space or venv.action_space)

    def step_async(self, actions): self.venv.step_async(actions)
    def close(self): return self.venv.close()
    def seed(self, seed=None): return self.venv.seed(seed)
    def get_attr(self, name, indices=None): return self.venv.get_attr(name, indices)

    def __getattr__(self, name):
        # only reached when normal lookup fails: forward to the wrapped env (set_cost_function, get_original_obs ...)
        if name.startswith("_") or name == "venv":
            raise AttributeError(name)
        return getattr(self.venv, name)


class VecCostWrapper(VecEnvWrapper):
    """Relabels every step with the learned cost: ``info[cost_info_str] = cost_function(previous_obs, action)``
    (vec_cost_wrapper.py:32-51). With ``ConstraintNet.cost_function`` plugged in, this is the per-step K1 caller."""

    def __init__(self, venv, cost_info_str='cost'):
        super().__init__(venv)
        self.cost_info_str = cost_info_str
        self.cost_function = None
        self.previous_obs = self.actions = None

    def step_async(self, actions: np.ndarray):
        self.actions = actions
        self.venv.step_async(actions)

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("class_attributes", "cost_function", "venv"):   # the function is re-attached by the driver
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.venv = None

    def set_venv(self, venv):
        if self.venv is not None:
            raise ValueError("Trying to set venv of already initialized VecNormalize wrapper.")
        VecEnvWrapper.__init__(self, venv)

    def step_wait(self):
        obs, rews, news, infos = self.venv.step_wait()
        if self.cost_function is not None:
            cost = self.cost_function(self.previous_obs.copy(), self.actions.copy())
            cost = np.atleast_1d(cost)    # GailDiscriminator.reward_function squeezes a single-env batch to 0-d (the reference
                                          # would raise on cost[i] there)
            for i in range(len(infos)):
                infos[i][self.cost_info_str] = cost[i]
        self.previous_obs = obs.copy()
        return obs, rews, news, infos

    def set_cost_function(self, cost_function):
        self.cost_function = cost_function

    def reset(self):
        obs = self.venv.reset()
        self.previous_obs = obs
        return obs

    @staticmethod
    def load(load_path: str, venv):
        with open(load_path, "rb") as f:
            w = pickle.load(f)
        w.set_venv(venv)
        return w

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            pickle.dump(self, f)


class VecNormalize(VecEnvWrapper):
    """Moving-average observation / reward normalisation (vec_normalize.py:11-181)."""

    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0,
                 gamma=0.99, epsilon=1e-8):
        super().__init__(venv)
        self.obs_rms = RunningMeanStd(shape=self.observation_space.shape)
        self.ret_rms = RunningMeanStd(shape=())
        self.clip_obs, self.clip_reward = clip_obs, clip_reward
        self.ret = np.zeros(self.num_envs)
        self.gamma, self.epsilon = gamma, epsilon
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.old_obs, self.old_reward = np.array([]), np.array([])

    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("venv", "class_attributes", "ret"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        assert "venv" not in state
        self.venv = None

    def set_venv(self, venv):
        if self.venv is not None:
            raise ValueError("Trying to set venv of already initialized VecNormalize wrapper.")
        VecEnvWrapper.__init__(self, venv)
        if self.obs_rms.mean.shape != self.observation_space.shape:
            raise ValueError("venv is incompatible with current statistics.")
        self.ret = np.zeros(self.num_envs)

    def step_wait(self):
        obs, rews, news, infos = self.venv.step_wait()
        self.old_obs, self.old_reward = obs, rews
        if self.training:
            self.obs_rms.update(obs)
        obs = self.normalize_obs(obs)
        if self.training:
            self._update_reward(rews)
        rews = self.normalize_reward(rews)
        self.ret[news] = 0
        return obs, rews, news, infos

    def _update_reward(self, reward: np.ndarray) -> None:
        self.ret = self.ret * self.gamma + reward
        self.ret_rms.update(self.ret)

    def normalize_obs(self, obs: np.ndarray) -> np.ndarray:
        if self.norm_obs:
            obs = np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs,
                          self.clip_obs)
        return obs

    def normalize_reward(self, reward: np.ndarray) -> np.ndarray:
        if self.norm_reward:
            reward = np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        return reward

    def unnormalize_obs(self, obs: np.ndarray) -> np.ndarray:
        if self.norm_obs:
            return (obs * np.sqrt(self.obs_rms.var + self.epsilon)) + self.obs_rms.mean
        return obs

    def unnormalize_reward(self, reward: np.ndarray) -> np.ndarray:
        if self.norm_reward:
            return reward * np.sqrt(self.ret_rms.var + self.epsilon)
        return reward

    def get_original_obs(self) -> np.ndarray:
        return self.old_obs.copy()

    def get_original_reward(self) -> np.ndarray:
        return self.old_reward.copy()

    def reset(self) -> np.ndarray:
        obs = self.venv.reset()
        self.old_obs = obs
        self.ret = np.zeros(self.num_envs)
        if self.training:
            self._update_reward(self.ret)
        return self.normalize_obs(obs)

    @staticmethod
    def load(load_path: str, venv):
        with open(load_path, "rb") as f:
            w = pickle.load(f)
        w.set_venv(venv)
        return w

    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            pickle.dump(self, f)


class VecNormalizeWithCost(VecNormalize):
    """Adds the cost stream: the cost read from ``info[cost_info_str]`` is divided by the running std of its discounted
    sum (no mean subtraction), clipped, and written back (vec_normalize.py:184-282)."""

    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, norm_cost=True, cost_info_str='cost',
                 clip_obs=10.0, clip_reward=10.0, clip_cost=10.0, reward_gamma=0.99, cost_gamma=0.99, epsilon=1e-8):
        super().__init__(venv=venv, training=training, norm_obs=norm_obs, norm_reward=norm_reward, clip_obs=clip_obs,
                         clip_reward=clip_reward, gamma=reward_gamma, epsilon=epsilon)
        self.norm_cost, self.cost_str, self.clip_cost = norm_cost, cost_info_str, clip_cost
        self.gamma, self.cost_gamma = reward_gamma, cost_gamma
        self.cost_rms = RunningMeanStd(shape=())
        self.cost_ret = np.zeros(self.num_envs)
        self.old_cost = np.array([])

    def __getstate__(self):
        state = super().__getstate__()
        state.pop("cost_ret", None)
        return state

    def set_venv(self, venv):
        super().set_venv(venv)
        self.cost_ret = np.zeros(self.num_envs)

    def step_wait(self):
        obs, rews, news, infos = super().step_wait()
        # the reference only inspects the first env's info to decide whether a cost is present
        if infos[0] is not None and self.cost_str in infos[0].keys():
            cost = np.array([infos[i][self.cost_str] for i in range(len(infos))])
            self.old_cost = cost
            if self.training:
                self._update_cost(cost)
            normalized_cost = self.normalize_cost(cost)
            for i in range(len(infos)):
                infos[i][self.cost_str] = normalized_cost[i]
            self.cost_ret[news] = 0
        return obs, rews, news, infos

    def _update_cost(self, cost):
        self.cost_ret = self.cost_ret * self.cost_gamma + cost
        self.cost_rms.update(self.cost_ret)

    def normalize_cost(self, cost):
        if self.norm_cost:
            cost = np.clip(cost / np.sqrt(self.cost_rms.var + self.epsilon), -self.clip_cost, self.clip_cost)
        return cost

    def unnormalize_cost(self, cost):
        if self.norm_cost:
            return cost * np.sqrt(self.cost_rms.var + self.epsilon)
        return cost

    def get_original_cost(self):
        return self.old_cost.copy()

    def reset(self):
        normalized_obs = super().reset()
        self.cost_ret = np.zeros(self.num_envs)
        if self.training:
            self._update_cost(self.cost_ret)
        return normalized_obs


def sync_envs_normalization(env, eval_env) -> None:
    """Copy the running statistics of every VecNormalize layer of `env` onto the matching layer of `eval_env`."""
    env_tmp, eval_env_tmp = env, eval_env
    while isinstance(env_tmp, VecEnvWrapper):
        if isinstance(env_tmp, VecNormalize):
            while isinstance(eval_env_tmp, VecEnvWrapper) and not isinstance(eval_env_tmp, VecNormalize):
                eval_env_tmp = eval_env_tmp.venv
            if not isinstance(eval_env_tmp, VecNormalize):
                return
            eval_env_tmp.obs_rms = deepcopy(env_tmp.obs_rms)
            eval_env_tmp.ret_rms = deepcopy(env_tmp.ret_rms)       # like the reference, cost_rms is not synced
            eval_env_tmp = eval_env_tmp.venv
        env_tmp = env_tmp.venv
