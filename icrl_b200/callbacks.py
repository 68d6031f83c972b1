"""The callbacks the reference's drivers attach to `PPOLagrangian.learn` (stable_baselines3/common/callbacks.py and
icrl/utils.py:516-620), reduced to the hooks `collect_rollouts` / `learn` actually call."""
import os

import numpy as np

from . import logger, vec_env
from .utils import evaluate_policy


class BaseCallback:
    def __init__(self, verbose: int = 0):
        self.model, self.training_env, self.verbose = None, None, verbose
        self.n_calls, self.num_timesteps = 0, 0
        self.locals, self.globals = {}, {}
        self.logger = logger

    def init_callback(self, model):
        self.model, self.training_env = model, model.env
        self._init_callback()

    def _init_callback(self): pass

    def on_training_start(self, locals_, globals_):
        self.locals, self.globals = locals_, globals_
        self._on_training_start()

    def _on_training_start(self): pass
    def on_rollout_start(self): self._on_rollout_start()
    def _on_rollout_start(self): pass
    def update_locals(self, locals_): self.locals.update(locals_)

    def on_step(self) -> bool:
        self.n_calls += 1
        self.num_timesteps = self.model.num_timesteps
        return self._on_step() is not False

    def _on_step(self): return True
    def on_rollout_end(self): self._on_rollout_end()
    def _on_rollout_end(self): pass
    def on_training_end(self): self._on_training_end()
    def _on_training_end(self): pass


class CheckpointCallback(BaseCallback):
    """callbacks.py:216-252: save every `save_freq` calls as <prefix>_<timesteps>_steps.zip."""

    def __init__(self, save_freq, save_path, name_prefix="rl_model", verbose=0, callback_on_new_save=None):
        super().__init__(verbose)
        self.save_freq, self.save_path, self.name_prefix, self.callback = save_freq, save_path, name_prefix, callback_on_new_save

    def _init_callback(self):
        if self.save_path is not None:
            os.makedirs(self.save_path, exist_ok=True)
        if self.callback is not None:
            self.callback.init_callback(self.model)

    def _on_step(self):
        if self.n_calls % self.save_freq == 0:
            self.model.save(os.path.join(self.save_path, f"{self.name_prefix}_{self.num_timesteps}_steps"))
            if self.callback is not None:
                self.callback.on_step()
        return True


class EvalCallback(BaseCallback):
    """callbacks.py:258-400: every `eval_freq` calls evaluate on `eval_env`, log eval/mean_reward, keep best_model."""

    def __init__(self, eval_env, callback_on_new_best=None, n_eval_episodes=5, eval_freq=10000, log_path=None,
                 best_model_save_path=None, deterministic=True, verbose=1):
        super().__init__(verbose)
        self.eval_env, self.callback = eval_env, callback_on_new_best
        self.n_eval_episodes, self.eval_freq, self.deterministic = n_eval_episodes, eval_freq, deterministic
        self.best_mean_reward, self.last_mean_reward = -np.inf, -np.inf
        self.best_model_save_path = best_model_save_path

    def _init_callback(self):
        if self.best_model_save_path is not None:
            os.makedirs(self.best_model_save_path, exist_ok=True)
        if self.callback is not None:
            self.callback.init_callback(self.model)

    def _on_step(self):
        if self.eval_freq > 0 and self.n_calls % self.eval_freq == 0:
            vec_env.sync_envs_normalization(self.training_env, self.eval_env)
            rewards, lengths = evaluate_policy(self.model, self.eval_env, n_eval_episodes=self.n_eval_episodes,
                                               deterministic=self.deterministic, return_episode_rewards=True)
            mean_reward = float(np.mean(rewards))
            self.last_mean_reward = mean_reward
            self.logger.record("eval/mean_reward", mean_reward)
            self.logger.record("eval/mean_ep_length", float(np.mean(lengths)))
            self.logger.record("eval/best_mean_reward", max(self.best_mean_reward, mean_reward))
            if mean_reward > self.best_mean_reward:
                if self.best_model_save_path is not None:
                    self.model.save(os.path.join(self.best_model_save_path, "best_model"))
                self.best_mean_reward = mean_reward
                if self.callback is not None:
                    self.callback.on_step()
        return True


class SaveEnvStatsCallback(BaseCallback):
    def __init__(self, env, save_path):
        super().__init__()
        self.env, self.save_path = env, save_path

    def _on_step(self):
        if isinstance(self.env, vec_env.VecNormalize):
            self.env.save(os.path.join(self.save_path, "train_env_stats.pkl"))
        return True


class AdjustedRewardCallback(BaseCallback):
    """rollout/adjusted_reward = mean(R - nu * C) and eval/true_cost on every rollout (icrl/utils.py:542-568)."""

    def __init__(self, cost_fn, verbose: int = 1):
        super().__init__(verbose)
        self.cost_fn = cost_fn

    def _on_rollout_end(self):
        buf = self.model.rollout_buffer
        rewards, costs = buf.rewards.copy(), buf.costs.copy()
        if isinstance(self.training_env, vec_env.VecNormalize):
            rewards = self.training_env.unnormalize_reward(rewards)
        self.logger.record("rollout/adjusted_reward", float(np.mean(rewards - self.model.dual.nu().item() * costs)))
        if self.cost_fn is not None:
            self.logger.record("eval/true_cost", float(np.mean(self.cost_fn(buf.orig_observations.copy(),
                                                                            buf.actions.copy()))))


class LogTorqueCallback(BaseCallback):
    def _on_rollout_end(self):
        actions_abs = np.abs(self.model.rollout_buffer.actions.copy())
        for thr in (0.5, 0.3, 0.25):
            self.logger.record(f'torque/greater_than_{thr}', np.sum(np.any(actions_abs > thr, axis=-1)))
        for i, m in enumerate(np.mean(actions_abs, axis=(0, 1))):
            self.logger.record('torque/mean_motor' + str(i), m)
