"""BENCH INFRASTRUCTURE (never imported by the product path) -- the UNMODIFIED reference timed on the host CPU.

`baseline/_ref/` holds the reference's own sources for the learner hot path (copied byte for byte by
tools/make_baseline_ref.sh; git-ignored, travels to the GPU box with the repo snapshot).  This module imports them through
oracle/ref_shim.py (stub gym / matplotlib) and drives exactly the calls SURVEY.md 8(d) / BASELINE.md 3 name:

    K1  ConstraintNet.cost_function per environment step, in the call order of VecCostWrapper.step_wait
        (stable_baselines3/common/vec_env/vec_cost_wrapper.py:51-66) followed by VecNormalizeWithCost's cost stream
        (vec_normalize.py:232-241: _update_cost, normalize_cost, cost_ret[news] = 0) on a real VecNormalizeWithCost object;
    K3  RolloutBufferWithCost.compute_returns_and_advantage (buffers.py:543-552);
    K4  PPOLagrangian.train() on a pre-filled buffer (ppo_lag.py:177-338, dual step included);
    K2  ConstraintNet.train(backward_iters, ...) (icrl/constraint_net.py:137-229), with the reference's N x N broadcast in
        per-step importance-sampling mode.

Every object is the reference's class with the reference's defaults (device "cpu"); the synthetic inputs are the ones the
B200 arm uses (icrl_b200.learner.synth_rollouts / synth_demos: plain numpy generators).  One `sample()` is a bounded piece
of an ICRL iteration (some of the K1 calls, one K3, some epochs of one K4 train(), some K2 iterations); the time of a
whole iteration is the linear extrapolation and the sample is described in the result.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "stable_baselines3", "__init__.py")) and \
        os.path.isfile(os.path.join(REF_DIR, "icrl", "constraint_net.py"))


def install():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["ICRL_REFERENCE_ROOT"] = REF_DIR
    os.environ.setdefault("TQDM_DISABLE", "1")          # ConstraintNet.train wraps its loop in tqdm (constraint_net.py:166)
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = REF_DIR
    ref_shim.install()
    # subproc_vec_env.py:8 does `import custom_envs` (gym registrations, MuJoCo envs): resolved as an empty namespace package
    # in a full checkout; the env package is outside the learner path and not copied, so an empty module stands in
    import types
    sys.modules.setdefault("custom_envs", types.ModuleType("custom_envs"))


class ReferenceIteration:
    """The reference's learner objects for one workload (icrl_b200.learner.Workload), pre-filled with synthetic data."""

    def __init__(self, w, seed=0):
        install()
        import gym
        import torch as th
        from icrl.constraint_net import ConstraintNet
        from stable_baselines3 import PPOLagrangian
        from stable_baselines3.common import logger
        from stable_baselines3.common.vec_env import DummyVecEnv, VecCostWrapper, VecNormalizeWithCost
        from icrl_b200.learner import synth_demos, synth_rollouts
        self.w, self.th = w, th
        logger.configure(folder=None, format_strings=[])
        T, E = w.n_steps, w.n_envs
        obs_space = gym.spaces.Box(-np.inf, np.inf, (w.obs_dim,), np.float32)
        act_space = gym.spaces.Discrete(w.act_dim) if w.is_discrete else gym.spaces.Box(-1, 1, (w.act_dim,), np.float32)

        class FakeEnv(gym.Env):
            observation_space, action_space = obs_space, act_space

            def reset(self):
                return np.zeros(obs_space.shape, np.float32)

            def step(self, a):
                return self.reset(), 0.0, False, {}

        one = type(w)(**{**w.__dict__, "rollouts": 1})
        host = synth_rollouts(one, seed)
        self.host = {k: v[0] for k, v in host.items()}
        th.manual_seed(seed)
        eo, ea, no, na, lengths = synth_demos(w, seed)
        low = high = None
        if not w.is_discrete:
            low, high = -np.ones(w.act_dim, np.float32), np.ones(w.act_dim, np.float32)
        self.cn = ConstraintNet(
            w.obs_dim, w.act_dim, w.cn_hidden, None, lambda _: w.cn_lr, eo if len(eo) else None, ea if len(ea) else None,
            w.is_discrete, w.cn_reg, per_step_importance_sampling=w.per_step_is, clip_obs=w.clip_obs, action_low=low,
            action_high=high, target_kl_old_new=-1, target_kl_new_old=-1, device="cpu")      # -1: no KL early stop (fixed work)
        self.nominal = (no, na, lengths)
        self.algo = PPOLagrangian(
            "TwoCriticsMlpPolicy", DummyVecEnv([FakeEnv for _ in range(E)]), n_steps=T, batch_size=w.batch_size,
            n_epochs=w.n_epochs, learning_rate=w.learning_rate, clip_range=w.clip_range,
            reward_gae_lambda=w.reward_gae_lambda, cost_gae_lambda=w.cost_gae_lambda, target_kl=None,
            penalty_initial_value=w.penalty_initial_value, penalty_learning_rate=w.penalty_learning_rate, seed=seed,
            device="cpu")
        self.algo._current_progress_remaining = 1.0
        self.vn = VecNormalizeWithCost(VecCostWrapper(DummyVecEnv([FakeEnv for _ in range(E)])), training=True,
                                       norm_obs=False, norm_reward=False, norm_cost=w.normalize_cost, cost_gamma=0.99)
        self.vn.venv.previous_obs = None
        self.vn.cost_ret = np.zeros(E)
        self.vn._update_cost(self.vn.cost_ret)                  # what reset() does (vec_normalize.py:278-282)
        # behaviour-policy values / log-probs on the stored (obs, action), as recorded at collection time
        buf, pol = self.algo.rollout_buffer, self.algo.policy
        n = T * E
        with th.no_grad():
            obs = th.tensor(self.host["obs"].reshape(n, w.obs_dim))
            acts = th.tensor(self.host["actions"].reshape(n, -1))
            v, cv, lp, _ = pol.evaluate_actions(obs, acts.long().flatten() if w.is_discrete else acts)
        self.pristine = dict(
            observations=self.host["obs"].copy(), orig_observations=self.host["orig_obs"].copy(),
            new_observations=self.host["obs"].copy(), new_orig_observations=self.host["orig_obs"].copy(),
            actions=self.host["actions"].copy(), rewards=self.host["rewards"].copy(), dones=self.host["dones"].copy(),
            reward_values=v.numpy().reshape(T, E).copy(), cost_values=cv.numpy().reshape(T, E).copy(),
            log_probs=lp.numpy().reshape(T, E).copy())
        self.last_values = (v[-E:].clone(), cv[-E:].clone())
        self.news = np.concatenate([self.host["dones"][1:], self.host["last_dones"][None].astype(np.float32)]) != 0

    def _refill(self):
        buf = self.algo.rollout_buffer
        buf.reset()
        for k, v in self.pristine.items():
            getattr(buf, k)[...] = v.reshape(getattr(buf, k).shape)
        buf.full, buf.pos, buf.generator_ready = True, self.w.n_steps, False

    def sample(self, k1_calls, k4_epochs, k2_iters):
        """Times a bounded piece of one ICRL iteration; returns (seconds per WHOLE iteration extrapolated linearly,
        per-part seconds already scaled to one rollout / one K2 call, description)."""
        w, T = self.w, self.w.n_steps
        t = {}
        self._refill()
        buf = self.algo.rollout_buffer
        oo, aa = self.host["orig_obs"].astype(np.float64), self.host["actions"]
        k1_calls = min(k1_calls, T)
        t0 = time.perf_counter()
        for s in range(k1_calls):                                    # per environment step, as the reference runs it
            cost = self.cn.cost_function(oo[s].copy(), aa[s].copy())
            self.vn.old_cost = cost
            if self.vn.training:
                self.vn._update_cost(cost)
            buf.orig_costs[s] = cost
            buf.costs[s] = self.vn.normalize_cost(cost)
            self.vn.cost_ret[self.news[s]] = 0
        t["k1"] = (time.perf_counter() - t0) / k1_calls * T
        reps = T // k1_calls + 1
        buf.costs[:] = np.tile(buf.costs[:k1_calls], (reps, 1))[:T]
        buf.orig_costs[:] = np.tile(buf.orig_costs[:k1_calls], (reps, 1))[:T]
        t0 = time.perf_counter()
        buf.compute_returns_and_advantage(self.last_values[0], self.last_values[1], dones=self.host["last_dones"])
        t["k3"] = time.perf_counter() - t0
        k4_epochs = max(1, min(k4_epochs, w.n_epochs))
        self.algo.n_epochs = k4_epochs
        t0 = time.perf_counter()
        self.algo.train()
        t["k4"] = (time.perf_counter() - t0) / k4_epochs * w.n_epochs
        self.algo.n_epochs = w.n_epochs
        t["k2"] = 0.0
        k2_done = 0
        if w.backward_iters > 0 and w.nominal_rows > 0:
            k2_done = max(1, min(k2_iters, w.backward_iters))
            no, na, lengths = self.nominal
            t0 = time.perf_counter()
            self.cn.train(k2_done, no, na, lengths)
            t["k2"] = (time.perf_counter() - t0) / k2_done * w.backward_iters
        est = w.rollouts * (t["k1"] + t["k3"] + t["k4"]) + t["k2"]
        steps_per_epoch = -(-T * w.n_envs // w.batch_size)
        desc = (f"unmodified reference classes: K1 {k1_calls} of {T} per-step cost_function calls (+ VecNormalizeWithCost cost "
                f"stream), K3 1 rollout, K4 PPOLagrangian.train() {k4_epochs} of {w.n_epochs} epochs "
                f"({k4_epochs * steps_per_epoch} optimiser steps) of 1 rollout, K2 ConstraintNet.train {k2_done} of "
                f"{w.backward_iters} iterations; x{w.rollouts} rollouts, linear extrapolation")
        return est, t, desc


def k4_seconds_per_epoch(it: ReferenceIteration, threads: int) -> float:
    """One-epoch PPOLagrangian.train() at the given torch thread count (thread-count calibration)."""
    it.th.set_num_threads(threads)
    it._refill()
    it.algo.rollout_buffer.compute_returns_and_advantage(it.last_values[0], it.last_values[1], dones=it.host["last_dones"])
    it.algo.n_epochs = 1
    t0 = time.perf_counter()
    it.algo.train()
    dt = time.perf_counter() - t0
    it.algo.n_epochs = it.w.n_epochs
    return dt
