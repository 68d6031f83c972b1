"""One ICRL learner iteration on device-resident data: the hot path named by BASELINE.json's north_star,
    for each of R rollouts:  K1 cost relabel -> K3 dual GAE -> K4 PPO-Lagrangian update -> dual step
    then once:               K2 constraint-net training against the expert batch
(call order of icrl/icrl.py:199-239 with env stepping removed).  Everything here talks to the C-ABI with device
pointers and never synchronises with the host inside `run()`; it is what `bench.py` times as the device-resident
`value`.  The reference-shaped classes (ConstraintNet / RolloutBufferWithCost / PPOLagrangian) cover the same path
with host buffers and are what `bench.py` times as `e2e`.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch as th

from . import _lib
from .constraint_net import ConstraintNet
from .device import resolve_device
from .dual_variable import DualVariable
from .policies import ActorTwoCriticsPolicy
from .vec_env import RunningMeanStd
from .spaces import Box, Discrete


@dataclass
class Workload:
    """Shapes + hyper-parameters of one named config (README.md:25,38,50,65 of the reference)."""
    name: str
    obs_dim: int
    act_dim: int               # action dims, or number of actions when discrete
    is_discrete: bool
    cn_hidden: Tuple[int, ...]
    n_steps: int = 2048
    n_envs: int = 5
    batch_size: int = 64
    n_epochs: int = 10
    rollouts: int = 20         # forward_timesteps / (n_steps * n_envs)
    backward_iters: int = 10
    expert_rows: int = 5000
    nominal_rows: int = 10000
    episode_len: int = 1000
    per_step_is: bool = True
    cn_reg: float = 0.5
    cn_lr: float = 0.05
    learning_rate: float = 3e-4
    clip_range: float = 0.2
    reward_gae_lambda: float = 0.95
    cost_gae_lambda: float = 0.95
    penalty_initial_value: float = 1.0
    penalty_learning_rate: float = 0.1
    clip_obs: float = 20.0
    normalize_cost: bool = True    # VecNormalizeWithCost cost stream (off with -dnc)

    @property
    def transitions_per_iteration(self) -> int:
        return self.rollouts * self.n_steps * self.n_envs


WORKLOADS = {
    # python run_me.py icrl ... -tei LGW-v0 -cl 20 -clr 0.003 -ft 0.5e5 -ni 10 -bi 20 -dno -dnr -dnc   (README.md:25)
    "lapgrid": Workload("LapGrid LGW-v0 ICRL (cl 20, ft 5e4, bi 20)", 1, 2, True, (20,), rollouts=5, backward_iters=20,
                        expert_rows=4000, nominal_rows=4000, episode_len=200, per_step_is=False, cn_reg=0.0, cn_lr=0.003,
                        normalize_cost=False),
    # ... -tei HCWithPos-v0 -cl 20 -bi 10 -ft 2e5 -clr 0.05 -crc 0.5 -psis                              (README.md:38)
    "halfcheetah": Workload("HalfCheetah HCWithPos-v0 ICRL (cl 20, ft 2e5, bi 10)", 18, 6, False, (20,)),
    # ... -tei AntWall-v0 -cl 40 40 -clr 0.005 -crc 0.6 -bi 5 -ft 2e5 --batch_size 128 --n_epochs 20 ... (README.md:50)
    "antwall": Workload("AntWall-v0 ICRL (cl 40 40, ft 2e5, bi 5, batch 128, n_epochs 20)", 113, 8, False, (40, 40),
                        batch_size=128, n_epochs=20, backward_iters=5, expert_rows=22500, nominal_rows=22500,
                        episode_len=500, cn_reg=0.6, cn_lr=0.005, learning_rate=3e-5, clip_range=0.4,
                        reward_gae_lambda=0.9, cost_gae_lambda=0.9, penalty_initial_value=0.1,
                        penalty_learning_rate=0.05),
    # cpg with a frozen constraint net: K1 relabel + K3 + K4 only                                        (README.md:65)
    "pointcircle": Workload("PointCircle-v0 cpg (frozen cn 40 40 on obs[0:2])", 6, 2, False, (40, 40), rollouts=20,
                            backward_iters=0, expert_rows=0, nominal_rows=0, episode_len=150, penalty_learning_rate=1.0),
}


def synth_rollouts(w: Workload, seed: int, n_envs: Optional[int] = None):
    """Host-side synthetic rollouts of one ICRL iteration (SURVEY §8(d) 'Synthetic inputs'), as numpy arrays
    [R, T, E, ...]; deterministic in `seed`."""
    rng = np.random.default_rng(seed)
    R, T, E = w.rollouts, w.n_steps, n_envs or w.n_envs
    scale = rng.uniform(0.5, 8.0, size=w.obs_dim).astype(np.float32)
    orig_obs = rng.standard_normal((R, T, E, w.obs_dim), dtype=np.float32) * scale
    obs = np.clip(orig_obs / scale, -10, 10).astype(np.float32)            # "VecNormalize-d" view
    if w.is_discrete:
        actions = rng.integers(0, w.act_dim, size=(R, T, E, 1)).astype(np.float32)
    else:
        actions = rng.standard_normal((R, T, E, w.act_dim), dtype=np.float32)
    rewards = rng.standard_normal((R, T, E), dtype=np.float32)
    phase = rng.integers(0, w.episode_len, size=E)
    t = np.arange(T)[:, None]
    dones = np.broadcast_to((((t + phase[None]) % w.episode_len) == 0).astype(np.float32), (R, T, E)).copy()
    last_dones = rng.random((R, E)) < 0.01
    return dict(orig_obs=orig_obs, obs=obs, actions=actions, rewards=rewards, dones=dones, last_dones=last_dones)


def synth_demos(w: Workload, seed: int):
    """Expert batch and nominal trajectories for K2 (shapes of icrl/icrl.py:25-43 and icrl/utils.py:323-357)."""
    rng = np.random.default_rng(seed + 1)
    scale = rng.uniform(0.5, 8.0, size=w.obs_dim)

    def draw(n):
        o = rng.standard_normal((n, w.obs_dim)) * scale
        a = (rng.integers(0, w.act_dim, size=(n, 1)).astype(np.float32) if w.is_discrete
             else rng.standard_normal((n, w.act_dim)).astype(np.float32))
        return o.astype(np.float32), a
    eo, ea = draw(w.expert_rows)
    no, na = draw(w.nominal_rows)
    lengths = np.full(max(w.nominal_rows // w.episode_len, 1), w.episode_len, dtype=np.int64)
    lengths[-1] = w.nominal_rows - lengths[:-1].sum()
    return eo, ea, no, na, lengths


def spaces_of(w: Workload):
    obs_space = Box(-np.inf, np.inf, (w.obs_dim,), np.float32)
    act_space = Discrete(w.act_dim) if w.is_discrete else Box(-1.0, 1.0, (w.act_dim,), np.float32)
    return obs_space, act_space


class DeviceLearner:
    """Device-resident learner state + data for one workload.  `run()` enqueues one full ICRL learner iteration."""

    def __init__(self, w: Workload, seed: int = 0, device="cuda", n_envs: Optional[int] = None, fixed_work: bool = True,
                 comm=None, param_seed: Optional[int] = None):
        """`comm` (icrl_b200.distributed.PpoComm): data-parallel mode -- this rank's rollouts (seed) differ per rank, the
        networks and the K2 batches (param_seed) are replicated."""
        self.w, self.dev = w, resolve_device(device)
        self.E = n_envs or w.n_envs
        self.comm = comm
        param_seed = seed if param_seed is None else param_seed
        th.manual_seed(param_seed)
        obs_space, act_space = spaces_of(w)
        self.policy = ActorTwoCriticsPolicy(obs_space, act_space, lambda _: w.learning_rate, device=self.dev)
        self.dual = DualVariable(0.0, w.penalty_learning_rate, w.penalty_initial_value, device=self.dev)
        eo, ea, no, na, lengths = synth_demos(w, param_seed)
        low = high = None
        if not w.is_discrete:
            low, high = -np.ones(w.act_dim, np.float32), np.ones(w.act_dim, np.float32)
        kl = -1 if fixed_work else 10     # fixed work: every backward iteration runs (no KL early stop)
        self.cn = ConstraintNet(w.obs_dim, w.act_dim, w.cn_hidden, None, lambda _: w.cn_lr, eo if len(eo) else None,
                                ea if len(ea) else None, w.is_discrete, w.cn_reg,
                                per_step_importance_sampling=w.per_step_is, clip_obs=w.clip_obs, action_low=low,
                                action_high=high, target_kl_old_new=kl, target_kl_new_old=kl if fixed_work else 2.5,
                                device=self.dev)
        self.fixed_work = fixed_work
        self.host = synth_rollouts(w, seed, self.E)
        d = self.dev
        self.data = {k: th.from_numpy(np.ascontiguousarray(v)).to(d) for k, v in self.host.items() if k != "last_dones"}
        self.data["last_dones"] = th.from_numpy(self.host["last_dones"].astype(np.uint8)).to(d)
        R, T, E = w.rollouts, w.n_steps, self.E
        self.n = T * E
        z = lambda *s: th.zeros(*s, device=d)
        # per-rollout learner-side arrays: values / log-probs of the behaviour policy, then K1/K3 outputs
        self.reward_values, self.cost_values, self.log_probs = z(R, T, E), z(R, T, E), z(R, T, E)
        self.last_rv, self.last_cv = z(R, E), z(R, E)
        self.costs = z(R, T, E)
        # K5: VecNormalizeWithCost's cost statistics {mean, var, count, cost_ret[E]} as they stand after reset()
        self.costs_norm = z(R, T, E) if w.normalize_cost else self.costs
        rms = RunningMeanStd(shape=())
        rms.update(np.zeros(E))
        self.cost_state = th.tensor([rms.mean, rms.var, rms.count] + [0.0] * E, dtype=th.float64, device=d)
        self.adv_r, self.ret_r, self.adv_c, self.ret_c = z(R, T, E), z(R, T, E), z(R, T, E), z(R, T, E)
        self.steps_per_epoch = (self.n + w.batch_size - 1) // w.batch_size
        self.stats = z(R, w.n_epochs * self.steps_per_epoch, _lib.PPO_STATS_PER_STEP)
        self.result = th.zeros(R, 4, dtype=th.int32, device=d)
        self.perm = th.empty(R, w.n_epochs, self.n, dtype=th.int32, device=d)
        if w.nominal_rows:
            self.nom_obs, self.nom_acs = th.from_numpy(no).to(d), th.from_numpy(
                na.reshape(-1) if w.is_discrete else na).to(d)
            self.exp_obs, self.exp_acs = th.from_numpy(eo).to(d), th.from_numpy(
                ea.reshape(-1) if w.is_discrete else ea).to(d)
            off = np.zeros(len(lengths) + 1, np.int32)
            off[1:] = np.cumsum(lengths)
            self.offsets, self.n_episodes = th.from_numpy(off).to(d), len(lengths)
        self.max_steps = 0
        if comm is not None and comm.world > 1:
            self.advsums = th.zeros(w.n_epochs * self.steps_per_epoch, 4, dtype=th.float64, device=d)
            self.cost_mean = th.zeros(1, device=d)
            self.shard_k2(seed)
        self.refresh_behaviour()
        self.new_permutations(seed)

    # ------------------------------------------------------------------ setup helpers (outside the timed region)
    def refresh_behaviour(self):
        """values / log-probs 'recorded at collection time': the current policy evaluated on the stored (obs, action)."""
        w = self.w
        for r in range(w.rollouts):
            obs = self.data["obs"][r].reshape(self.n, w.obs_dim)
            acts = self.data["actions"][r].reshape(self.n, -1)
            v, cv, lp, _ = self.policy.evaluate_actions(obs, acts.reshape(-1) if w.is_discrete else acts)
            self.reward_values[r] = v.reshape(w.n_steps, self.E)
            self.cost_values[r] = cv.reshape(w.n_steps, self.E)
            self.log_probs[r] = lp.reshape(w.n_steps, self.E)
            self.last_rv[r], self.last_cv[r] = self.reward_values[r, -1], self.cost_values[r, -1]

    def new_permutations(self, seed):
        """numpy's permutations for every (rollout, epoch), as RolloutBufferWithCost.get draws them (buffers.py:596)."""
        rs = np.random.RandomState(seed)
        w = self.w
        perms = np.stack([np.stack([rs.permutation(self.n) for _ in range(w.n_epochs)]) for _ in range(w.rollouts)])
        self.perm.copy_(th.from_numpy(perms.astype(np.int32)))

    # ------------------------------------------------------------------ the hot path
    def ppo_cfg(self):
        w = self.w
        return self.policy.make_cfg(
            T=w.n_steps, E=self.E, batch_size=w.batch_size, n_epochs=w.n_epochs, has_target_kl=0, target_kl=0.0,
            clip_range=w.clip_range, ent_coef=0.0, reward_vf_coef=0.5, cost_vf_coef=0.5, max_grad_norm=0.5,
            nu=0.0, max_steps=self.max_steps)

    def run(self, stream=None):
        """Enqueue one ICRL learner iteration on the current stream.  No host synchronisation."""
        w, L = self.w, _lib.lib()
        st = _lib.current_stream()
        desc = self.cn._get_desc()
        cfg = self.ppo_cfg()
        pol = self.policy
        for r in range(w.rollouts):
            # K1: relabel the whole rollout with the current constraint net (VecCostWrapper.step_wait, batched)
            _lib.check(L.icrl_cn_forward(C.byref(desc), _lib.ptr(self.data["orig_obs"][r]), 0,
                                         _lib.ptr(self.data["actions"][r]), self.n, _lib.ptr(self.costs[r]), 0, st))
            # K5: the cost stream of VecNormalizeWithCost replayed over the whole rollout (float64 statistics on device)
            if w.normalize_cost:
                _lib.check(L.icrl_cost_normalize(_lib.ptr(self.costs[r]), _lib.ptr(self.data["dones"][r]),
                                                 _lib.ptr(self.data["last_dones"][r]), w.n_steps, self.E, 0.99, 1e-8, 10.0,
                                                 1, 1, _lib.ptr(self.cost_state), _lib.ptr(self.costs_norm[r]), st))
            # K3: dual GAE
            _lib.check(L.icrl_dual_gae(
                _lib.ptr(self.data["rewards"][r]), _lib.ptr(self.reward_values[r]), _lib.ptr(self.costs_norm[r]),
                _lib.ptr(self.cost_values[r]), _lib.ptr(self.data["dones"][r]), _lib.ptr(self.last_rv[r]),
                _lib.ptr(self.last_cv[r]), _lib.ptr(self.data["last_dones"][r]), w.n_steps, self.E, 0.99,
                w.reward_gae_lambda, 0.99, w.cost_gae_lambda, _lib.ptr(self.adv_r[r]), _lib.ptr(self.ret_r[r]),
                _lib.ptr(self.adv_c[r]), _lib.ptr(self.ret_c[r]), st))
            # K4: every minibatch of every epoch in one persistent launch
            data = _lib.PpoData()
            data.nu_device = self.dual.nu.state[4:5].data_ptr()      # nu stays on the device between launches
            data.observations, data.actions = self.data["obs"][r].data_ptr(), self.data["actions"][r].data_ptr()
            data.old_log_prob = self.log_probs[r].data_ptr()
            data.old_reward_values, data.old_cost_values = self.reward_values[r].data_ptr(), self.cost_values[r].data_ptr()
            data.reward_advantages, data.reward_returns = self.adv_r[r].data_ptr(), self.ret_r[r].data_ptr()
            data.cost_advantages, data.cost_returns = self.adv_c[r].data_ptr(), self.ret_c[r].data_ptr()
            data.perm = self.perm[r].data_ptr()
            dp = self.comm is not None and self.comm.world > 1
            if not dp:
                _lib.check(L.icrl_ppo_train(C.byref(cfg), C.byref(data), _lib.ptr(pol._params), _lib.ptr(pol._adam_m),
                                            _lib.ptr(pol._adam_v), pol.optimizer.step_count, _lib.ptr(self.stats[r]),
                                            _lib.ptr(self.result[r]), st))
            else:
                # global-minibatch advantage statistics: local partial sums -> one small NCCL all-reduce
                _lib.check(L.icrl_ppo_local_advsums(C.byref(cfg), C.byref(data), _lib.ptr(self.advsums), st))
                self.comm.all_reduce_sum(self.advsums)
                desc_d = self.comm.descriptor(self.advsums)
                _lib.check(L.icrl_ppo_train_dist(C.byref(cfg), C.byref(data), _lib.ptr(pol._params), _lib.ptr(pol._adam_m),
                                                 _lib.ptr(pol._adam_v), pol.optimizer.step_count, _lib.ptr(self.stats[r]),
                                                 _lib.ptr(self.result[r]), C.byref(desc_d), _lib.current_stream()))
                self.comm.advance(self.steps_taken_per_rollout())
            pol.optimizer.step_count += self.steps_taken_per_rollout()
            # dual step on the rollout's (relabelled) costs -- over every rank's environments in data-parallel mode
            if not dp:
                _lib.check(L.icrl_dual_update(_lib.ptr(self.dual.nu.state), _lib.ptr(self.costs[r]), self.n, 0.0,
                                              float(w.penalty_learning_rate), self.dual.steps, self.dual.nu.clamp_min(), st))
            else:
                self.cost_mean.copy_(self.costs[r].mean().reshape(1))
                self.comm.all_reduce_sum(self.cost_mean)
                self.cost_mean /= self.comm.world
                _lib.check(L.icrl_dual_update(_lib.ptr(self.dual.nu.state), _lib.ptr(self.cost_mean), 1, 0.0,
                                              float(w.penalty_learning_rate), self.dual.steps, self.dual.nu.clamp_min(),
                                              _lib.current_stream()))
            self.dual.steps += 1
        if w.backward_iters > 0 and w.nominal_rows > 0:
            self._cn_train_device()

    def check(self):
        """Synchronises and raises if any K4 launch of the last run() reported an exchange time-out (result[2])."""
        res = self.result.cpu().numpy()
        if (res[:, 2] != 0).any():
            raise _lib.IcrlError(f"PPO kernel exchange time-out in rollouts {np.nonzero(res[:, 2])[0].tolist()}: the "
                                 "replicated parameters are no longer valid")
        want = self.steps_taken_per_rollout()
        if (res[:, 1] != want).any():
            raise _lib.IcrlError(f"PPO kernel took {res[:, 1].tolist()} optimiser steps, expected {want} per rollout")

    def steps_taken_per_rollout(self):
        full = self.w.n_epochs * self.steps_per_epoch
        return min(full, self.max_steps) if self.max_steps > 0 else full

    def _cn_train_device(self):
        w, cn = self.w, self.cn
        g = cn.optimizer.param_groups[0]
        cfg = _lib.CnTrainCfg(
            iterations=w.backward_iters, importance_sampling=1, per_step_is=int(w.per_step_is), train_gail_lambda=0,
            eps=float(cn.eps), regularizer_coeff=float(w.cn_reg), target_kl_old_new=float(cn.target_kl_old_new),
            target_kl_new_old=float(cn.target_kl_new_old), lr=float(g["lr"]), adam_beta1=0.9, adam_beta2=0.999,
            adam_eps=float(g["eps"]), batch_size=0, perm=None)
        metrics = _lib.CnTrainMetrics()
        step = C.c_int64(cn.optimizer.step_count)
        args = (C.byref(cn._get_desc()), C.byref(cfg), _lib.ptr(self.nom_obs), 0, _lib.ptr(self.nom_acs), w.nominal_rows,
                _lib.ptr(self.offsets), self.n_episodes, _lib.ptr(self.exp_obs), 0, _lib.ptr(self.exp_acs),
                self.exp_obs.shape[0], _lib.ptr(cn._adam_m), _lib.ptr(cn._adam_v), C.byref(step), C.byref(metrics))
        cc = getattr(cn, "comm", None)
        if cc is None or cc.world == 1:
            _lib.check(_lib.lib().icrl_cn_train(*args, _lib.current_stream()))
        else:
            # data parallel: this rank's own nominal episodes, its slice of the expert batch (set up in shard_k2)
            d = cc.descriptor(w.nominal_rows * cc.world, self.exp_rows_global, self.n_episodes * cc.world,
                              self.n_episodes * cc.rank)
            _lib.check(_lib.lib().icrl_cn_train_dist(*args, C.byref(d), _lib.current_stream()))
            cc.advance(w.backward_iters)
        cn.optimizer.step_count = int(step.value)
        self.last_cn_metrics = metrics

    def shard_k2(self, seed: int):
        """Data-parallel K2: every rank keeps nominal episodes of its own (seed) and an even slice of the
        replicated expert batch; the constraint net's exchange buffers are set up here (outside the timed region)."""
        w, cn, comm = self.w, self.cn, self.comm
        if comm is None or comm.world == 1 or not w.nominal_rows:
            return
        cn.enable_data_parallel(max_episodes=max(1024, 2 * self.n_episodes * comm.world))
        _, _, no, na, _ = synth_demos(w, seed)
        self.nom_obs = th.from_numpy(no).to(self.dev)
        self.nom_acs = th.from_numpy(na.reshape(-1) if w.is_discrete else na).to(self.dev)
        n = self.exp_obs.shape[0]
        self.exp_rows_global = n
        lo, hi = n * comm.rank // comm.world, n * (comm.rank + 1) // comm.world
        self.exp_obs, self.exp_acs = self.exp_obs[lo:hi].contiguous(), self.exp_acs[lo:hi].contiguous()
