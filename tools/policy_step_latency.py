"""Wall-clock latency of the rollout-time policy call (ActorTwoCriticsPolicy.forward on [n_envs, obs_dim] host rows, once per
environment step in collect_rollouts).  Usage: python tools/policy_step_latency.py [workload]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from icrl_b200.learner import WORKLOADS, spaces_of  # noqa: E402
from icrl_b200.policies import ActorTwoCriticsPolicy  # noqa: E402

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "halfcheetah"]
obs_space, act_space = spaces_of(w)
th.manual_seed(0)
pol = ActorTwoCriticsPolicy(obs_space, act_space, lambda _: w.learning_rate, device=th.device("cuda", 0))
obs = np.random.default_rng(0).standard_normal((w.n_envs, w.obs_dim)).astype(np.float32)
for _ in range(200):
    pol.forward(th.as_tensor(obs))
th.cuda.synchronize()
n = 3000
t0 = time.perf_counter()
for _ in range(n):
    pol.forward(th.as_tensor(obs))
dt = (time.perf_counter() - t0) / n
print(f"{w.name}: policy.forward([{w.n_envs}, {w.obs_dim}]) {dt * 1e6:.1f} us per call -> {dt * w.n_steps * 1e3:.1f} ms per {w.n_steps}-step rollout")
