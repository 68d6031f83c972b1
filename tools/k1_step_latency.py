"""Wall-clock latency of the per-environment-step cost call (VecCostWrapper.step_wait -> ConstraintNet.cost_function on
[n_envs, .] host rows): the drivers' default mode makes 2048 of these per rollout.  Usage: python tools/k1_step_latency.py [workload]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from icrl_b200.learner import WORKLOADS, DeviceLearner  # noqa: E402

w = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "halfcheetah"]
w = type(w)(**{**w.__dict__, "rollouts": 1})
learner = DeviceLearner(w, seed=0)
cn = learner.cn
rng = np.random.default_rng(0)
E = learner.E
obs = rng.standard_normal((E, w.obs_dim))                      # float64, as the env wrappers hand them over
acs = (rng.integers(0, w.act_dim, (E,)) if w.is_discrete else rng.standard_normal((E, w.act_dim)).astype(np.float32))
for _ in range(200):
    cn.cost_function(obs, acs)
th.cuda.synchronize()
n = 4000
t0 = time.perf_counter()
for _ in range(n):
    cn.cost_function(obs, acs)
dt = (time.perf_counter() - t0) / n
print(f"{w.name}: cost_function([{E}, {w.obs_dim}]) {dt * 1e6:.1f} us per call -> {dt * w.n_steps * 1e3:.1f} ms per {w.n_steps}-step rollout"
      f" (zero-copy {'off' if os.environ.get('ICRL_K1_NO_ZEROCOPY') else 'on'})")
