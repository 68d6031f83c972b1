"""Per-phase cycles of the many-cluster K4 kernel (ICRL_PPO_TIMING=1) and its launch time, for an env shape and buffer size.
Usage: ICRL_PPO_TIMING=1 python tools/k4_wide_time.py [antwall|halfcheetah] [rows] [batch] [epochs]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from icrl_b200 import _lib  # noqa: E402
from icrl_b200.learner import WORKLOADS, spaces_of  # noqa: E402
from icrl_b200.policies import ActorTwoCriticsPolicy  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "antwall"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
w = WORKLOADS[name]
T = 2048
E = n // T
n = T * E
B = int(sys.argv[3]) if len(sys.argv) > 3 else max(2048, n // 80)
n_epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = th.device("cuda")
L = _lib.lib()
th.manual_seed(0)
obs_space, act_space = spaces_of(w)
pol = ActorTwoCriticsPolicy(obs_space, act_space, lambda _: w.learning_rate, device=dev)
obs = th.randn(n, w.obs_dim, device=dev)
acs = (th.randint(0, w.act_dim, (n, 1), device=dev).float() if w.is_discrete else th.randn(n, w.act_dim, device=dev))
sc = [th.randn(T, E, device=dev) for _ in range(6)]
logp = th.randn(T, E, device=dev) * 0.1 - 11.3
perm = th.stack([th.randperm(n, device=dev) for _ in range(n_epochs)]).to(th.int32)
spe = -(-n // B)
cfg = pol.make_cfg(T=T, E=E, batch_size=B, n_epochs=n_epochs, has_target_kl=0, target_kl=0.0, clip_range=w.clip_range,
                   ent_coef=0.0, reward_vf_coef=0.5, cost_vf_coef=0.5, max_grad_norm=0.5, nu=0.1, max_steps=0)
data = _lib.PpoData()
data.observations, data.actions, data.old_log_prob = obs.data_ptr(), acs.data_ptr(), logp.data_ptr()
data.old_reward_values, data.reward_advantages, data.reward_returns = (x.data_ptr() for x in sc[:3])
data.old_cost_values, data.cost_advantages, data.cost_returns = (x.data_ptr() for x in sc[3:])
data.perm = perm.data_ptr()
stats = th.zeros(n_epochs * spe, 8, device=dev)
result = th.zeros(4, dtype=th.int32, device=dev)
for rep in range(2):
    s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    s.record()
    _lib.check(L.icrl_ppo_train(C.byref(cfg), C.byref(data), _lib.ptr(pol._params), _lib.ptr(pol._adam_m), _lib.ptr(pol._adam_v),
                                pol.optimizer.step_count, _lib.ptr(stats), _lib.ptr(result), _lib.current_stream()))
    e.record()
    th.cuda.synchronize()
    pol.optimizer.step_count += n_epochs * spe
    ms = s.elapsed_time(e)
    print(f"{name} rows={n} batch={B} epochs={n_epochs}: {ms:.2f} ms, {ms * 1e3 / (n_epochs * spe):.1f} us/step, "
          f"{n * n_epochs / ms / 1e3:.1f} M sample-passes/s, result={result.cpu().tolist()}")
