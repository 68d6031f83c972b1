"""DualVariable / Nu / PIDLagrangian -- mirrors of stable_baselines3/common/dual_variable.py:9-122.

`DualVariable.update_parameter` runs as one tiny CUDA launch (K4 dual step): when given the rollout's original
costs as a device tensor it also does the `np.mean(orig_costs)` reduction there; given a host scalar it wraps it.
The PID variant is scalar Python bookkeeping and stays on the host.
"""
from collections import deque

import numpy as np
import torch as th

from . import _lib
from .device import resolve_device


class Nu:
    """nu = softplus(log_nu); state lives in a device float32[6] (see include/icrl_b200.h, icrl_dual_update)."""

    def __init__(self, penalty_init=1., clamp_at=None, device="cuda"):
        self.penalty_init = penalty_init
        penalty_init = np.log(max(np.exp(penalty_init) - 1, 1e-8))          # dual_variable.py:17-19
        self.state = th.zeros(6, dtype=th.float32, device=resolve_device(device))
        self.state[0] = float(penalty_init)
        self.state[4] = float(th.nn.functional.softplus(th.tensor(float(penalty_init), dtype=th.float32)))
        self.clamp_at = penalty_init if clamp_at is None else clamp_at

    @property
    def log_nu(self) -> th.Tensor:
        return self.state[0:1]

    def forward(self) -> th.Tensor:
        return th.nn.functional.softplus(self.state[0:1])

    __call__ = forward

    def clamp_min(self) -> float:
        return float(np.log(max(np.exp(self.clamp_at) - 1, 1e-8)))           # dual_variable.py:27-29


class DualVariable:
    def __init__(self, alpha=0, learning_rate=10, penalty_init=1, clamp_at=None, device="cuda"):
        self.nu = Nu(penalty_init, clamp_at, device)
        self.alpha = alpha
        self.learning_rate = learning_rate
        self.loss = th.tensor(0)
        self.steps = 0

    def update_parameter(self, cost) -> None:
        """cost: python/numpy scalar (already averaged, as the reference passes) or a device float32 tensor of the
        rollout's original costs (averaged on the device)."""
        st = self.nu.state
        if isinstance(cost, th.Tensor) and cost.is_cuda:
            costs = cost.reshape(-1).to(th.float32).contiguous()
        else:
            costs = th.full((1,), float(np.float32(cost)), dtype=th.float32, device=st.device)
        with th.cuda.device(st.device):
            _lib.check(_lib.lib().icrl_dual_update(_lib.ptr(st), _lib.ptr(costs), costs.numel(), float(self.alpha),
                                                   float(self.learning_rate), self.steps, self.nu.clamp_min(),
                                                   _lib.current_stream()))
        self.steps += 1
        self.loss = st[3]


class PIDLagrangian:
    """dual_variable.py:60-122 (scalar host arithmetic, kept in Python as in the reference)."""

    def __init__(self, alpha=0, penalty_init=1, Kp=0, Kd=0, Ki=1, pid_delay=10, delta_d_ema_alpha=0.95,
                 delta_p_ema_alpha=0.95):
        self.budget, self.Kp, self.Ki, self.Kd, self.pid_delay = alpha, Kp, Ki, Kd, pid_delay
        self.pid_i = self.cost_penalty = penalty_init
        self.cost_deltas = deque([0], maxlen=pid_delay)
        self._delta_p = self._cost_delta = 0
        self.delta_d_ema_alpha, self.delta_p_ema_alpha = delta_d_ema_alpha, delta_p_ema_alpha

    def update_parameter(self, cost):
        cost = float(cost.mean()) if isinstance(cost, th.Tensor) else float(cost)
        self.loss = th.tensor(cost)
        delta = cost - self.budget
        self.pid_i = max(0, self.pid_i + self.Ki * delta)
        self._delta_p = self.delta_p_ema_alpha * self._delta_p + (1 - self.delta_p_ema_alpha) * delta
        self._cost_delta = self.delta_d_ema_alpha * self._cost_delta + (1 - self.delta_d_ema_alpha) * cost
        pid_d = max(0, self._cost_delta - self.cost_deltas[0])
        self.cost_penalty = max(0, self.Kp * self._delta_p + self.Kd * pid_d + self.pid_i)
        self.cost_deltas.append(self._cost_delta)

    def nu(self):
        return th.tensor(self.cost_penalty)
