"""Shared helpers for the parity tests: rebuild oracle-side objects from golden fixtures."""
import numpy as np
import torch as th

from oracle import cn as ocn
from oracle import ppo as oppo

SHAPES = {
    "lgw": dict(obs_dim=1, acs_dim=2, is_discrete=True, hidden=(20,)),
    "hc": dict(obs_dim=18, acs_dim=6, is_discrete=False, hidden=(20,)),
    "ant": dict(obs_dim=113, acs_dim=8, is_discrete=False, hidden=(40, 40)),
    "point": dict(obs_dim=6, acs_dim=2, is_discrete=False, hidden=(40, 40)),
}

# tolerances (BASELINE.json north_star): costs / advantages rel <= 1e-5, params after one update <= 1e-4.
# `1 - sigmoid` in the reference is itself only resolved to 1 ulp(1.0) = 6e-8 absolute, hence the atol.
COST_RTOL, COST_ATOL = 1e-5, 2e-7
PARAM_RTOL = 1e-4


def cn_params(d, prefix="p."):
    """[W0, b0, W1, b1, ...] float32 tensors from a golden dict holding `<prefix>{0,2,4}.{weight,bias}`."""
    out, i = [], 0
    while f"{prefix}{i}.weight" in d:
        out += [th.tensor(d[f"{prefix}{i}.weight"]), th.tensor(d[f"{prefix}{i}.bias"])]
        i += 2
    return out


def cn_spec(shape, d, **kw):
    s = SHAPES[shape]
    low = high = None
    if not s["is_discrete"]:
        low, high = -np.ones(s["acs_dim"], np.float32), np.ones(s["acs_dim"], np.float32)
    base = dict(obs_dim=s["obs_dim"], acs_dim=s["acs_dim"], hidden_sizes=s["hidden"], is_discrete=s["is_discrete"],
                clip_obs=20., obs_mean=d.get("obs_mean"), obs_var=d.get("obs_var"), action_low=low, action_high=high)
    base.update(kw)
    return ocn.CNSpec(**base)


def policy_params(d, prefix="p0."):
    names = [str(n) for n in d["param_order"]]
    from collections import OrderedDict
    return OrderedDict((n, th.tensor(d[prefix + n])) for n in names)


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor + 1e-300)))


def max_param_err(pa, pb):
    """max over tensors of ||a-b||_inf / max(||b||_inf, 1e-3)  (per-tensor scale, so zero-init biases don't divide by 0)."""
    worst = 0.0
    for a, b in zip(pa, pb):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        worst = max(worst, float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-3)))
    return worst
