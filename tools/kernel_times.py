"""Device-time of the streaming kernels (K1 relabel, K3 dual GAE, K5 cost normalisation) at the rollout size and at sweep sizes,
per env shape, against the measured HBM peak.  Usage: python tools/kernel_times.py [k1,k3,k5] [sizes...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch as th  # noqa: E402

from icrl_b200 import _lib  # noqa: E402
from icrl_b200.constraint_net import ConstraintNet  # noqa: E402
from icrl_b200.learner import WORKLOADS  # noqa: E402

which = (sys.argv[1] if len(sys.argv) > 1 else "k1,k3,k5").split(",")
sizes = [int(x) for x in sys.argv[2:]] or [10240, 1 << 20, 1 << 22, 1 << 24]
L = _lib.lib()
peak = 6549.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = th.zeros(64 * 1024 * 1024, device="cuda")


def timeit(fn, reps=5):
    fn(); th.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)
        s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); th.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    return float(np.median(ts))


for n in sizes:
    T = 2048
    E = max(n // T, 1)
    n = T * E
    if "k3" in which:
        arrs = [th.randn(T, E, device="cuda") for _ in range(4)] + [(th.rand(T, E, device="cuda") < 0.002).float()]
        lv = [th.randn(E, device="cuda") for _ in range(2)] + [th.zeros(E, dtype=th.uint8, device="cuda")]
        outs = [th.empty(T, E, device="cuda") for _ in range(4)]
        d = timeit(lambda: _lib.check(L.icrl_dual_gae(*[_lib.ptr(x) for x in arrs + lv], T, E, 0.99, 0.95, 0.99, 0.95,
                                                      *[_lib.ptr(o) for o in outs], _lib.current_stream())))
        print(f"k3 T={T} E={E} rows={n}: {d * 1e6:.1f} us  {36 * n / d / 1e9:.0f} GB/s  {36 * n / d / 1e9 / peak:.3f} of HBM peak")
        if "k5" in which and E <= 64:
            state = th.tensor([0.0, 1.0, 1e-4] + [0.0] * E, dtype=th.float64, device="cuda")
            d = timeit(lambda: _lib.check(L.icrl_cost_normalize(_lib.ptr(arrs[0]), _lib.ptr(arrs[4]), _lib.ptr(lv[2]), T, E, 0.99,
                                                                1e-8, 10.0, 1, 1, _lib.ptr(state), _lib.ptr(outs[0]),
                                                                _lib.current_stream())))
            print(f"k5 T={T} E={E}: {d * 1e6:.1f} us")
    if "k1" in which:
        for name in ("halfcheetah", "lapgrid", "antwall", "pointcircle"):
            w = WORKLOADS[name]
            low = high = None
            if not w.is_discrete:
                low, high = -np.ones(w.act_dim, np.float32), np.ones(w.act_dim, np.float32)
            th.manual_seed(0)
            kw = {}
            if name == "pointcircle":
                kw = dict(obs_select_dim=[0, 1], acs_select_dim=[-1])
            cn = ConstraintNet(w.obs_dim, w.act_dim, w.cn_hidden, None, lambda _: 0.01, None, None, w.is_discrete, 0.5,
                               clip_obs=20., action_low=low, action_high=high, **kw)
            obs = th.randn(n, w.obs_dim, device="cuda") * 3
            acs = (th.randint(0, w.act_dim, (n,), device="cuda").float() if w.is_discrete else th.randn(n, w.act_dim, device="cuda"))
            cost = th.empty(n, device="cuda")
            desc = cn._get_desc()
            d = timeit(lambda: _lib.check(L.icrl_cn_forward(C.byref(desc), _lib.ptr(obs), 0, _lib.ptr(acs), n, _lib.ptr(cost), 0,
                                                            _lib.current_stream())))
            b = (cn.input_dims + 1) * 4
            fl = 2 * sum(a * bb for a, bb in zip([cn.input_dims, *w.cn_hidden], [*w.cn_hidden, 1]))
            print(f"k1 {name:12s} rows={n}: {d * 1e6:.1f} us  {b * n / d / 1e9:.0f} GB/s  {b * n / d / 1e9 / peak:.3f} of HBM peak  "
                  f"{fl * n / d / 1e12:.2f} TFLOP/s")
