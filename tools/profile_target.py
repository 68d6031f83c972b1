"""Small, fixed workloads for ncu captures (one process, one GPU).  Usage: python tools/profile_target.py k4|k1|k3|k2|iter [workload]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402

from icrl_b200 import _lib  # noqa: E402
from icrl_b200.learner import WORKLOADS, DeviceLearner  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "iter"
wname = sys.argv[2] if len(sys.argv) > 2 else "halfcheetah"
w = WORKLOADS[wname]
w = type(w)(**{**w.__dict__, "rollouts": 2})
learner = DeviceLearner(w, seed=0)
L = _lib.lib()
th.cuda.synchronize()
if what == "iter":
    for _ in range(2):
        learner.run()
elif what == "k4":
    learner.max_steps = int(os.environ.get("K4_STEPS", "320"))
    w2 = type(w)(**{**w.__dict__, "backward_iters": 0})
    learner.w = w2
    for _ in range(2):
        learner.run()
elif what in ("k1", "k3"):
    n_rep = 3
    desc = learner.cn._get_desc()
    for T, E in ((w.n_steps, learner.E), (2048, 2048)):
        n = T * E
        obs = th.randn(n, w.obs_dim, device="cuda")
        acs = th.randint(0, w.act_dim, (n,), device="cuda").float() if w.is_discrete else th.randn(n, w.act_dim, device="cuda")
        cost = th.empty(n, device="cuda")
        arrs = [th.randn(T, E, device="cuda") for _ in range(4)] + [(th.rand(T, E, device="cuda") < 0.002).float()]
        lv = [th.randn(E, device="cuda") for _ in range(2)] + [th.zeros(E, dtype=th.uint8, device="cuda")]
        outs = [th.empty(T, E, device="cuda") for _ in range(4)]
        for _ in range(n_rep):
            if what == "k1":
                _lib.check(L.icrl_cn_forward(C.byref(desc), _lib.ptr(obs), 0, _lib.ptr(acs), n, _lib.ptr(cost), 0, _lib.current_stream()))
            else:
                _lib.check(L.icrl_dual_gae(*[_lib.ptr(x) for x in arrs + lv], T, E, 0.99, 0.95, 0.99, 0.95,
                                           *[_lib.ptr(o) for o in outs], _lib.current_stream()))
elif what == "k2":
    learner._cn_train_device()
th.cuda.synchronize()
print("done", what, wname)
