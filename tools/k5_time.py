"""ICRL_K5_TIMING=1 python tools/k5_time.py -- per-phase cycles of the fused cost-normalisation kernel."""
import numpy as np
import torch as th
from icrl_b200 import _lib
T, E = 2048, 5
d = "cuda"
o = th.rand(T, E, device=d); dn = (th.rand(T, E, device=d) < 0.002).float(); last = th.zeros(E, dtype=th.uint8, device=d)
st = th.tensor([0.0, 1.0, 1e-4] + [0.0] * E, dtype=th.float64, device=d); out = th.empty(T, E, device=d)
for _ in range(3):
    _lib.check(_lib.lib().icrl_cost_normalize(_lib.ptr(o), _lib.ptr(dn), _lib.ptr(last), T, E, 0.99, 1e-8, 10.0, 1, 1,
                                              _lib.ptr(st), _lib.ptr(out), _lib.current_stream()))
th.cuda.synchronize()
