"""GPU parity: K3 (dual GAE) -- bit exact against the reference goldens and the oracle, plus properties at full size."""
import glob
import os

import numpy as np
import pytest
import torch as th

from conftest import GOLDEN, load_golden
from oracle import gae as ogae

pytestmark = pytest.mark.gpu
K3_CASES = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "k3_*.npz")))
OUT = ("reward_returns", "reward_advantages", "cost_returns", "cost_advantages")


def run_buffer(d, T, E, g):
    from icrl_b200.buffers import RolloutBufferWithCost
    from icrl_b200.spaces import Box
    buf = RolloutBufferWithCost(T, Box(-1, 1, (3,)), Box(-1, 1, (2,)), "cuda", reward_gamma=g[0], reward_gae_lambda=g[1],
                                cost_gamma=g[2], cost_gae_lambda=g[3], n_envs=E)
    for k in ("rewards", "reward_values", "costs", "cost_values", "dones"):
        getattr(buf, k)[:] = d[k]
    buf.compute_returns_and_advantage(th.tensor(d["reward_last_value"]), th.tensor(d["cost_last_value"]), d["last_dones"])
    return buf


@pytest.mark.parametrize("case", K3_CASES)
def test_gae_bit_exact_vs_reference_golden(case):
    d = load_golden(f"k3_{case}")
    T, E = d["rewards"].shape
    buf = run_buffer(d, T, E, [float(x) for x in d["gammas"]])
    for k in OUT:
        got = getattr(buf, k)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, d[k], err_msg=k)


def synth(rng, T, E, ep_len):
    d = {k: rng.standard_normal((T, E)).astype(np.float32) for k in ("rewards", "reward_values", "costs", "cost_values")}
    phase = rng.integers(0, ep_len, size=E)
    d["dones"] = (((np.arange(T)[:, None] + phase[None]) % ep_len) == 0).astype(np.float32)
    d["last_dones"] = rng.random(E) < 0.5
    d["reward_last_value"] = rng.standard_normal((E, 1)).astype(np.float32)
    d["cost_last_value"] = rng.standard_normal((E, 1)).astype(np.float32)
    return d


@pytest.mark.parametrize("T,E,ep", [(1, 1, 5), (2, 33, 3), (63, 5, 7), (64, 5, 200), (257, 31, 150), (1000, 64, 1000),
                                    (2048, 96, 500), (300, 1, 1)])
def test_gae_bit_exact_vs_oracle_ragged(T, E, ep):
    rng = np.random.default_rng(T * 1000 + E)
    d = synth(rng, T, E, ep)
    g = [0.99, 0.95, 0.97, 0.9]
    want = ogae.dual_gae(d["rewards"], d["reward_values"], d["costs"], d["cost_values"], d["dones"],
                         d["reward_last_value"], d["cost_last_value"], d["last_dones"], *g)
    buf = run_buffer(d, T, E, g)
    for k in OUT:
        np.testing.assert_array_equal(getattr(buf, k), want[k], err_msg=k)


def test_gae_full_size_properties():
    """4M transitions (T=2048, E=2048) on device-resident buffers: (i) lambda=1, no dones => advantage + value equals the
    discounted return with bootstrap (buffers.py:508-511 docstring); (ii) returns == advantages + values exactly;
    (iii) a `done` at t+1 cuts every dependence on the future; (iv) columns are independent."""
    import ctypes as C
    from icrl_b200 import _lib
    T, E = 2048, 2048
    g = th.Generator(device="cuda").manual_seed(1)
    r, vr, c, vc = (th.randn(T, E, device="cuda", generator=g) for _ in range(4))
    dones = th.zeros(T, E, device="cuda")
    lvr, lvc = th.randn(E, device="cuda", generator=g), th.randn(E, device="cuda", generator=g)
    last = th.zeros(E, dtype=th.uint8, device="cuda")
    outs = [th.empty(T, E, device="cuda") for _ in range(4)]

    def run(dn):
        _lib.check(_lib.lib().icrl_dual_gae(*[_lib.ptr(x) for x in (r, vr, c, vc, dn, lvr, lvc, last)], T, E, 0.99, 1.0,
                                            0.9, 1.0, *[_lib.ptr(o) for o in outs], _lib.current_stream()))
        th.cuda.synchronize()
        return [o.clone() for o in outs]

    ar, rr, ac, rc = run(dones)
    assert th.equal(rr, ar + vr) and th.equal(rc, ac + vc)
    # discounted return in float64 on a few columns
    cols = [0, 1, 777, E - 1]
    ret = lvr[cols].double()
    R = th.empty(T, len(cols), dtype=th.float64, device="cuda")
    for t in range(T - 1, -1, -1):
        ret = r[t, cols].double() + 0.99 * ret
        R[t] = ret
    assert th.allclose(rr[:, cols].double(), R, rtol=2e-5, atol=2e-5)
    # a done at t0+1 makes [0, t0] independent of everything after t0
    t0 = 1000
    d2 = dones.clone()
    d2[t0 + 1] = 1.0
    ar2 = run(d2)[0]
    r_saved = r[t0 + 1:].clone()
    r[t0 + 1:] += 3.0
    ar3 = run(d2)[0]
    r[t0 + 1:] = r_saved
    assert th.equal(ar2[:t0 + 1], ar3[:t0 + 1]) and not th.equal(ar2[t0 + 1:], ar3[t0 + 1:])
