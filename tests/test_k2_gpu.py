"""GPU parity: K2 (IS-weighted constraint-net training) through ConstraintNet.train() against goldens produced by the
unmodified reference's ConstraintNet.train() (two consecutive calls, so Adam state carries over)."""
import glob
import os

import numpy as np
import pytest
import torch as th

from conftest import GOLDEN, load_golden
from helpers import PARAM_RTOL, SHAPES, max_param_err

pytestmark = pytest.mark.gpu
K2_CASES = sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "k2_*.npz")))


def build(d, shape):
    from icrl_b200.constraint_net import ConstraintNet
    s = SHAPES[shape]
    low = high = None
    if not s["is_discrete"]:
        low, high = -np.ones(s["acs_dim"], np.float32), np.ones(s["acs_dim"], np.float32)
    lr = float(d["lr"])
    batch_size = int(d["batch_size"]) if int(d.get("batch_size", 0)) > 0 else None
    cn = ConstraintNet(s["obs_dim"], s["acs_dim"], s["hidden"], batch_size, lambda x: lr, d["expert_obs"], d["expert_acs"],
                       s["is_discrete"], float(d["reg"]), no_importance_sampling=bool(d["no_is"]),
                       per_step_importance_sampling=bool(d["per_step"]), clip_obs=20., initial_obs_mean=d.get("obs_mean"),
                       initial_obs_var=d.get("obs_var"), action_low=low, action_high=high,
                       target_kl_old_new=float(d["tkon"]), target_kl_new_old=float(d["tkno"]),
                       train_gail_lambda=bool(d.get("gail", False)))
    cn.load_network_state_dict({k[3:]: th.tensor(v) for k, v in d.items() if k.startswith("p0.")})
    return cn


def net_params(cn):
    return [v.numpy() for v in cn.network.state_dict().values()]


@pytest.mark.parametrize("case", K2_CASES)
def test_train_matches_reference(case):
    d = load_golden(f"k2_{case}")
    cn = build(d, case.split("_")[0])
    if "numpy_seed" in d:                 # -cbs fixtures: the minibatch permutations come from the global numpy RNG
        np.random.seed(int(d["numpy_seed"]))
    for call in (1, 2):
        if f"p{call}.0.weight" not in d:
            break
        m = cn.train(int(d["iters"]), d["nominal_obs"], d["nominal_acs"], d["lengths"], d.get("obs_mean"), d.get("obs_var"), 1.0)
        if f"rng_pos{call}" in d:         # ... and is left where the reference leaves it (early stops included)
            st = np.random.get_state()
            assert st[2] == int(d[f"rng_pos{call}"]) and (st[1][:4].astype(np.int64) == d[f"rng_probe{call}"]).all()
        ref = [d[k] for k in sorted(k for k in d if k.startswith(f"p{call}."))]
        # sorted() orders 0.bias before 0.weight; state_dict order is weight, bias -> compare by key instead
        sd = cn.network.state_dict()
        err = max_param_err([sd[k[3:]].numpy() for k in d if k.startswith(f"p{call}.")],
                            [d[k] for k in d if k.startswith(f"p{call}.")])
        assert err <= PARAM_RTOL, f"call {call}: params {err}"
        assert int(m.get("backward/early_stop_itr", d["iters"])) == int(d.get(f"m{call}.backward/early_stop_itr", d["iters"]))
        assert cn.optimizer.step_count == int(d[f"c{call}.adam.0.step"])
        for k, v in m.items():
            r = float(d[f"m{call}.{k}"])
            if np.isnan(r) or np.isinf(r):
                assert (np.isnan(v) and np.isnan(r)) or v == r, (k, v, r)
            else:
                assert abs(v - r) <= 2e-4 * max(abs(r), 1e-2), (k, v, r)


def test_iteration_zero_known_answers():
    """at backward iteration 0 current_preds == start_preds, so ratio == 1, kl_old_new == -log(1+eps),
    kl_new_old == 0 and every IS weight is 1 (fp32)."""
    d = load_golden("k2_hc_perstep_mild")
    cn = build(d, "hc")
    m = cn.train(1, d["nominal_obs"], d["nominal_acs"], d["lengths"], None, None, 1.0)
    assert m["backward/is_mean"] == 1.0 and m["backward/is_max"] == 1.0 and m["backward/is_min"] == 1.0
    assert abs(m["backward/kl_old_new"] - (-np.log(np.float32(1.0) + np.float32(1e-5)))) < 1e-9
    assert m["backward/kl_new_old"] == 0.0


def test_adam_state_roundtrip_and_save_load(tmp_path):
    """cn.pt written by save() has the reference's keys and loads back (constraint_net.py:323-402)."""
    from icrl_b200.constraint_net import ConstraintNet
    d = load_golden("k2_hc_nois")
    cn = build(d, "hc")
    cn.train(int(d["iters"]), d["nominal_obs"], d["nominal_acs"], d["lengths"], None, None, 1.0)
    path = str(tmp_path / "cn.pt")
    cn.save(path)
    sd = th.load(path, weights_only=False)
    assert set(sd) == {"cn_network", "cn_optimizer", "obs_dim", "acs_dim", "is_discrete", "obs_select_dim",
                       "acs_select_dim", "clip_obs", "obs_mean", "obs_var", "action_low", "action_high", "device",
                       "hidden_sizes"}
    m = [sd["cn_optimizer"]["state"][i]["exp_avg"].numpy() for i in range(4)]
    assert max_param_err(m, [d[f"c1.adam.{i}.exp_avg"] for i in range(4)]) <= 1e-4
    cn2 = ConstraintNet.load(path)
    obs, acs = d["nominal_obs"][:50], d["nominal_acs"][:50]
    # the loaded net does not clip (load() quirk); compare against an un-clipped forward of the trained net
    cn.clip_obs, cn.action_low, cn.action_high = None, None, None
    np.testing.assert_array_equal(cn2.cost_function(obs, acs), cn.cost_function(obs, acs))
