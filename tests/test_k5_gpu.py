"""K5 (whole-rollout cost normalisation) against the oracle restatement of VecNormalizeWithCost's online
arithmetic -- bit-exact, float64 statistics included -- and the whole-buffer relabel mode of collect_rollouts against
the reference's per-step mode."""
import os
import sys

import numpy as np
import pytest
import torch as th

from conftest import load_golden
from icrl_b200 import _lib

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def _run_k5(orig, news, state, gamma=0.99, eps=1e-8, clip=10.0, norm=True, training=True):
    T, E = orig.shape
    dones = np.zeros((T, E), np.float32)
    dones[1:] = news[:-1]
    d = "cuda"
    st = th.from_numpy(np.concatenate([[state["mean"], state["var"], state["count"]], state["cost_ret"]])).to(d)
    o, dn = th.from_numpy(orig.astype(np.float32)).to(d), th.from_numpy(dones).to(d)
    last = th.from_numpy(news[-1].astype(np.uint8)).to(d)
    out = th.empty(T, E, device=d)
    _lib.check(_lib.lib().icrl_cost_normalize(_lib.ptr(o), _lib.ptr(dn), _lib.ptr(last), T, E, gamma, eps, clip, int(norm),
                                              int(training), _lib.ptr(st), _lib.ptr(out), _lib.current_stream()))
    st = st.cpu().numpy()
    return out.cpu().numpy(), dict(mean=st[0], var=st[1], count=st[2], cost_ret=st[3:])


@pytest.mark.parametrize("name,norm", [("all", True), ("nocost", False)])
def test_matches_reference_wrapper_fixture(name, norm):
    from oracle import costnorm
    g = load_golden(f"venv_{name}")
    E = g["orig_cost"].shape[1]
    out, st = _run_k5(g["orig_cost"], g["done"], costnorm.reset(costnorm.initial_state(E)), gamma=0.97, norm=norm)
    np.testing.assert_array_equal(out, g["cost"].astype(np.float32))
    assert st["mean"] == g["cost_rms_mean"] and st["var"] == g["cost_rms_var"] and st["count"] == float(g["cost_rms_count"])


@pytest.mark.parametrize("T,E", [(1, 1), (7, 3), (64, 5), (2048, 5), (33, 8), (50, 13), (40, 40), (20, 128), (9, 130),
                                 (5, 300), (3, 1000)])
@pytest.mark.parametrize("training", [True, False])
def test_bit_exact_vs_oracle(T, E, training):
    from oracle import costnorm
    rng = np.random.default_rng(T * 1000 + E)
    orig = rng.uniform(0, 1, (T, E)).astype(np.float32) ** 3
    news = rng.random((T, E)) < 0.05
    s0 = costnorm.reset(costnorm.initial_state(E))
    # a used state: two earlier rollouts' worth of statistics
    costnorm.normalize_rollout(rng.uniform(0, 1, (11, E)).astype(np.float32), rng.random((11, E)) < 0.1, s0)
    s_ref = {k: np.copy(v) for k, v in s0.items()}
    ref = costnorm.normalize_rollout(orig, news, s_ref, training=training)
    out, st = _run_k5(orig, news, s0, training=training)
    np.testing.assert_array_equal(out, ref)
    for k in ("mean", "var", "count"):
        assert st[k] == s_ref[k], k
    if training:
        np.testing.assert_array_equal(st["cost_ret"], s_ref["cost_ret"])


def _collect(whole_buffer):
    from scripted_env import ScriptedEnv
    from icrl_b200 import vec_env
    from icrl_b200.constraint_net import ConstraintNet
    from icrl_b200.ppo_lag import PPOLagrangian
    th.manual_seed(0)
    np.random.seed(0)
    rng = np.random.default_rng(0)
    cn = ConstraintNet(5, 2, (20,), None, lambda _: 1e-3, rng.standard_normal((50, 5)), rng.standard_normal((50, 2)),
                       False, 0.5, clip_obs=20, action_low=-np.ones(2, np.float32), action_high=np.ones(2, np.float32))
    env = vec_env.DummyVecEnv([(lambda i=i: ScriptedEnv(100 + i)) for i in range(4)])
    env = vec_env.VecCostWrapper(env)
    env = vec_env.VecNormalizeWithCost(env, training=True, cost_info_str="cost", reward_gamma=0.99, cost_gamma=0.99)
    env.set_cost_function(None if whole_buffer else cn.cost_function)
    model = PPOLagrangian("TwoCriticsMlpPolicy", env, n_steps=96, batch_size=64, n_epochs=1, seed=4)
    for _ in range(2):   # two rollouts: the statistics carry over
        model.learn(total_timesteps=96 * 4, cost_function=cn if whole_buffer else "cost", reset_num_timesteps=False)
    b = model.rollout_buffer
    return {k: b.time_major(k).copy() for k in ("costs", "orig_costs", "cost_advantages", "actions")}, env


def test_whole_buffer_relabel_equals_per_step_mode():
    a, env_a = _collect(False)
    b, env_b = _collect(True)
    np.testing.assert_array_equal(a["actions"], b["actions"])
    np.testing.assert_array_equal(a["orig_costs"], b["orig_costs"])
    np.testing.assert_array_equal(a["costs"], b["costs"])
    np.testing.assert_array_equal(a["cost_advantages"], b["cost_advantages"])
    assert env_a.cost_rms.var == env_b.cost_rms.var and env_a.cost_rms.mean == env_b.cost_rms.mean
    assert env_a.cost_rms.count == env_b.cost_rms.count
    np.testing.assert_array_equal(env_a.cost_ret, env_b.cost_ret)
