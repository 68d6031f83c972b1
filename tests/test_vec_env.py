"""Host-side callers of the hot path: VecCostWrapper / VecNormalizeWithCost against fixtures produced by the
reference's own classes on the scripted environment (tests/golden/make_golden.py::golden_venv), bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden
from icrl_b200 import vec_env

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from scripted_env import ScriptedEnv, scripted_actions, scripted_cost  # noqa: E402

CASES = {"all": dict(norm_obs=True, norm_reward=True, norm_cost=True),
         "nocost": dict(norm_obs=True, norm_reward=True, norm_cost=False),
         "raw": dict(norm_obs=False, norm_reward=False, norm_cost=False)}


def _train_env(kw, n_envs=3):
    env = vec_env.DummyVecEnv([(lambda i=i: ScriptedEnv(100 + i)) for i in range(n_envs)])
    env = vec_env.VecCostWrapper(env)
    env = vec_env.VecNormalizeWithCost(env, training=True, cost_info_str="cost", reward_gamma=0.99, cost_gamma=0.97,
                                       **kw)
    env.set_cost_function(scripted_cost)
    return env


@pytest.mark.parametrize("name", sorted(CASES))
def test_wrappers_match_reference(name):
    g = load_golden(f"venv_{name}")
    env = _train_env(CASES[name])
    steps = g["obs"].shape[0]
    acts = scripted_actions(7, steps, 3)
    np.testing.assert_array_equal(env.reset(), g["obs0"])
    for t in range(steps):
        o, r, d, infos = env.step(acts[t])
        np.testing.assert_array_equal(o, g["obs"][t])
        np.testing.assert_array_equal(env.get_original_obs(), g["orig_obs"][t])
        np.testing.assert_array_equal(r, g["rew"][t])
        np.testing.assert_array_equal(d, g["done"][t])
        np.testing.assert_array_equal(np.array([i["cost"] for i in infos]), g["cost"][t])
        np.testing.assert_array_equal(env.get_original_cost(), g["orig_cost"][t])
    for rms in ("obs_rms", "ret_rms", "cost_rms"):
        r = getattr(env, rms)
        np.testing.assert_array_equal(r.mean, g[rms + "_mean"])
        np.testing.assert_array_equal(r.var, g[rms + "_var"])
        assert r.count == float(g[rms + "_count"])
    ev = vec_env.VecNormalizeWithCost(vec_env.DummyVecEnv([lambda: ScriptedEnv(555)]), training=False,
                                      norm_obs=CASES[name]["norm_obs"], norm_reward=False, norm_cost=False)
    vec_env.sync_envs_normalization(env, ev)
    np.testing.assert_array_equal(ev.reset(), g["eval_obs0"])
    np.testing.assert_array_equal(ev.step(acts[0][:1])[0], g["eval_obs1"])


def test_cost_wrapper_uses_previous_obs():
    """vec_cost_wrapper.py:37-41: the cost of step t is evaluated on the observation the action was taken FROM."""
    seen = []

    def cost_fn(obs, acs):
        seen.append(obs.copy())
        return np.zeros(obs.shape[0], np.float32)
    env = vec_env.VecCostWrapper(vec_env.DummyVecEnv([lambda: ScriptedEnv(3)]))
    env.set_cost_function(cost_fn)
    o0 = env.reset()
    o1, *_ = env.step(np.zeros((1, 2), np.float32))
    env.step(np.zeros((1, 2), np.float32))
    np.testing.assert_array_equal(seen[0], o0)
    np.testing.assert_array_equal(seen[1], o1)


def test_vecnormalize_pickle_roundtrip(tmp_path):
    env = _train_env(CASES["all"])
    env.reset()
    acts = scripted_actions(7, 10, 3)
    for t in range(10):
        env.step(acts[t])
    path = str(tmp_path / "train_env_stats.pkl")
    env.save(path)
    fresh = vec_env.VecCostWrapper(vec_env.DummyVecEnv([(lambda i=i: ScriptedEnv(100 + i)) for i in range(3)]))
    loaded = vec_env.VecNormalize.load(path, fresh)
    assert isinstance(loaded, vec_env.VecNormalizeWithCost)
    np.testing.assert_array_equal(loaded.obs_rms.mean, env.obs_rms.mean)
    np.testing.assert_array_equal(loaded.cost_rms.var, env.cost_rms.var)
    assert loaded.venv is fresh and loaded.ret.shape == (3,)


def test_synthetic_envs_shapes():
    from icrl_b200 import envs
    from icrl_b200.true_constraint_net import get_true_cost_function
    e = envs.make("SynthHCWithPosTest-v0")
    o = e.reset()
    assert o.shape == (18,) and e.action_space.shape == (6,)
    o, r, d, info = e.step(np.ones(6, np.float32))
    assert o.shape == (18,) and isinstance(r, float) and d is False
    cost = get_true_cost_function("SynthHCWithPosTest-v0")
    assert cost(np.array([[-3.5] + [0] * 17, [0.0] + [0] * 17]), None).tolist() == [True, False]
    g = envs.make("SynthCLGW-v0")
    assert g.reset().shape == (1,) and g.action_space.n == 2
    with pytest.raises(ImportError):
        envs.make("HCWithPos-v0")


def test_dummy_vec_env_reports_episode_summaries():
    """The reference wraps every env in Monitor: at episode end info['episode'] = {r: raw reward sum, l: length, t: time}."""
    env = vec_env.DummyVecEnv([lambda: ScriptedEnv(21), lambda: ScriptedEnv(22)])
    env.reset()
    acc, n, seen = [0.0, 0.0], [0, 0], 0
    for _ in range(200):
        _, rews, dones, infos = env.step(np.zeros((2, 2), np.float32))
        for i in range(2):
            acc[i] += float(rews[i]); n[i] += 1
            if dones[i]:
                ep = infos[i]["episode"]
                assert ep["l"] == n[i] and abs(ep["r"] - acc[i]) < 1e-4 and ep["t"] >= 0 and "terminal_observation" in infos[i]
                acc[i], n[i], seen = 0.0, 0, seen + 1
            else:
                assert "episode" not in infos[i]
    assert seen > 5
