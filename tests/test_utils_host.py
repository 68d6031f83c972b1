"""Host-side driver helpers with a scripted agent (no GPU): the nominal sampler, evaluate_policy and the environment factory
keep the reference's conventions (icrl/utils.py:247-360, stable_baselines3/common/evaluation.py)."""
import os
import sys

import numpy as np

from icrl_b200 import utils, vec_env

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from scripted_env import ScriptedEnv  # noqa: E402


class _Agent:
    def __init__(self):
        self.calls = 0

    def predict(self, obs, state=None, deterministic=False):
        self.calls += 1
        return np.zeros((obs.shape[0], 2), np.float32), state


def _eval_env(seed=9):
    env = vec_env.DummyVecEnv([lambda: ScriptedEnv(seed)])
    return vec_env.VecNormalizeWithCost(env, training=False, norm_obs=True, norm_reward=False, norm_cost=False)


def test_sample_from_agent_shapes_and_episode_accounting():
    env, agent = _eval_env(), _Agent()
    orig, obs, acs, rewards, lengths = utils.sample_from_agent(agent, env, 4)
    n = int(lengths.sum())
    assert orig.shape == (n, 5) and obs.shape == (n, 5) and acs.shape == (n, 2)
    assert rewards.shape == (4,) and lengths.shape == (4,) and agent.calls == n
    # normalised observations are the clipped standardisation of the originals with the (frozen) statistics
    want = np.clip((orig - env.obs_rms.mean) / np.sqrt(env.obs_rms.var + env.epsilon), -env.clip_obs, env.clip_obs)
    np.testing.assert_allclose(obs, want, rtol=1e-6)
    assert env.obs_rms.count == 1e-4            # training=False: statistics untouched


def test_evaluate_policy_counts_whole_episodes():
    env, agent = _eval_env(3), _Agent()
    rewards, lengths = utils.evaluate_policy(agent, env, n_eval_episodes=3, return_episode_rewards=True)
    assert len(rewards) == 3 and len(lengths) == 3 and sum(lengths) == agent.calls
    mean, std = utils.evaluate_policy(_Agent(), _eval_env(3), n_eval_episodes=3)
    assert np.isclose(mean, np.mean(rewards)) and np.isclose(std, np.std(rewards))


def test_make_train_env_wrapper_stack():
    env = utils.make_train_env("SynthHCWithPos-v0", None, True, base_seed=1, num_threads=3, cost_info_str="cost",
                               reward_gamma=0.99, cost_gamma=0.98)
    assert isinstance(env, vec_env.VecNormalizeWithCost) and isinstance(env.venv, vec_env.VecCostWrapper)
    assert env.num_envs == 3 and env.cost_gamma == 0.98 and env.training
    env.set_cost_function(lambda o, a: np.full(o.shape[0], 0.25, np.float32))       # forwarded to the cost wrapper
    env.reset()
    _, _, _, infos = env.step(np.zeros((3, 6), np.float32))
    assert all("cost" in i for i in infos) and np.allclose(env.get_original_cost(), 0.25)
    ev = utils.make_eval_env("SynthHCWithPosTest-v0", use_cost_wrapper=False)
    assert ev.num_envs == 1 and not ev.training and not ev.norm_reward
    vec_env.sync_envs_normalization(env, ev)
    np.testing.assert_array_equal(ev.obs_rms.mean, env.obs_rms.mean)


def test_speculated_permutations_follow_the_numpy_stream():
    """PPOLagrangian pre-draws the next train()'s permutations while its kernel runs; they are used only when the global
    RNG is found exactly where the speculation assumed, and then leave the stream where drawing them late would have."""
    import types
    from icrl_b200.ppo_lag import PPOLagrangian
    algo = types.SimpleNamespace(n_epochs=3)
    for name in ("_draw_permutations", "_speculate_permutations", "_take_speculated_permutations"):
        setattr(algo, name, types.MethodType(getattr(PPOLagrangian, name), algo))
    n = 50
    np.random.seed(7)
    want1, _ = algo._draw_permutations(n)
    want2, _ = algo._draw_permutations(n)
    end_state = np.random.get_state()

    np.random.seed(7)
    perms1, states1 = algo._draw_permutations(n)
    algo._speculate_permutations(n, states1[-1])
    np.random.set_state(states1[-1])                       # train() ran every epoch
    perms2, states2 = algo._take_speculated_permutations(n)
    np.testing.assert_array_equal(perms1, want1)
    np.testing.assert_array_equal(perms2, want2)
    assert np.array_equal(np.random.get_state()[1], end_state[1]) and np.random.get_state()[2] == end_state[2]

    # an early stop (state rewound to after epoch 1) or any other consumer of np.random invalidates the speculation
    algo._speculate_permutations(n, states1[-1])
    np.random.set_state(states1[0])
    assert algo._take_speculated_permutations(n) == (None, None)
    algo._speculate_permutations(n, states1[-1])
    np.random.set_state(states1[-1])
    np.random.random()
    assert algo._take_speculated_permutations(n) == (None, None)
    algo._speculate_permutations(n, states1[-1])
    np.random.set_state(states1[-1])
    assert algo._take_speculated_permutations(n + 1) == (None, None)      # another buffer size
