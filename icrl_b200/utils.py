"""Driver-side helpers mirroring the parts of the reference's icrl/utils.py that the ICRL / CPG loops call around the
hot path.  Same names and argument meaning; the W&B run object is replaced by a local run directory."""
import json
import math
import os
import shutil
import time
from typing import Optional

import numpy as np
import torch as th

from icrl_b200 import envs as _envs
from icrl_b200 import vec_env


# ---------------------------------------------------------------- misc (icrl/utils.py:30-75)
_COLORS = dict(gray=30, red=31, green=32, yellow=33, blue=34, magenta=35, cyan=36, white=37, crimson=38)


def colorize(string, color, bold=False, highlight=False):
    num = _COLORS[color] + (10 if highlight else 0)
    attr = [str(num)] + (['1'] if bold else [])
    return '\x1b[%sm%s\x1b[0m' % (';'.join(attr), string)


def del_and_make(d):
    if os.path.isdir(d):
        shutil.rmtree(d)
    os.makedirs(d)


def save_dict_as_json(dic, save_dir, name=None):
    path = os.path.join(save_dir, name + ".json") if name is not None else save_dir
    with open(path, 'w') as out:
        out.write(json.dumps(dic, separators=(',\n', '\t:\t'), sort_keys=True, default=str))


def load_dict_from_json(load_from, name=None):
    path = os.path.join(load_from, name + ".json") if name is not None else load_from
    with open(path, "rb") as f:
        return json.load(f)


class Config(dict):
    """Attribute-style view of the merged configuration (what ``wandb.config`` is to the reference drivers)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    def as_dict(self):
        return dict(self)


# ---------------------------------------------------------------- configuration merging (icrl/utils.py:139-222)
def get_sl_map(parser):
    """long option name -> short option name for every option of `parser`."""
    sl_map = {}
    for action in parser._actions:
        longs = [o[2:] for o in action.option_strings if o.startswith('--')]
        shorts = [o[1:] for o in action.option_strings if not o.startswith('--')]
        for name in longs:
            sl_map[name] = shorts[0] if shorts else name
        if not longs:
            for name in shorts:
                sl_map[name] = name
    return sl_map


def key_was_specified(key1, key2, sys_argv):
    return any(arg[0] == '-' and arg.strip('-') in (key1, key2) for arg in sys_argv if arg)


def merge_configs(config, parser, sys_argv):
    """Priority: command line > config file > parser default."""
    parser_dict = vars(parser.parse_args(sys_argv))
    sl_map = get_sl_map(parser)
    merged = {}
    for key in list(config.keys()) + list(parser_dict.keys()):
        if key in parser_dict:
            if key_was_specified(key, sl_map.get(key, key), sys_argv) or key not in config:
                merged[key] = parser_dict[key]
            else:
                merged[key] = config[key]
        else:
            merged[key] = config[key]
    return merged


def get_name(parser, default_config, actual_config, mod_name=''):
    """Run name: env ids, then `short_value` for every argument that differs from its default, then the seed."""
    name = actual_config.get("name")
    if name is None:
        sl_map = get_sl_map(parser)
        ignore = {"config_file", "train_env_id", "eval_env_id", "seed", "timesteps", "save_every", "eval_every",
                  "n_iters", "sync_wandb", "file_to_run", "project", "group", "name"}
        parts = []
        for key, value in actual_config.items():
            if key in ignore or key not in sl_map:
                continue
            default = default_config.get(key, parser.get_default(key))
            if value == default:
                continue
            if key == "expert_path" and value is not None:
                value = str(value).rstrip('/').split('/')[-1]
            if type(value) not in (bool, int) and hasattr(value, "__float__") and value != 0:
                value = round(value, 4 - int(math.floor(math.log10(abs(value)))) - 1)
            parts.append('%s_%s' % (sl_map[key], value))
        name = '_'.join([actual_config["train_env_id"], actual_config["eval_env_id"]] +
                        ([mod_name.split('.')[-1]] if mod_name else []) + parts)
    return '%s_s_%s' % (name, actual_config["seed"])


def make_save_dir(config) -> str:
    """The reference saves under the W&B run directory (./icrl/wandb/<run>/files); without W&B the run directory is
    ``$ICRL_SAVE_ROOT`` (default ./icrl/runs) / <time>-<name> / files."""
    root = os.environ.get("ICRL_SAVE_ROOT", os.path.join(".", "icrl", "runs"))
    name = str(config["name"]).replace(os.sep, "_").replace(" ", "")[:120]
    path = os.path.join(root, "%s-%s" % (time.strftime("%Y%m%d_%H%M%S"), name), "files")
    os.makedirs(path, exist_ok=True)
    return path


def get_net_arch(config):
    """icrl/utils.py:636-655."""
    separate_layers = dict(pi=config.policy_layers, vf=config.reward_vf_layers, cvf=config.cost_vf_layers)
    if config.shared_layers is not None:
        return [*config.shared_layers, separate_layers]
    return [separate_layers]


# ---------------------------------------------------------------- environments (icrl/utils.py:247-303)
def set_random_seed(seed: int) -> None:
    """stable_baselines3/common/utils.py:23-40: python, numpy and torch global generators."""
    import random
    random.seed(seed)
    np.random.seed(seed)
    th.manual_seed(seed)


def make_env(env_id, rank, log_dir, seed=0):
    def _init():
        env = _envs.make(env_id)
        if hasattr(env, "seed"):
            env.seed(seed + rank)
        return env
    # like the reference (icrl/utils.py:247-256) building the factory seeds the GLOBAL generators: this is what makes the
    # constraint net's initial weights (created right after the environments) a function of --seed
    set_random_seed(seed)
    return _init


def make_train_env(env_id, save_dir, use_cost_wrapper, base_seed=0, num_threads=1, normalize_obs=True,
                   normalize_reward=True, normalize_cost=True, **kwargs):
    env = vec_env.DummyVecEnv([make_env(env_id, i, save_dir, base_seed) for i in range(num_threads)])
    if use_cost_wrapper:
        env = vec_env.VecCostWrapper(env)
    if normalize_reward and normalize_cost:
        assert all(key in kwargs for key in ['cost_info_str', 'reward_gamma', 'cost_gamma'])
        return vec_env.VecNormalizeWithCost(env, training=True, norm_obs=normalize_obs, norm_reward=normalize_reward,
                                            norm_cost=normalize_cost, cost_info_str=kwargs['cost_info_str'],
                                            reward_gamma=kwargs['reward_gamma'], cost_gamma=kwargs['cost_gamma'])
    if normalize_reward:
        assert 'reward_gamma' in kwargs
        return vec_env.VecNormalizeWithCost(env, training=True, norm_obs=normalize_obs, norm_reward=normalize_reward,
                                            norm_cost=normalize_cost, reward_gamma=kwargs['reward_gamma'])
    return vec_env.VecNormalizeWithCost(env, training=True, norm_obs=normalize_obs, norm_reward=normalize_reward,
                                        norm_cost=normalize_cost)


def make_eval_env(env_id, use_cost_wrapper, normalize_obs=True):
    env = vec_env.DummyVecEnv([lambda: _envs.make(env_id)])
    if use_cost_wrapper:
        env = vec_env.VecCostWrapper(env)
    return vec_env.VecNormalizeWithCost(env, training=False, norm_obs=normalize_obs, norm_reward=False, norm_cost=False)


# ---------------------------------------------------------------- sampling / evaluation
def sample_from_agent(agent, env, rollouts):
    """icrl/utils.py:325-360: `rollouts` full episodes from a single-env VecEnv; returns (original observations,
    normalised observations, actions, episode rewards, episode lengths) -- the nominal data K2 trains on."""
    if isinstance(env, vec_env.VecEnv):
        assert env.num_envs == 1, "You must pass only one environment when using this function"
    orig_observations, observations, actions, rewards, lengths = [], [], [], [], []
    for i in range(rollouts):
        if not isinstance(env, vec_env.VecEnv) or i == 0:
            obs = env.reset()
        done, state = False, None
        episode_reward, episode_length = 0.0, 0
        while not done:
            action, state = agent.predict(obs, state=state, deterministic=False)
            obs, reward, done, _info = env.step(action)
            observations.append(obs)
            orig_observations.append(env.get_original_obs() if isinstance(env, vec_env.VecNormalize) else obs)
            actions.append(action)
            episode_reward += reward
            episode_length += 1
        rewards.append(episode_reward)
        lengths.append(episode_length)
    return (np.squeeze(np.array(orig_observations), axis=1), np.squeeze(np.array(observations), axis=1),
            np.squeeze(np.array(actions), axis=1), np.squeeze(np.array(rewards), axis=1), np.array(lengths))


def evaluate_policy(model, env, n_eval_episodes=10, deterministic=True, return_episode_rewards=False):
    """stable_baselines3/common/evaluation.py:10-67 (single-env VecEnv)."""
    if isinstance(env, vec_env.VecEnv):
        assert env.num_envs == 1, "You must pass only one environment when using this function"
    episode_rewards, episode_lengths = [], []
    for i in range(n_eval_episodes):
        if not isinstance(env, vec_env.VecEnv) or i == 0:
            obs = env.reset()
        done, state = False, None
        episode_reward, episode_length = 0.0, 0
        while not done:
            action, state = model.predict(obs, state=state, deterministic=deterministic)
            obs, reward, done, _info = env.step(action)
            episode_reward += reward
            episode_length += 1
        episode_rewards.append(episode_reward)
        episode_lengths.append(episode_length)
    if return_episode_rewards:
        return episode_rewards, episode_lengths
    return np.mean(episode_rewards), np.std(episode_rewards)


def compute_kl(agent_2, observations, actions, agent_1=None, index=1):
    """icrl/utils.py:420-438: Monte-Carlo KL(agent_1 || agent_2) on samples of agent_1 (uniform when agent_1 is None).

    Reference quirk kept on purpose: it reads element [1] of ``evaluate_actions``, which for the two-critics policy
    (values, cost_values, log_prob, entropy) is the COST VALUE, not the log-likelihood, so the logged
    ``true/forward_kl`` / ``true/reverse_kl`` are differences of cost values.  Pass ``index=2`` for the real KL."""
    observations = th.tensor(observations, dtype=th.float32)
    actions = th.tensor(actions, dtype=th.float32)
    log_prob = lambda agent: agent.policy.evaluate_actions(observations, actions)[index]
    kl = -log_prob(agent_2)
    if agent_1 is not None:
        kl = kl + log_prob(agent_1)
    return (kl.sum() / observations.shape[0]).item()
