"""`python run_me.py cpg ...` -- constrained policy gradient (PPO-Lagrangian against a fixed cost: the true one, a
saved constraint net, or none), with the reference's flags (icrl/cpg.py:24-212 body, 214-298 flags).  K1 runs per
env step through VecCostWrapper when a constraint net is loaded; K3 + K4 run per rollout."""
import argparse
import os
import sys
import time

from icrl_b200 import callbacks, utils
from icrl_b200.constraint_net import ConstraintNet
from icrl_b200.icrl import resolve_config
from icrl_b200.ppo_lag import PPOLagrangian
from icrl_b200.spaces import is_discrete as _is_discrete
from icrl_b200.true_constraint_net import get_true_cost_function, null_cost
from icrl_b200.vec_env import sync_envs_normalization


def cpg(config):
    train_env = utils.make_train_env(env_id=config.train_env_id, save_dir=config.save_dir, use_cost_wrapper=True,
                                     base_seed=config.seed, num_threads=config.num_threads,
                                     normalize_obs=not config.dont_normalize_obs,
                                     normalize_reward=not config.dont_normalize_reward,
                                     normalize_cost=not config.dont_normalize_cost,
                                     cost_info_str=config.cost_info_str, reward_gamma=config.reward_gamma,
                                     cost_gamma=config.cost_gamma)
    eval_env = utils.make_eval_env(env_id=config.eval_env_id, use_cost_wrapper=True,
                                   normalize_obs=not config.dont_normalize_obs)
    is_discrete = _is_discrete(train_env.action_space)
    obs_dim = train_env.observation_space.shape[0]
    acs_dim = train_env.action_space.n if is_discrete else train_env.action_space.shape[0]

    if config.use_null_cost:
        cost_function = null_cost
    elif config.cn_path is None:
        cost_function = get_true_cost_function(config.eval_env_id)
    elif config.load_gail:
        from icrl_b200.gail_utils import GailDiscriminator
        action_low = action_high = None
        if not is_discrete:
            action_low, action_high = train_env.action_space.low, train_env.action_space.high
        gail = GailDiscriminator.load(config.cn_path, obs_dim=obs_dim, acs_dim=acs_dim, is_discrete=is_discrete,
                                      obs_select_dim=config.cn_obs_select_dim, acs_select_dim=config.cn_acs_select_dim,
                                      clip_obs=None, obs_mean=None, obs_var=None, action_low=action_low,
                                      action_high=action_high, device=config.cn_device or "auto")

        def cost_function(obs, acs):
            return gail.reward_function(obs, acs, apply_log=False)
    else:
        action_low = action_high = None
        if not is_discrete:
            action_low, action_high = train_env.action_space.low, train_env.action_space.high
        # ConstraintNet.load reproduces the reference's positional-argument shift INSIDE load() (constraint_net.py:394-399):
        # the loaded net neither clips observations nor actions, whatever is passed here
        constraint_net = ConstraintNet.load(config.cn_path, obs_dim=obs_dim, acs_dim=acs_dim, is_discrete=is_discrete,
                                            obs_select_dim=config.cn_obs_select_dim,
                                            acs_select_dim=config.cn_acs_select_dim, clip_obs=None, obs_mean=None,
                                            obs_var=None, action_low=action_low, action_high=action_high,
                                            device=config.cn_device or "auto")
        cost_function = constraint_net.cost_function
    train_env.set_cost_function(cost_function)
    eval_env.set_cost_function(cost_function)

    from icrl_b200.icrl import ppo_lagrangian_kwargs
    model = PPOLagrangian(env=train_env, algo_type='pidlagrangian' if config.use_pid else 'lagrangian',
                          update_penalty_after=config.update_penalty_after, verbose=config.verbose,
                          **ppo_lagrangian_kwargs(config))

    save_periodically = callbacks.CheckpointCallback(config.save_every, os.path.join(config.save_dir, "models"),
                                                     verbose=0)
    save_env_stats = callbacks.SaveEnvStatsCallback(train_env, config.save_dir)
    save_best = callbacks.EvalCallback(eval_env, eval_freq=config.eval_every, best_model_save_path=config.save_dir,
                                       verbose=0, deterministic=False, callback_on_new_best=save_env_stats)
    adjusted_reward = callbacks.AdjustedRewardCallback(get_true_cost_function(config.eval_env_id))
    all_callbacks = [save_periodically, save_best, adjusted_reward]
    if config.use_curiosity_driven_exploration or config.use_lambda_shaping:
        raise NotImplementedError("exploration / lambda-shaping callbacks (icrl/exploration.py) are outside the hot path")
    if any(env in config.train_env_id for env in ['Ant', 'HalfCheetah', 'Point', 'Swimmer', 'Walker', 'HC']):
        all_callbacks.append(callbacks.LogTorqueCallback())

    cost_info_str = config.cost_info_str if config.cost_info_str is not None else cost_function
    model.learn(total_timesteps=int(config.timesteps), cost_function=cost_info_str, callback=all_callbacks)
    sync_envs_normalization(train_env, eval_env)
    return model


def build_parser():
    from icrl_b200.cli import COMMON, CPG_ONLY, make_parser
    return make_parser(COMMON, CPG_ONLY)


def main(argv=None):
    start = time.time()
    config = resolve_config(build_parser(), sys.argv[1:] if argv is None else argv)
    cpg(config)
    print(utils.colorize("Time taken: %05.2f hours" % ((time.time() - start) / 3600), color="green", bold=True))


if __name__ == '__main__':
    main()
