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
        # NB: keyword call, so the reference's positional-shift quirk in ConstraintNet.load is not triggered here
        constraint_net = ConstraintNet.load(config.cn_path, obs_dim=obs_dim, acs_dim=acs_dim, is_discrete=is_discrete,
                                            obs_select_dim=config.cn_obs_select_dim,
                                            acs_select_dim=config.cn_acs_select_dim, clip_obs=None, obs_mean=None,
                                            obs_var=None, action_low=action_low, action_high=action_high,
                                            device=config.cn_device or "auto")
        cost_function = constraint_net.cost_function
    train_env.set_cost_function(cost_function)
    eval_env.set_cost_function(cost_function)

    model = PPOLagrangian(
        policy=config.policy_name, env=train_env, algo_type='pidlagrangian' if config.use_pid else 'lagrangian',
        learning_rate=config.learning_rate, n_steps=config.n_steps, batch_size=config.batch_size,
        n_epochs=config.n_epochs, reward_gamma=config.reward_gamma, reward_gae_lambda=config.reward_gae_lambda,
        cost_gamma=config.cost_gamma, cost_gae_lambda=config.cost_gae_lambda, clip_range=config.clip_range,
        clip_range_reward_vf=config.clip_range_reward_vf, clip_range_cost_vf=config.clip_range_cost_vf,
        ent_coef=config.ent_coef, reward_vf_coef=config.reward_vf_coef, cost_vf_coef=config.cost_vf_coef,
        max_grad_norm=config.max_grad_norm, use_sde=config.use_sde, sde_sample_freq=config.sde_sample_freq,
        target_kl=config.target_kl, penalty_initial_value=config.penalty_initial_value,
        penalty_learning_rate=config.penalty_learning_rate, update_penalty_after=config.update_penalty_after,
        budget=config.budget, seed=config.seed, device=config.device, verbose=config.verbose,
        pid_kwargs=dict(alpha=config.budget, penalty_init=config.penalty_initial_value,
                        Kp=config.proportional_control_coeff, Ki=config.integral_control_coeff,
                        Kd=config.derivative_control_coeff, pid_delay=config.pid_delay,
                        delta_p_ema_alpha=config.proportional_cost_ema_alpha,
                        delta_d_ema_alpha=config.derivative_cost_ema_alpha),
        policy_kwargs=dict(net_arch=utils.get_net_arch(config)))

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
    parser = argparse.ArgumentParser()
    parser.add_argument("file_to_run", type=str)
    # setup
    parser.add_argument("--config_file", "-cf", type=str, default=None)
    parser.add_argument("--project", "-p", type=str, default="ABC")
    parser.add_argument("--name", "-n", type=str, default=None)
    parser.add_argument("--group", "-g", type=str, default=None)
    parser.add_argument("--message", "-m", type=str, default=None)
    parser.add_argument("--device", "-d", type=str, default="cpu")
    parser.add_argument("--verbose", "-v", type=int, default=2)
    parser.add_argument("--wandb_sweep", "-ws", type=bool, default=False)
    parser.add_argument("--sync_wandb", "-sw", action="store_true")
    parser.add_argument("--cost_info_str", "-cis", type=lambda x: None if str(x).lower() == "none" else str(x),
                        default="cost")
    # environment
    parser.add_argument("--train_env_id", "-tei", type=str, default="HalfCheetah-v3")
    parser.add_argument("--eval_env_id", "-eei", type=str, default="HalfCheetah-v3")
    parser.add_argument("--dont_normalize_obs", "-dno", action="store_true")
    parser.add_argument("--dont_normalize_reward", "-dnr", action="store_true")
    parser.add_argument("--dont_normalize_cost", "-dnc", action="store_true")
    parser.add_argument("--seed", "-s", type=int, default=None)
    # networks
    parser.add_argument("--policy_name", "-pn", type=str, default="TwoCriticsMlpPolicy")
    parser.add_argument("--shared_layers", "-sl", type=int, default=None, nargs='*')
    parser.add_argument("--policy_layers", "-pl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--reward_vf_layers", "-rl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--cost_vf_layers", "-cl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--cnn_features_dim", "-cfd", type=int, default=512)
    # training
    parser.add_argument("--timesteps", "-t", type=lambda x: int(float(x)), default=1e6)
    parser.add_argument("--n_steps", "-ns", type=int, default=2048)
    parser.add_argument("--batch_size", "-bs", type=int, default=64)
    parser.add_argument("--n_epochs", "-ne", type=int, default=10)
    parser.add_argument("--num_threads", "-nt", type=int, default=5)
    parser.add_argument("--save_every", "-se", type=float, default=5e5)
    parser.add_argument("--eval_every", "-ee", type=float, default=2048)
    parser.add_argument("--plot_every", "-pe", type=float, default=2048)
    # MDP
    parser.add_argument("--reward_gamma", "-rg", type=float, default=0.99)
    parser.add_argument("--reward_gae_lambda", "-rgl", type=float, default=0.95)
    parser.add_argument("--cost_gamma", "-cg", type=float, default=0.99)
    parser.add_argument("--cost_gae_lambda", "-cgl", type=float, default=0.95)
    # losses
    parser.add_argument("--clip_range", "-cr", type=float, default=0.2)
    parser.add_argument("--clip_range_reward_vf", "-crv", type=float, default=None)
    parser.add_argument("--clip_range_cost_vf", "-ccv", type=float, default=None)
    parser.add_argument("--ent_coef", "-ec", type=float, default=0.)
    parser.add_argument("--reward_vf_coef", "-rvc", type=float, default=0.5)
    parser.add_argument("--cost_vf_coef", "-cvc", type=float, default=0.5)
    parser.add_argument("--target_kl", "-tk", type=float, default=None)
    parser.add_argument("--max_grad_norm", "-mgn", type=float, default=0.5)
    parser.add_argument("--learning_rate", "-lr", type=float, default=3e-4)
    # Lagrangian
    parser.add_argument("--use_pid", "-upid", action="store_true")
    parser.add_argument("--penalty_initial_value", "-piv", type=float, default=1)
    parser.add_argument("--budget", "-b", type=float, default=0.0)
    parser.add_argument("--update_penalty_after", "-upa", type=int, default=1)
    parser.add_argument("--proportional_control_coeff", "-kp", type=float, default=10)
    parser.add_argument("--derivative_control_coeff", "-kd", type=float, default=0)
    parser.add_argument("--integral_control_coeff", "-ki", type=float, default=0.0001)
    parser.add_argument("--proportional_cost_ema_alpha", "-pema", type=float, default=0.5)
    parser.add_argument("--derivative_cost_ema_alpha", "-dema", type=float, default=0.5)
    parser.add_argument("--pid_delay", "-pidd", type=int, default=1)
    parser.add_argument("--penalty_learning_rate", "-plr", type=float, default=0.1,
                        help="Sets Learning Rate of Dual Variables if not using PID Lagrangian.")
    # exploration
    parser.add_argument("--use_sde", "-us", action="store_true")
    parser.add_argument("--use_curiosity_driven_exploration", "-ucde", action="store_true")
    parser.add_argument("--use_lambda_shaping", "-uls", action="store_true")
    parser.add_argument("--sde_sample_freq", "-ssf", type=int, default=-1)
    # constraint net
    parser.add_argument("--use_null_cost", "-unc", action="store_true")
    parser.add_argument("--cn_path", "-cp", type=str, default=None)
    parser.add_argument('--cn_obs_select_dim', '-cosd', type=int, default=None, nargs='+')
    parser.add_argument('--cn_acs_select_dim', '-casd', type=int, default=None, nargs='+')
    parser.add_argument('--cn_device', '-cd', type=str, default=None)
    parser.add_argument("--load_gail", "-lg", action="store_true")
    return parser


def main(argv=None):
    start = time.time()
    config = resolve_config(build_parser(), sys.argv[1:] if argv is None else argv)
    cpg(config)
    print(utils.colorize("Time taken: %05.2f hours" % ((time.time() - start) / 3600), color="green", bold=True))


if __name__ == '__main__':
    main()
