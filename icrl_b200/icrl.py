"""`python run_me.py icrl ...` -- the ICRL outer loop with the reference's flags (icrl/icrl.py:45-312 loop,
316-417 flags).  Every learner-side call lands on the CUDA path: `nominal_agent.learn` -> K1 relabel through
VecCostWrapper, K3, K4; `constraint_net.train` -> K2.  Plotting, video and W&B are out of scope (the run directory
and the metric table replace them)."""
import argparse
import importlib
import os
import pickle
import sys
import time

import numpy as np

from icrl_b200 import logger, utils
from icrl_b200.constraint_net import ConstraintNet
from icrl_b200.ppo_lag import PPOLagrangian
from icrl_b200.spaces import is_discrete as _is_discrete
from icrl_b200.true_constraint_net import get_true_cost_function, null_cost
from icrl_b200.vec_env import VecNormalize, sync_envs_normalization


def load_expert_data(expert_path, num_rollouts):
    """files/EXPERT/rollouts/<i>.pkl, each a dict with observations / actions / rewards (icrl/icrl.py:26-43)."""
    expert_mean_reward, obs, acs = [], [], []
    for i in range(num_rollouts):
        with open(os.path.join(expert_path, "files/EXPERT/rollouts", "%s.pkl" % str(i)), "rb") as f:
            data = pickle.load(f)
        obs.append(data['observations'])
        acs.append(data['actions'])
        expert_mean_reward.append(data['rewards'])
    return (np.concatenate(obs, axis=0), np.concatenate(acs, axis=0)), np.mean(expert_mean_reward)


def icrl(config):
    train_env = utils.make_train_env(env_id=config.train_env_id, save_dir=config.save_dir, use_cost_wrapper=True,
                                     base_seed=config.seed, num_threads=config.num_threads,
                                     normalize_obs=not config.dont_normalize_obs,
                                     normalize_reward=not config.dont_normalize_reward,
                                     normalize_cost=not config.dont_normalize_cost,
                                     cost_info_str=config.cost_info_str, reward_gamma=config.reward_gamma,
                                     cost_gamma=config.cost_gamma)
    sampling_env = utils.make_eval_env(env_id=config.train_env_id, use_cost_wrapper=False,
                                       normalize_obs=not config.dont_normalize_obs)
    eval_env = utils.make_eval_env(env_id=config.eval_env_id, use_cost_wrapper=False,
                                   normalize_obs=not config.dont_normalize_obs)

    is_discrete = _is_discrete(train_env.action_space)
    obs_dim = train_env.observation_space.shape[0]
    acs_dim = train_env.action_space.n if is_discrete else train_env.action_space.shape[0]
    action_low = action_high = None
    if not is_discrete:
        action_low, action_high = sampling_env.action_space.low, sampling_env.action_space.high

    (expert_obs, expert_acs), expert_mean_reward = load_expert_data(config.expert_path, config.expert_rollouts)
    expert_agent = PPOLagrangian.load(os.path.join(config.expert_path, "files/best_model.zip"), device=config.device)

    icrl_logger = logger.HumanOutputFormat(sys.stdout)

    cn_lr_schedule = lambda x: (config.anneal_clr_by_factor ** (config.n_iters * (1 - x))) * config.cn_learning_rate
    constraint_net = ConstraintNet(
        obs_dim, acs_dim, config.cn_layers, config.cn_batch_size, cn_lr_schedule, expert_obs, expert_acs, is_discrete,
        config.cn_reg_coeff, config.cn_obs_select_dim, config.cn_acs_select_dim,
        no_importance_sampling=config.no_importance_sampling,
        per_step_importance_sampling=config.per_step_importance_sampling, clip_obs=config.clip_obs,
        initial_obs_mean=None if not config.cn_normalize else np.zeros(obs_dim),
        initial_obs_var=None if not config.cn_normalize else np.ones(obs_dim),
        action_low=action_low, action_high=action_high, target_kl_old_new=config.cn_target_kl_old_new,
        target_kl_new_old=config.cn_target_kl_new_old, train_gail_lambda=config.train_gail_lambda, eps=config.cn_eps,
        device=config.device)
    # ICRL_WHOLE_BUFFER_RELABEL=1: relabel + cost-normalise each rollout on the device after collection instead of
    # calling the cost function at every environment step (same numbers, T fewer launches per rollout)
    whole_buffer = os.environ.get("ICRL_WHOLE_BUFFER_RELABEL", "0") == "1"
    train_env.set_cost_function(None if whole_buffer else constraint_net.cost_function)
    true_cost_function = get_true_cost_function(config.eval_env_id)

    create_nominal_agent = lambda: PPOLagrangian(
        policy=config.policy_name, env=train_env, learning_rate=config.learning_rate, n_steps=config.n_steps,
        batch_size=config.batch_size, n_epochs=config.n_epochs, reward_gamma=config.reward_gamma,
        reward_gae_lambda=config.reward_gae_lambda, cost_gamma=config.cost_gamma,
        cost_gae_lambda=config.cost_gae_lambda, clip_range=config.clip_range,
        clip_range_reward_vf=config.clip_range_reward_vf, clip_range_cost_vf=config.clip_range_cost_vf,
        ent_coef=config.ent_coef, reward_vf_coef=config.reward_vf_coef, cost_vf_coef=config.cost_vf_coef,
        max_grad_norm=config.max_grad_norm, use_sde=config.use_sde, sde_sample_freq=config.sde_sample_freq,
        target_kl=config.target_kl, penalty_initial_value=config.penalty_initial_value,
        penalty_learning_rate=config.penalty_learning_rate, budget=config.budget, seed=config.seed,
        device=config.device, verbose=0,
        pid_kwargs=dict(alpha=config.budget, penalty_init=config.penalty_initial_value,
                        Kp=config.proportional_control_coeff, Ki=config.integral_control_coeff,
                        Kd=config.derivative_control_coeff, pid_delay=config.pid_delay,
                        delta_p_ema_alpha=config.proportional_cost_ema_alpha,
                        delta_d_ema_alpha=config.derivative_cost_ema_alpha),
        policy_kwargs=dict(net_arch=utils.get_net_arch(config)))
    nominal_agent = create_nominal_agent()

    if config.use_curiosity_driven_exploration:
        raise NotImplementedError("curiosity-driven exploration (icrl/exploration.py) is outside the hot path")

    timesteps = 0.
    if config.warmup_timesteps is not None:
        print(utils.colorize("\nWarming up", color="green", bold=True))
        nominal_agent.learn(total_timesteps=config.warmup_timesteps, cost_function=null_cost)
        timesteps += nominal_agent.num_timesteps

    start_time = time.time()
    print(utils.colorize("\nBeginning training", color="green", bold=True), flush=True)
    best_true_reward, best_true_cost, best_forward_kl, best_reverse_kl = -np.inf, np.inf, np.inf, np.inf
    metrics = {}
    for itr in range(config.n_iters):
        if config.reset_policy and itr != 0:
            print(utils.colorize("Resetting agent", color="green", bold=True), flush=True)
            nominal_agent = create_nominal_agent()
        current_progress_remaining = 1 - float(itr) / float(config.n_iters)

        # forward step: PPO-Lagrangian on the current constraint (K1 per env step, K3 + K4 per rollout)
        nominal_agent.learn(total_timesteps=config.forward_timesteps,
                            cost_function=constraint_net if whole_buffer else "cost")
        forward_metrics = dict(logger.Logger.CURRENT.name_to_value)
        timesteps += nominal_agent.num_timesteps

        sync_envs_normalization(train_env, sampling_env)
        orig_observations, observations, actions, rewards, lengths = utils.sample_from_agent(
            nominal_agent, sampling_env, config.expert_rollouts)

        # backward step: constraint-net update (K2)
        mean, var = None, None
        if config.cn_normalize:
            mean, var = sampling_env.obs_rms.mean, sampling_env.obs_rms.var
        backward_metrics = constraint_net.train(config.backward_iters, orig_observations, actions, lengths, mean, var,
                                                current_progress_remaining)
        train_env.set_cost_function(None if whole_buffer else constraint_net.cost_function)

        average_true_cost = np.mean(true_cost_function(orig_observations, actions))
        samples_behind = np.mean(orig_observations[..., 0] < -3)
        samples_infront = np.mean(orig_observations[..., 0] > 3)
        sync_envs_normalization(train_env, eval_env)
        average_true_reward, std_true_reward = utils.evaluate_policy(nominal_agent, eval_env, n_eval_episodes=10,
                                                                     deterministic=False)
        forward_kl = utils.compute_kl(nominal_agent, expert_obs, expert_acs, expert_agent)
        reverse_kl = utils.compute_kl(expert_agent, orig_observations, actions, nominal_agent)

        if itr % config.save_every == 0:
            path = os.path.join(config.save_dir, f"models/icrl_{itr}_itrs")
            utils.del_and_make(path)
            nominal_agent.save(os.path.join(path, "nominal_agent"))
            constraint_net.save(os.path.join(path, "cn.pt"))
            if isinstance(train_env, VecNormalize):
                train_env.save(os.path.join(path, f"{itr}_train_env_stats.pkl"))
        if average_true_reward > best_true_reward:
            print(utils.colorize("Saving new best model", color="green", bold=True), flush=True)
            nominal_agent.save(os.path.join(config.save_dir, "best_nominal_model"))
            constraint_net.save(os.path.join(config.save_dir, "best_cn_model.pt"))
            if isinstance(train_env, VecNormalize):
                train_env.save(os.path.join(config.save_dir, "train_env_stats.pkl"))

        best_true_reward = max(best_true_reward, average_true_reward)
        best_true_cost = min(best_true_cost, average_true_cost)
        best_forward_kl = min(best_forward_kl, forward_kl)
        best_reverse_kl = min(best_reverse_kl, reverse_kl)

        metrics = {
            "time(m)": (time.time() - start_time) / 60, "iteration": itr, "timesteps": timesteps,
            "true/reward": average_true_reward, "true/reward_std": std_true_reward, "true/cost": average_true_cost,
            "true/samples_infront": samples_infront, "true/samples_behind": samples_behind,
            "true/forward_kl": forward_kl, "true/reverse_kl": reverse_kl,
            "best_true/best_reward": best_true_reward, "best_true/best_cost": best_true_cost,
            "best_true/best_forward_kl": best_forward_kl, "best_true/best_reverse_kl": best_reverse_kl,
        }
        metrics.update({k.replace("train/", "forward/"): v for k, v in forward_metrics.items()})
        metrics.update(backward_metrics)
        if config.verbose > 0:
            icrl_logger.write(metrics, {k: None for k in metrics.keys()}, step=itr)
        with open(os.path.join(config.save_dir, "metrics.jsonl"), "a") as f:
            import json
            f.write(json.dumps({k: (v.item() if hasattr(v, "item") else v) for k, v in metrics.items()},
                               default=float) + "\n")
    return metrics


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("file_to_run", type=str)
    # setup
    parser.add_argument("--config_file", "-cf", type=str, default=None)
    parser.add_argument("--project", "-p", type=str, default="ABC")
    parser.add_argument("--name", "-n", type=str, default=None)
    parser.add_argument("--group", "-g", type=str, default=None)
    parser.add_argument("--device", "-d", type=str, default="cpu")
    parser.add_argument("--verbose", "-v", type=int, default=2)
    parser.add_argument("--sync_wandb", "-sw", action="store_true")
    parser.add_argument("--wandb_sweep", "-ws", type=bool, default=False)
    # environments
    parser.add_argument("--train_env_id", "-tei", type=str, default="HalfCheetah-v3")
    parser.add_argument("--eval_env_id", "-eei", type=str, default="HalfCheetah-v3")
    parser.add_argument("--dont_normalize_obs", "-dno", action="store_true")
    parser.add_argument("--dont_normalize_reward", "-dnr", action="store_true")
    parser.add_argument("--dont_normalize_cost", "-dnc", action="store_true")
    parser.add_argument("--seed", "-s", type=int, default=None)
    parser.add_argument("--clip_obs", "-co", type=int, default=20)
    parser.add_argument("--cost_info_str", "-cis", type=str, default="cost")
    # networks
    parser.add_argument("--policy_name", "-pn", type=str, default="TwoCriticsMlpPolicy")
    parser.add_argument("--shared_layers", "-sl", type=int, default=None, nargs='*')
    parser.add_argument("--policy_layers", "-pl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--reward_vf_layers", "-rvl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--cost_vf_layers", "-cvl", type=int, default=[64, 64], nargs='*')
    # training
    parser.add_argument("--n_steps", "-ns", type=int, default=2048)
    parser.add_argument("--batch_size", "-bs", type=int, default=64)
    parser.add_argument("--n_epochs", "-ne", type=int, default=10)
    parser.add_argument("--num_threads", "-nt", type=int, default=5)
    parser.add_argument("--save_every", "-se", type=float, default=1)
    parser.add_argument("--eval_every", "-ee", type=float, default=2048)
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
                        help="Sets Learning Rate of Dual Variables if use_pid is not true.")
    # exploration
    parser.add_argument("--use_sde", "-us", action="store_true")
    parser.add_argument("--use_curiosity_driven_exploration", "-ucde", action="store_true")
    parser.add_argument("--sde_sample_freq", "-ssf", type=int, default=-1)
    # ICRL
    parser.add_argument('--train_gail_lambda', '-tgl', action='store_true')
    parser.add_argument("--n_iters", "-ni", type=int, default=100)
    parser.add_argument("--warmup_timesteps", "-wt", type=lambda x: int(float(x)), default=None)
    parser.add_argument("--forward_timesteps", "-ft", type=lambda x: int(float(x)), default=1e6)
    parser.add_argument("--backward_iters", "-bi", type=int, default=10)
    parser.add_argument('--no_importance_sampling', '-nis', action='store_true')
    parser.add_argument('--per_step_importance_sampling', '-psis', action='store_true')
    parser.add_argument('--reset_policy', '-rp', action='store_true')
    # constraint net
    parser.add_argument("--cn_layers", "-cl", type=int, default=[64, 64], nargs='*')
    parser.add_argument("--anneal_clr_by_factor", "-aclr", type=float, default=1.0)
    parser.add_argument("--cn_learning_rate", "-clr", type=float, default=3e-4)
    parser.add_argument("--cn_reg_coeff", "-crc", type=float, default=0)
    parser.add_argument("--cn_batch_size", "-cbs", type=int, default=None)
    parser.add_argument('--cn_obs_select_dim', '-cosd', type=int, default=None, nargs='+')
    parser.add_argument('--cn_acs_select_dim', '-casd', type=int, default=None, nargs='+')
    parser.add_argument('--cn_plot_every', '-cpe', type=int, default=1)
    parser.add_argument('--cn_normalize', '-cn', action='store_true')
    parser.add_argument("--cn_target_kl_old_new", "-ctkon", type=float, default=10)
    parser.add_argument("--cn_target_kl_new_old", "-ctkno", type=float, default=10)
    parser.add_argument("--cn_eps", "-ce", type=float, default=1e-5)
    # expert data
    parser.add_argument('--expert_path', '-ep', type=str, default='icrl/expert_data/HCWithPos-vm0')
    parser.add_argument('--expert_rollouts', '-er', type=int, default=20)
    return parser


def resolve_config(parser, argv):
    """Merge config file and command line (icrl/icrl.py:419-447), pick a seed, name the run, make its directory."""
    args = vars(parser.parse_args(argv))
    default_config, mod_name = {}, ''
    if args["config_file"] is not None:
        if args["config_file"].endswith(".py"):
            mod_name = args["config_file"].replace('/', '.')[:-3]
            default_config = importlib.import_module(mod_name).config
        elif args["config_file"].endswith(".json"):
            default_config = utils.load_dict_from_json(args["config_file"])
        else:
            raise ValueError("Invalid type of config file")
    config = utils.merge_configs(default_config, parser, argv)
    if config["seed"] is None:
        config["seed"] = np.random.randint(0, 100)
    config["name"] = utils.get_name(parser, default_config, config, mod_name)
    config = utils.Config(config)
    config.save_dir = utils.make_save_dir(config)
    print(utils.colorize("Configured folder %s for saving" % config.save_dir, color="green", bold=True))
    print(utils.colorize("Name: %s" % config.name, color="green", bold=True))
    utils.save_dict_as_json(config.as_dict(), config.save_dir, "config")
    return config


def main(argv=None):
    start = time.time()
    config = resolve_config(build_parser(), sys.argv[1:] if argv is None else argv)
    icrl(config)
    print(utils.colorize("Time taken: %05.2f hours" % ((time.time() - start) / 3600), color="green", bold=True))


if __name__ == '__main__':
    main()
