"""`python run_me.py icrl ...` -- the ICRL outer loop with the reference's flags (icrl/icrl.py:45-312 loop,
316-417 flags).  Every learner-side call lands on the CUDA path: `nominal_agent.learn` -> K1 relabel through
VecCostWrapper, K3, K4; `constraint_net.train` -> K2.  Plotting, video and W&B are out of scope (the run directory
and the metric table replace them)."""
import argparse
import importlib
import json
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


def ppo_lagrangian_kwargs(config):
    """Constructor arguments both drivers hand to PPOLagrangian (icrl/icrl.py:139-173, icrl/cpg.py:117-154)."""
    kw = {k: config[k] for k in (
        "learning_rate", "n_steps", "batch_size", "n_epochs", "reward_gamma", "reward_gae_lambda", "cost_gamma",
        "cost_gae_lambda", "clip_range", "clip_range_reward_vf", "clip_range_cost_vf", "ent_coef", "reward_vf_coef",
        "cost_vf_coef", "max_grad_norm", "use_sde", "sde_sample_freq", "target_kl", "penalty_initial_value",
        "penalty_learning_rate", "budget", "seed", "device")}
    kw["policy"] = config.policy_name
    kw["pid_kwargs"] = dict(alpha=config.budget, penalty_init=config.penalty_initial_value,
                            Kp=config.proportional_control_coeff, Ki=config.integral_control_coeff,
                            Kd=config.derivative_control_coeff, pid_delay=config.pid_delay,
                            delta_p_ema_alpha=config.proportional_cost_ema_alpha,
                            delta_d_ema_alpha=config.derivative_cost_ema_alpha)
    kw["policy_kwargs"] = dict(net_arch=utils.get_net_arch(config))
    return kw


def _say(text):
    print(utils.colorize(text, color="green", bold=True), flush=True)


def icrl(config):
    """The outer loop of icrl/icrl.py:45-312: forward step (PPO-Lagrangian under the current constraint: K1 per env step or
    per rollout, K3, K4), nominal sampling, backward step (K2), evaluation, checkpoints, metrics."""
    norm = dict(normalize_obs=not config.dont_normalize_obs)
    train_env = utils.make_train_env(env_id=config.train_env_id, save_dir=config.save_dir, use_cost_wrapper=True,
                                     base_seed=config.seed, num_threads=config.num_threads,
                                     normalize_reward=not config.dont_normalize_reward,
                                     normalize_cost=not config.dont_normalize_cost, cost_info_str=config.cost_info_str,
                                     reward_gamma=config.reward_gamma, cost_gamma=config.cost_gamma, **norm)
    sampling_env = utils.make_eval_env(env_id=config.train_env_id, use_cost_wrapper=False, **norm)   # no cost needed
    eval_env = utils.make_eval_env(env_id=config.eval_env_id, use_cost_wrapper=False, **norm)

    discrete = _is_discrete(train_env.action_space)
    obs_dim = train_env.observation_space.shape[0]
    acs_dim = train_env.action_space.n if discrete else train_env.action_space.shape[0]
    bounds = (None, None) if discrete else (sampling_env.action_space.low, sampling_env.action_space.high)

    (expert_obs, expert_acs), _expert_reward = load_expert_data(config.expert_path, config.expert_rollouts)
    expert_agent = PPOLagrangian.load(os.path.join(config.expert_path, "files/best_model.zip"), device=config.device)
    table = logger.HumanOutputFormat(sys.stdout)

    def cn_lr(progress_remaining):        # annealed per ICRL iteration (icrl/icrl.py:89)
        return config.cn_learning_rate * config.anneal_clr_by_factor ** (config.n_iters * (1 - progress_remaining))

    unit_stats = (np.zeros(obs_dim), np.ones(obs_dim)) if config.cn_normalize else (None, None)
    constraint_net = ConstraintNet(
        obs_dim, acs_dim, config.cn_layers, config.cn_batch_size, cn_lr, expert_obs, expert_acs, discrete,
        config.cn_reg_coeff, config.cn_obs_select_dim, config.cn_acs_select_dim,
        no_importance_sampling=config.no_importance_sampling,
        per_step_importance_sampling=config.per_step_importance_sampling, clip_obs=config.clip_obs,
        initial_obs_mean=unit_stats[0], initial_obs_var=unit_stats[1], action_low=bounds[0], action_high=bounds[1],
        target_kl_old_new=config.cn_target_kl_old_new, target_kl_new_old=config.cn_target_kl_new_old,
        train_gail_lambda=config.train_gail_lambda, eps=config.cn_eps, device=config.device)

    # ICRL_WHOLE_BUFFER_RELABEL=1: relabel + cost-normalise each rollout on the device after collection instead of
    # calling the cost function at every environment step (T fewer launches per rollout).  The costs, cost advantages
    # and cost_rms / cost_ret statistics are bit-identical to the per-step path (tests/test_drivers_gpu.py); what
    # differs: callbacks see all-zero `costs` / `orig_costs` in update_locals() at every step (the buffer is filled
    # after collection), and info['cost'] is absent.  During warm-up (null_cost) the per-step wrapper stays plugged in,
    # as in the reference, so that cost_rms keeps seeing constraint-net costs.
    whole_buffer = os.environ.get("ICRL_WHOLE_BUFFER_RELABEL", "0") == "1"

    def plug_cost():
        train_env.set_cost_function(None if whole_buffer else constraint_net.cost_function)
    plug_cost()
    true_cost = get_true_cost_function(config.eval_env_id)

    def new_agent():
        return PPOLagrangian(env=train_env, verbose=0, **ppo_lagrangian_kwargs(config))
    agent = new_agent()
    if config.use_curiosity_driven_exploration:
        raise NotImplementedError("curiosity-driven exploration (icrl/exploration.py) is outside the hot path")

    total_steps = 0.
    if config.warmup_timesteps is not None:             # no cost during warm-up
        _say("\nWarming up")
        # the reference's wrapper keeps calling the constraint net during warm-up (vec_cost_wrapper.py:62) and
        # VecNormalizeWithCost keeps updating cost_rms from it: same here in both relabel modes
        train_env.set_cost_function(constraint_net.cost_function)
        agent.learn(total_timesteps=config.warmup_timesteps, cost_function=null_cost)
        plug_cost()
        total_steps += agent.num_timesteps

    def checkpoint(folder, agent_name, cn_name, stats_name):
        agent.save(os.path.join(folder, agent_name))
        constraint_net.save(os.path.join(folder, cn_name))
        if isinstance(train_env, VecNormalize):
            train_env.save(os.path.join(folder, stats_name))

    t_start = time.time()
    _say("\nBeginning training")
    best = {"reward": -np.inf, "cost": np.inf, "forward_kl": np.inf, "reverse_kl": np.inf}
    row = {}
    for it in range(config.n_iters):
        if it and config.reset_policy:
            _say("Resetting agent")
            agent = new_agent()
        progress_remaining = 1 - float(it) / float(config.n_iters)

        # ---- forward step
        agent.learn(total_timesteps=config.forward_timesteps, cost_function=constraint_net if whole_buffer else "cost")
        forward = dict(logger.Logger.CURRENT.name_to_value)
        total_steps += agent.num_timesteps

        # ---- nominal trajectories of the updated policy
        sync_envs_normalization(train_env, sampling_env)
        nom_obs, _nom_obs_normalised, nom_acs, _nom_rewards, nom_lengths = utils.sample_from_agent(
            agent, sampling_env, config.expert_rollouts)

        # ---- backward step
        stats = (sampling_env.obs_rms.mean, sampling_env.obs_rms.var) if config.cn_normalize else (None, None)
        backward = constraint_net.train(config.backward_iters, nom_obs, nom_acs, nom_lengths, stats[0], stats[1],
                                        progress_remaining)
        plug_cost()

        # ---- evaluation: true cost of the nominal samples, reward in the true environment, KLs to the expert
        seen = {"cost": np.mean(true_cost(nom_obs, nom_acs)),
                "behind": np.mean(nom_obs[..., 0] < -3), "infront": np.mean(nom_obs[..., 0] > 3)}
        sync_envs_normalization(train_env, eval_env)
        seen["reward"], seen["reward_std"] = utils.evaluate_policy(agent, eval_env, n_eval_episodes=10, deterministic=False)
        seen["forward_kl"] = utils.compute_kl(agent, expert_obs, expert_acs, expert_agent)
        seen["reverse_kl"] = utils.compute_kl(expert_agent, nom_obs, nom_acs, agent)

        # ---- checkpoints: periodic, and the best-reward model so far
        if it % config.save_every == 0:
            folder = os.path.join(config.save_dir, f"models/icrl_{it}_itrs")
            utils.del_and_make(folder)
            checkpoint(folder, "nominal_agent", "cn.pt", f"{it}_train_env_stats.pkl")
        if seen["reward"] > best["reward"]:
            _say("Saving new best model")
            checkpoint(config.save_dir, "best_nominal_model", "best_cn_model.pt", "train_env_stats.pkl")
        best["reward"] = max(best["reward"], seen["reward"])
        for key in ("cost", "forward_kl", "reverse_kl"):
            best[key] = min(best[key], seen[key])

        row = {"time(m)": (time.time() - t_start) / 60, "iteration": it, "timesteps": total_steps,
               "true/reward": seen["reward"], "true/reward_std": seen["reward_std"], "true/cost": seen["cost"],
               "true/samples_infront": seen["infront"], "true/samples_behind": seen["behind"],
               "true/forward_kl": seen["forward_kl"], "true/reverse_kl": seen["reverse_kl"]}
        row.update({f"best_true/best_{k}": v for k, v in best.items()})
        row.update({k.replace("train/", "forward/"): v for k, v in forward.items()})
        row.update(backward)
        if config.verbose > 0:
            table.write(row, {k: None for k in row}, step=it)
        with open(os.path.join(config.save_dir, "metrics.jsonl"), "a") as f:
            f.write(json.dumps({k: (v.item() if hasattr(v, "item") else v) for k, v in row.items()}, default=float) + "\n")
    return row


def build_parser():
    from icrl_b200.cli import COMMON, ICRL_ONLY, make_parser
    return make_parser(COMMON, ICRL_ONLY)


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
