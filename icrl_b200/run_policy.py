"""`python run_me.py run_policy -l <run dir> -s EXPERT -nr 20` -- roll a saved policy out and store the trajectories
in the layout `load_expert_data` reads (icrl/run_policy.py:21-123; video and W&B restore dropped)."""
import argparse
import os
import pickle
import sys

import numpy as np

from icrl_b200 import utils
from icrl_b200.ppo_lag import PPOLagrangian
from icrl_b200.vec_env import VecNormalize


def run_policy(args):
    if args.is_icrl:
        f = f"models/icrl_{args.load_itr}_itrs/nominal_agent" if args.load_itr is not None else "best_nominal_model"
    else:
        f = f"models/rl_model_{args.load_itr}_steps" if args.load_itr is not None else "best_model"
    load_dir = os.path.join(args.load_dir, "files")
    config = utils.Config(utils.load_dict_from_json(load_dir, "config"))
    save_dir = os.path.join(load_dir, args.save_dir)
    utils.del_and_make(save_dir)
    model = PPOLagrangian.load(os.path.join(load_dir, f))

    env = utils.make_eval_env(args.env_id or config.eval_env_id, use_cost_wrapper=False, normalize_obs=False)
    if not config.dont_normalize_obs:
        env = VecNormalize.load(os.path.join(load_dir, "train_env_stats.pkl"), env)
        env.norm_reward = False
        env.training = False

    if args.dont_save_trajs:
        return
    rollouts_dir = os.path.join(save_dir, "rollouts")
    utils.del_and_make(rollouts_dir)
    idx = 0
    while idx < args.n_rollouts:
        observations, _, actions, rewards, lengths = utils.sample_from_agent(model, env, 1)
        d = dict(observations=observations, actions=actions, rewards=rewards, lengths=lengths, save_scheme='not_airl')
        if ((args.reward_threshold is None or np.mean(rewards) >= args.reward_threshold) and
                (args.length_threshold is None or np.mean(lengths) >= args.length_threshold)):
            print(f"{idx}. Mean reward: {np.mean(rewards)} | Mean length: {np.mean(lengths)}")
            with open(os.path.join(rollouts_dir, f"{idx}.pkl"), "wb") as fh:
                pickle.dump(d, fh)
            idx += 1


def main(argv=None):
    from icrl_b200.cli import RUN_POLICY, make_parser
    args = make_parser(RUN_POLICY).parse_args(sys.argv[1:] if argv is None else argv)
    if args.remote or args.save_using_airl_scheme:
        raise NotImplementedError("W&B restore and the AIRL save scheme are outside the ICRL hot path")
    run_policy(args)
