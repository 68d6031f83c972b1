"""The command-line surface of the drivers as data.  One row per option: long name -> (short name, kind, default); the
kinds spell the argparse behaviour (`flag` = store_true, `ints*` / `ints+` = nargs lists, `sci` = an int written like 2e5,
`opt_str` = the string "none" means None).  `icrl/icrl.py:316-417`, `icrl/cpg.py:216-298` and `icrl/run_policy.py:102-118`
of the reference define the same options one add_argument call at a time; tests/test_cli.py checks every name, short name,
type, default and nargs against fixtures read from those files."""
import argparse

_KINDS = {
    "str": dict(type=str), "int": dict(type=int), "float": dict(type=float), "bool": dict(type=bool),
    "flag": dict(action="store_true"),
    "ints*": dict(type=int, nargs='*'), "ints+": dict(type=int, nargs='+'),
    "sci": dict(type=lambda x: int(float(x))),
    "opt_str": dict(type=lambda x: None if str(x).lower() == "none" else str(x)),
}

# options shared by `icrl` and `cpg` (same names and defaults in both reference drivers)
COMMON = {
    "config_file": ("cf", "str", None), "project": ("p", "str", "ABC"), "name": ("n", "str", None),
    "group": ("g", "str", None), "device": ("d", "str", "cpu"), "verbose": ("v", "int", 2),
    "sync_wandb": ("sw", "flag", None), "wandb_sweep": ("ws", "bool", False),
    "train_env_id": ("tei", "str", "HalfCheetah-v3"), "eval_env_id": ("eei", "str", "HalfCheetah-v3"),
    "dont_normalize_obs": ("dno", "flag", None), "dont_normalize_reward": ("dnr", "flag", None),
    "dont_normalize_cost": ("dnc", "flag", None), "seed": ("s", "int", None),
    "policy_name": ("pn", "str", "TwoCriticsMlpPolicy"), "shared_layers": ("sl", "ints*", None),
    "policy_layers": ("pl", "ints*", [64, 64]),
    "n_steps": ("ns", "int", 2048), "batch_size": ("bs", "int", 64), "n_epochs": ("ne", "int", 10),
    "num_threads": ("nt", "int", 5), "eval_every": ("ee", "float", 2048),
    "reward_gamma": ("rg", "float", 0.99), "reward_gae_lambda": ("rgl", "float", 0.95),
    "cost_gamma": ("cg", "float", 0.99), "cost_gae_lambda": ("cgl", "float", 0.95),
    "clip_range": ("cr", "float", 0.2), "clip_range_reward_vf": ("crv", "float", None),
    "clip_range_cost_vf": ("ccv", "float", None), "ent_coef": ("ec", "float", 0.),
    "reward_vf_coef": ("rvc", "float", 0.5), "cost_vf_coef": ("cvc", "float", 0.5), "target_kl": ("tk", "float", None),
    "max_grad_norm": ("mgn", "float", 0.5), "learning_rate": ("lr", "float", 3e-4),
    "use_pid": ("upid", "flag", None), "penalty_initial_value": ("piv", "float", 1), "budget": ("b", "float", 0.0),
    "update_penalty_after": ("upa", "int", 1), "proportional_control_coeff": ("kp", "float", 10),
    "derivative_control_coeff": ("kd", "float", 0), "integral_control_coeff": ("ki", "float", 0.0001),
    "proportional_cost_ema_alpha": ("pema", "float", 0.5), "derivative_cost_ema_alpha": ("dema", "float", 0.5),
    "pid_delay": ("pidd", "int", 1), "penalty_learning_rate": ("plr", "float", 0.1),
    "use_sde": ("us", "flag", None), "use_curiosity_driven_exploration": ("ucde", "flag", None),
    "sde_sample_freq": ("ssf", "int", -1),
    "cn_obs_select_dim": ("cosd", "ints+", None), "cn_acs_select_dim": ("casd", "ints+", None),
}

ICRL_ONLY = {
    "clip_obs": ("co", "int", 20), "cost_info_str": ("cis", "str", "cost"),
    "reward_vf_layers": ("rvl", "ints*", [64, 64]), "cost_vf_layers": ("cvl", "ints*", [64, 64]),
    "save_every": ("se", "float", 1),
    "train_gail_lambda": ("tgl", "flag", None), "n_iters": ("ni", "int", 100), "warmup_timesteps": ("wt", "sci", None),
    "forward_timesteps": ("ft", "sci", 1e6), "backward_iters": ("bi", "int", 10),
    "no_importance_sampling": ("nis", "flag", None), "per_step_importance_sampling": ("psis", "flag", None),
    "reset_policy": ("rp", "flag", None),
    "cn_layers": ("cl", "ints*", [64, 64]), "anneal_clr_by_factor": ("aclr", "float", 1.0),
    "cn_learning_rate": ("clr", "float", 3e-4), "cn_reg_coeff": ("crc", "float", 0), "cn_batch_size": ("cbs", "int", None),
    "cn_plot_every": ("cpe", "int", 1), "cn_normalize": ("cn", "flag", None),
    "cn_target_kl_old_new": ("ctkon", "float", 10), "cn_target_kl_new_old": ("ctkno", "float", 10),
    "cn_eps": ("ce", "float", 1e-5),
    "expert_path": ("ep", "str", "icrl/expert_data/HCWithPos-vm0"), "expert_rollouts": ("er", "int", 20),
}

CPG_ONLY = {
    "message": ("m", "str", None), "cost_info_str": ("cis", "opt_str", "cost"),
    "reward_vf_layers": ("rl", "ints*", [64, 64]), "cost_vf_layers": ("cl", "ints*", [64, 64]),
    "cnn_features_dim": ("cfd", "int", 512), "timesteps": ("t", "sci", 1e6), "save_every": ("se", "float", 5e5),
    "plot_every": ("pe", "float", 2048), "use_lambda_shaping": ("uls", "flag", None),
    "use_null_cost": ("unc", "flag", None), "cn_path": ("cp", "str", None), "cn_device": ("cd", "str", None),
    "load_gail": ("lg", "flag", None),
}

RUN_POLICY = {
    "load_dir": ("l", "str", "icrl/wandb/latest-run/"), "is_icrl": ("ii", "flag", None), "remote": ("r", "flag", None),
    "save_dir": ("s", "str", "run_policy"), "env_id": ("e", "str", None), "load_itr": ("li", "int", None),
    "n_rollouts": ("nr", "int", 3), "dont_make_video": ("dmv", "flag", None), "dont_save_trajs": ("dst", "flag", None),
    "save_using_airl_scheme": ("suas", "flag", None), "reward_threshold": ("rt", "float", None),
    "length_threshold": ("lt", "int", None),
}


def make_parser(*tables) -> argparse.ArgumentParser:
    """`python run_me.py <file_to_run> [options]`: the first positional is consumed by the dispatcher."""
    parser = argparse.ArgumentParser()
    parser.add_argument("file_to_run", type=str)
    for table in tables:
        for long_name, (short, kind, default) in table.items():
            kw = dict(_KINDS[kind])
            if kind != "flag":
                kw["default"] = default
            parser.add_argument("--" + long_name, "-" + short, **kw)
    return parser
