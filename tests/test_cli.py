"""The `run_me.py icrl|cpg|run_policy` flag surface against the reference's (fixtures read from the reference source by
tests/golden/make_golden.py::golden_flags), plus the config-merging rules of icrl/utils.py:176-222."""
import json
import os

import pytest

from icrl_b200 import utils

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _parser(name):
    if name == "icrl":
        from icrl_b200.icrl import build_parser
        return build_parser()
    if name == "cpg":
        from icrl_b200.cpg import build_parser
        return build_parser()
    from icrl_b200.cli import RUN_POLICY, make_parser
    return make_parser(RUN_POLICY)


@pytest.mark.parametrize("name", ["icrl", "cpg", "run_policy"])
def test_flags_match_reference(name):
    ref = json.load(open(os.path.join(GOLD, f"flags_{name}.json")))
    parser = _parser(name)
    ours = {tuple(a.option_strings) or (a.dest,): a for a in parser._actions if a.dest != "help"}
    assert len(ours) == len(ref)
    for flag in ref:
        key = tuple(flag["opts"])
        assert key in ours, f"missing flag {key}"
        a = ours[key]
        if "default" in flag:
            assert a.default == flag["default"], key
        if flag.get("action") == "store_true":
            assert a.const is True and a.default is False and a.nargs == 0, key
        if "nargs" in flag:
            assert a.nargs == flag["nargs"], key
        if flag.get("type") in ("int", "float", "str", "bool"):
            assert a.type.__name__ == flag["type"], key


def test_merge_priority_cli_over_file_over_default():
    from icrl_b200.icrl import build_parser
    parser = build_parser()
    file_cfg = {"n_iters": 7, "cn_learning_rate": 0.5, "extra_key": "kept"}
    argv = ["icrl", "-clr", "0.05", "--batch_size", "128"]
    merged = utils.merge_configs(file_cfg, parser, argv)
    assert merged["cn_learning_rate"] == 0.05            # command line (short name) beats the file
    assert merged["n_iters"] == 7                        # file beats the parser default
    assert merged["batch_size"] == 128 and merged["n_epochs"] == 10
    assert merged["extra_key"] == "kept"
    merged["seed"] = 3
    name = utils.get_name(parser, file_cfg, merged)
    assert name.startswith("HalfCheetah-v3_HalfCheetah-v3") and name.endswith("_s_3")
    assert "clr_0.05" in name and "bs_128" in name and "ni_" not in name


README_ICRL = {
    # README.md:25, 38, 50 of the reference
    "lapgrid": "icrl -p ICRL-FE2 --group LapGrid-ICRL -er 20 -ep icrl/expert_data/LGW -tei LGW-v0 -eei CLGW-v0 -tk 0.01 "
               "-cl 20 -clr 0.003 -ft 0.5e5 -ni 10 -bi 20 -dno -dnr -dnc",
    "halfcheetah": "icrl -p ICRL-FE2 --group HC-ICRL -er 10 -ep icrl/expert_data/HCWithPos-New -tk 0.01 -cl 20 -bi 10 "
                   "-ft 2e5 -ni 30 -tei HCWithPos-v0 -eei HCWithPosTest-v0 -clr 0.05 -aclr 0.9 -crc 0.5 -psis -ctkno 2.5",
    "antwall": "icrl -p ICRL-FE2 --group AntWall-ICRL -ep icrl/expert_data/AntWall -er 45 -cl 40 40 -clr 0.005 -aclr 0.9 "
               "-crc 0.6 -bi 5 -ft 2e5 -ni 20 -tei AntWall-v0 -eei AntWallTest-v0 --batch_size 128 "
               "--reward_gae_lambda 0.9 --cost_gae_lambda 0.9 --n_epochs 20 --learning_rate 3e-5 --clip_range 0.4 "
               "-piv 0.1 -plr 0.05 -psis -tk 0.02 -ctkno 2.5",
}
README_CPG = ("cpg -p ICRL-FE2 --group Point-CT-ICRL --cn_path ./icrl/expert_data/ConstraintTransfer/ICRL/Point/files/"
              "best_cn_model.pt -cosd 0 1 -casd -1 -tei PointCircle-v0 -eei PointCircleTestBack-v0 -tk 0.01 -t 1.5e6 "
              "-plr 1.0")


def test_readme_commands_parse_to_the_bench_workloads():
    """The README command lines parse, and give the hyper-parameters bench.py's WORKLOADS table claims for them."""
    from icrl_b200.cpg import build_parser as cpg_parser
    from icrl_b200.icrl import build_parser
    from icrl_b200.learner import WORKLOADS
    for key, line in README_ICRL.items():
        c = utils.Config(vars(build_parser().parse_args(line.split())))
        w = WORKLOADS[key]
        assert tuple(c.cn_layers) == tuple(w.cn_hidden), key
        assert (c.batch_size, c.n_epochs, c.n_steps, c.num_threads) == (w.batch_size, w.n_epochs, w.n_steps, w.n_envs), key
        assert c.backward_iters == w.backward_iters and c.cn_learning_rate == w.cn_lr and c.cn_reg_coeff == w.cn_reg, key
        assert c.per_step_importance_sampling == w.per_step_is, key
        assert (c.learning_rate, c.clip_range, c.reward_gae_lambda, c.cost_gae_lambda) == \
               (w.learning_rate, w.clip_range, w.reward_gae_lambda, w.cost_gae_lambda), key
        assert (c.penalty_initial_value, c.penalty_learning_rate) == (w.penalty_initial_value, w.penalty_learning_rate)
        assert utils.get_net_arch(c) == [dict(pi=[64, 64], vf=[64, 64], cvf=[64, 64])]
    c = utils.Config(vars(cpg_parser().parse_args(README_CPG.split())))
    assert c.cn_obs_select_dim == [0, 1] and c.cn_acs_select_dim == [-1] and c.timesteps == 1500000
    assert c.penalty_learning_rate == WORKLOADS["pointcircle"].penalty_learning_rate == 1.0


def test_resolve_config_with_json_file_and_run_directory(tmp_path, monkeypatch):
    """Config file < command line (icrl/icrl.py:419-447); the run directory stands in for wandb.run.dir and holds config.json."""
    import json as _json
    from icrl_b200.icrl import build_parser, resolve_config
    cfg_file = tmp_path / "hc.json"
    cfg_file.write_text(_json.dumps({"n_iters": 3, "cn_layers": [20], "train_env_id": "SynthHCWithPos-v0", "seed": 11}))
    monkeypatch.setenv("ICRL_SAVE_ROOT", str(tmp_path / "runs"))
    config = resolve_config(build_parser(), ["icrl", "-cf", str(cfg_file), "-ni", "5", "-bs", "128"])
    assert config.n_iters == 5 and config.cn_layers == [20] and config.batch_size == 128 and config.seed == 11
    assert config.train_env_id == "SynthHCWithPos-v0" and config.name.endswith("_s_11")
    assert os.path.isdir(config.save_dir) and config.save_dir.endswith("files")
    saved = utils.load_dict_from_json(config.save_dir, "config")
    assert saved["n_iters"] == 5 and saved["batch_size"] == 128 and saved["save_dir"] == config.save_dir


def test_load_expert_data_layout(tmp_path):
    """files/EXPERT/rollouts/<i>.pkl with observations / actions / rewards (icrl/icrl.py:26-43)."""
    import pickle
    import numpy as np
    from icrl_b200.icrl import load_expert_data
    d = tmp_path / "files" / "EXPERT" / "rollouts"
    d.mkdir(parents=True)
    for i in range(3):
        with open(d / f"{i}.pkl", "wb") as f:
            pickle.dump(dict(observations=np.full((4 + i, 2), i, np.float64), actions=np.zeros((4 + i, 1), np.float32),
                             rewards=np.array([10.0 * i]), lengths=np.array([4 + i]), save_scheme="not_airl"), f)
    (obs, acs), mean_reward = load_expert_data(str(tmp_path), 3)
    assert obs.shape == (15, 2) and acs.shape == (15, 1) and mean_reward == 10.0
    assert obs[4, 0] == 1 and obs[-1, 0] == 2
