"""The reference-facing drivers end to end on the GPU: checkpoints in the reference's zip layout (including the
reference's own expert zips), and `run_me.py cpg -> run_policy -> icrl` on the built-in synthetic environment."""
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch as th

from conftest import load_golden

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["lgw", "hc"])
def test_load_reference_expert_zip(name, tmp_path):
    from icrl_b200.ppo_lag import PPOLagrangian
    g = load_golden(f"ckpt_{name}")
    model = PPOLagrangian.load(os.path.join(GOLD, f"ref_{name}_best_model.zip"))
    v, cv, lp, ent = (x.cpu().numpy() for x in model.policy.evaluate_actions(th.tensor(g["obs"]), th.tensor(g["acts"])))
    np.testing.assert_allclose(v, g["values"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(cv, g["cost_values"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(lp, g["log_prob"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ent, g["entropy"], rtol=1e-5, atol=1e-6)
    osd = model.policy.optimizer.state_dict()
    np.testing.assert_array_equal(osd["state"][0]["exp_avg"].numpy(), g["adam_m0"])
    assert int(osd["state"][0]["step"]) == int(g["adam_step"])
    # our own save -> load round trip keeps parameters, optimiser moments and the dual variable bit for bit
    model.dual.update_parameter(0.3)
    path = str(tmp_path / "again")
    model.save(path)
    again = PPOLagrangian.load(path)
    assert th.equal(again.policy.parameters_flat(), model.policy.parameters_flat())
    assert th.equal(again.policy._adam_m, model.policy._adam_m) and th.equal(again.policy._adam_v, model.policy._adam_v)
    assert th.equal(again.dual.nu.state, model.dual.nu.state) and again.dual.steps == 1
    assert again.n_steps == model.n_steps and again.target_kl == model.target_kl


def _run(args, env):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run_me.py")] + args, cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_run_me_cpg_run_policy_icrl(tmp_path):
    env = dict(os.environ, ICRL_SAVE_ROOT=str(tmp_path / "runs"))
    common = ["-tei", "SynthHCWithPos-v0", "-eei", "SynthHCWithPosTest-v0", "-ns", "256", "-nt", "4", "-s", "1"]
    # 1. expert: PPO-Lagrangian on the true cost
    _run(["cpg"] + common + ["-t", "4096", "-ee", "256", "-se", "512", "-tk", "0.01"], env)
    run_dir = os.path.dirname(glob.glob(str(tmp_path / "runs" / "*" / "files"))[0])
    assert os.path.exists(os.path.join(run_dir, "files", "best_model.zip"))
    assert os.path.exists(os.path.join(run_dir, "files", "train_env_stats.pkl"))
    assert glob.glob(os.path.join(run_dir, "files", "models", "rl_model_*_steps.zip"))
    # 2. expert rollouts in the layout load_expert_data reads
    _run(["run_policy", "-l", run_dir, "-s", "EXPERT", "-nr", "3"], env)
    assert len(glob.glob(os.path.join(run_dir, "files", "EXPERT", "rollouts", "*.pkl"))) == 3
    # 3. ICRL: two outer iterations with the HalfCheetah README hyper-parameters, scaled down
    out = _run(["icrl"] + common + ["-ep", run_dir, "-er", "3", "-ft", "2048", "-ni", "2", "-bi", "3", "-cl", "20",
                                    "-clr", "0.05", "-crc", "0.5", "-aclr", "0.9", "-psis", "-ctkno", "2.5", "-tk",
                                    "0.01"], env)
    icrl_dir = [d for d in glob.glob(str(tmp_path / "runs" / "*" / "files")) if os.path.dirname(d) != run_dir][0]
    rows = [json.loads(l) for l in open(os.path.join(icrl_dir, "metrics.jsonl"))]
    assert len(rows) == 2 and rows[1]["iteration"] == 1
    for k in ("true/reward", "true/cost", "backward/cn_loss", "backward/is_mean", "forward/nu", "forward/policy_gradient_loss"):
        assert k in rows[1] and np.isfinite(rows[1][k]), (k, rows[1].get(k))
    assert os.path.exists(os.path.join(icrl_dir, "best_cn_model.pt"))
    assert os.path.exists(os.path.join(icrl_dir, "models", "icrl_1_itrs", "nominal_agent.zip"))
    assert "Beginning training" in out
    # 4. the same ICRL run with whole-rollout relabelling on the device (K1 + K5 after collection instead of one cost call
    #    per environment step): identical seeds -> identical learned costs, so the metrics of iteration 0 agree
    env_wb = dict(env, ICRL_WHOLE_BUFFER_RELABEL="1")
    _run(["icrl"] + common + ["-ep", run_dir, "-er", "3", "-ft", "2048", "-ni", "1", "-bi", "3", "-cl", "20", "-clr", "0.05", "-crc",
                              "0.5", "-aclr", "0.9", "-psis", "-ctkno", "2.5", "-tk", "0.01"], env_wb)
    dirs = sorted(d for d in glob.glob(str(tmp_path / "runs" / "*" / "files")) if os.path.dirname(d) != run_dir)
    wb_dir = [d for d in dirs if d != icrl_dir][0]
    wb_row = json.loads(open(os.path.join(wb_dir, "metrics.jsonl")).readline())
    for k in ("forward/average_cost", "forward/nu", "backward/cn_loss", "true/cost"):
        assert np.isclose(wb_row[k], rows[0][k], rtol=1e-5, atol=1e-7), (k, wb_row[k], rows[0][k])


def test_run_me_cpg_with_frozen_constraint_net_and_gail(tmp_path):
    """`cpg --cn_path <cn.pt>` (K1 per env step through VecCostWrapper) and `cpg --load_gail --cn_path <gail.pt>`."""
    import torch as th
    from icrl_b200.constraint_net import ConstraintNet
    env = dict(os.environ, ICRL_SAVE_ROOT=str(tmp_path / "runs"))
    rng = np.random.default_rng(0)
    cn = ConstraintNet(18, 6, (20,), None, lambda _: 1e-3, rng.standard_normal((20, 18)), rng.standard_normal((20, 6)), False,
                       0.5, clip_obs=20)
    cn_path = str(tmp_path / "cn.pt")
    cn.save(cn_path)
    common = ["-tei", "SynthHCWithPos-v0", "-eei", "SynthHCWithPosTest-v0", "-ns", "128", "-nt", "2", "-s", "2", "-t", "512",
              "-ee", "128", "-se", "256"]
    _run(["cpg"] + common + ["--cn_path", cn_path], env)
    # a discriminator in the reference's gail_discriminator.pt schema with this env's shape
    sd = th.load(os.path.join(GOLD, "ref_gail_point.pt"), weights_only=False)
    net = {k: v.clone() for k, v in sd["network"].items()}
    sd.update(obs_dim=18, acs_dim=6, obs_select_dim=[0, 1], acs_select_dim=[-1], action_low=-np.ones(6, np.float32),
              action_high=np.ones(6, np.float32), network=net)
    gail_path = str(tmp_path / "gail_discriminator.pt")
    th.save(sd, gail_path)
    _run(["cpg"] + common + ["--load_gail", "--cn_path", gail_path, "-cosd", "0", "1", "-casd", "-1"], env)
    assert len(glob.glob(str(tmp_path / "runs" / "*" / "files" / "best_model.zip"))) == 2
