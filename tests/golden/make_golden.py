"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference) on seeded inputs, through oracle/ref_shim.py.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # everything: k1 k2 k2x k3 k4 k4x venv flags ckpt gail
    python tests/golden/make_golden.py venv gail  # or a subset

Outputs are small .npz files committed next to this script; the tests (CPU and GPU) read
only those, never the reference.  torch 2.11.0 / numpy 2.3.5 CPU, torch.set_num_threads(1)
for run-to-run determinism.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

ref_shim.install()
import gym  # noqa: E402  (the stub)
import torch as th  # noqa: E402
from icrl.constraint_net import ConstraintNet  # noqa: E402
from stable_baselines3 import PPOLagrangian  # noqa: E402
from stable_baselines3.common import logger  # noqa: E402
from stable_baselines3.common.buffers import RolloutBufferWithCost  # noqa: E402
from stable_baselines3.common.vec_env import DummyVecEnv  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = ref_shim.REFERENCE_ROOT
th.set_num_threads(1)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KB)")


def sd_arrays(prefix, state_dict):
    return {f"{prefix}{k}": v.detach().cpu().numpy().copy() for k, v in state_dict.items()}


# ------------------------------------------------------------------------------------------- K1

SHAPES = {
    # name: obs_dim, acs_dim, is_discrete, cn hidden, policy batch, epochs
    "lgw": dict(obs_dim=1, acs_dim=2, is_discrete=True, hidden=(20,)),
    "hc": dict(obs_dim=18, acs_dim=6, is_discrete=False, hidden=(20,)),
    "ant": dict(obs_dim=113, acs_dim=8, is_discrete=False, hidden=(40, 40)),
    "point": dict(obs_dim=6, acs_dim=2, is_discrete=False, hidden=(40, 40)),
}


def synth_obs_acs(rng, n, obs_dim, acs_dim, is_discrete, obs_dtype=np.float64):
    scale = rng.uniform(0.5, 8.0, size=obs_dim)
    obs = (rng.standard_normal((n, obs_dim)) * scale).astype(obs_dtype)
    if is_discrete:
        acs = rng.integers(0, acs_dim, size=(n, 1)).astype(np.float32)
    else:
        acs = rng.standard_normal((n, acs_dim)).astype(np.float32)
    return obs, acs


def make_cn(shape, seed, *, normalize=False, clip_obs=20., reg=0.0, no_is=False, per_step=False,
            expert=None, tkon=-1, tkno=-1, lr=3e-3, obs_select=None, acs_select=None, rng=None, batch_size=None,
            gail=False):
    th.manual_seed(seed)
    s = SHAPES[shape]
    low = high = None
    if not s["is_discrete"]:
        low, high = -np.ones(s["acs_dim"], np.float32), np.ones(s["acs_dim"], np.float32)
    mean = var = None
    if normalize:
        mean = rng.standard_normal(s["obs_dim"])
        var = rng.uniform(0.3, 9.0, size=s["obs_dim"])
    eo, ea = expert if expert is not None else (None, None)
    return ConstraintNet(
        s["obs_dim"], s["acs_dim"], s["hidden"], batch_size, lambda x: lr, eo, ea, s["is_discrete"], reg,
        obs_select, acs_select, no_importance_sampling=no_is, per_step_importance_sampling=per_step,
        clip_obs=clip_obs, initial_obs_mean=mean, initial_obs_var=var, action_low=low, action_high=high,
        target_kl_old_new=tkon, target_kl_new_old=tkno, train_gail_lambda=gail, eps=1e-5, device="cpu")


def golden_k1():
    rng = np.random.default_rng(1234)
    for shape in SHAPES:
        s = SHAPES[shape]
        for variant, normalize in (("raw", False), ("norm", True)):
            cn = make_cn(shape, seed=0, normalize=normalize, rng=rng)
            for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
                obs, acs = synth_obs_acs(rng, 257, s["obs_dim"], s["acs_dim"], s["is_discrete"], dt)
                cost = cn.cost_function(obs, acs)
                x = cn.prepare_data(obs, acs).numpy()
                extra = {}
                if normalize:
                    extra = dict(obs_mean=cn.current_obs_mean, obs_var=cn.current_obs_var)
                save(f"k1_{shape}_{variant}_{tag}", obs=obs, acs=acs, cost=cost, x=x, clip_obs=np.float64(20.),
                     **sd_arrays("p.", cn.network.state_dict()), **extra)
    # 3-D input as AdjustedRewardCallback passes it (icrl/utils.py:565-567): [T, E, .]
    cn = make_cn("hc", seed=1, rng=rng)
    obs, acs = synth_obs_acs(rng, 6 * 5, 18, 6, False, np.float32)
    obs, acs = obs.reshape(6, 5, 18), acs.reshape(6, 5, 6)
    save("k1_hc_3d", obs=obs, acs=acs, cost=cn.cost_function(obs, acs), **sd_arrays("p.", cn.network.state_dict()))

    # obs/acs select dims + the real frozen checkpoint shipped with the reference (cpg path)
    path = os.path.join(REF, "icrl/expert_data/ConstraintTransfer/ICRL/Point/files/best_cn_model.pt")
    raw = th.load(path)
    cn = ConstraintNet.load(path, obs_dim=6, acs_dim=2, is_discrete=False, obs_select_dim=[0, 1], acs_select_dim=[-1],
                            clip_obs=None, obs_mean=None, obs_var=None, action_low=-0.25 * np.ones(2, np.float32),
                            action_high=0.25 * np.ones(2, np.float32), device="cpu")
    obs, acs = synth_obs_acs(rng, 300, 6, 2, False, np.float64)
    obs[:4, :2] = [[0, 0], [5, 5], [-8, 3], [100, -100]]
    save("k1_point_ckpt", obs=obs, acs=acs, cost=cn.cost_function(obs, acs),
         loaded_clip_obs_is_none=np.array(cn.clip_obs is None), loaded_action_high_is_none=np.array(cn.action_high is None),
         hidden_sizes=np.array(raw["hidden_sizes"]), select_dim=np.array(cn.select_dim),
         **sd_arrays("p.", raw["cn_network"]))
    path = os.path.join(REF, "icrl/expert_data/ConstraintTransfer/ICRL/AntBroken/files/best_cn_model.pt")
    raw = th.load(path)
    cn = ConstraintNet.load(path, obs_dim=113, acs_dim=8, is_discrete=False, device="cpu")
    obs, acs = synth_obs_acs(rng, 128, 113, 8, False, np.float64)
    save("k1_antbroken_ckpt", obs=obs, acs=acs, cost=cn.cost_function(obs, acs),
         hidden_sizes=np.array(raw["hidden_sizes"]), select_dim=np.array(cn.select_dim),
         **sd_arrays("p.", raw["cn_network"]))

    # a slice of the real AntWall expert rollouts (icrl/icrl.py:25-43 schema)
    import pickle
    with open(os.path.join(REF, "icrl/expert_data/AntWall/files/EXPERT/rollouts/0.pkl"), "rb") as f:
        data = pickle.load(f)
    cn = make_cn("ant", seed=3, rng=rng)
    obs, acs = data["observations"][:96], data["actions"][:96]
    save("k1_ant_expert_slice", obs=obs, acs=acs, cost=cn.cost_function(obs, acs),
         **sd_arrays("p.", cn.network.state_dict()))


# ------------------------------------------------------------------------------------------- K2

def opt_arrays(cn):
    out = {}
    st = cn.optimizer.state_dict()["state"]
    for i, (k, v) in enumerate(sorted(st.items())):
        out[f"adam.{i}.step"] = np.asarray(float(v["step"]))
        out[f"adam.{i}.exp_avg"] = v["exp_avg"].numpy().copy()
        out[f"adam.{i}.exp_avg_sq"] = v["exp_avg_sq"].numpy().copy()
    return out


def golden_k2():
    rng = np.random.default_rng(99)
    cases = {
        # name: shape, episode lengths, n_expert, kwargs, train calls [(iters, lr-progress)]
        "hc_perstep": ("hc", [500, 500, 400], 900, dict(per_step=True, reg=0.5, tkno=2.5, tkon=10, lr=0.05), 3),
        "ant_perstep": ("ant", [150, 150], 250, dict(per_step=True, reg=0.6, tkno=2.5, tkon=10, lr=0.005), 3),
        "hc_perstep_mild": ("hc", [50, 60, 40], 140, dict(per_step=True, reg=0.5, tkno=2.5, tkon=10, lr=0.002), 4),
        "lgw_perepisode": ("lgw", [40, 35, 50, 25], 160, dict(tkno=10, tkon=10, lr=0.003, clip_obs=20.), 4),
        "hc_perepisode_norm": ("hc", [30, 20, 25, 40], 100, dict(reg=0.1, lr=0.01, normalize=True), 3),
        "hc_nois": ("hc", [200, 100], 250, dict(no_is=True, reg=0.5, lr=0.05), 2),
        "hc_earlystop": ("hc", [60, 50, 40], 120, dict(tkno=1e-4, tkon=10, lr=0.05), 6),
    }
    for name, (shape, lengths, n_exp, kw, iters) in cases.items():
        s = SHAPES[shape]
        n_nom = int(np.sum(lengths))
        eo, ea = synth_obs_acs(rng, n_exp, s["obs_dim"], s["acs_dim"], s["is_discrete"])
        no, na = synth_obs_acs(rng, n_nom, s["obs_dim"], s["acs_dim"], s["is_discrete"])
        no = no * 1.5 + 0.5
        cn = make_cn(shape, seed=7, expert=(eo, ea), rng=rng, **kw)
        mean = var = None
        if kw.get("normalize"):
            mean, var = cn.current_obs_mean, cn.current_obs_var
        arrays = dict(expert_obs=eo, expert_acs=ea, nominal_obs=no, nominal_acs=na, lengths=np.array(lengths),
                      iters=np.array(iters), lr=np.array(kw["lr"]), reg=np.array(kw.get("reg", 0.0)),
                      per_step=np.array(kw.get("per_step", False)), no_is=np.array(kw.get("no_is", False)),
                      tkon=np.array(kw.get("tkon", -1.0)), tkno=np.array(kw.get("tkno", -1.0)))
        if mean is not None:
            arrays.update(obs_mean=mean, obs_var=var)
        arrays.update(sd_arrays("p0.", cn.network.state_dict()))
        # two consecutive train() calls: Adam state persists across ICRL iterations (icrl/icrl.py:235)
        for call in (1, 2):
            m = cn.train(iters, no, na, np.array(lengths), mean, var, 1.0)
            arrays.update(sd_arrays(f"p{call}.", cn.network.state_dict()))
            arrays.update({f"m{call}.{k}": np.asarray(v, dtype=np.float64) for k, v in m.items()})
            arrays.update({f"c{call}.{k}": v for k, v in opt_arrays(cn).items()})
            if name == "hc_earlystop":
                break
        save(f"k2_{name}", **arrays)



def golden_k2x():
    """Round-2 additions: the --train_gail_lambda BCE variant (constraint_net.py:193-197) and the -cbs minibatch mode
    (constraint_net.py:304-317; numpy permutation per backward iteration, seed stored)."""
    rng = np.random.default_rng(199)
    cases = {
        # name: shape, episode lengths, n_expert, kwargs, iters
        "hc_gail": ("hc", [70, 50, 60], 200, dict(gail=True, lr=0.01, tkno=10, tkon=10), 4),
        "ant_gail_nois": ("ant", [90, 70], 150, dict(gail=True, no_is=True, lr=0.003), 3),
        "hc_mb_perstep": ("hc", [50, 60, 40], 140, dict(per_step=True, reg=0.5, tkno=2.5, tkon=10, lr=0.001, batch_size=64), 3),
        "hc_mb_perepisode": ("hc", [50, 60, 40, 30], 150, dict(reg=0.2, tkno=10, tkon=10, lr=0.0005, batch_size=48), 3),
        "lgw_mb_big": ("lgw", [40, 35, 50], 160, dict(lr=0.003, clip_obs=20., batch_size=500), 3),   # one minibatch per iteration
        "hc_mb_earlystop": ("hc", [60, 50, 40], 120, dict(tkno=1e-4, tkon=10, lr=0.05, batch_size=50), 6),
    }
    for name, (shape, lengths, n_exp, kw, iters) in cases.items():
        s = SHAPES[shape]
        n_nom = int(np.sum(lengths))
        eo, ea = synth_obs_acs(rng, n_exp, s["obs_dim"], s["acs_dim"], s["is_discrete"])
        no, na = synth_obs_acs(rng, n_nom, s["obs_dim"], s["acs_dim"], s["is_discrete"])
        no = no * 1.5 + 0.5
        cn = make_cn(shape, seed=7, expert=(eo, ea), rng=rng, **kw)
        arrays = dict(expert_obs=eo, expert_acs=ea, nominal_obs=no, nominal_acs=na, lengths=np.array(lengths),
                      iters=np.array(iters), lr=np.array(kw["lr"]), reg=np.array(kw.get("reg", 0.0)),
                      per_step=np.array(kw.get("per_step", False)), no_is=np.array(kw.get("no_is", False)),
                      tkon=np.array(kw.get("tkon", -1.0)), tkno=np.array(kw.get("tkno", -1.0)),
                      gail=np.array(kw.get("gail", False)), batch_size=np.array(kw.get("batch_size", 0)),
                      numpy_seed=np.array(321))
        arrays.update(sd_arrays("p0.", cn.network.state_dict()))
        np.random.seed(321)
        for call in (1, 2):
            m = cn.train(iters, no, na, np.array(lengths), None, None, 1.0)
            arrays.update(sd_arrays(f"p{call}.", cn.network.state_dict()))
            arrays.update({f"m{call}.{k}": np.asarray(v, dtype=np.float64) for k, v in m.items()})
            arrays.update({f"c{call}.{k}": v for k, v in opt_arrays(cn).items()})
            arrays[f"rng_probe{call}"] = np.array(np.random.get_state()[1][:4].astype(np.int64))   # where the global RNG stands
            arrays[f"rng_pos{call}"] = np.array(np.random.get_state()[2])
            if name == "hc_mb_earlystop":
                break
        save(f"k2_{name}", **arrays)

# ------------------------------------------------------------------------------------------- K3

def golden_k3():
    rng = np.random.default_rng(5)
    for name, T, E, ep_len, lam_r, lam_c in (("small", 37, 3, 11, 0.95, 0.9), ("hc", 2048, 5, 1000, 0.95, 0.95),
                                             ("ant", 2048, 5, 500, 0.9, 0.9), ("lam1", 64, 2, 1000, 1.0, 1.0),
                                             ("wide", 256, 67, 50, 0.95, 0.97)):
        buf = RolloutBufferWithCost(T, gym.spaces.Box(-1, 1, (3,)), gym.spaces.Box(-1, 1, (2,)), "cpu",
                                    reward_gamma=0.99, reward_gae_lambda=lam_r, cost_gamma=0.98, cost_gae_lambda=lam_c,
                                    n_envs=E)
        for k in ("rewards", "reward_values", "costs", "cost_values"):
            getattr(buf, k)[:] = rng.standard_normal((T, E)).astype(np.float32)
        phase = rng.integers(0, ep_len, size=E)
        t = np.arange(T)[:, None]
        buf.dones[:] = (((t + phase[None]) % ep_len) == 0).astype(np.float32)
        last_dones = rng.random(E) < 0.4
        rlv = th.tensor(rng.standard_normal((E, 1)).astype(np.float32))
        clv = th.tensor(rng.standard_normal((E, 1)).astype(np.float32))
        ins = {k: getattr(buf, k).copy() for k in ("rewards", "reward_values", "costs", "cost_values", "dones")}
        buf.compute_returns_and_advantage(rlv, clv, dones=last_dones)
        save(f"k3_{name}", **ins, last_dones=last_dones, reward_last_value=rlv.numpy(), cost_last_value=clv.numpy(),
             gammas=np.array([0.99, lam_r, 0.98, lam_c]),
             reward_returns=buf.reward_returns, reward_advantages=buf.reward_advantages,
             cost_returns=buf.cost_returns, cost_advantages=buf.cost_advantages)


# ------------------------------------------------------------------------------------------- K4

class FakeEnv(gym.Env):
    def __init__(self, obs_dim, acs_dim, is_discrete):
        self.observation_space = gym.spaces.Box(-np.inf, np.inf, (obs_dim,), np.float32)
        self.action_space = gym.spaces.Discrete(acs_dim) if is_discrete else gym.spaces.Box(-1, 1, (acs_dim,), np.float32)

    def reset(self):
        return np.zeros(self.observation_space.shape, np.float32)

    def step(self, a):
        return self.reset(), 0.0, False, {}


def golden_k4():
    rng = np.random.default_rng(11)
    cases = {
        # name: shape, T, E, batch, epochs, kwargs
        "hc": ("hc", 64, 4, 64, 3, dict(target_kl=0.01)),
        "ant": ("ant", 128, 5, 128, 2, dict(learning_rate=3e-5, clip_range=0.4, penalty_initial_value=0.1,
                                            penalty_learning_rate=0.05, target_kl=0.02)),
        "lgw": ("lgw", 50, 4, 64, 2, dict(target_kl=0.01)),          # last minibatch is short (200 = 3*64 + 8)
        "point": ("point", 48, 4, 64, 2, dict(target_kl=0.01, penalty_learning_rate=1.0, ent_coef=0.01)),
        "hc_vfclip": ("hc", 32, 4, 32, 2, dict(clip_range_reward_vf=0.2, clip_range_cost_vf=0.3, ent_coef=0.02)),
        "hc_fullbatch": ("hc", 32, 4, None, 1, dict()),              # exactly one optimiser step
        "hc_klstop": ("hc", 64, 4, 64, 6, dict(target_kl=1e-6, learning_rate=3e-3)),
    }
    for name, (shape, T, E, bs, ne, kw) in cases.items():
        s = SHAPES[shape]
        env = DummyVecEnv([lambda s=s: FakeEnv(s["obs_dim"], s["acs_dim"], s["is_discrete"]) for _ in range(E)])
        th.manual_seed(5)
        algo = PPOLagrangian("TwoCriticsMlpPolicy", env, n_steps=T, batch_size=bs, n_epochs=ne, seed=3, device="cpu",
                             **kw)
        pol, buf = algo.policy, algo.rollout_buffer
        obs, acs = synth_obs_acs(rng, T * E, s["obs_dim"], s["acs_dim"], s["is_discrete"], np.float32)
        obs = np.clip(obs / 4.0, -10, 10).astype(np.float32)
        buf.observations[:] = obs.reshape(T, E, -1)
        buf.orig_observations[:] = obs.reshape(T, E, -1) * 2
        buf.actions[:] = acs.reshape(T, E, -1)
        for k in ("rewards", "costs", "orig_costs"):
            getattr(buf, k)[:] = np.abs(rng.standard_normal((T, E))).astype(np.float32) * (0.2 if "cost" in k else 1.0)
        buf.dones[:] = (rng.random((T, E)) < 0.02).astype(np.float32)
        # values / log-probs from the policy itself on the stored (obs, action): ratio == 1 at the first step
        with th.no_grad():
            a = th.tensor(acs).long().flatten() if s["is_discrete"] else th.tensor(acs)
            v, cv, lp, _ = pol.evaluate_actions(th.tensor(obs), a)
        buf.reward_values[:] = v.numpy().reshape(T, E)
        buf.cost_values[:] = cv.numpy().reshape(T, E)
        buf.log_probs[:] = lp.numpy().reshape(T, E) + 0.05 * rng.standard_normal((T, E)).astype(np.float32)
        buf.full, buf.pos = True, T
        last_dones = rng.random(E) < 0.3
        buf.compute_returns_and_advantage(v[-E:], cv[-E:], dones=last_dones)
        arrays = {f"buf.{k}": getattr(buf, k).copy() for k in (
            "observations", "actions", "log_probs", "reward_values", "reward_advantages",
            "reward_returns", "cost_values", "cost_advantages", "cost_returns", "orig_costs")}
        arrays.update(sd_arrays("p0.", pol.state_dict()))
        arrays["param_order"] = np.array([n for n, _ in pol.named_parameters()])
        arrays["nu0"] = algo.dual.nu().detach().numpy()
        arrays["log_nu0"] = algo.dual.nu.log_nu.detach().numpy().copy()
        logger.configure(folder=None, format_strings=[])
        algo._current_progress_remaining = 1.0
        np.random.seed(17)
        algo.train()
        arrays.update(sd_arrays("p1.", pol.state_dict()))
        arrays.update({f"log.{k}": np.asarray(v, dtype=np.float64) for k, v in logger.Logger.CURRENT.name_to_value.items()})
        arrays["log_nu1"] = algo.dual.nu.log_nu.detach().numpy().copy()
        st = pol.optimizer.state_dict()["state"]
        for i, k in enumerate(sorted(st)):
            arrays[f"adam.{i}.exp_avg"] = st[k]["exp_avg"].numpy().copy()
            arrays[f"adam.{i}.exp_avg_sq"] = st[k]["exp_avg_sq"].numpy().copy()
            arrays[f"adam.{i}.step"] = np.asarray(float(st[k]["step"]))
        hp = dict(batch_size=-1 if bs is None else bs, n_epochs=ne, numpy_seed=17, T=T, E=E,
                  learning_rate=kw.get("learning_rate", 3e-4), clip_range=kw.get("clip_range", 0.2),
                  target_kl=kw.get("target_kl", -1.0), ent_coef=kw.get("ent_coef", 0.0),
                  penalty_initial_value=kw.get("penalty_initial_value", 1.0),
                  penalty_learning_rate=kw.get("penalty_learning_rate", 0.01),
                  clip_range_reward_vf=kw.get("clip_range_reward_vf", -1.0),
                  clip_range_cost_vf=kw.get("clip_range_cost_vf", -1.0))
        arrays.update({f"hp.{k}": np.asarray(v, dtype=np.float64) for k, v in hp.items()})
        # a second train() on the same buffer: Adam moments / step counts / nu carry over
        algo.train()
        arrays.update(sd_arrays("p2.", pol.state_dict()))
        arrays["log_nu2"] = algo.dual.nu.log_nu.detach().numpy().copy()
        save(f"k4_{name}", **arrays)



def k4x_inputs(seed, T, E, obs_dim, acs_dim, is_discrete):
    """Seeded (obs, actions, rewards-like scalars) of the round-2 K4 fixtures; tests/helpers.py regenerates the SAME arrays
    (numpy Generator streams are stable), so only the torch-dependent arrays and the outputs are stored."""
    rng = np.random.default_rng(seed)
    n = T * E
    scale = rng.uniform(0.5, 8.0, size=obs_dim)
    obs = np.clip((rng.standard_normal((n, obs_dim)) * scale).astype(np.float32) / 4.0, -10, 10).astype(np.float32)
    acs = (rng.integers(0, acs_dim, size=(n, 1)).astype(np.float32) if is_discrete
           else rng.standard_normal((n, acs_dim)).astype(np.float32))
    rewards = np.abs(rng.standard_normal((T, E))).astype(np.float32)
    costs = (np.abs(rng.standard_normal((T, E))) * 0.2).astype(np.float32)
    dones = (rng.random((T, E)) < 0.02).astype(np.float32)
    lp_noise = (0.05 * rng.standard_normal((T, E))).astype(np.float32)
    last_dones = rng.random(E) < 0.3
    return obs, acs, rewards, costs, dones, lp_noise, last_dones


def golden_k4x():
    """Round-2 K4 fixtures: the full-size HalfCheetah train() (2048 x 5, batch 64, 10 epochs = 1 600 dependent optimiser
    steps: drift over a whole launch) and the large-batch regime (batch >= 2048: the many-cluster kernel)."""
    cases = {
        # name: shape, T, E, batch, epochs, trains, kwargs
        "hc_full": ("hc", 2048, 5, 64, 10, 1, dict()),
        "hc_wide": ("hc", 1024, 8, 4096, 3, 2, dict(target_kl=0.05)),
        "ant_wide_ragged": ("ant", 512, 8, 3000, 3, 2, dict(learning_rate=3e-5, clip_range=0.4, penalty_initial_value=0.1,
                                                            penalty_learning_rate=0.05)),
        "lgw_wide": ("lgw", 1024, 4, 2048, 2, 1, dict(ent_coef=0.01)),
        "hc_wide_fullbatch": ("hc", 1024, 8, None, 2, 1, dict(clip_range_reward_vf=0.2)),
    }
    for ci, (name, (shape, T, E, bs, ne, trains, kw)) in enumerate(cases.items()):
        s = SHAPES[shape]
        seed = 7000 + ci
        env = DummyVecEnv([lambda s=s: FakeEnv(s["obs_dim"], s["acs_dim"], s["is_discrete"]) for _ in range(E)])
        th.manual_seed(5)
        algo = PPOLagrangian("TwoCriticsMlpPolicy", env, n_steps=T, batch_size=bs, n_epochs=ne, seed=3, device="cpu", **kw)
        pol, buf = algo.policy, algo.rollout_buffer
        obs, acs, rewards, costs, dones, lp_noise, last_dones = k4x_inputs(seed, T, E, s["obs_dim"], s["acs_dim"], s["is_discrete"])
        buf.observations[:] = obs.reshape(T, E, -1)
        buf.orig_observations[:] = obs.reshape(T, E, -1) * 2
        buf.actions[:] = acs.reshape(T, E, -1)
        buf.rewards[:], buf.costs[:], buf.orig_costs[:], buf.dones[:] = rewards, costs, costs, dones
        with th.no_grad():
            a = th.tensor(acs).long().flatten() if s["is_discrete"] else th.tensor(acs)
            v, cv, lp, _ = pol.evaluate_actions(th.tensor(obs), a)
        buf.reward_values[:] = v.numpy().reshape(T, E)
        buf.cost_values[:] = cv.numpy().reshape(T, E)
        buf.log_probs[:] = lp.numpy().reshape(T, E) + lp_noise
        buf.full, buf.pos = True, T
        buf.compute_returns_and_advantage(v[-E:], cv[-E:], dones=last_dones)
        arrays = {f"buf.{k}": getattr(buf, k).copy() for k in (
            "log_probs", "reward_values", "reward_advantages", "reward_returns", "cost_values", "cost_advantages",
            "cost_returns")}
        arrays["input_seed"] = np.array(seed)
        arrays["obs_sum"] = np.array(np.float64(obs.astype(np.float64).sum()))        # guards the regenerated inputs
        arrays["acs_sum"] = np.array(np.float64(acs.astype(np.float64).sum()))
        arrays.update(sd_arrays("p0.", pol.state_dict()))
        arrays["param_order"] = np.array([n for n, _ in pol.named_parameters()])
        arrays["log_nu0"] = algo.dual.nu.log_nu.detach().numpy().copy()
        logger.configure(folder=None, format_strings=[])
        algo._current_progress_remaining = 1.0
        np.random.seed(17)
        for call in range(1, trains + 1):
            algo.train()
            arrays.update(sd_arrays(f"p{call}.", pol.state_dict()))
            arrays[f"log_nu{call}"] = algo.dual.nu.log_nu.detach().numpy().copy()
            if call == 1:
                arrays.update({f"log.{k}": np.asarray(v, dtype=np.float64)
                               for k, v in logger.Logger.CURRENT.name_to_value.items()})
        hp = dict(batch_size=-1 if bs is None else bs, n_epochs=ne, numpy_seed=17, T=T, E=E, trains=trains,
                  learning_rate=kw.get("learning_rate", 3e-4), clip_range=kw.get("clip_range", 0.2),
                  target_kl=kw.get("target_kl", -1.0), ent_coef=kw.get("ent_coef", 0.0),
                  penalty_initial_value=kw.get("penalty_initial_value", 1.0),
                  penalty_learning_rate=kw.get("penalty_learning_rate", 0.01),
                  clip_range_reward_vf=kw.get("clip_range_reward_vf", -1.0),
                  clip_range_cost_vf=kw.get("clip_range_cost_vf", -1.0))
        arrays.update({f"hp.{k}": np.asarray(v, dtype=np.float64) for k, v in hp.items()})
        save(f"k4x_{name}", **arrays)

# ------------------------------------------------------------------------------------------- VecEnv wrappers
def golden_venv():
    """Reference VecCostWrapper + VecNormalizeWithCost driven by the scripted env (tests/golden/scripted_env.py)."""
    sys.path.insert(0, OUT)
    from scripted_env import ScriptedEnv, scripted_actions, scripted_cost
    from stable_baselines3.common.vec_env import VecCostWrapper, VecNormalizeWithCost, sync_envs_normalization
    n_envs, steps = 3, 120
    for name, kw in {"all": dict(norm_obs=True, norm_reward=True, norm_cost=True),
                     "nocost": dict(norm_obs=True, norm_reward=True, norm_cost=False),
                     "raw": dict(norm_obs=False, norm_reward=False, norm_cost=False)}.items():
        env = DummyVecEnv([(lambda i=i: ScriptedEnv(100 + i)) for i in range(n_envs)])
        env = VecCostWrapper(env)
        env = VecNormalizeWithCost(env, training=True, cost_info_str="cost", reward_gamma=0.99, cost_gamma=0.97, **kw)
        env.set_cost_function(scripted_cost)
        acts = scripted_actions(7, steps, n_envs)
        out = dict(obs0=env.reset(), obs=[], orig_obs=[], rew=[], done=[], cost=[], orig_cost=[])
        for t in range(steps):
            o, r, d, infos = env.step(acts[t])
            out["obs"].append(o), out["orig_obs"].append(env.get_original_obs()), out["rew"].append(r)
            out["done"].append(d), out["cost"].append([i["cost"] for i in infos])
            out["orig_cost"].append(env.get_original_cost())
        arrays = {k: np.asarray(v) for k, v in out.items()}
        for rms in ("obs_rms", "ret_rms", "cost_rms"):
            r = getattr(env, rms)
            arrays[rms + "_mean"], arrays[rms + "_var"], arrays[rms + "_count"] = r.mean, r.var, np.float64(r.count)
        # eval env synced from the training env: normalised obs of a fresh episode
        ev = VecNormalizeWithCost(DummyVecEnv([lambda: ScriptedEnv(555)]), training=False, norm_obs=kw["norm_obs"],
                                  norm_reward=False, norm_cost=False)
        sync_envs_normalization(env, ev)
        arrays["eval_obs0"] = ev.reset()
        arrays["eval_obs1"] = ev.step(acts[0][:1])[0]
        save(f"venv_{name}", **arrays)


# ------------------------------------------------------------------------------------------- CLI flag surface
def golden_flags():
    """The reference drivers' argparse surface, read from their source with `ast` (they import wandb / MuJoCo envs
    and cannot be executed here): option strings, action, nargs and literal defaults -> tests/golden/flags_*.json."""
    import ast
    import json
    for name, rel in {"icrl": "icrl/icrl.py", "cpg": "icrl/cpg.py", "run_policy": "icrl/run_policy.py"}.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        flags = []
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
                opts = [a.value for a in node.args]
                kw = {}
                for k in node.keywords:
                    if k.arg in ("default", "action", "nargs"):
                        kw[k.arg] = ast.literal_eval(k.value)
                    elif k.arg == "type":
                        kw["type"] = ast.unparse(k.value)
                flags.append(dict(opts=opts, **kw))
        with open(os.path.join(OUT, f"flags_{name}.json"), "w") as f:
            json.dump(flags, f, indent=1)
        print(f"wrote flags_{name}.json ({len(flags)} flags)")


# ------------------------------------------------------------------------------------------- reference checkpoints
def golden_ckpt():
    """Copy two of the reference's expert `best_model.zip` DATA files (SB3 zip checkpoints, not source) and record what
    the reference's own `PPOLagrangian.load(...).policy.evaluate_actions` returns for seeded inputs."""
    import shutil
    for name, sub, disc in (("lgw", "LGW", True), ("hc", "HCWithPos-New", False)):
        src = os.path.join(REF, "icrl/expert_data", sub, "files/best_model.zip")
        dst = os.path.join(OUT, f"ref_{name}_best_model.zip")
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        # the zip's `data` cloudpickles real gym spaces (not importable here), so the reference policy is built on
        # a stub env of the same shape and the zip's tensors are loaded into it with the reference's own classes
        import io
        import zipfile
        zf = zipfile.ZipFile(src)
        sd = th.load(io.BytesIO(zf.read("policy.pth")), map_location="cpu", weights_only=False)
        osd = th.load(io.BytesIO(zf.read("policy.optimizer.pth")), map_location="cpu", weights_only=False)
        obs_dim, act_out = sd["mlp_extractor.policy_net.0.weight"].shape[1], sd["action_net.weight"].shape[0]
        env = DummyVecEnv([lambda: FakeEnv(obs_dim, act_out, disc)])
        model = PPOLagrangian("TwoCriticsMlpPolicy", env, device="cpu")
        model.policy.load_state_dict(sd)
        pol = model.policy
        rng = np.random.default_rng(11)
        obs = rng.standard_normal((64, obs_dim)).astype(np.float32) * 2
        acts = (rng.integers(0, 2, (64,)).astype(np.float32) if disc
                else rng.uniform(-1, 1, (64, act_out)).astype(np.float32))
        with th.no_grad():
            v, cv, lp, ent = pol.evaluate_actions(th.tensor(obs), th.tensor(acts))
        save(f"ckpt_{name}", obs=obs, acts=acts, values=v.numpy(), cost_values=cv.numpy(), log_prob=lp.numpy(),
             entropy=ent.numpy(), adam_m0=osd["state"][osd["param_groups"][0]["params"][0]]["exp_avg"].numpy(),
             adam_step=np.int64(osd["state"][osd["param_groups"][0]["params"][0]]["step"]))


# ------------------------------------------------------------------------------------------- GAIL discriminator as a cost
def golden_gail():
    """The reference's GailDiscriminator.load + reward_function on its two shipped `gail_discriminator.pt` files (copied
    as DATA fixtures), for seeded inputs."""
    import shutil
    from icrl.gail_utils import GailDiscriminator
    rng = np.random.default_rng(23)
    for name in ("AntBroken", "Point"):
        src = os.path.join(REF, "icrl/expert_data/ConstraintTransfer/GAIL", name, "files/gail_discriminator.pt")
        dst = os.path.join(OUT, f"ref_gail_{name.lower()}.pt")
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        gail = GailDiscriminator.load(src, device="cpu")
        obs = (rng.standard_normal((257, gail.obs_dim)) * 3).astype(np.float32)
        acs = rng.uniform(-1.5, 1.5, (257, gail.acs_dim)).astype(np.float32)
        obs3, acs3 = obs[:60].reshape(12, 5, -1), acs[:60].reshape(12, 5, -1)
        save(f"gail_{name.lower()}", obs=obs, acs=acs, d=gail.reward_function(obs, acs, apply_log=False),
             logd=gail.reward_function(obs, acs, apply_log=True), d3=gail.reward_function(obs3, acs3, apply_log=False))


if __name__ == "__main__":
    which = sys.argv[1:] or ["k1", "k2", "k2x", "k3", "k4", "k4x", "venv", "flags", "ckpt", "gail"]
    for w in which:
        globals()[f"golden_{w}"]()
