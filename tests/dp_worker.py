"""Data-parallel parity checks, one process per GPU (launched by tests/test_dp_gpu.py under torch.distributed.run, and
called by bench.py --gpus N before it times anything).

  ppo_parity   K4: every rank trains on its own environment columns with the in-kernel NVLink all-reduce (single-cluster
               kernel, or the many-cluster one with `wide=True`); rank 0 checks the result against the CPU oracle run on the
               equivalent single-process problem (global batch = world * local batch), and all ranks check that their
               replicated parameters are bit-identical.
  cn_parity    K2: nominal episodes sharded by whole episodes, expert rows evenly, the three in-kernel exchanges per backward
               iteration; same two checks against oracle/cn.py::train on the global problem.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _bit_identical(t: th.Tensor, world: int) -> bool:
    allp = [th.empty_like(t) for _ in range(world)]
    dist.all_gather(allp, t.contiguous())
    return all(th.equal(allp[r], allp[0]) for r in range(world))


def ppo_parity(rank, world, dev, shape="hc", wide=False, comm=None):
    """Returns {"ok", "max_param_err", "replicas_identical", "scheme", ...}; raises nothing (bench prints the dict)."""
    from icrl_b200 import _lib
    from icrl_b200.distributed import PpoComm, global_minibatch_rows, shard_envs
    from icrl_b200.policies import ActorTwoCriticsPolicy
    from icrl_b200.spaces import Box
    from oracle import cn as ocn, gae as ogae, ppo as oppo

    D, A, T, E_local, B_local, n_epochs = 18, 6, 32, 2, 32, 3
    if shape == "ant":                                    # the widest layer-1 tiling (NT1 = 8) through the exchange
        D, A = 113, 8
    if wide:                                              # several chunks per cluster and a ragged last minibatch
        T, E_local, B_local, n_epochs = 96, 4, 160, 2
    E = E_local * world
    rng = np.random.default_rng(7)                       # identical global data on every rank
    g = {"observations": rng.standard_normal((T, E, D)).astype(np.float32),
         "actions": rng.standard_normal((T, E, A)).astype(np.float32)}
    for k in ("old_log_prob", "old_reward_values", "reward_advantages", "reward_returns", "old_cost_values",
              "cost_advantages", "cost_returns"):
        g[k] = rng.standard_normal((T, E)).astype(np.float32)
    g["old_log_prob"] = g["old_log_prob"] * 0.1 - (8.5 if shape == "hc" else 11.3)
    lo, hi = shard_envs(E, rank, world)
    assert (lo, hi) == (rank * E_local, (rank + 1) * E_local)
    th.manual_seed(3)                                     # replicated initial parameters
    pol = ActorTwoCriticsPolicy(Box(-np.inf, np.inf, (D,)), Box(-1, 1, (A,)), lambda _: 3e-4, device=dev)
    P0 = {k: v.clone() for k, v in pol.state_dict().items()}
    n_local = T * E_local
    perm_local = [[np.random.RandomState(100 * e + r).permutation(n_local) for r in range(world)] for e in range(n_epochs)]
    perm_dev = th.from_numpy(np.stack([perm_local[e][rank] for e in range(n_epochs)]).astype(np.int32)).to(dev)
    loc = {k: th.from_numpy(np.ascontiguousarray(v[:, lo:hi])).to(dev) for k, v in g.items()}
    own_comm = comm is None
    if own_comm:
        comm = PpoComm()
    steps_per_epoch = -(-n_local // B_local)
    cfg = pol.make_cfg(T=T, E=E_local, batch_size=B_local, n_epochs=n_epochs, has_target_kl=0, clip_range=0.2, ent_coef=0.01,
                       reward_vf_coef=0.5, cost_vf_coef=0.5, max_grad_norm=0.5, nu=0.7)
    data = _lib.PpoData()
    for k, v in loc.items():
        setattr(data, k, v.data_ptr())
    data.perm = perm_dev.data_ptr()
    advsums = th.zeros(n_epochs * steps_per_epoch, 4, dtype=th.float64, device=dev)
    stats = th.zeros(n_epochs * steps_per_epoch, 8, device=dev)
    result = th.zeros(4, dtype=th.int32, device=dev)
    L = _lib.lib()
    saved = {k: os.environ.get(k) for k in ("ICRL_PPO_WIDE", "ICRL_PPO_WIDE_CLUSTERS")}
    if wide:
        os.environ["ICRL_PPO_WIDE"], os.environ["ICRL_PPO_WIDE_CLUSTERS"] = "1", "6"
    timed_out = False
    try:
        for launch in range(2):                               # two launches: flags / Adam state carry over
            _lib.check(L.icrl_ppo_local_advsums(C.byref(cfg), C.byref(data), _lib.ptr(advsums), _lib.current_stream()))
            comm.all_reduce_sum(advsums)
            d = comm.descriptor(advsums)
            _lib.check(L.icrl_ppo_train_dist(C.byref(cfg), C.byref(data), _lib.ptr(pol._params), _lib.ptr(pol._adam_m),
                                             _lib.ptr(pol._adam_v), pol.optimizer.step_count, _lib.ptr(stats), _lib.ptr(result),
                                             C.byref(d), _lib.current_stream()))
            th.cuda.synchronize()
            res = result.cpu().numpy()
            timed_out |= bool(res[2] != 0)
            assert res[1] == n_epochs * steps_per_epoch, res
            pol.optimizer.step_count += int(res[1])
            comm.advance(int(res[1]))
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    identical = _bit_identical(pol.parameters_flat().clone(), world)
    mode = {"0": "auto", "1": "direct", "2": "rsag", "3": "bcast"}.get(os.environ.get("ICRL_PPO_DIST_MODE", "0"), "auto")
    scheme = "wide(6 clusters)+slice exchange" if wide else (
        mode if mode != "auto" else ("bcast" if world <= 4 else "rsag") + str(world))
    out = {"kernel": "K4", "world": world, "shape": shape, "scheme": scheme, "replicas_identical": bool(identical),
           "peer_timeout": bool(timed_out), "max_param_err": None, "stats_ok": None}
    if rank == 0:       # per-step stats are global already (the loss sums ride along with the gradient all-reduce)
        th.set_num_threads(1)
        flat = {k: ogae.env_major(v) for k, v in g.items()}
        P = {k: v.clone() for k, v in P0.items()}
        adam = ocn.adam_init(list(P.values()))
        for launch in range(2):
            perms = [np.concatenate(global_minibatch_rows(perm_local[e], T, E_local, world, B_local)) for e in range(n_epochs)]
            ref = oppo.train(P, adam, flat, perms, is_discrete=False, batch_size=B_local * world, n_epochs=n_epochs, lr=3e-4,
                             clip_range=0.2, nu=0.7, ent_coef=0.01)
        got = pol.state_dict()
        worst = 0.0
        for k in P:
            a, b = got[k].numpy(), P[k].numpy()
            worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-3)))
        st = stats.cpu().numpy()
        stats_ok = True
        for i, key in ((0, "pg_loss"), (2, "reward_value_loss"), (3, "cost_value_loss"), (5, "approx_kl")):
            want = np.array(ref["per_step"][key])
            stats_ok &= bool(np.allclose(st[:, i], want, rtol=2e-4, atol=2e-6))
        out.update(max_param_err=worst, stats_ok=stats_ok)
    flag = th.tensor([0.0 if rank != 0 else float(out["max_param_err"] <= 1e-4 and out["stats_ok"])], device=dev)
    dist.broadcast(flag, 0)
    out["ok"] = bool(flag.item() == 1.0) and identical and not timed_out
    if own_comm:
        comm.close()
    return out


def cn_parity(rank, world, dev, shape="hc", per_step=True):
    """K2 sharded over `world` ranks vs oracle/cn.py::train on the global problem."""
    from icrl_b200.constraint_net import ConstraintNet
    from oracle import cn as ocn
    from oracle.cn_dp import shard_episodes
    obs_dim, acs_dim, hidden = (18, 6, (20,)) if shape == "hc" else (113, 8, (40, 40))
    rng = np.random.default_rng(11)
    lengths = [int(x) for x in rng.integers(40, 90, size=2 * world + 1)]
    n_nom, n_exp, iters, lr = int(np.sum(lengths)), 37 * world + 5, 4, 0.004
    scale = rng.uniform(0.5, 4.0, size=obs_dim)
    no = rng.standard_normal((n_nom, obs_dim)) * scale + 0.3
    na = rng.standard_normal((n_nom, acs_dim)).astype(np.float32)
    eo = rng.standard_normal((n_exp, obs_dim)) * scale
    ea = rng.standard_normal((n_exp, acs_dim)).astype(np.float32)
    low, high = -np.ones(acs_dim, np.float32), np.ones(acs_dim, np.float32)
    th.manual_seed(5)                                     # replicated initial parameters
    cn = ConstraintNet(obs_dim, acs_dim, hidden, None, lambda _: lr, eo, ea, False, 0.5,
                       per_step_importance_sampling=per_step, clip_obs=20., action_low=low, action_high=high,
                       target_kl_old_new=10, target_kl_new_old=10, device=dev)
    p0 = [p.clone() for p in cn.network.state_dict().values()]
    cn.enable_data_parallel(max_episodes=1024)
    ep_lo, ep_hi, r_lo, r_hi = shard_episodes(lengths, world)[rank]
    metrics = None
    for call in range(2):                                 # Adam state and exchange sequence numbers carry over
        metrics = cn.train(iters, no[r_lo:r_hi], na[r_lo:r_hi], np.array(lengths[ep_lo:ep_hi]))
    identical = _bit_identical(cn.parameters_flat().clone(), world)
    out = {"kernel": "K2", "world": world, "shape": shape, "is_mode": "per-step" if per_step else "per-episode",
           "replicas_identical": bool(identical), "max_param_err": None, "metrics_ok": None}
    if rank == 0:
        th.set_num_threads(1)
        spec = ocn.CNSpec(obs_dim, acs_dim, hidden, False, clip_obs=20., action_low=low, action_high=high,
                          regularizer_coeff=0.5, per_step_importance_sampling=per_step, target_kl_old_new=10,
                          target_kl_new_old=10)
        adam = ocn.adam_init(p0)
        for call in range(2):
            ref = ocn.train(p0, adam, spec, iters, no, na, lengths, eo, ea, lr=lr)
        worst = 0.0
        for a, b in zip(cn.network.state_dict().values(), p0):
            a, b = a.numpy(), b.numpy()
            worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-3)))
        ok = True
        for k, v in ref.items():
            gv = metrics[k]
            ok &= bool((np.isnan(v) and np.isnan(gv)) or abs(gv - v) <= 2e-4 * max(abs(v), 1e-2))
        out.update(max_param_err=worst, metrics_ok=ok)
    flag = th.tensor([0.0 if rank != 0 else float(out["max_param_err"] <= 1e-4 and out["metrics_ok"])], device=dev)
    dist.broadcast(flag, 0)
    out["ok"] = bool(flag.item() == 1.0) and identical
    cn.comm.close()
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    what = os.environ.get("ICRL_DP_TEST", "ppo")
    shape = os.environ.get("ICRL_DP_TEST_SHAPE", "hc")
    if what == "ppo":
        out = ppo_parity(rank, world, dev, shape)
    elif what == "ppo_wide":
        out = ppo_parity(rank, world, dev, shape, wide=True)
    elif what == "cn":
        out = cn_parity(rank, world, dev, shape, per_step=True)
    elif what == "cn_episode":
        out = cn_parity(rank, world, dev, shape, per_step=False)
    else:
        raise SystemExit(f"unknown ICRL_DP_TEST={what}")
    if rank == 0:
        print(out)
    assert out["ok"], out
    if rank == 0:
        print(f"dp parity ok: {what} world={world} max param err {out['max_param_err']:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
