"""Worker for the data-parallel K4 parity test (launched by tests/test_dp_gpu.py under torch.distributed.run, one
process per GPU).  Every rank trains on its own environment columns with the in-kernel NVLink all-reduce; rank 0 checks
the result against the CPU oracle run on the equivalent single-process problem (global batch = world * local batch)."""
import ctypes as C
import os
import sys

import numpy as np
import torch as th
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    th.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=th.device("cuda", local))
    from icrl_b200 import _lib
    from icrl_b200.distributed import PpoComm, global_minibatch_rows, shard_envs
    from icrl_b200.policies import ActorTwoCriticsPolicy
    from icrl_b200.spaces import Box
    from oracle import cn as ocn, gae as ogae, ppo as oppo

    D, A, T, E_local, B_local, n_epochs = 18, 6, 32, 2, 32, 3
    if os.environ.get("ICRL_DP_TEST_SHAPE") == "ant":      # the widest layer-1 tiling (NT1 = 8) through the exchange
        D, A = 113, 8
    E = E_local * world
    rng = np.random.default_rng(7)                       # identical global data on every rank
    g = {"observations": rng.standard_normal((T, E, D)).astype(np.float32),
         "actions": rng.standard_normal((T, E, A)).astype(np.float32)}
    for k in ("old_log_prob", "old_reward_values", "reward_advantages", "reward_returns", "old_cost_values",
              "cost_advantages", "cost_returns"):
        g[k] = rng.standard_normal((T, E)).astype(np.float32)
    g["old_log_prob"] = g["old_log_prob"] * 0.1 - 8.5
    lo, hi = shard_envs(E, rank, world)
    assert (lo, hi) == (rank * E_local, (rank + 1) * E_local)
    dev = th.device("cuda", local)
    th.manual_seed(3)                                     # replicated initial parameters
    pol = ActorTwoCriticsPolicy(Box(-np.inf, np.inf, (D,)), Box(-1, 1, (A,)), lambda _: 3e-4, device=dev)
    P0 = {k: v.clone() for k, v in pol.state_dict().items()}
    n_local = T * E_local
    perm_local = [[np.random.RandomState(100 * e + r).permutation(n_local) for r in range(world)] for e in range(n_epochs)]
    perm_dev = th.from_numpy(np.stack([perm_local[e][rank] for e in range(n_epochs)]).astype(np.int32)).to(dev)
    loc = {k: th.from_numpy(np.ascontiguousarray(v[:, lo:hi])).to(dev) for k, v in g.items()}
    comm = PpoComm()
    steps_per_epoch = n_local // B_local
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
    for launch in range(2):                               # two launches: flags / Adam state carry over
        _lib.check(L.icrl_ppo_local_advsums(C.byref(cfg), C.byref(data), _lib.ptr(advsums), _lib.current_stream()))
        comm.all_reduce_sum(advsums)
        d = comm.descriptor(advsums)
        _lib.check(L.icrl_ppo_train_dist(C.byref(cfg), C.byref(data), _lib.ptr(pol._params), _lib.ptr(pol._adam_m),
                                         _lib.ptr(pol._adam_v), pol.optimizer.step_count, _lib.ptr(stats), _lib.ptr(result),
                                         C.byref(d), _lib.current_stream()))
        th.cuda.synchronize()
        res = result.cpu().numpy()
        assert res[2] == 0, "peer time-out"
        assert res[1] == n_epochs * steps_per_epoch
        pol.optimizer.step_count += int(res[1])
        comm.advance(int(res[1]))
    # replicated parameters must be bit-identical on every rank
    mine = pol.parameters_flat().clone()
    allp = [th.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    for r in range(world):
        assert th.equal(allp[r], allp[0]), f"rank {r} parameters drifted"
    # per-step stats are global already (the loss sums ride along with the gradient all-reduce)
    if rank == 0:
        th.set_num_threads(1)
        flat = {k: ogae.env_major(v) for k, v in g.items()}
        P = {k: v.clone() for k, v in P0.items()}
        adam = ocn.adam_init(list(P.values()))
        for launch in range(2):
            perms = [np.concatenate(global_minibatch_rows(perm_local[e], T, E_local, world, B_local)) for e in range(n_epochs)]
            out = oppo.train(P, adam, flat, perms, is_discrete=False, batch_size=B_local * world, n_epochs=n_epochs, lr=3e-4,
                             clip_range=0.2, nu=0.7, ent_coef=0.01)
        got = pol.state_dict()
        worst = 0.0
        for k in P:
            a, b = got[k].numpy(), P[k].numpy()
            worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-3)))
        assert worst <= 1e-4, f"data-parallel parameters differ from the oracle: {worst}"
        st = stats.cpu().numpy()
        for i, key in ((0, "pg_loss"), (2, "reward_value_loss"), (3, "cost_value_loss"), (5, "approx_kl")):
            want = np.array(out["per_step"][key])
            assert np.allclose(st[:, i], want, rtol=2e-4, atol=2e-6), (key, st[:3, i], want[:3])
        print(f"dp parity ok: world={world} max param err {worst:.2e}")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
