"""CPU: host-side logic of the data-parallel path -- env sharding, the global-minibatch index mapping, and the
all-reduced advantage-statistics table (world_size-2 gloo processes)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from icrl_b200.distributed import global_minibatch_rows, shard_envs


def test_shard_envs_partitions_exactly():
    for total in (5, 8, 10, 40, 41):
        for world in (1, 2, 4, 8):
            spans = [shard_envs(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_global_minibatch_rows_cover_every_row_once():
    T, E_local, world, B = 16, 3, 2, 8
    perms = [np.random.RandomState(r).permutation(T * E_local) for r in range(world)]
    mbs = global_minibatch_rows(perms, T, E_local, world, B)
    allrows = np.concatenate(mbs)
    assert sorted(allrows.tolist()) == list(range(T * E_local * world))
    # rank r's rows of every minibatch live in rank r's env columns
    for mb in mbs:
        for r in range(world):
            part = mb[r * B:(r + 1) * B]
            assert np.all((part // T) // E_local == r)


def test_advsums_allreduce_gloo_world2(tmp_path):
    """Two gloo ranks: summing the per-rank (sum x, sum x^2, sum c, n) tables gives the global minibatch statistics the
    reference computes with .mean() / .std() on the concatenated minibatch (ppo_lag.py:218-222)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent('''
        import os, sys, numpy as np, torch as th, torch.distributed as dist
        sys.path.insert(0, %r)
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        rng = np.random.default_rng(0)
        adv = rng.standard_normal((world, 6, 32)).astype(np.float32) * 3 + 1     # [rank][step][row]
        cadv = rng.standard_normal((world, 6, 32)).astype(np.float32)
        x = adv[rank].astype(np.float64)
        tab = th.tensor(np.stack([x.sum(1), (x * x).sum(1), cadv[rank].astype(np.float64).sum(1), np.full(6, 32.0)], 1))
        dist.all_reduce(tab)
        sx, sxx, sc, n = tab.numpy().T
        mean = sx / n
        std = np.sqrt(np.maximum(sxx - sx * mean, 0) / (n - 1))
        g = th.tensor(adv).permute(1, 0, 2).reshape(6, -1)
        assert np.allclose(mean, g.mean(1).numpy(), rtol=1e-6)
        assert np.allclose(std, g.std(1).numpy(), rtol=1e-6)
        assert np.allclose(sc / n, th.tensor(cadv).permute(1, 0, 2).reshape(6, -1).mean(1).numpy(), rtol=1e-5, atol=1e-7)
        dist.destroy_process_group()
        print("ok", rank)
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ------------------------------------------------------------------------------------------- K2 sharding blueprint
import pytest  # noqa: E402


@pytest.mark.parametrize("per_step", [True, False])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_constraint_net_update_equals_single_process(world, per_step):
    """oracle/cn_dp.py (episodes sharded over ranks, two all-reduces per Adam step) reproduces oracle/cn.py::train."""
    import numpy as np
    import torch as th
    from oracle import cn as ocn, cn_dp
    rng = np.random.default_rng(5)
    lengths = [30, 41, 25, 50, 37, 44, 29]
    n = sum(lengths)
    nom_obs, nom_acs = rng.standard_normal((n, 18)) * 3, rng.standard_normal((n, 6)).astype(np.float32)
    exp_obs, exp_acs = rng.standard_normal((150, 18)) * 3, rng.standard_normal((150, 6)).astype(np.float32)
    spec = ocn.CNSpec(18, 6, (20,), False, clip_obs=20., regularizer_coeff=0.5, per_step_importance_sampling=per_step,
                      target_kl_old_new=10, target_kl_new_old=10)
    th.manual_seed(0)
    dims = [spec.input_dims, 20, 1]
    params = []
    for i in range(2):
        lin = th.nn.Linear(dims[i], dims[i + 1])
        params += [lin.weight.detach().clone(), lin.bias.detach().clone()]
    single = [p.clone() for p in params]
    m1 = ocn.train(single, ocn.adam_init(single), spec, 4, nom_obs, nom_acs, lengths, exp_obs, exp_acs, lr=0.01)
    reps, ms = cn_dp.train_simulated(world, params, spec, 4, nom_obs, nom_acs, lengths, exp_obs, exp_acs, lr=0.01)
    for r in range(world):
        for a, b in zip(reps[r], single):
            assert th.allclose(a, b, rtol=2e-5, atol=2e-7), float((a - b).abs().max())
        for a, b in zip(reps[r], reps[0]):
            assert th.equal(a, b)                       # replicas stay bit identical
        for k in ("backward/cn_loss", "backward/kl_old_new", "backward/kl_new_old", "backward/is_max", "backward/is_min"):
            assert abs(ms[r][k] - m1[k]) <= 1e-4 * max(1.0, abs(m1[k])), (k, ms[r][k], m1[k])
    covered = cn_dp.shard_episodes(lengths, world)
    assert covered[0][0] == 0 and covered[-1][1] == len(lengths) and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_sharded_constraint_net_update_gloo_world2(tmp_path):
    """The same blueprint over real collectives: two gloo processes, torch.distributed all_reduce (sum / min / max)."""
    script = tmp_path / "k2.py"
    script.write_text(textwrap.dedent('''
        import sys, numpy as np, torch as th, torch.distributed as dist
        sys.path.insert(0, %r)
        from oracle import cn as ocn, cn_dp
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        th.set_num_threads(1)
        rng = np.random.default_rng(5)
        lengths = [30, 41, 25, 50, 37, 44]
        n = sum(lengths)
        nom_obs, nom_acs = rng.standard_normal((n, 18)) * 3, rng.standard_normal((n, 6)).astype(np.float32)
        exp_obs, exp_acs = rng.standard_normal((120, 18)) * 3, rng.standard_normal((120, 6)).astype(np.float32)
        spec = ocn.CNSpec(18, 6, (20,), False, clip_obs=20., regularizer_coeff=0.5, per_step_importance_sampling=True,
                          target_kl_old_new=10, target_kl_new_old=10)
        th.manual_seed(0)
        params = []
        for a, b in ((spec.input_dims, 20), (20, 1)):
            lin = th.nn.Linear(a, b)
            params += [lin.weight.detach().clone(), lin.bias.detach().clone()]
        single = [p.clone() for p in params]
        ocn.train(single, ocn.adam_init(single), spec, 3, nom_obs, nom_acs, lengths, exp_obs, exp_acs, lr=0.01)

        def reduce(op):
            def run(t):
                t = t.clone()
                dist.all_reduce(t, op=op)
                return t
            return run
        e0, e1, r0, r1 = cn_dp.shard_episodes(lengths, world)[rank]
        x0, x1 = 120 * rank // world, 120 * (rank + 1) // world
        mine = [p.clone() for p in params]
        cn_dp.train_rank(mine, ocn.adam_init(mine), spec, 3, nom_obs[r0:r1], nom_acs[r0:r1], lengths[e0:e1], exp_obs[x0:x1],
                         exp_acs[x0:x1], 0.01, reduce(dist.ReduceOp.SUM), reduce(dist.ReduceOp.MIN), reduce(dist.ReduceOp.MAX))
        for a, b in zip(mine, single):
            assert th.allclose(a, b, rtol=2e-5, atol=2e-7), float((a - b).abs().max())
        dist.destroy_process_group()
        print("ok", rank)
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_bench_replica_digest_gloo_world2(tmp_path):
    """bench.py's post-run soak check: `replicas_identical` is true only if every rank holds bit-identical tensors (a one-ulp
    difference in one element on one rank, or two swapped elements, must be caught)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent('''
        import os, sys, torch as th, torch.distributed as dist
        sys.path.insert(0, %r)
        import bench
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        th.manual_seed(0)
        a, b = th.randn(1000), th.randn(37, 3)
        assert bench.replicas_identical([a, b], world)
        a2 = a.clone()
        if rank == 1:
            a2[123] = th.nextafter(a2[123], th.tensor(float("inf")))
        assert not bench.replicas_identical([a2, b], world)
        a3 = a.clone()
        if rank == 0:
            a3[[5, 6]] = a3[[6, 5]]
        assert not bench.replicas_identical([a3, b], world)
        dist.destroy_process_group()
        print("ok", rank)
    ''' % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
