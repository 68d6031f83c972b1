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
