"""GPU (>= 2 devices): the data-parallel paths against the single-process oracle -- K4 with the in-kernel NVLink gradient
all-reduce (every exchange scheme, single-cluster and many-cluster kernels) and the sharded K2."""
import os
import subprocess
import sys

import pytest
import torch as th

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(world, what, mode="auto", shape="hc"):
    if th.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, ICRL_PPO_DIST_MODE={"auto": "0", "direct": "1", "rsag": "2", "bcast": "3"}[mode],
               ICRL_DP_TEST_SHAPE=shape, ICRL_DP_TEST=what)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp parity ok" in r.stdout


@pytest.mark.parametrize("world,mode,shape", [(2, "auto", "hc"), (2, "auto", "ant"), (2, "direct", "hc"), (2, "rsag", "hc"),
                                              (2, "direct", "ant"), (2, "rsag", "ant"), (4, "auto", "hc"), (4, "rsag", "hc"),
                                              (8, "auto", "hc"), (8, "auto", "ant")])
def test_data_parallel_ppo_matches_oracle(world, mode, shape):
    """mode: the in-kernel exchange -- "auto" (2 / 4 ranks: one-hop tagged broadcast + sum; 8 ranks: reduce-scatter +
    all-gather), "direct" ({value, seq} words, any world size) or "rsag"; ICRL_PPO_DIST_MODE forces the latter two so that
    all three run on a 2-GPU box."""
    _run(world, "ppo", mode, shape)


@pytest.mark.parametrize("world,shape", [(2, "hc"), (2, "ant"), (4, "hc"), (8, "ant")])
def test_data_parallel_wide_ppo_matches_oracle(world, shape):
    """The many-cluster kernel (6 clusters forced, some without rows) with the per-reducer slice exchange between ranks."""
    _run(world, "ppo_wide", "auto", shape)


@pytest.mark.parametrize("world,what,shape", [(2, "cn", "hc"), (2, "cn_episode", "hc"), (2, "cn", "ant"), (4, "cn", "hc"),
                                              (8, "cn_episode", "ant")])
def test_sharded_constraint_net_matches_oracle(world, what, shape):
    """K2: nominal episodes sharded by whole episodes, expert rows evenly; three in-kernel exchanges per backward iteration."""
    _run(world, what, "auto", shape)
