"""GPU (>= 2 devices): data-parallel K4 with the in-kernel NVLink gradient all-reduce vs the single-process oracle."""
import os
import subprocess
import sys

import pytest
import torch as th

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,mode,shape", [(2, "auto", "hc"), (2, "auto", "ant"), (2, "direct", "hc"), (2, "rsag", "hc"),
                                              (2, "direct", "ant"), (2, "rsag", "ant"), (4, "auto", "hc"), (4, "rsag", "hc"),
                                              (8, "auto", "hc")])
def test_data_parallel_ppo_matches_oracle(world, mode, shape):
    """mode: the in-kernel exchange -- "auto" (2 ranks: one-hop tagged broadcast + sum; 4 / 8 ranks: reduce-scatter +
    all-gather), "direct" ({value, seq} words, any world size) or "rsag"; ICRL_PPO_DIST_MODE forces the latter two so that
    all three run on a 2-GPU box."""
    if th.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, ICRL_PPO_DIST_MODE={"auto": "0", "direct": "1", "rsag": "2"}[mode], ICRL_DP_TEST_SHAPE=shape)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp parity ok" in r.stdout
