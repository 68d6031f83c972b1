"""GPU (>= 2 devices): data-parallel K4 with the in-kernel NVLink gradient all-reduce vs the single-process oracle."""
import os
import subprocess
import sys

import pytest
import torch as th

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_data_parallel_ppo_matches_oracle(world):
    if th.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp parity ok" in r.stdout
