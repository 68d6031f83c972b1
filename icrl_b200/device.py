"""Device selection.  The learner hot path exists only as sm_100a CUDA kernels: the reference's `--device cpu`
default (icrl/icrl.py:324) and `"auto"` both resolve to this process's CUDA device (cuda:LOCAL_RANK); there is no
CPU fallback, so a missing GPU is an error at first use, not a silent slow path."""
import os

import torch as th


def resolve_device(device="cuda") -> th.device:
    if isinstance(device, th.device):
        dev = device
    elif device in (None, "auto", "cpu", "cuda"):
        dev = th.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) if th.cuda.is_available() else 0)
    else:
        dev = th.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"icrl_b200 runs the learner on CUDA only (got device={device!r})")
    if not th.cuda.is_available():
        raise RuntimeError("icrl_b200: no CUDA device available -- the learner hot path has no CPU fallback")
    return dev
