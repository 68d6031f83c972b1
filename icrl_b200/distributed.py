"""Data-parallel plumbing for the learner hot path across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink for the two tiny host-visible collectives).  What shards:
  * K1 (cost relabel) and K3 (dual GAE): by environment column -- no collective at all;
  * K4 (PPO-Lagrangian): every rank trains on `batch_size` rows of ITS OWN environments per optimiser step; the
    three trunks' gradients are all-reduced INSIDE the persistent kernel through CUDA-IPC mapped peer buffers
    (`PpoComm`), the global-minibatch advantage statistics through one small NCCL all-reduce per train() call;
  * the dual step: mean cost over all ranks' environments (one scalar all-reduce);
  * K2 (constraint-net train): replicated -- the nominal / expert batches are identical on every rank and the call is
    < 1 % of an iteration.
Parameters are replicated; identical reduced gradients (summed in rank order) keep them bit-identical.
"""
import ctypes as C
from typing import List, Tuple

import numpy as np
import torch as th

from . import _lib


def shard_envs(n_envs_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the environment columns owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_envs_total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_minibatch_rows(perm_local: List[np.ndarray], T: int, E_local: int, world: int, batch_local: int):
    """The global-buffer rows (env-major numbering over all world*E_local environments) that make up every global
    minibatch when rank r contributes minibatch m of its local permutation perm_local[r].  Used by tests to build the
    single-process equivalent of a data-parallel run."""
    n_local = T * E_local
    steps = (n_local + batch_local - 1) // batch_local
    out = []
    for m in range(steps):
        rows = []
        for r in range(world):
            loc = perm_local[r][m * batch_local:(m + 1) * batch_local]
            e_loc, t = loc // T, loc % T
            rows.append((r * E_local + e_loc) * T + t)
        out.append(np.concatenate(rows))
    return out


class PpoComm:
    """Peer receive buffers + flags for the in-kernel gradient all-reduce of `icrl_ppo_train_dist`."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise NotImplementedError("the fused all-reduce covers one NVSwitch node (<= 8 GPUs)")
        L = _lib.lib()
        self._local, handles = [], []
        for nbytes in (_lib.PPO_RECV_BYTES, _lib.PPO_FLAG_BYTES):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(L.icrl_comm_alloc(nbytes, C.byref(ptr), handle))
            self._local.append(ptr.value)
            handles.append(handle.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self.recv, self.flags, self._opened = [0] * self.world, [0] * self.world, []
        for r, (h_recv, h_flag) in enumerate(gathered):
            if r == self.rank:
                self.recv[r], self.flags[r] = self._local
                continue
            for i, h in enumerate((h_recv, h_flag)):
                ptr = C.c_void_p()
                _lib.check(L.icrl_comm_open(h, C.byref(ptr)))
                self._opened.append(ptr.value)
                (self.recv if i == 0 else self.flags)[r] = ptr.value
        self.flag_base = 0
        dist.barrier(group=group)

    def descriptor(self, advsums: th.Tensor) -> _lib.PpoDist:
        d = _lib.PpoDist()
        d.rank, d.world = self.rank, self.world
        for r in range(self.world):
            d.recv[r], d.flags[r] = self.recv[r], self.flags[r]
        d.flag_base = self.flag_base & 0xFFFFFFFF
        d.advsums = advsums.data_ptr()
        return d

    def advance(self, steps: int):
        """Every rank calls this after each launch with the same value: keeps the step flags monotonic across launches."""
        self.flag_base = (self.flag_base + steps + 1) & 0x7FFFFFFF

    def all_reduce_sum(self, t: th.Tensor) -> th.Tensor:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def close(self):
        L = _lib.lib()
        for p in self._opened:
            L.icrl_comm_close(C.c_void_p(p))
        for p in self._local:
            L.icrl_comm_free(C.c_void_p(p))
        self._opened, self._local = [], []
