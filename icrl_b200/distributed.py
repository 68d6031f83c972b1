"""Data-parallel plumbing for the learner hot path across the GPUs of one node.

One process per GPU (torch.distributed, NCCL over NVLink for the two tiny host-visible collectives).  What shards:
  * K1 (cost relabel) and K3 (dual GAE): by environment column -- no collective at all;
  * K4 (PPO-Lagrangian): every rank trains on `batch_size` rows of ITS OWN environments per optimiser step; the
    three trunks' gradients are all-reduced INSIDE the persistent kernel through CUDA-IPC mapped peer buffers
    (`PpoComm`), the global-minibatch advantage statistics through one small NCCL all-reduce per train() call;
  * the dual step: mean cost over all ranks' environments (one scalar all-reduce);
  * K2 (constraint-net train): nominal rows sharded by whole episodes (every rank samples its own), expert rows
    evenly; the three reductions of a backward iteration are exchanged inside the kernels over peer memory (`CnComm`).
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


class _PeerBuffers:
    """Device buffers of the given sizes on every rank, each mapped into every other rank through CUDA IPC handles
    (exchanged over torch.distributed).  `ptrs[i][r]` is the address of rank r's i-th buffer in THIS process."""

    def __init__(self, sizes, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise NotImplementedError("the peer-memory exchanges cover one NVSwitch node (<= 8 GPUs)")
        L = _lib.lib()
        self._local, handles = [], []
        for nbytes in sizes:
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(L.icrl_comm_alloc(int(nbytes), C.byref(ptr), handle))
            self._local.append(ptr.value)
            handles.append(handle.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        self.ptrs, self._opened = [[0] * self.world for _ in sizes], []
        for r, hs in enumerate(gathered):
            for i, h in enumerate(hs):
                if r == self.rank:
                    self.ptrs[i][r] = self._local[i]
                    continue
                ptr = C.c_void_p()
                _lib.check(L.icrl_comm_open(h, C.byref(ptr)))
                self._opened.append(ptr.value)
                self.ptrs[i][r] = ptr.value
        dist.barrier(group=group)

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


class PpoComm(_PeerBuffers):
    """Peer receive buffers + flags for the in-kernel gradient all-reduce of `icrl_ppo_train_dist`."""

    def __init__(self, group=None):
        super().__init__((_lib.PPO_RECV_BYTES, _lib.PPO_FLAG_BYTES), group)
        self.recv, self.flags = self.ptrs
        self.flag_base = 0

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


class CnComm(_PeerBuffers):
    """Exchange buffers of the data-parallel constraint-net update (`icrl_cn_train_dist`): every rank trains on its own
    nominal episodes and its slice of the expert batch; three small in-kernel exchanges per backward iteration."""

    def __init__(self, desc: _lib.CnDesc, max_episodes: int = 65536, group=None):
        nbytes = int(_lib.lib().icrl_cn_dist_bytes(C.byref(desc), int(max_episodes)))
        if nbytes <= 0:
            raise _lib.IcrlError("icrl_cn_dist_bytes rejected the constraint-net description")
        super().__init__((nbytes,), group)
        self.buffer_bytes, self.max_episodes, self.seq_base = nbytes, int(max_episodes), 0

    def shapes(self, n_nominal: int, n_expert: int, n_episodes: int, device):
        """(global nominal rows, global expert rows, global episodes, index of this rank's first episode)."""
        mine = th.tensor([n_nominal, n_expert, n_episodes], dtype=th.int64, device=device)
        allc = [th.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allc, mine, group=self.group)
        allc = th.stack(allc).cpu().numpy()
        tot = allc.sum(0)
        if int(tot[2]) > self.max_episodes:
            raise _lib.IcrlError(f"{int(tot[2])} nominal episodes over all ranks > max_episodes={self.max_episodes}")
        return int(tot[0]), int(tot[1]), int(tot[2]), int(allc[:self.rank, 2].sum())

    def descriptor(self, n_nominal_global, n_expert_global, n_episodes_global, episode_base) -> _lib.CnDist:
        d = _lib.CnDist()
        d.rank, d.world = self.rank, self.world
        for r in range(self.world):
            d.recv[r] = self.ptrs[0][r]
        d.buffer_bytes, d.seq_base = self.buffer_bytes, self.seq_base & 0xFFFFFFFF
        d.n_nominal_global, d.n_expert_global = int(n_nominal_global), int(n_expert_global)
        d.n_episodes_global, d.episode_base = int(n_episodes_global), int(episode_base)
        return d

    def advance(self, iterations: int):
        self.seq_base = (self.seq_base + iterations + 1) & 0x7FFFFFFF
