"""TEST INFRASTRUCTURE ONLY -- data-parallel restatement of the constraint-net update (K2), SURVEY §8 (e):
"shard nominal rows by whole episodes and expert rows evenly; one exchange per Adam step".

It spells out WHICH quantities cross ranks so that the sharded update equals the single-process one
(`oracle/cn.py::train`, i.e. icrl/constraint_net.py:137-256) -- the blueprint for sharding the CUDA K2 kernels:

  reduce 1 (importance sampling, constraint_net.py:231-256), all sums over ranks:
      [ sum ratio, sum prod, sum log(prod+eps), sum prod*log(prod+eps), n_episodes ]  (+ min / max of ratio, prod for the
      logged is_min / is_max)  ->  mean(ratio), normed_j, kl_old_new, kl_new_old  (identical on every rank, so the
      early-stop decision needs no further exchange)
  reduce 2 (one flat buffer): [ grad of the rank's (expert + regulariser + per-episode nominal) partial sums, already
      divided by the GLOBAL counts | per-step mode: grad of sum log(p_N + eps) / N, kept separate because its
      multiplier mean(w) is itself a global mean | sum w | the loss partial sums for the metrics ]
  then every rank applies the same Adam step to its replica.

`all_reduce` is injected: the in-process list version below simulates the ranks; tests/test_dp_host.py also runs it
over torch.distributed (gloo, world size 2)."""
from itertools import accumulate

import numpy as np
import torch as th

from . import cn as ocn


def shard_episodes(episode_lengths, world):
    """Contiguous blocks of whole episodes per rank, balanced by row count; returns [(ep_lo, ep_hi, row_lo, row_hi)]."""
    lengths = [int(l) for l in episode_lengths]
    cum = [0] + list(accumulate(lengths))
    total, out, ep = cum[-1], [], 0
    for r in range(world):
        target = total * (r + 1) / world
        hi = ep
        while hi < len(lengths) and (cum[hi + 1] <= target or hi == ep) and len(lengths) - hi > world - r - 1:
            hi += 1
        if r == world - 1:
            hi = len(lengths)
        out.append((ep, hi, cum[ep], cum[hi]))
        ep = hi
    return out


def train_rank(params, adam_state, spec, iterations, nominal_obs, nominal_acs, episode_lengths, expert_obs, expert_acs,
               lr, all_reduce_sum, all_reduce_min, all_reduce_max, adam_eps=1e-5):
    """One rank's view: ITS nominal episodes / expert rows, the replicated parameters, and the three collectives."""
    eps = spec.eps
    nominal = th.tensor(ocn.prepare_data(spec, nominal_obs, nominal_acs), dtype=th.float32)
    expert = th.tensor(ocn.prepare_data(spec, expert_obs, expert_acs), dtype=th.float32)
    counts = all_reduce_sum(th.tensor([nominal.shape[0], expert.shape[0], len(episode_lengths)], dtype=th.float64))
    N, Ne, M = (float(c) for c in counts)
    cum = [0] + list(accumulate(int(l) for l in episode_lengths))
    for p in params:
        p.requires_grad_(True)
    if spec.importance_sampling:
        with th.no_grad():
            start = ocn.forward(params, nominal)
    early_stop_itr, metrics = iterations, {}
    for itr in range(iterations):
        mean_w = 1.0
        if spec.importance_sampling:
            with th.no_grad():
                cur = ocn.forward(params, nominal)
                ratio = (cur + eps) / (start + eps)
                prod = th.stack([th.prod(ratio[cum[j]:cum[j + 1]]) for j in range(len(episode_lengths))]) \
                    if len(episode_lengths) else th.zeros(0)
                logp = th.log(prod + eps)
                s = all_reduce_sum(th.stack([ratio.sum().double(), prod.sum().double(), logp.sum().double(),
                                             (prod * logp).sum().double()]))
                mean_ratio, sum_prod = float(s[0]) / N, float(s[1])
                kl_old_new = -float(s[2]) / M
                prod_mean = sum_prod / M
                kl_new_old = (float(s[3]) - prod_mean * float(s[2])) / M / (prod_mean + eps)
                if spec.per_step_importance_sampling:
                    w = (ratio / mean_ratio).squeeze(-1)
                else:
                    normed = M * prod / (sum_prod + eps)
                    w = th.repeat_interleave(normed, th.as_tensor([int(l) for l in episode_lengths]))
            if ((spec.target_kl_old_new != -1 and kl_old_new > spec.target_kl_old_new) or
                    (spec.target_kl_new_old != -1 and kl_new_old > spec.target_kl_new_old)):
                early_stop_itr = itr
                break
        else:
            w = th.ones(nominal.shape[0])
        pn, pe = ocn.forward(params, nominal), ocn.forward(params, expert)
        log_n = th.log(pn + eps).squeeze(-1)
        # partial sums already divided by the GLOBAL counts, so that a plain sum over ranks gives the global loss terms
        expert_part = th.log(pe + eps).sum() / Ne
        reg_part = spec.regularizer_coeff * ((1 - pe).sum() / Ne + (1 - pn).sum() / N)
        per_step = spec.importance_sampling and spec.per_step_importance_sampling
        if per_step:
            nominal_unscaled = log_n.sum() / N              # multiplied by the global mean(w) after the exchange
            other = -expert_part + reg_part
        else:
            nominal_unscaled = th.zeros(())
            other = -expert_part + (w * log_n).sum() / N + reg_part
        g_other = th.autograd.grad(other, params, retain_graph=per_step)
        flat = [g.reshape(-1) for g in g_other]
        if per_step:
            flat += [g.reshape(-1) for g in th.autograd.grad(nominal_unscaled, params)]
        tail = th.stack([w.sum(), expert_part.detach(), reg_part.detach(), nominal_unscaled.detach(),
                         ((w * log_n).sum() / N).detach(), (log_n.sum() / N).detach()])
        buf = all_reduce_sum(th.cat(flat + [tail]).double()).float()
        P = sum(p.numel() for p in params)
        mean_w = float(buf[-6]) / N
        grads, off = [], 0
        for p in params:
            g = buf[off:off + p.numel()]
            if per_step:
                g = g + mean_w * buf[P + off:P + off + p.numel()]
            grads.append(g.reshape(p.shape))
            off += p.numel()
        ocn.adam_step(params, grads, adam_state, lr, eps=adam_eps)
        nominal_loss = mean_w * float(buf[-3]) if per_step else float(buf[-2])
        metrics = {"backward/expert_loss": float(buf[-5]), "backward/regularizer_loss": float(buf[-4]),
                   "backward/nominal_loss": nominal_loss, "backward/unweighted_nominal_loss": float(buf[-1]),
                   "backward/cn_loss": -float(buf[-5]) + nominal_loss + float(buf[-4]), "backward/is_mean": mean_w,
                   "backward/is_min": float(all_reduce_min(w.min().reshape(1))), "backward/is_max": float(all_reduce_max(w.max().reshape(1)))}
        if spec.importance_sampling:
            metrics.update({"backward/kl_old_new": kl_old_new, "backward/kl_new_old": kl_new_old})
    for p in params:
        p.requires_grad_(False)
    metrics["backward/early_stop_itr"] = early_stop_itr
    return metrics


def train_simulated(world, params, spec, iterations, nominal_obs, nominal_acs, episode_lengths, expert_obs, expert_acs, lr):
    """All ranks in one process, run in lock step with python threads; returns (per-rank params, per-rank metrics)."""
    import threading
    shards = shard_episodes(episode_lengths, world)
    ex_bounds = [len(expert_obs) * r // world for r in range(world + 1)]
    barrier = threading.Barrier(world)
    slots, results = [None] * world, [None] * world
    lock = threading.Lock()

    def make_collective(rank, op):
        def run(t):
            slots[rank] = t.clone()
            barrier.wait()
            stacked = th.stack(slots)
            out = stacked.sum(0) if op == "sum" else stacked.min(0).values if op == "min" else stacked.max(0).values
            barrier.wait()
            return out
        return run

    def worker(rank):
        e0, e1, r0, r1 = shards[rank]
        P = [p.clone() for p in params]
        adam = ocn.adam_init(P)
        m = train_rank(P, adam, spec, iterations, nominal_obs[r0:r1], nominal_acs[r0:r1], list(episode_lengths[e0:e1]),
                       expert_obs[ex_bounds[rank]:ex_bounds[rank + 1]], expert_acs[ex_bounds[rank]:ex_bounds[rank + 1]], lr,
                       make_collective(rank, "sum"), make_collective(rank, "min"), make_collective(rank, "max"))
        with lock:
            results[rank] = (P, m)

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    return [r[0] for r in results], [r[1] for r in results]
