"""Data-parallel plumbing for the GAC step: one process per GPU, the replay
batch sharded by state, parameters replicated, ONE all-reduce(SUM) of the flat gradient arena
(+ the [loss, sum_adv, M] stats tail) per train step.  Critic/value losses are sums over the
batch (F5) and the actor gradient is exchanged un-normalised, so SUM -- never mean -- is exact;
every rank then applies the same 1/(M_g * A * sum_adv_g) scale, Adam and Polyak."""
from __future__ import annotations

from typing import Dict

import torch
import torch.distributed as dist


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_grads(flat: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks of the gradient arena + stats tail (NCCL on GPUs, gloo on CPU)."""
    if world()[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def broadcast_state(arena, src: int = 0):
    """Make every replica start from rank `src`'s parameters, targets, Adam moments and step counters.  Replicas only
    stay identical if they start identical: the train step exchanges gradients, never parameters."""
    if world()[1] > 1:
        for t in (arena.online, arena.target, arena.adam_m, arena.adam_v, arena.adam_t):
            dist.broadcast(t, src=src)


def shard_rows(n: int, rank: int, world_size: int) -> slice:
    """Contiguous shard of `n` replay states for `rank` (n must divide evenly)."""
    if n % world_size:
        raise ValueError(f"global batch {n} is not divisible by world size {world_size}")
    per = n // world_size
    return slice(rank * per, (rank + 1) * per)


def shard_step_inputs(batch: Dict, noise: Dict, K: int, rank: int, world_size: int):
    """Shard one step's global replay batch and injected RNG buffers by state index b.

    Global layouts (B states, P = B*K rows): value_tau is state-major (row = b*K + k,
    networks.py:464-467); filter_* are sample-major (row = k*B + b, agent.py:186); actor_tau is
    indexed by compacted row, which is rank-local, so each rank takes a disjoint 2*P_local block.
    Works on numpy arrays and torch tensors."""
    B = batch["s"].shape[0]
    sl = shard_rows(B, rank, world_size)
    Bl = sl.stop - sl.start
    b = {k: v[sl] for k, v in batch.items()}
    A = noise["value_tau"].shape[1]
    n = {"critic_u": noise["critic_u"][sl],
         "value_tau": noise["value_tau"].reshape(B, K, A)[sl].reshape(Bl * K, A)}
    for k in ("filter_tau", "filter_normal", "filter_rand"):
        n[k] = noise[k].reshape(K, B, A)[:, sl].reshape(K * Bl, A)
    n["actor_tau"] = noise["actor_tau"][rank * 2 * Bl * K:(rank + 1) * 2 * Bl * K]
    return b, n


def local_to_global_candidate(idx_local, B_local: int, K: int, rank: int, world_size: int):
    """Map a rank-local candidate index r (over [K*B_l target rows ; K*B_l random rows]) to its
    index in the single-process candidate order of agent.py:198-205 (SURVEY.md §8e index parity)."""
    P_l = B_local * K
    B = B_local * world_size
    blk = idx_local // P_l
    rem = idx_local % P_l
    k, b = rem // B_local, rem % B_local
    return blk * (B * K) + k * B + (rank * B_local + b)
