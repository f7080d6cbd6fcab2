"""Multi-GPU plumbing for the N-diverse-futures rollout: one process per GPU (torchrun), rollouts sharded
across ranks, weights replicated, NO data-path collective during the rollout.  The only exchange is the
best-of-N selection at the end (generate_frames.py:185-190 picks ``argsort(mean ssim)[-1]`` per sequence):
every rank scores its own rollouts, ONE all-gather of the score matrix, a global arg-best per sequence, and
only the winners' frames travel (never the full frame tensor -- SURVEY 8e).

Pure host logic on torch tensors: works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rollouts(n_rollouts: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of the S rollouts: returns (first, count) for ``rank``.
    The first ``S % world_size`` ranks get one extra rollout."""
    base, extra = divmod(n_rollouts, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_scores(local_scores: torch.Tensor, n_rollouts: int, group=None) -> torch.Tensor:
    """All-gather per-rollout scores.  ``local_scores`` [S_local, B] (this rank's shard, any S_local given by
    ``shard_rollouts``) -> [S, B] in global rollout order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_scores
    world = dist.get_world_size(group)
    B = local_scores.shape[1]
    max_local = -(-n_rollouts // world)
    pad = torch.full((max_local, B), float("nan"), dtype=local_scores.dtype, device=local_scores.device)
    pad[: local_scores.shape[0]] = local_scores
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        _, cnt = shard_rollouts(n_rollouts, world, r)
        parts.append(out[r][:cnt])
    return torch.cat(parts, 0)


def select_best(scores: torch.Tensor, higher_is_better: bool = True) -> torch.Tensor:
    """Global best rollout per batch sequence: scores [S, B] -> int64 [B] (ties -> lowest rollout index,
    NaN never wins)."""
    s = torch.nan_to_num(scores, nan=float("-inf") if higher_is_better else float("inf"))
    return s.argmax(0) if higher_is_better else s.argmin(0)


def owner_of(rollout: int, n_rollouts: int, world_size: int) -> Tuple[int, int]:
    """(rank, local index) that holds global rollout ``rollout``."""
    for r in range(world_size):
        first, cnt = shard_rollouts(n_rollouts, world_size, r)
        if first <= rollout < first + cnt:
            return r, rollout - first
    raise IndexError(rollout)


def gather_winners(local_frames: torch.Tensor, best: torch.Tensor, n_rollouts: int, group=None) -> torch.Tensor:
    """Collect the winning rollout of every sequence on all ranks.
    ``local_frames`` [S_local, B, ...] ; ``best`` [B] global rollout ids -> [B, ...].
    Each rank contributes only the (rollout, sequence) pairs it owns; one all-reduce(sum) of a [B, ...]
    buffer (B * frame bytes, independent of S)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    first, cnt = shard_rollouts(n_rollouts, world, rank)
    B = best.shape[0]
    # static-shape masked gather (no nonzero / host sync: usable inside CUDA-graph capture)
    mine = (best >= first) & (best < first + cnt)
    local = (best - first).clamp(0, max(cnt - 1, 0))
    cols = torch.arange(B, device=best.device)
    if cnt == 0:          # more ranks than rollouts: this rank owns nothing and contributes zeros
        out = torch.zeros((B,) + tuple(local_frames.shape[2:]), dtype=local_frames.dtype, device=local_frames.device)
        if world > 1:
            dist.all_reduce(out, group=group)
        return out
    out = local_frames[local, cols]
    out = torch.where(mine.reshape((B,) + (1,) * (out.dim() - 1)), out, torch.zeros((), dtype=out.dtype, device=out.device))
    if world > 1:
        dist.all_reduce(out, group=group)
    return out


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that pinned host buffers allocated
    afterwards are node-local (first touch) and the launch thread does not cross sockets.  With 8 ranks streaming
    tens of MB per rollout each, buffers that all sit on one node made the host memory system the limiter of the
    end-to-end numbers (round 1: e2e scaling 0.68 at 8 GPUs).  Best effort: returns what it did."""
    import os
    info = {"local_rank": local_rank}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info.update(pci=bdf, numa_node=node)
        if node < 0:
            return info
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"
    except Exception as e:                                        # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"[:200]
    return info
