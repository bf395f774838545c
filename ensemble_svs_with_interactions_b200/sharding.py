"""Track-sharded multi-GPU inference: one process per GPU, independent (song, track) work items, NO data-path collective.

The reference synthesises sequentially (nnsvs/bin/synthesis_multitrack.py:113-118, one B=1 pair at a time); the mgc/bap
diffusion and the vocoder see one track each (multistream.py:1684-1696), so every (song, track) item is independent once
its conditioning exists.  This module only decides WHO does WHAT and reduces the timing:

* ``assign(lengths, world_size)``  longest-processing-time-first static partition (deterministic, identical on every rank);
* ``batches(items, lengths, max_frames)`` groups a rank's items into padded batches under a frame budget
  (the inference analogue of ``batch_by_size``, nnsvs/train_util.py:190-246);
* ``max_over_ranks(seconds)`` the only communication: a scalar MAX all-reduce for the benchmark clock.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import torch
import torch.distributed as dist


def assign(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Returns per-rank lists of item indices.  Greedy LPT: items sorted by length (desc, index as tie-break) go to the
    currently least-loaded rank; within a rank, items keep descending-length order (good for batching)."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + int(lengths[i]), r))
    return out


def batches(items: Sequence[int], lengths: Sequence[int], max_frames: int) -> List[List[int]]:
    """Group items (already sorted by descending length) so that len(batch) * longest <= max_frames."""
    out: List[List[int]] = []
    cur: List[int] = []
    longest = 0
    for i in items:
        L = int(lengths[i])
        new_longest = max(longest, L)
        if cur and (len(cur) + 1) * new_longest > max_frames:
            out.append(cur)
            cur, new_longest = [], L
        cur.append(i)
        longest = new_longest
    if cur:
        out.append(cur)
    return out


def max_over_ranks(seconds: float, device=None) -> float:
    """MAX all-reduce of a scalar (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
