"""Batch sharding of independent utterances across ranks: one process per GPU, no data-path collective.

The generator forward never mixes batch elements, so multi-GPU inference is a partition of the utterance list.
``shard_slice`` gives rank r a contiguous block (sizes differ by at most one); ``gather_batch`` reassembles the
per-rank outputs in the original order (one all_gather of the padded shards; only needed when a single rank must
hold all audio - the benchmark does not gather).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of rank's contiguous block; the first (n_items % world) ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_slice(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    s, e = shard_bounds(x.shape[0], rank, world)
    return x[s:e]


def gather_batch(y_local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All ranks receive the full [n_items, ...] tensor, utterances in their original order."""
    world = dist.get_world_size(group)
    max_n = -(-n_items // world)
    pad = torch.zeros((max_n,) + tuple(y_local.shape[1:]), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, p in enumerate(parts):
        s, e = shard_bounds(n_items, r, world)
        out.append(p[: e - s])
    return torch.cat(out, dim=0)


def sharded_forward(fn: Callable[[torch.Tensor], torch.Tensor], x_full: torch.Tensor, gather: bool = True,
                    group=None) -> Optional[torch.Tensor]:
    """Run ``fn`` on this rank's utterances of ``x_full``; optionally reassemble the full output everywhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s, e = shard_bounds(x_full.shape[0], rank, world)
    y_local = fn(x_full[s:e]) if e > s else None
    if not gather:
        return y_local
    probe = torch.zeros(8, dtype=torch.int64, device=x_full.device)
    if y_local is not None:  # ranks without utterances learn the trailing output shape from the others
        probe[0] = y_local.ndim
        for i, d in enumerate(y_local.shape[1:]):
            probe[1 + i] = d
    shapes = [torch.empty_like(probe) for _ in range(world)]
    dist.all_gather(shapes, probe, group=group)
    ref = next(sh for sh in shapes if int(sh[0]) > 0)
    trailing = tuple(int(v) for v in ref[1:int(ref[0])])
    if y_local is None:
        y_local = torch.zeros((0,) + trailing, dtype=torch.float32, device=x_full.device)
    return gather_batch(y_local, x_full.shape[0], group)
