"""Request sharding across the GPUs of one box: one process per GPU, utterance i -> rank i mod N, no collective
on the data path; ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in CPU tests) only gathers ragged outputs
and timing counters afterwards.  The reference is single-process, batch 1 (generation.py:124,156)."""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_indices(n_requests: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_requests, world))


def gather_ragged(local: Sequence[torch.Tensor], n_requests: int, device: Optional[torch.device] = None
                  ) -> Optional[List[torch.Tensor]]:
    """Each rank holds the outputs of ``shard_indices(n_requests, rank, world)`` (1-D or (F, C) tensors of one dtype).
    Returns, on rank 0, the list of all ``n_requests`` outputs in request order; ``None`` elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = device if device is not None else (local[0].device if len(local) else torch.device("cpu"))
    per_rank = (n_requests + world - 1) // world
    trailing = tuple(local[0].shape[1:]) if len(local) else ()
    tinfo = torch.tensor([len(trailing)] + list(trailing) + [0] * (4 - len(trailing)), dtype=torch.int64, device=dev)
    dist.all_reduce(tinfo, op=dist.ReduceOp.MAX)
    trailing = tuple(int(x) for x in tinfo[1:1 + int(tinfo[0])])
    lens = torch.zeros((per_rank,), dtype=torch.int64, device=dev)
    for j, t in enumerate(local):
        lens[j] = t.shape[0]
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens)
    max_len = max(int(l.max()) for l in all_lens)
    dtype = local[0].dtype if len(local) else torch.float32
    pad = torch.zeros((per_rank, max_len) + trailing, dtype=dtype, device=dev)
    for j, t in enumerate(local):
        pad[j, : t.shape[0]] = t.to(dev)
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    out: List[torch.Tensor] = []
    for i in range(n_requests):
        r, j = i % world, i // world
        out.append(bufs[r][j, : int(all_lens[r][j])].clone())
    return out


def reduce_max(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
