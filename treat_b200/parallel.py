"""Multi-GPU plumbing: reads shard across ranks with no data-path collective.

One process per GPU (torchrun); every rank aligns a contiguous range of the batch against the
same template; the only cross-GPU traffic is one all-reduce of the uint64 summary counters.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(N: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of reads owned by `rank`: keeps output order trivial."""
    return (N * rank) // world, (N * (rank + 1)) // world


def fasta_shard(data, rank: int, world: int):
    """The slice of a FASTA file's bytes owned by `rank`: cut at record starts (treat_b200_fasta_split), so every rank
    parses whole records and the ranks' records, in rank order, are the file's records in order."""
    from .api import fasta_split
    if isinstance(data, (bytes, bytearray)):
        data = np.frombuffer(data, dtype=np.uint8)
    off = fasta_split(data, world)
    return data[int(off[rank]):int(off[rank + 1])], int(off[rank])


class _DevMem:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr: int, n_u64: int):
        self.__cuda_array_interface__ = {"shape": (n_u64,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def summary_tensor(engine):
    """torch int64 view (no copy) of the context's device-resident summary counters."""
    import torch
    ptr, n = engine.summary_device_ptr()
    return torch.as_tensor(_DevMem(ptr, n), device=f"cuda:{engine.device}")


def heatmap_tensor(engine):
    """torch int64 view (no copy) of the device-resident ESS x junction-length heat map, flattened [dim * dim]."""
    import torch
    ptr, n = engine.heatmap_device_ptr()
    return torch.as_tensor(_DevMem(ptr, n), device=f"cuda:{engine.device}")


def allreduce_summary(t, group=None):
    """Sum the counters over all ranks.  Exact integers: reduction order cannot change the result.
    `t` is an int64 tensor (device tensor -> NCCL over NVLink; CPU tensor -> gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def reduced_summary(engine, group=None) -> np.ndarray:
    """All-GPU summary as uint64 numpy (engine's own counters are left untouched)."""
    engine.synchronize()
    t = summary_tensor(engine).clone()
    allreduce_summary(t, group)
    return t.cpu().numpy().view(np.uint64)
