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


# ---- digests of a batch's results: the same bytes whatever the number of GPUs (bench.py's parity leg, tests) -------------
_CANON_FIELDS = ("edit_stop", "junc_start", "junc_end", "junc_len", "read_count", "has_mutation", "alt_editing",
                 "mismatches", "indel", "score", "junc_seq_off", "junc_seq_len", "aln_len", "n_bases", "status")


def canonical_records(rec: np.ndarray) -> np.ndarray:
    """The per-read record fields the reference defines (Alignment: align.go:41-55, score, JuncSeq window, lengths) as a
    packed array: the JuncSeq offset only where a JuncSeq exists, of the status byte only the reference-panic bit."""
    from .cabi import RECORD_DTYPE
    out = np.zeros(len(rec), dtype=RECORD_DTYPE)
    for f in _CANON_FIELDS:
        out[f] = rec[f]
    out["junc_seq_off"] = np.where(rec["junc_seq_len"] > 0, rec["junc_seq_off"], 0)
    out["status"] = rec["status"] & 1
    return out


def compact_ops(rec: np.ndarray, ops: np.ndarray) -> np.ndarray:
    """The bytes of every read's ops row that hold alignment columns (ceil(aln_len / 4) per read), back to back: the
    traceback of the batch independent of the row stride."""
    nbytes = (rec["aln_len"].astype(np.int64) + 3) // 4
    keep = np.arange(ops.shape[1], dtype=np.int64)[None, :] < nbytes[:, None]
    return ops[keep]


def sha256_hex(*chunks) -> str:
    import hashlib
    h = hashlib.sha256()
    for c in chunks:
        h.update(memoryview(np.ascontiguousarray(c)).cast("B"))
    return h.hexdigest()
