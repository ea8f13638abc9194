"""N>1 path on CPU: world_size-2 gloo run of the host-side sharding + summary all-reduce logic
(treat_b200/parallel.py).  The per-read records come from the oracle here (no GPU in this
container); the code under test is the partitioning and the exact-integer reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from treat_b200 import parallel, synth
    from oracle import oracle as O
    from helpers import summary_from_records
    import tempfile

    t = synth.rps12_template()
    with tempfile.TemporaryDirectory() as d:
        om = O.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), O.FORWARD, "T")
    lo, hi = parallel.shard_range(N, rank, world)
    # counter-based generator: a shard is generated independently of the others
    seqs, off, rc = synth.reads(t, hi - lo, seed=synth.SEED_BASE + 9, first=lo)
    rec, _, _ = O.align_batch(om, seqs, off, rc, want_ops=False)
    local = summary_from_records(rec, t.m, 0, om.EditStop)
    red = torch.from_numpy(local.view(np.int64).copy())
    parallel.allreduce_summary(red)
    np.save(os.path.join(out_dir, "rec_%d.npy" % rank), rec)
    np.save(os.path.join(out_dir, "red_%d.npy" % rank), red.numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from treat_b200 import parallel
    for N in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_summary_reduce(tmp_path):
    from treat_b200 import synth
    from oracle import oracle as O
    from helpers import summary_from_records
    N, world = 3001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, N, str(tmp_path)), nprocs=world, join=True)
    # single-process ground truth over the whole batch
    t = synth.rps12_template()
    om = O.NewTemplateFromFasta(t.write_fasta(str(tmp_path / "t.fa")), O.FORWARD, "T")
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 9)
    rec, _, _ = O.align_batch(om, seqs, off, rc, want_ops=False)
    want = summary_from_records(rec, t.m, 0, om.EditStop)
    shards = [np.load(tmp_path / ("rec_%d.npy" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate(shards), rec)  # same per-read records, same order, any rank count
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("red_%d.npy" % r)), want)  # every rank holds the reduced summary


def _fasta_worker(rank, world, port, fasta_path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import treat_b200 as T
    from treat_b200 import parallel, synth
    from oracle import oracle as O
    from helpers import summary_from_records
    import tempfile

    t = synth.rps12_template()
    with tempfile.TemporaryDirectory() as d:
        om = O.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "t.fa")), O.FORWARD, "T")
    data = np.fromfile(fasta_path, dtype=np.uint8)
    mine, base = parallel.fasta_shard(data, rank, world)  # whole records, cut at record starts
    fb = T.parse_fasta(mine)
    rec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, want_ops=False)
    red = torch.from_numpy(summary_from_records(rec, t.m, 0, om.EditStop).view(np.int64).copy())
    parallel.allreduce_summary(red)
    np.save(os.path.join(out_dir, "frec_%d.npy" % rank), rec)
    np.save(os.path.join(out_dir, "fred_%d.npy" % rank), red.numpy().view(np.uint64))
    with open(os.path.join(out_dir, "fids_%d.txt" % rank), "w") as f:
        f.write("\n".join(fb.ids))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_fasta_sharding(tmp_path):
    """A FASTA file shared between ranks by treat_b200_fasta_split: the ranks' records, in rank order, are the file's."""
    import treat_b200 as T
    from treat_b200 import synth
    from oracle import oracle as O
    from helpers import summary_from_records
    N, world = 2001, 2
    t = synth.rps12_template()
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 11)
    text = b"".join(b">q%d_%d\n" % (i, rc[i]) + seqs[int(off[i]):int(off[i + 1])].tobytes() + b"\n" for i in range(N))
    path = tmp_path / "reads.fa"
    path.write_bytes(text)
    mp.spawn(_fasta_worker, args=(world, _free_port(), str(path), str(tmp_path)), nprocs=world, join=True)
    om = O.NewTemplateFromFasta(t.write_fasta(str(tmp_path / "t.fa")), O.FORWARD, "T")
    whole = T.read_fasta(str(path))
    assert len(whole) == N and np.array_equal(whole.read_counts, rc)
    rec, _, _ = O.align_batch(om, whole.seqs, whole.offsets, whole.read_counts, want_ops=False)
    want = summary_from_records(rec, t.m, 0, om.EditStop)
    shards = [np.load(tmp_path / ("frec_%d.npy" % r)) for r in range(world)]
    assert all(len(s) > 0 for s in shards) and np.array_equal(np.concatenate(shards), rec)
    ids = sum(((tmp_path / ("fids_%d.txt" % r)).read_text().split("\n") for r in range(world)), [])
    assert ids == whole.ids
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("fred_%d.npy" % r)), want)
