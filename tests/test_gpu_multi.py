"""Several GPUs behind the C ABI (treat_b200_group_* / treat_b200_comm_*): the same batch at 1, 2, 4, 8 GPUs must give the
oracle's per-read records and ops rows (cmd/treat/align.go:114-120, storage.go:738-785) and identical reduced summaries
(cmd/treat/stats.go:140-187, handlers.go:203-215) on every device.  Run on the B200 box with -m gpu; the cases that need
more GPUs than the box has are skipped."""
import os
import socket
import sys

import numpy as np
import pytest

import treat_b200 as T
from treat_b200 import cabi, synth
from oracle import oracle as O
from helpers import compare_records, heatmap_from_records, summary_from_records

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _templates(t, tmp_path, offset=0):
    path = t.write_fasta(str(tmp_path / "tmpl.fa"))
    om = O.NewTemplateFromFasta(path, O.FORWARD, t.edit_base)
    om.SetOffset(offset)
    tm = T.NewTemplateFromFasta(path, T.FORWARD, t.edit_base)
    tm.SetOffset(offset)
    return om, tm


@pytest.fixture(scope="module")
def batch(tmp_path_factory):
    """70,003 RPS12 reads (not a multiple of any GPU count; several pipelined chunks per device) + the oracle's answers."""
    tmp = tmp_path_factory.mktemp("multi")
    t = synth.rps12_template()
    om, tm = _templates(t, tmp, offset=3)
    N = 70003
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 41, p_trunc=0.05)
    stride = ((t.m + int(np.diff(off.astype(np.int64)).max()) + 3) // 4 + 15) // 16 * 16
    orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=8, ops_stride=stride)
    return dict(t=t, om=om, tm=tm, N=N, seqs=seqs, off=off, rc=rc, orec=orec, oops=oops, stride=stride,
                summary=summary_from_records(orec, t.m, 3, om.EditStop), heat=heatmap_from_records(orec, t.m + 1, 3))


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_group_equals_oracle_at_every_gpu_count(batch, n):
    if _gpus() < n:
        pytest.skip("needs %d GPUs" % n)
    b = batch
    g = T.DeviceGroup(n=n)
    try:
        g.SetTemplate(b["tm"])
        assert g.ops_stride(int(np.diff(b["off"].astype(np.int64)).max())) == b["stride"]
        for flags in (cabi.WANT_OPS | cabi.HEATMAP, cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY):
            g.reset_summary()
            l0 = g.launch_count()
            ops = np.zeros((b["N"], b["stride"]), dtype=np.uint8)
            rec, ops = g.align_batch(b["seqs"], b["off"], b["rc"], flags, ops=ops)
            assert g.launch_count() - l0 >= 2 * n  # every device ran its kernels
            assert compare_records(rec, b["orec"]) == []
            gapped = b["orec"]["indel"] != 0
            rows = slice(None) if not flags & cabi.OPS_GAPPED_ONLY else gapped
            assert np.array_equal(ops[rows], b["oops"][rows])
            for r in range(n):  # every device holds the same reduced counters
                assert np.array_equal(g.summary(r), b["summary"]), "summary read from device %d" % r
            if flags & cabi.HEATMAP:
                assert np.array_equal(g.heatmap(n - 1), b["heat"])
        # the devices' own counters are untouched by the reduction: reducing twice gives the same
        assert np.array_equal(g.summary(0), b["summary"])
        # a batch smaller than the group (some devices get no read) and an empty one
        k = max(1, n - 1)
        rec, ops = g.align_batch(b["seqs"][:int(b["off"][k])], b["off"][:k + 1], b["rc"][:k], cabi.WANT_OPS | cabi.NO_SUMMARY)
        assert compare_records(rec, b["orec"][:k]) == []
        rec, _ = g.align_batch(b["seqs"][:0], b["off"][:1], b["rc"][:0], cabi.NO_SUMMARY)
        assert len(rec) == 0
    finally:
        g.close()


def test_group_errors():
    g = T.DeviceGroup(n=1)
    try:
        with pytest.raises(T.TreatError):  # no template yet
            g.align_batch(np.frombuffer(b"ACGT", dtype=np.uint8), np.array([0, 4], dtype=np.uint64))
    finally:
        g.close()
    with pytest.raises(T.TreatError):
        T.DeviceGroup(devices=[0, 0])  # one context per device
    with pytest.raises(T.TreatError):
        T.DeviceGroup(devices=[_gpus() + 3])


def test_comm_single_rank(batch):
    """nranks = 1: the reduction is the context's own counters (no NCCL needed)."""
    b = batch
    eng = T.Aligner(0)
    try:
        eng.SetTemplate(b["tm"])
        k = 5000
        eng.align_batch(b["seqs"][:int(b["off"][k])], b["off"][:k + 1], b["rc"][:k], cabi.HEATMAP)
        c = T.Comm(eng, 1, 0)
        assert np.array_equal(c.reduce_summary(), eng.summary())
        assert np.array_equal(c.reduce_heatmap(), eng.heatmap())
        c.close()
    finally:
        eng.close()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_worker(rank, world, uid_path, out_dir):
    """One process per GPU: the library's own communicator, the id handed over through a file."""
    import time
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import treat_b200 as T2
    from treat_b200 import cabi as cb, parallel, synth as sy
    t = sy.rps12_template()
    path = t.write_fasta(os.path.join(out_dir, "tmpl_%d.fa" % rank))
    tm = T2.NewTemplateFromFasta(path, T2.FORWARD, "T")
    tm.SetOffset(3)
    eng = T2.Aligner(rank)
    eng.SetTemplate(tm)
    if rank == 0:
        uid = T2.Comm.unique_id()
        with open(uid_path + ".tmp", "wb") as fh:
            fh.write(uid)
        os.rename(uid_path + ".tmp", uid_path)
    else:
        for _ in range(600):
            if os.path.exists(uid_path):
                break
            time.sleep(0.1)
        uid = open(uid_path, "rb").read()
    comm = T2.Comm(eng, world, rank, uid)
    N = 70003
    lo, hi = parallel.shard_range(N, rank, world)
    seqs, off, rc = sy.reads(t, hi - lo, seed=sy.SEED_BASE + 41, first=lo, p_trunc=0.05)
    rec, _ = eng.align_batch(seqs, off, rc, cb.HEATMAP)
    np.save(os.path.join(out_dir, "rec_%d.npy" % rank), rec)
    np.save(os.path.join(out_dir, "red_%d.npy" % rank), comm.reduce_summary())
    np.save(os.path.join(out_dir, "heat_%d.npy" % rank), comm.reduce_heatmap())
    comm.close()
    eng.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_comm_one_process_per_gpu(batch, tmp_path, world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    b = batch
    mp.spawn(_rank_worker, args=(world, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=world, join=True)
    shards = [np.load(tmp_path / ("rec_%d.npy" % r)) for r in range(world)]
    assert compare_records(np.concatenate(shards), b["orec"]) == []
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("red_%d.npy" % r)), b["summary"])
        assert np.array_equal(np.load(tmp_path / ("heat_%d.npy" % r)), b["heat"])
