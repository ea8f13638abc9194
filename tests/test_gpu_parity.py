"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact records, traceback
ops and `treat align` text.  Run on the B200 box with -m gpu."""
import os
import random

import numpy as np
import pytest

os.environ.setdefault("TREAT_B200_FASTA_PIECE", "300000")  # small pieces: treat_b200_fasta_stream cuts even the test files
os.environ.setdefault("TREAT_B200_TEXT_PIECE", "200000")   # treat_b200_fasta_text: several pieces per test file
os.environ.setdefault("TREAT_B200_FASTA_CHUNK", "30000")   # treat_b200_fasta_align: several chunks for the larger test files

import treat_b200 as T
from treat_b200 import cabi, synth
from oracle import oracle as O
from oracle import pyref as P
from helpers import compare_records, heatmap_from_records, summary_from_records

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = T.Aligner(0)
    yield e
    e.close()


def _oracle_template(t: synth.TemplateArrays, tmp_path, offset=0):
    path = t.write_fasta(str(tmp_path / "tmpl.fa"))
    om = O.NewTemplateFromFasta(path, O.FORWARD, t.edit_base)
    om.SetOffset(offset)
    tm = T.NewTemplateFromFasta(path, T.FORWARD, t.edit_base)
    tm.SetOffset(offset)
    return om, tm


def _check_attrs(a, attr, rid):
    # align_test.go:86-133
    assert a.EditStop == attr["ess"], rid
    assert a.JuncStart == attr["jss"], rid
    assert a.JuncEnd == attr["jes"], rid
    assert a.JuncLen == attr["jes"] - attr["ess"], rid
    assert a.AltEditing == attr["alt"], rid
    assert a.HasMutation == attr["has_mutation"], rid
    assert a.Mismatches == attr["mismatches"], rid
    assert a.Indel == attr["indel"], rid


def test_align_fragments(eng, fx, examples):
    """TestAlignFragments (align_test.go:68-137) on the GPU, read by read as the reference does."""
    ts = fx["test_sample"]
    tmpl = T.NewTemplateFromFasta(examples[ts["templates"]], T.FORWARD, ts["base"])
    assert tmpl.Size() == ts["size"]
    for r in ts["reads"]:
        frag = T.NewFragment(r["id"], r["seq"], T.FORWARD, ts["base"])
        aln = eng.NewAlignment(frag, tmpl, False)
        _check_attrs(aln, r["attr"], r["id"])


def test_align_smoke(eng, fx):
    """TestAlign (align_test.go:32-52): must not fail; text equals the oracle's."""
    s = fx["align_smoke"]
    a, b, c = (T.NewFragment(n, s[k], T.FORWARD, s["base"]) for n, k in (("a", "fe"), ("b", "pe"), ("c", "read")))
    tmpl = T.NewTemplate(a, b, None, None)
    aln = eng.NewAlignment(c, tmpl, False)
    oa, ob, oc = (O.NewFragment(n, s[k], O.FORWARD, s["base"]) for n, k in (("a", "fe"), ("b", "pe"), ("c", "read")))
    ot = O.NewTemplate(oa, ob)
    assert aln.WriteTo(c, tmpl, 80) == O.NewAlignment(oc, ot).WriteTo(oc, ot, 80)


def test_readme_goldens(eng, fx, examples):
    rb = fx["readme_block"]
    assert eng.cli_align(template=examples[rb["templates"]], fragment=examples[rb["reads"]], base=rb["base"]) == rb["stdout"]
    rp = fx["readme_pairwise"]
    assert eng.cli_align(s1=rp["s1"], s2=rp["s2"], base=rp["base"]) == rp["stdout"]
    rc = fx["readme_csv"]
    tmpl = T.NewTemplateFromFasta(examples[rc["templates"]], T.FORWARD, "T")
    reads = T.read_fasta(examples[rc["reads"]])
    alns = {rid: a for rid, a in zip(reads.ids, eng.NewAlignments(reads, tmpl, False))}
    for row in rc["rows"]:
        a = alns[row["id"]].fields()
        for k in ("read_count", "alt_editing", "has_mutation", "edit_stop", "junc_end", "junc_len", "junc_seq"):
            assert a[k] == row[k], (row["id"], k)


def test_derived_fixtures(eng, fx, examples):
    """Every frozen example set: fields, gapped strings, WriteTo text and MarshalBinary bytes."""
    for s in fx["derived"]["sets"]:
        tmpl = T.NewTemplateFromFasta(examples[s["templates"]], T.FORWARD, s["base"])
        tmpl.SetOffset(s["offset"])
        reads = T.read_fasta(examples[s["reads"]])
        alns = eng.NewAlignments(reads, tmpl, s["exclude_snps"])
        assert len(alns) == len(s["results"])
        for i, (a, want) in enumerate(zip(alns, s["results"])):
            assert reads.ids[i] == want["id"]
            assert a.fields() == want["fields"], want["id"]
            frag = T.NewFragment(reads.ids[i], reads.seq(i), T.FORWARD, s["base"])
            assert a.gapped(tmpl.Bases, frag.Bases) == (want["aln1"], want["aln2"]), want["id"]
            assert a.WriteTo(frag, tmpl, 80) == want["text"], want["id"]
            assert a.MarshalBinary().hex() == want["marshal_hex"], want["id"]


def test_cli_matches_oracle(eng, examples):
    for t, f, base, off in (("templates.fa", "clones.fa", "T", 0), ("templates.fa", "sample-1.fa", "T", 10),
                            ("test-templates.fa", "test-sample.fa", "t", 0)):
        assert eng.cli_align(template=examples[t], fragment=examples[f], base=base, offset=off) == \
            O.cli_align(template=examples[t], fragment=examples[f], base=base, offset=off)
    # two-record file mode (cmd/treat/align.go:95-111)
    assert eng.cli_align(fragment=examples["sample-1.fa"], base="T") == O.cli_align(fragment=examples["sample-1.fa"], base="T")


def _mutate(rng, s):
    s = list(s.upper())
    for _ in range(rng.randint(0, 5)):
        k = rng.randint(0, 5)
        pos = rng.randrange(len(s) + 1)
        if k == 0 and s:
            del s[min(pos, len(s) - 1)]
        elif k == 1:
            s.insert(pos, rng.choice("ACGTTTN"))
        elif k == 2 and s:
            s[min(pos, len(s) - 1)] = rng.choice("ACGTTNRY")
        elif k == 3:
            s[pos:pos] = "T" * rng.randint(1, 9)
        elif k == 4:
            lo = rng.randrange(len(s) + 1)
            del s[lo:lo + rng.randint(0, 30)]
        else:
            s[pos:pos] = [rng.choice("ACG") for _ in range(rng.randint(1, 12))]
    return "".join(s)


@pytest.mark.parametrize("tname,base", [("templates.fa", "T"), ("test-templates.fa", "t"), ("simple-templates.fa", "T")])
def test_fuzz_vs_oracle(eng, examples, tname, base):
    """Differential fuzz: mutated template strings (indels, substitutions, N/IUPAC, long edit-base
    runs, truncations), both excludeSnps settings; records, ops and text must equal the oracle's."""
    rng = random.Random(hash(tname) & 0xffff)
    srcs = [s for _, s in P.read_fasta(examples[tname])]
    recs = [("f-%d" % (i + 1), _mutate(rng, rng.choice(srcs))) for i in range(400)]
    recs += [("empty-1", ""), ("allT-2", "TTTTTTT"), ("one-3", "A"), ("lower-4", srcs[0].lower())]
    fb = T.FastaBatch.from_records(recs)
    tm = T.NewTemplateFromFasta(examples[tname], T.FORWARD, base)
    om = O.NewTemplateFromFasta(examples[tname], O.FORWARD, base)
    for ex in (False, True):
        alns = eng.NewAlignments(fb, tm, ex)
        orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, exclude_snps=ex)
        grec = np.concatenate([a._rec for a in alns])
        assert compare_records(grec, orec) == []
        for i, a in enumerate(alns):
            n = int(orec[i]["aln_len"])
            assert a.ops() == [(int(oops[i, c >> 2]) >> (2 * (c & 3))) & 3 for c in range(n)], recs[i]
        for i in range(0, len(recs), 23):
            of = O.NewFragment(recs[i][0], recs[i][1], O.FORWARD, base)
            tf = T.NewFragment(recs[i][0], recs[i][1], T.FORWARD, base)
            oa = O.NewAlignment(of, om, ex)
            assert alns[i].WriteTo(tf, tm, 80) == oa.WriteTo(of, om, 80)
            assert alns[i].JuncSeq == oa.JuncSeq
            assert alns[i].MarshalBinary() == oa.MarshalBinary()
    # the same reads as a FASTA file through the device-side FASTA parse + WriteTo: every byte of `treat align`'s output
    # (the empty sequence is written as a header-only record)
    text = "".join(">%s\n%s\n" % (rid, sq) if sq else ">%s\n" % rid for rid, sq in recs).encode()
    eng.SetTemplate(tm)
    got, n = eng.fasta_text(text)
    assert n == len(recs)
    want = "".join(O.NewAlignment(O.NewFragment(rid, sq, O.FORWARD, base), om).WriteTo(O.NewFragment(rid, sq, O.FORWARD, base), om, 80)
                   for rid, sq in recs)
    assert got.decode("latin-1") == want


def _batch_vs_oracle(eng, t, tmp_path, N, seed, p_trunc=0.0, offset=0, exclude=False, nthreads=8):
    om, tm = _oracle_template(t, tmp_path, offset)
    eng.SetTemplate(tm)
    seqs, off, rc = synth.reads(t, N, seed=seed, p_trunc=p_trunc)
    flags = cabi.WANT_OPS | (cabi.EXCLUDE_SNPS if exclude else 0)
    eng.reset_summary()
    grec, gops = eng.align_batch(seqs, off, rc, flags)
    orec, oops, _ = O.align_batch(om, seqs, off, rc, exclude_snps=exclude, nthreads=nthreads, ops_stride=gops.shape[1])
    assert compare_records(grec, orec) == []
    assert np.array_equal(gops, oops)
    lens = np.diff(off.astype(np.int64))
    assert np.array_equal(grec["raw_len"], lens)
    want = summary_from_records(orec, t.m, offset, om.EditStop)
    assert np.array_equal(eng.summary(), want)
    # the one-read-per-warp kernel, the full (unbanded) traceback store and the DP for reads whose bases equal
    # the template's (finished without a DP by default) must all give the same bytes
    k = min(N, 3001)
    E = cabi.DP_EVERY_READ
    # (the first two also run k_band -- pure deletions, banded fill -- in front of the full-matrix kernels; NO_BAND does not)
    for extra in (E, cabi.NO_FRONTIER, cabi.NO_BAND, E | cabi.NO_BAND, cabi.ONE_READ_PER_WARP, cabi.FULL_TRACEBACK_STORE, E | cabi.ONE_READ_PER_WARP,
                  E | cabi.FULL_TRACEBACK_STORE, E | cabi.ONE_READ_PER_WARP | cabi.FULL_TRACEBACK_STORE):
        srec, sops = eng.align_batch(seqs[:int(off[k])], off[:k + 1], rc[:k], flags | extra | cabi.NO_SUMMARY)
        assert np.array_equal(srec, grec[:k]) and np.array_equal(sops[:, :gops.shape[1]], gops[:k, :sops.shape[1]])
    return grec, orec, seqs, off


def test_synthetic_rps12_vs_oracle(eng, tmp_path):
    """Config 3 shape (RPS12, K=5) at a size the oracle finishes in seconds: all records + all ops."""
    grec, orec, seqs, off = _batch_vs_oracle(eng, synth.rps12_template(), tmp_path, 20001, synth.SEED_BASE + 3)
    # JuncSeq is a window of the upper-cased raw read
    for i in np.nonzero(orec["junc_seq_len"] > 0)[0][:200]:
        s = seqs[int(off[i]):int(off[i + 1])].tobytes().upper()
        lo = int(grec[i]["junc_seq_off"])
        assert len(s[lo:lo + int(grec[i]["junc_seq_len"])]) == int(orec[i]["junc_seq_len"])


def test_synthetic_rps12_offset_exclude(eng, tmp_path):
    _batch_vs_oracle(eng, synth.rps12_template(), tmp_path, 4000, synth.SEED_BASE + 31, offset=10, exclude=True)


def test_synthetic_long_template_vs_oracle(eng, tmp_path):
    """Config 5 shape: ~600-nt gene (m=300, K=10), variable-length reads with 5'/3' truncation."""
    _batch_vs_oracle(eng, synth.long_template(300, 8), tmp_path, 3000, synth.SEED_BASE + 5, p_trunc=0.2)


def test_band_fallback(eng, tmp_path):
    """Reads whose optimal path leaves the shared-memory band (long truncations, big deletions, a
    deliberately tiny band) must be re-filled into the global scratch and stay exact."""
    t = synth.rps12_template()
    _batch_vs_oracle(eng, t, tmp_path, 4001, synth.SEED_BASE + 77, p_trunc=0.6)
    from treat_b200.cabi import lib, check
    for rows in (9, 20, 1000):
        check(lib().treat_b200_set_band_rows(eng._ctx, rows), eng._ctx)
        _batch_vs_oracle(eng, t, tmp_path, 2001, synth.SEED_BASE + 78 + rows, p_trunc=0.3)
        _batch_vs_oracle(eng, synth.long_template(300, 8), tmp_path, 1001, synth.SEED_BASE + 79 + rows, p_trunc=0.3)
    check(lib().treat_b200_set_band_rows(eng._ctx, 0), eng._ctx)


def test_other_template_sizes(eng, tmp_path):
    """Exercise every columns-per-lane instantiation (m = 20 .. 470)."""
    for m in (20, 40, 100, 150, 200, 230, 300, 330, 470):
        _batch_vs_oracle(eng, synth.long_template(m, 2, seed=99 + m), tmp_path, 300, 1000 + m, p_trunc=0.3)


def test_templates_longer_than_one_pass(eng, tmp_path):
    """Templates of 481 .. 1023 bases run in column strips of 480 (k_align_strips): strip boundaries (481, 960, 961),
    the largest template (Len = 1024 sites), truncated reads, offset + exclude-snps, and the same on the wide path
    (a template base outside the packed alphabet)."""
    for m, n_alt, N in ((481, 2, 300), (505, 2, 300), (511, 3, 300), (512, 2, 300), (600, 8, 1500), (960, 2, 300), (961, 3, 300), (1023, 4, 600)):
        _batch_vs_oracle(eng, synth.long_template(m, n_alt, seed=7 + m), tmp_path, N, 5000 + m, p_trunc=0.3)
    _batch_vs_oracle(eng, synth.long_template(700, 4, seed=11), tmp_path, 500, 6001, p_trunc=0.2, offset=7, exclude=True)
    # wide path: one template base becomes N (every read then differs from it in that base)
    t = synth.long_template(700, 2, seed=12)
    wb = bytearray(t.bases)
    wb[wb.index(b"G", 500)] = ord("N")
    tw = synth.TemplateArrays(bytes(wb), t.edit_site, t.alt_start, t.alt_end, "T")
    wide_path = tw.write_fasta(str(tmp_path / "long_wide_n.fa"))
    recs = list(zip(["fe", "pe"], tw.raw_sequences()))
    tm = T.NewTemplateFromFasta(wide_path, T.FORWARD, "T")
    om = O.NewTemplateFromFasta(wide_path, O.FORWARD, "T")
    assert "N" in tm.Bases and len(tm.Bases) == 700
    eng.SetTemplate(tm)
    seqs, off, rc = synth.reads(t, 400, seed=6002, p_trunc=0.2)
    grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=8, ops_stride=gops.shape[1])
    assert compare_records(grec, orec) == []
    assert np.array_equal(gops, oops)
    # single reads through the reference-shaped API, text and all
    fe = T.NewFragment("fe", recs[0][1])
    for name, sq in (("same-3", recs[0][1]), ("pe-1", recs[1][1]), ("cut-2", recs[0][1][40:-300]), ("none-1", "TTTT")):
        a = eng.NewAlignment(T.NewFragment(name, sq), tm, False)
        oa = O.NewAlignment(O.NewFragment(name, sq), om)
        assert a.fields() == oa.fields() and a.WriteTo(T.NewFragment(name, sq), tm, 80) == oa.WriteTo(O.NewFragment(name, sq), om, 80)
    assert fe.Len() > 700


def test_wide_paths(eng, tmp_path):
    """Template with a non-ACG base, and a read with an edit-base run >= 255: both must take the
    wide kernels and stay exact."""
    fe, pe = "ACNGTTAGC" + "AG" * 20, "ACNGAGC" + "AG" * 20
    p = tmp_path / "wide.fa"
    p.write_text(">fe\n%s\n>pe\n%s\n" % (fe, pe))
    tm = T.NewTemplateFromFasta(str(p), T.FORWARD, "T")
    om = O.NewTemplateFromFasta(str(p), O.FORWARD, "T")
    recs = [("a-1", fe), ("b-2", pe), ("c-3", fe.replace("N", "R")), ("d-4", "ACNG" + "T" * 300 + "AGC" + "AG" * 20),
            ("e-5", "ACNGTTAGC" + "AG" * 19 + "A" + "T" * 700 + "G"), ("f-6", "N" * 30)]
    fb = T.FastaBatch.from_records(recs)
    alns = eng.NewAlignments(fb, tm, False)
    orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts)
    assert compare_records(np.concatenate([a._rec for a in alns]), orec) == []
    for i, a in enumerate(alns):
        of, tf = O.NewFragment(*recs[i]), T.NewFragment(*recs[i])
        oa = O.NewAlignment(of, om)
        assert a.JuncSeq == oa.JuncSeq
        assert a.WriteTo(tf, tm, 80) == oa.WriteTo(of, om, 80)
    # canonical template, saturated read -> automatic re-run on the wide path
    t = synth.rps12_template()
    om2, tm2 = _oracle_template(t, tmp_path)
    raw = t.raw_sequences()[0]
    recs = [("x-1", raw), ("y-2", raw[:100] + "T" * 400 + raw[100:]), ("z-3", "T" * 260 + raw)]
    fb = T.FastaBatch.from_records(recs)
    alns = eng.NewAlignments(fb, tm2, False)
    orec, _, _ = O.align_batch(om2, fb.seqs, fb.offsets, fb.read_counts)
    assert compare_records(np.concatenate([a._rec for a in alns]), orec) == []
    assert all(a.Status & cabi.ST_SATURATED for a in alns[1:])
    for i, a in enumerate(alns):
        assert a.JuncSeq == O.NewAlignment(O.NewFragment(*recs[i]), om2).JuncSeq
    # the same through the FASTA entry points (upload + align, and streamed): the saturated pieces fall back to
    # compacted reads + the wide kernels, the wide template never reads in place
    seqs, off, rcs = synth.reads(t, 5000, seed=31337)
    body = [(("r%d_%d" % (i, rcs[i])), seqs[int(off[i]):int(off[i + 1])].tobytes().decode()) for i in range(5000)]
    body[1234] = ("sat-7", raw[:80] + "T" * 300 + raw[80:])
    body[4321] = ("sat-9", "T" * 999 + raw)
    text = "".join(">%s\n%s\n" % b for b in body).encode()
    _fasta_device_vs_host(eng, tmp_path, text, "saturated_reads")
    eng.SetTemplate(tm)
    text = "".join(">%s\n%s\n" % r for r in [("a-1", fe), ("b-2", pe), ("c-3", fe.replace("N", "R")), ("f-6", "N" * 30)] * 50).encode()
    _fasta_device_vs_host(eng, tmp_path, text, "wide_template")


def test_edge_cases(eng, tmp_path):
    """Reference panic case, uint8 mismatch wrap, empty batch, argument errors."""
    tm = T.NewTemplate(T.NewFragment("fe", "AT"), T.NewFragment("pe", "A"), None, None)
    a = eng.NewAlignment(T.NewFragment("r-1", "TA"), tm, False)
    ot = O.NewTemplate(O.NewFragment("fe", "AT"), O.NewFragment("pe", "A"))
    oa = O.NewAlignment(O.NewFragment("r-1", "TA"), ot)
    assert a.WouldPanic and oa.WouldPanic and a.fields() == oa.fields()
    tm = T.NewTemplate(T.NewFragment("fe", "A" * 300), T.NewFragment("pe", "A" * 300), None, None)
    a = eng.NewAlignment(T.NewFragment("x-1", "G" * 300), tm, False)
    assert (a.Mismatches, a.HasMutation, a.Indel) == (300 % 256, 1, 0)
    rec, ops = eng.align_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert len(rec) == 0
    with pytest.raises(T.TreatError):
        T.Aligner(0).align_batch(np.zeros(4, np.uint8), np.array([0, 4], np.uint64))  # no template
    big = T.NewFragment("fe", "AC" * 512)
    with pytest.raises(T.TreatError):
        eng.SetTemplate(T.NewTemplate(big, big, None, None))  # 1024 bases > 1023 (Len <= 1024 sites)
    with pytest.raises(T.TreatError):
        tm = T.NewTemplate(T.NewFragment("fe", "ACG"), T.NewFragment("pe", "ACG"), None, None)
        eng.NewAlignment(T.NewFragment("long-1", "A" * 20000), tm, False)


def test_device_resident_entry_point(eng, tmp_path):
    """treat_b200_align_batch_device with torch-owned buffers on a torch stream == host-buffer call."""
    import torch
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 5000
    seqs, off, rc = synth.reads(t, N, seed=77)
    hrec, hops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    d_seqs = torch.from_numpy(seqs.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_rc = torch.from_numpy(rc.astype(np.int32)).cuda()
    stride = hops.shape[1]
    d_rec = torch.zeros(N * 48, dtype=torch.uint8, device="cuda")
    d_ops = torch.zeros(N * stride, dtype=torch.uint8, device="cuda")
    max_len = int(np.diff(off.astype(np.int64)).max())
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, cabi.WANT_OPS,
                               d_rec.data_ptr(), d_ops.data_ptr(), stride, st.cuda_stream)
    st.synchronize()
    drec = d_rec.cpu().numpy().view(cabi.RECORD_DTYPE)
    assert np.array_equal(drec, hrec)
    assert np.array_equal(d_ops.cpu().numpy().reshape(N, stride), hops)


def test_pack_lengths_and_alignments(eng, tmp_path):
    """k_pack16 takes 16-byte aligned blocks: every start misalignment, lengths around the block and the 512-character
    pass boundaries, lower case, non-alphabet characters, long edit-base runs (below the saturation limit)."""
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    rng = random.Random(20261020)
    reads = []
    lens = [0, 1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 255, 256, 257, 495, 496, 497, 511, 512, 513, 527, 528, 1023, 1024, 1025, 3000,
            16383, 16384]
    for L in lens * 6:
        kind = rng.randrange(4)
        if kind == 0:
            s = "".join(rng.choice("ACGT") for _ in range(L))
        elif kind == 1:
            s = "".join(rng.choice("acgtACGTTTTT") for _ in range(L))
        elif kind == 2:
            s = "".join(rng.choice("ACGTTTTTTTTTTTTTTTTTNn-") for _ in range(L))
        else:  # edit-base runs that cross many blocks, still < 255
            s = ""
            while len(s) < L:
                s += "T" * rng.randrange(0, 200) + rng.choice("ACG") * rng.randrange(1, 3)
            s = s[:L]
        reads.append(s)
    with pytest.raises(T.TreatError):  # longer than the 16,384-character limit
        eng.align_batch(np.full(16385, ord("A"), dtype=np.uint8), np.array([0, 16385], dtype=np.uint64), None, cabi.NO_SUMMARY)
    for pad in range(17):  # shift the whole batch through every misalignment
        batch = ["A" * pad] + reads
        seqs = np.frombuffer("".join(batch).encode(), dtype=np.uint8).copy()
        off = np.concatenate([[0], np.cumsum([len(x) for x in batch])]).astype(np.uint64)
        rc = np.ones(len(batch), dtype=np.uint32)
        grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.NO_SUMMARY)
        orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=8, ops_stride=gops.shape[1])
        assert compare_records(grec, orec) == [], pad
        assert np.array_equal(gops, oops), pad


def _fasta_device_vs_host(eng, tmp_path, text: bytes, tag: str, flags=None, tm=None):
    """upload_fasta/align_fasta (parse on the GPU) == read_fasta (host parser) + align_batch, field by field."""
    path = os.path.join(str(tmp_path), tag + ".fa")
    with open(path, "wb") as f:
        f.write(text)
    host = T.read_fasta(path)
    n, ml = eng.upload_fasta(text)
    assert n == len(host), tag
    lens = np.diff(host.offsets.astype(np.int64))
    assert ml == (int(lens.max()) if n else 0), tag
    flags = cabi.WANT_OPS | cabi.NO_SUMMARY if flags is None else flags
    rec, ops, ids, rcs = eng.align_fasta(flags)
    assert ids == host.ids, tag
    assert np.array_equal(rcs, host.read_counts), tag
    if n:
        hrec, hops = eng.align_batch(host.seqs, host.offsets, host.read_counts, flags)
        assert np.array_equal(rec, hrec), tag
        if hops is not None and flags & cabi.OPS_GAPPED_ONLY:  # rows of the gapped reads only
            g = hrec["indel"] != 0
            frec, fops = eng.align_batch(host.seqs, host.offsets, host.read_counts, flags & ~cabi.OPS_GAPPED_ONLY)
            used = (len(tm.Bases) + int(hrec["n_bases"].max()) + 3) // 4
            assert g.any() and np.array_equal(ops[g][:, :used], fops[g][:, :used]), tag
        elif hops is not None:
            assert np.array_equal(ops[:, :hops.shape[1]], hops), tag
    # the streaming form: same records, ids, counts and ops, delivered piece by piece in file order
    pieces, total = eng.stream_fasta(text, flags)
    assert total == n and sum(p["n"] for p in pieces) == n, tag
    if n:
        assert [p["first_record"] for p in pieces] == list(np.cumsum([0] + [p["n"] for p in pieces[:-1]])), tag
        srec = np.concatenate([p["rec"] for p in pieces])
        assert np.array_equal(srec, rec), tag
        assert np.array_equal(np.concatenate([p["read_count"] for p in pieces]), rcs), tag
        sid = [text[int(o):int(o) + int(l)].decode("latin-1") for p in pieces for o, l in zip(p["id_off"], p["id_len"])]
        assert sid == host.ids, tag
        slen = np.concatenate([p["seq_len"] for p in pieces])
        assert np.array_equal(slen, lens), tag
        for p in pieces:  # single-line sequences sit in the file exactly where seq_off says
            if p["seq_in_place"]:
                for i in range(0, p["n"], max(1, p["n"] // 50)):
                    r = p["first_record"] + i
                    assert text[int(p["seq_off"][i]):int(p["seq_off"][i]) + int(p["seq_len"][i])] == host.seq(r), tag
        if flags & cabi.WANT_OPS:
            if flags & cabi.OPS_GAPPED_ONLY:
                for p in pieces:
                    g = np.nonzero(p["rec"]["indel"] != 0)[0]
                    assert sorted(p["ops_index"].tolist()) == g.tolist(), tag
                    w = p["ops_stride"]
                    for q, i in enumerate(p["ops_index"]):
                        assert np.array_equal(p["ops"][q], ops[p["first_record"] + int(i)][:w]), tag
            else:
                for p in pieces:
                    w = min(p["ops_stride"], ops.shape[1])
                    assert np.array_equal(p["ops"][:, :w], ops[p["first_record"]:p["first_record"] + p["n"], :w]), tag
    # `treat load` in one call: the sink also gets every record's MarshalBinary blob, built on the device
    pieces_b, total_b = eng.stream_fasta(text, flags | cabi.WANT_BLOBS)
    assert total_b == n
    if n:
        hblob, hboff = T.marshal_batch(rec, host.seqs, host.offsets, threads=2)
        assert np.array_equal(np.concatenate([p["blob"] for p in pieces_b]), hblob), tag
        at = 0
        for p in pieces_b:
            assert np.array_equal(p["blob_off"] + np.uint64(at), hboff[p["first_record"]:p["first_record"] + p["n"] + 1]), tag
            at += int(p["blob_off"][-1])
        assert np.array_equal(np.concatenate([p["rec"] for p in pieces_b]), rec), tag
    return n


def test_fasta_ingest_on_device(eng, tmp_path):
    """gofasta record loop + parseMergeCount on the device, against the host reader."""
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    rng = random.Random(77)
    seqs, off, rc = synth.reads(t, 3000, seed=515)
    reads = [seqs[int(off[i]):int(off[i + 1])].tobytes() for i in range(3000)]
    names = ["r%d_%d" % (i, rc[i]) for i in range(3000)]
    odd_ids = ["plain", "a_b_007", "x-3 tail_9", "_5", "5_", "q_12x", "big_99999999999999999999", "z_0", "  padded_4  \t", "",
               "tab\tsep_8", "dash-17", "m_n-o_21 second_3", "r_4294967297"]
    # (a) two-line records, odd ids
    a = b"".join(b">" + (odd_ids[i % len(odd_ids)] if i % 7 == 0 else names[i]).encode() + b"\n" + reads[i] + b"\n" for i in range(3000))
    _fasta_device_vs_host(eng, tmp_path, a, "two_line")
    # (b) CRLF, no newline at the end / a lone CR at the end
    b = b"".join(b">" + names[i].encode() + b"\r\n" + reads[i] + b"\r\n" for i in range(500))
    _fasta_device_vs_host(eng, tmp_path, b, "crlf")
    _fasta_device_vs_host(eng, tmp_path, b[:-2], "crlf_no_eol")
    _fasta_device_vs_host(eng, tmp_path, b[:-1], "crlf_cr_at_eof")
    # (c) wrapped sequences, blank lines, junk in front, the first '>' inside a line, '>' inside a sequence line
    parts = [b"junk line\nmore junk>first_2 in a line\n"]
    for i in range(700):
        w = rng.choice([1, 7, 60, 61, 80, 1000])
        r = reads[i]
        body = b"\n".join(r[j:j + w] for j in range(0, len(r), w))
        if i % 11 == 0:
            body = body[:10] + b"\n\n" + body[10:]
        if i % 13 == 0:
            body = body[:5] + b">" + body[5:]
        if i % 17 == 0:
            body = body.replace(b"\n", b"\r\n", 2)
        if i % 19 == 0:
            body = b""
        parts.append(body + b"\n>" + names[i].encode() + (b"  \n" if i % 3 else b"\n"))
    parts.append(reads[700])
    _fasta_device_vs_host(eng, tmp_path, b"".join(parts), "wrapped")
    # (d) degenerate files
    for tag, text in (("empty", b""), ("no_records", b"just text\nno records\n"), ("only_header", b">h_3"), ("only_gt", b">"),
                      ("gt_newline", b">\n"), ("header_then_seq_no_eol", b">x_2\nACGTTTG")):
        _fasta_device_vs_host(eng, tmp_path, text, tag)
    # (e) a file that spans many tiles and scan blocks, records-only and gapped-only flags
    seqs, off, rc = synth.reads(t, 200000, seed=616)
    buf = bytearray()
    for i in range(200000):
        buf += b">s%d-%d\n" % (i, rc[i])
        buf += seqs[int(off[i]):int(off[i + 1])].tobytes()
        buf += b"\n"
    assert _fasta_device_vs_host(eng, tmp_path, bytes(buf), "large", cabi.NO_SUMMARY) == 200000
    _fasta_device_vs_host(eng, tmp_path, bytes(buf), "large_gapped_rows", cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY | cabi.NO_SUMMARY, tm)
    # summaries accumulate like the host-buffer call's
    eng.reset_summary()
    eng.align_fasta(cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY, want_ids=False)
    s1 = eng.summary().copy()
    eng.reset_summary()
    eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    assert np.array_equal(s1, eng.summary())


def _text_vs_oracle(eng, tmpl_path, text: bytes, tmp_path, tag, base="T", offset=0):
    """treat_b200_fasta_text (parse + alignment + WriteTo on the GPU) == the oracle's `treat align` stdout."""
    tm = T.NewTemplateFromFasta(tmpl_path, T.FORWARD, base)
    tm.SetOffset(offset)
    eng.SetTemplate(tm)
    fpath = str(tmp_path / ("text_%s.fa" % tag))
    with open(fpath, "wb") as fh:
        fh.write(text)
    got, n = eng.fasta_text(text)
    want = O.cli_align(template=tmpl_path, fragment=fpath, base=base, offset=offset).encode("latin-1")
    if got != want:
        i = next((k for k in range(min(len(got), len(want))) if got[k] != want[k]), min(len(got), len(want)))
        raise AssertionError("%s: text differs at byte %d of %d/%d: %r vs %r" % (tag, i, len(got), len(want), got[max(0, i - 200):i + 80], want[max(0, i - 200):i + 80]))
    return n


def test_write_to_on_device(eng, examples, tmp_path):
    """Alignment.WriteTo (align.go:388-520) for whole files on the device: k_text via treat_b200_fasta_text and
    treat_b200_text_device, byte for byte against the oracle's `treat align` output and the host formatter."""
    # the reference's own example files (C1 / C2 of SURVEY 8d), with and without an offset, lower-case edit base
    for t, f, base, off in (("simple-templates.fa", "simple-sequences.fa", "T", 0), ("templates.fa", "clones.fa", "T", 0),
                            ("templates.fa", "sample-1.fa", "T", 10), ("test-templates.fa", "test-sample.fa", "t", 0)):
        with open(examples[f], "rb") as fh:
            _text_vs_oracle(eng, examples[t], fh.read(), tmp_path, f.replace(".", "_"), base, off)
    # synthetic RPS12 reads (indels, substitutions, N) in several pieces; CRLF, wrapped sequences, odd headers, empty read
    t = synth.rps12_template()
    tpath = t.write_fasta(str(tmp_path / "rps12_text.fa"))
    seqs, off, rc = synth.reads(t, 2500, seed=4242, p_trunc=0.05)
    reads = [seqs[int(off[i]):int(off[i + 1])].tobytes().decode() for i in range(2500)]
    body = "".join(">r%d_%d\n%s\n" % (i, rc[i], r) for i, r in enumerate(reads))
    assert _text_vs_oracle(eng, tpath, body.encode(), tmp_path, "rps12") == 2500
    # sharded the way ranks would take it (treat_b200_fasta_split): the slices' texts in rank order are the file's text
    whole, _ = eng.fasta_text(body.encode())
    cuts = T.fasta_split(body.encode(), 3)
    assert len(cuts) == 4 and cuts[0] == 0 and cuts[-1] == len(body)
    assert b"".join(eng.fasta_text(body.encode()[int(cuts[i]):int(cuts[i + 1])])[0] for i in range(3)) == whole
    odd = ">  spaced name_7  \r\n%s\r\n>wrapped-3\n%s\n%s\n\n>empty-1\n\n>lower_2\n%s\n>tt-1\nTTTTT\n>last-9\n%s" % (
        reads[0], reads[1][:100], reads[1][100:], reads[2].lower(), reads[3])
    assert _text_vs_oracle(eng, tpath, odd.encode(), tmp_path, "odd") == 6
    sat = ">sat-7\n%s\n>plain-1\n%s\n" % (reads[4][:80] + "T" * 300 + reads[4][80:], reads[5])
    assert _text_vs_oracle(eng, tpath, sat.encode(), tmp_path, "saturated") == 2
    # very long reads: the position tables of k_text no longer fit shared memory (column-wise writer, fewer warps per CTA)
    rng = random.Random(5)
    big = "".join(rng.choice("ACGTTTT") for _ in range(15000))
    mid = reads[6][:150] + "".join(rng.choice("ACGTT") for _ in range(5800)) + reads[6][150:]
    longf = ">big-2\n%s\n>plain-1\n%s\n>mid-3\n%s\n" % (big, reads[7], mid)
    assert _text_vs_oracle(eng, tpath, longf.encode(), tmp_path, "long_reads") == 3
    # long template, K = 10 (labels A1..A8) and K = 14 (labels A10..A12 are three characters wide), truncated reads
    # ... and K = 32 on a 600-base template: labels up to A30, EditSite rows too large for k_text's shared memory
    for m, n_alt, N in ((300, 8, 600), (90, 12, 300), (700, 3, 200), (600, 30, 150)):
        tl = synth.long_template(m, n_alt, seed=31 + m)
        lp = tl.write_fasta(str(tmp_path / ("long_text_%d.fa" % m)))
        seqs, off, rc = synth.reads(tl, N, seed=99 + m, p_trunc=0.3)
        body = "".join(">q%d-%d\n%s\n" % (i, rc[i], seqs[int(off[i]):int(off[i + 1])].tobytes().decode()) for i in range(N))
        assert _text_vs_oracle(eng, lp, body.encode(), tmp_path, "long%d" % m) == N
    # wide template (a base outside the packed alphabet)
    wb = bytearray(t.bases)
    wb[wb.index(b"G", 60)] = ord("N")
    tw_ = synth.TemplateArrays(bytes(wb), t.edit_site, t.alt_start, t.alt_end, "T")
    wpath = tw_.write_fasta(str(tmp_path / "wide_text.fa"))
    body = "".join(">w%d-1\n%s\n" % (i, r) for i, r in enumerate(reads[:300]))
    assert _text_vs_oracle(eng, wpath, body.encode(), tmp_path, "wide") == 300
    # treat_b200_text_device on device buffers, other terminal widths, against Alignment.WriteTo of the host mirror
    import torch
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 700
    seqs, off, rc = synth.reads(t, N, seed=777, p_trunc=0.1)
    names = [("read %d_%d" % (i, rc[i])).encode() for i in range(N)]
    noff = np.zeros(N + 1, dtype=np.int64)
    noff[1:] = np.cumsum([len(x) for x in names])
    max_len = int(np.diff(off.astype(np.int64)).max())
    stride = eng.ops_stride(max_len)
    d_seqs = torch.from_numpy(seqs.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_names = torch.from_numpy(np.frombuffer(b"".join(names), dtype=np.uint8).copy()).cuda()
    d_noff = torch.from_numpy(noff).cuda()
    d_rec = torch.zeros(N * 48, dtype=torch.uint8, device="cuda")
    d_ops = torch.zeros(N * stride, dtype=torch.uint8, device="cuda")
    d_toff = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), 0, N, max_len, cabi.WANT_OPS | cabi.NO_SUMMARY, d_rec.data_ptr(),
                           d_ops.data_ptr(), stride, st)
    alns = eng.NewAlignments([(nm.decode(), seqs[int(off[i]):int(off[i + 1])].tobytes().decode()) for i, nm in enumerate(names)], tm)
    for tw in (80, 0, 20, 133, 5):
        args = (d_seqs.data_ptr(), d_off.data_ptr(), d_names.data_ptr(), d_noff.data_ptr(), d_rec.data_ptr(), d_ops.data_ptr(), stride, N,
                max_len, tw, d_toff.data_ptr())
        total = eng.text_device(*args, 0, 0, st)
        d_text = torch.zeros(total, dtype=torch.uint8, device="cuda")
        assert eng.text_device(*args, d_text.data_ptr(), total, st) == total
        torch.cuda.synchronize()
        text = d_text.cpu().numpy().tobytes()
        toff = d_toff.cpu().numpy()
        assert toff[0] == 0 and toff[-1] == total
        for i in range(0, N, 7 if tw in (80, 0) else 37):
            frag = T.NewFragment(names[i].decode(), seqs[int(off[i]):int(off[i + 1])].tobytes().decode())
            assert text[int(toff[i]):int(toff[i + 1])].decode() == alns[i].WriteTo(frag, tm, tw), (tw, i)
    with pytest.raises(T.TreatError):
        eng.text_device(*args[:9], 4, d_toff.data_ptr(), 0, 0, st)  # tw = 4: the reference divides by zero
    # treat_b200_text_batch: the same from host buffers in chunks (40,000 reads: two chunks), names as the oracle builds them
    N2 = 40000
    seqs2, off2, rc2 = synth.reads(t, N2, seed=778, p_trunc=0.02)
    got = eng.text_batch(seqs2, off2, ["r%d-%d" % (i, rc2[i]) for i in range(N2)], cabi.NO_SUMMARY)
    want, _ = O.text_batch(om, seqs2, off2, rc2, prefix="r", nthreads=8)
    assert got == want.tobytes()
    assert eng.text_batch(seqs2[:0], off2[:1], []) == b""


def test_heatmap_consumer(eng, tmp_path):
    """the ESS x junction-length heat map accumulated on the device == the same from the oracle's records,
    through the host-buffer call (speculative chunks), the device entry point and the FASTA stream; with an edit offset."""
    import torch
    t = synth.rps12_template()
    for offset in (0, 7):
        om, tm = _oracle_template(t, tmp_path, offset)
        eng.SetTemplate(tm)
        N = 150000
        seqs, off, rc = synth.reads(t, N, seed=2024 + offset)
        rc = (rc.astype(np.uint64) * 40000 % 4000000000 + 1).astype(np.uint32)  # large counts: sums pass 32 bits
        orec, _, _ = O.align_batch(om, seqs, off, rc, want_ops=False, nthreads=8)
        want = heatmap_from_records(orec, t.m + 1, offset)
        assert want.sum() > 0 and int(want.max()) > 2 ** 32
        eng.reset_summary()
        eng.align_batch(seqs, off, rc, cabi.HEATMAP)
        assert np.array_equal(eng.heatmap(), want)
        eng.align_batch(seqs, off, rc, 0)  # without the flag nothing is added
        assert np.array_equal(eng.heatmap(), want)
        eng.reset_summary()
        assert eng.heatmap().sum() == 0
        # device entry point
        k = 40000
        d_seqs = torch.from_numpy(seqs[:int(off[k])].copy()).cuda()
        d_off = torch.from_numpy(off[:k + 1].astype(np.int64)).cuda()
        d_rc = torch.from_numpy(rc[:k].view(np.int32).copy()).cuda()
        d_rec = torch.zeros(k * 48, dtype=torch.uint8, device="cuda")
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), k, int(np.diff(off[:k + 1].astype(np.int64)).max()),
                               cabi.HEATMAP, d_rec.data_ptr(), 0, 0)
        assert np.array_equal(eng.heatmap(), heatmap_from_records(orec[:k], t.m + 1, offset))
    # FASTA stream (pieces aligned speculatively, added when validated)
    eng.reset_summary()
    text = b"".join(b">h%d_%d\n" % (i, rc[i]) + seqs[int(off[i]):int(off[i + 1])].tobytes() + b"\n" for i in range(60000))
    pieces, total = eng.stream_fasta(text, cabi.HEATMAP)
    assert total == 60000
    assert np.array_equal(eng.heatmap(), heatmap_from_records(orec[:60000], t.m + 1, 7))


def test_stage_calls_and_pack_marks(eng, tmp_path):
    """pack_device + align_packed_device: with and without USE_PACK_MARKS the records equal the one-call result;
    the mark (bit 7 of the flag byte) is set exactly for reads whose stripped bases equal the template's."""
    import ctypes as C
    import torch
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 6001
    seqs, off, rc = synth.reads(t, N, seed=4242)
    hrec, hops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.NO_SUMMARY)
    max_len = int(np.diff(off.astype(np.int64)).max())
    pk = cabi.Packed()
    cabi.check(cabi.lib().treat_b200_packed_layout(eng._ctx, N, max_len, 0, C.byref(pk)), eng._ctx)
    dev = dict(n=torch.zeros(N, dtype=torch.int16, device="cuda"), flags=torch.zeros(N, dtype=torch.uint8, device="cuda"),
               bases=torch.zeros(N * pk.bases_stride, dtype=torch.uint8, device="cuda"),
               counts=torch.zeros(N * pk.counts_stride, dtype=torch.uint8, device="cuda"),
               raw_len=torch.zeros(N, dtype=torch.int32, device="cuda"))
    pk.d_n, pk.d_flags, pk.d_bases, pk.d_counts, pk.d_raw_len = (dev[k].data_ptr() for k in ("n", "flags", "bases", "counts", "raw_len"))
    d_seqs = torch.from_numpy(seqs.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_rc = torch.from_numpy(rc.astype(np.int32)).cuda()
    stride = hops.shape[1]
    cabi.check(cabi.lib().treat_b200_pack_device(eng._ctx, d_seqs.data_ptr(), d_off.data_ptr(), C.byref(pk), None), eng._ctx)
    eng.synchronize()
    marks = (dev["flags"].cpu().numpy() & 0x80) != 0
    m = t.m
    want = np.array([int(r["n_bases"]) == m and int(r["score"]) == m for r in hrec])
    assert np.array_equal(marks, want) and 0 < marks.sum() < N
    # bit 6: same length, exactly one base differs (score m - 2: the diagonal is the unique optimum, no DP either)
    one = (dev["flags"].cpu().numpy() & 0x40) != 0
    want1 = np.array([int(r["n_bases"]) == m and int(r["score"]) == m - 2 for r in hrec])
    assert np.array_equal(one, want1) and one.sum() > 0
    assert (hrec["mismatches"][one] == 1).all() and (hrec["indel"][one] == 0).all()
    for fl in (0, cabi.USE_PACK_MARKS, cabi.USE_PACK_MARKS | cabi.ONE_READ_PER_WARP):
        d_rec = torch.zeros(N * 48, dtype=torch.uint8, device="cuda")
        d_ops = torch.full((N * stride,), 0xee, dtype=torch.uint8, device="cuda")
        cabi.check(cabi.lib().treat_b200_align_packed_device(eng._ctx, C.byref(pk), d_rc.data_ptr(),
                                                             cabi.WANT_OPS | cabi.NO_SUMMARY | fl, d_rec.data_ptr(),
                                                             d_ops.data_ptr(), stride, None), eng._ctx)
        eng.synchronize()
        assert np.array_equal(d_rec.cpu().numpy().view(cabi.RECORD_DTYPE), hrec), fl
        assert np.array_equal(d_ops.cpu().numpy().reshape(N, stride), hops), fl


def test_ops_for_gapped_reads_only(eng, tmp_path):
    """OPS_GAPPED_ONLY: rows of reads with a gap equal the full output, every other row is left untouched; also with
    many chunks, more gapped reads than the ahead-of-time copy holds, and on the device entry point."""
    import torch
    t = synth.long_template(300, 8)
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 150000  # three chunks
    seqs, off, rc = synth.reads(t, N, seed=991, p_trunc=0.5)  # truncation: about half of the reads are gapped
    frec, fops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.NO_SUMMARY)
    gapped = frec["indel"] != 0
    assert 0.2 * N < gapped.sum() < 0.8 * N
    srec = np.zeros(N, dtype=cabi.RECORD_DTYPE)
    sops = np.full(fops.shape, 0xAB, dtype=np.uint8)
    eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.NO_SUMMARY | cabi.OPS_GAPPED_ONLY, rec=srec, ops=sops)
    assert np.array_equal(srec, frec)
    used0 = (t.m + int(frec["n_bases"].max()) + 3) // 4  # bytes of a row that can hold alignment columns
    assert np.array_equal(sops[gapped][:, :used0], fops[gapped][:, :used0])
    # rows of gapless reads: untouched (compacted copy) or zero (the call copies every row once many reads have gaps)
    rest = sops[~gapped][:, :fops.shape[1] // 2]
    assert ((rest == 0xAB).all(axis=1) | (rest == 0).all(axis=1)).all()
    # few gapped reads: their rows come back compacted and nothing else is written
    t2 = synth.rps12_template()
    om2, tm2 = _oracle_template(t2, tmp_path)
    eng2 = T.Aligner(0)
    try:
        eng2.SetTemplate(tm2)
        s2, o2, r2 = synth.reads(t2, 140000, seed=992)
        frec2, fops2 = eng2.align_batch(s2, o2, r2, cabi.WANT_OPS | cabi.NO_SUMMARY)
        g2 = frec2["indel"] != 0
        srec2 = np.zeros(len(frec2), dtype=cabi.RECORD_DTYPE)
        sops2 = np.full(fops2.shape, 0xAB, dtype=np.uint8)
        eng2.align_batch(s2, o2, r2, cabi.WANT_OPS | cabi.NO_SUMMARY | cabi.OPS_GAPPED_ONLY, rec=srec2, ops=sops2)
        used = (t2.m + int(frec2["n_bases"].max()) + 3) // 4
        assert np.array_equal(srec2, frec2) and 0 < g2.sum() < 0.05 * len(g2)
        assert np.array_equal(sops2[g2][:, :used], fops2[g2][:, :used]) and (sops2[~g2] == 0xAB).all()
    finally:
        eng2.close()
    # device buffers: same contract, rows stay at their read's index
    k = 20000
    d_seqs = torch.from_numpy(seqs[:int(off[k])].copy()).cuda()
    d_off = torch.from_numpy(off[:k + 1].astype(np.int64)).cuda()
    d_rc = torch.from_numpy(rc[:k].astype(np.int32)).cuda()
    stride = fops.shape[1]
    d_rec = torch.zeros(k * 48, dtype=torch.uint8, device="cuda")
    d_ops = torch.full((k * stride,), 0xAB, dtype=torch.uint8, device="cuda")
    max_len = int(np.diff(off[:k + 1].astype(np.int64)).max())
    eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), k, max_len,
                           cabi.WANT_OPS | cabi.NO_SUMMARY | cabi.OPS_GAPPED_ONLY, d_rec.data_ptr(), d_ops.data_ptr(), stride)
    eng.synchronize()
    dops = d_ops.cpu().numpy().reshape(k, stride)
    assert np.array_equal(d_rec.cpu().numpy().view(cabi.RECORD_DTYPE), frec[:k])
    assert np.array_equal(dops[gapped[:k]][:, :used0], fops[:k][gapped[:k]][:, :used0]) and (dops[~gapped[:k]] == 0xAB).all()


def test_speculative_chunks_are_validated(eng, tmp_path):
    """Multi-chunk host batches launch the align kernel without waiting for the pack kernel's max n.
    Chunks whose reads turn out longer than the guess, or hold a saturated edit-base run, must be
    re-run: results and summaries stay exact."""
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 300_000
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 91)
    raw = t.raw_sequences()[0]
    special = {140_000: raw[:200] + "ACGGCA" * 9 + raw[200:],      # n = 163 + 54: longer than any earlier read
               280_000: raw[:50] + "T" * 300 + raw[50:],            # saturated run (>= 255)
               299_999: "ACG" * 150}
    parts, lens, prev = [], np.diff(off.astype(np.int64)).copy(), 0
    for idx in sorted(special):
        parts.append(seqs[int(off[prev]):int(off[idx])])
        parts.append(np.frombuffer(special[idx].encode(), dtype=np.uint8))
        lens[idx] = len(special[idx])
        prev = idx + 1
    parts.append(seqs[int(off[prev]):int(off[N])])
    seqs2 = np.concatenate(parts)
    off2 = np.zeros(N + 1, dtype=np.uint64)
    off2[1:] = np.cumsum(lens)
    eng.reset_summary()
    rec, ops = eng.align_batch(seqs2, off2, rc, cabi.WANT_OPS)
    summ = eng.summary()
    idx = np.unique(np.concatenate([np.arange(0, N, 37), np.array(sorted(special)), np.arange(139_990, 140_010)]))
    sub = np.concatenate([seqs2[int(off2[i]):int(off2[i + 1])] for i in idx])
    sub_off = np.zeros(len(idx) + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum(lens[idx])
    orec, oops, _ = O.align_batch(om, sub, sub_off, rc[idx], nthreads=8, ops_stride=ops.shape[1])
    assert compare_records(rec[idx], orec) == []
    assert np.array_equal(ops[idx], oops)
    assert rec[280_000]["status"] & cabi.ST_SATURATED
    assert np.array_equal(summ, summary_from_records(rec, t.m, 0, om.EditStop))
    rec2, ops2 = eng.align_batch(seqs2, off2, rc, cabi.WANT_OPS)   # second call: hint already large
    assert np.array_equal(rec, rec2) and np.array_equal(ops, ops2)


def test_overlapping_chunks_with_fallbacks(eng, tmp_path):
    """Several host chunks in flight, many reads on the global-scratch fallback (truncated reads against
    a long template): launches of different chunks overlap on the GPU and must not share scratch."""
    t = synth.long_template(300, 8)
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 400_000
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 5, p_trunc=0.2)
    rec, ops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    rec2, ops2 = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.FULL_TRACEBACK_STORE)
    assert np.array_equal(rec, rec2) and np.array_equal(ops, ops2)
    idx = np.arange(0, N, 61)
    lens = np.diff(off.astype(np.int64))
    sub = np.concatenate([seqs[int(off[i]):int(off[i + 1])] for i in idx])
    sub_off = np.zeros(len(idx) + 1, dtype=np.uint64)
    sub_off[1:] = np.cumsum(lens[idx])
    orec, oops, _ = O.align_batch(om, sub, sub_off, rc[idx], nthreads=8, ops_stride=ops.shape[1])
    assert compare_records(rec[idx], orec) == []
    assert np.array_equal(ops[idx], oops)


def test_full_size_properties(eng, tmp_path):
    """Config 3 at full size (1M RPS12 reads): size-independent properties on every read, oracle
    parity on a stride sample, determinism across two runs and across batch splits."""
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 1_000_000
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 3)
    eng.reset_summary()
    rec, ops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    summ = eng.summary()
    # (1) ops consume exactly m template bases and n read bases, and re-scoring them gives `score`
    alen = rec["aln_len"].astype(np.int64)
    cols = np.arange(ops.shape[1] * 4)
    o = (ops[:, cols >> 2] >> (2 * (cols & 3)).astype(np.uint8)) & 3
    valid = cols[None, :] < alen[:, None]
    n_ins = ((o == 2) & valid).sum(1)
    n_del = ((o == 1) & valid).sum(1)
    n_mat = ((o == 0) & valid).sum(1)
    assert np.array_equal(n_mat + n_del, np.full(N, t.m))
    assert np.array_equal(n_mat + n_ins, rec["n_bases"].astype(np.int64))
    assert np.array_equal((n_ins + n_del > 0).astype(np.uint8), rec["indel"])
    # score = matches - mismatches - gaps, and mismatches (mod 256) is what the record holds
    mism = (n_mat - rec["score"] - n_ins - n_del) // 2
    assert np.array_equal((mism & 0xff).astype(np.uint8), rec["mismatches"])
    assert np.all((o == 0)[~valid])  # padding is zero
    # (2) summary == reference definition applied to the records
    assert np.array_equal(summ, summary_from_records(rec, t.m, 0, om.EditStop))
    # (3) oracle parity on a 1/64 stride sample
    idx = np.arange(0, N, 64)
    sub_off = np.zeros(len(idx) + 1, dtype=np.uint64)
    lens = (off[idx + 1] - off[idx]).astype(np.uint64)
    sub_off[1:] = np.cumsum(lens)
    sub = np.concatenate([seqs[int(off[i]):int(off[i + 1])] for i in idx])
    orec, oops, _ = O.align_batch(om, sub, sub_off, rc[idx], nthreads=8, ops_stride=ops.shape[1])
    assert compare_records(rec[idx], orec) == []
    assert np.array_equal(ops[idx], oops)
    # (4) idempotence / independence of batch boundaries
    rec2, ops2 = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    assert np.array_equal(rec, rec2) and np.array_equal(ops, ops2)
    lo, hi = 300_000, 300_000 + 70_001
    rec3, ops3 = eng.align_batch(seqs[int(off[lo]):int(off[hi])], off[lo:hi + 1] - off[lo], rc[lo:hi], cabi.WANT_OPS)
    assert np.array_equal(rec3, rec[lo:hi]) and np.array_equal(ops3[:, :ops.shape[1]], ops[lo:hi, :ops3.shape[1]])


def test_load_and_mutant_rows_from_device_results(eng, tmp_path):
    """SURVEY 8(f) rows on top of the device results: `treat load` blobs (MarshalBinary per read) and the
    `treat mutant` census computed from the device's records + ops equal those from the oracle's."""
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 6000
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 55, p_trunc=0.05)
    grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
    orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=8, ops_stride=gops.shape[1])
    blob, boff = T.marshal_batch(grec, seqs, off, threads=4)
    for i in range(0, N, 97):
        s = seqs[int(off[i]):int(off[i + 1])].tobytes().decode()
        f = O.NewFragment("r-%d" % rc[i], s)
        assert blob[int(boff[i]):int(boff[i + 1])].tobytes() == O.NewAlignment(f, om).MarshalBinary()
    orec_p = np.zeros(N, dtype=cabi.RECORD_DTYPE)
    for f in orec.dtype.names:
        if f in orec_p.dtype.names and f != "pad":
            orec_p[f] = orec[f]
    got = T.mutant_report(tm, grec, gops, seqs, off, 20)
    assert got == T.mutant_report(tm, orec_p, oops, seqs, off, 20)
    assert got.count("\n") > 10


def test_marshal_on_device(eng, tmp_path):
    """`treat load` blobs (Alignment.MarshalBinary, align.go:316-335) built on the device == the host's marshal_batch == the
    oracle's MarshalBinary, for reads with junction sequences, negative fields (offset 0, EditStop -1) and lower-case bases."""
    import torch
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    N = 5000
    seqs, off, rc = synth.reads(t, N, seed=synth.SEED_BASE + 91, p_trunc=0.05)
    seqs = seqs.copy()
    lo = int(off[10])
    seqs[lo:int(off[400])] = np.frombuffer(seqs[lo:int(off[400])].tobytes().lower(), dtype=np.uint8)  # JuncSeq is upper-cased
    max_len = int(np.diff(off.astype(np.int64)).max())
    d_seqs = torch.from_numpy(seqs).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_rc = torch.from_numpy(rc.astype(np.int32)).cuda()
    d_rec = torch.zeros(N * 48, dtype=torch.uint8, device="cuda")
    d_boff = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, cabi.NO_SUMMARY, d_rec.data_ptr(), 0, 0, st)
    total = eng.marshal_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rec.data_ptr(), N, d_boff.data_ptr(), 0, 0, st)
    d_blob = torch.zeros(total, dtype=torch.uint8, device="cuda")
    assert eng.marshal_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rec.data_ptr(), N, d_boff.data_ptr(), d_blob.data_ptr(), total, st) == total
    torch.cuda.synchronize()
    blob, boff = d_blob.cpu().numpy(), d_boff.cpu().numpy()
    rec = d_rec.cpu().numpy().view(cabi.RECORD_DTYPE)
    hblob, hboff = T.marshal_batch(rec, seqs, off, threads=4)
    assert np.array_equal(boff.astype(np.uint64), hboff) and np.array_equal(blob, hblob)
    assert (rec["junc_seq_len"] > 0).sum() > 100
    for i in list(range(0, N, 61)) + list(range(10, 60)):
        f = O.NewFragment("r-%d" % rc[i], seqs[int(off[i]):int(off[i + 1])].tobytes().decode())
        assert blob[int(boff[i]):int(boff[i + 1])].tobytes() == O.NewAlignment(f, om).MarshalBinary(), i
    with pytest.raises(T.TreatError):
        eng.marshal_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rec.data_ptr(), N, d_boff.data_ptr(), d_blob.data_ptr(), total - 1, st)
    # Storage.Search's record filter on the device (cmd/treat/storage.go:162-188, 257-286) == the host's search_records
    d_idx = torch.zeros(N, dtype=torch.int64, device="cuda")
    es0 = int(np.bincount(rec["edit_stop"] - rec["edit_stop"].min()).argmax() + rec["edit_stop"].min())
    for f in (dict(), dict(all=True), dict(has_mutation=True), dict(all=True, offset=17, limit=300), dict(edit_stop=es0), dict(all=True, has_alt=True),
              dict(all=True, junc_len=0, limit=40), dict(has_mutation=True, offset=10 ** 6), dict(all=True, edit_stop=es0, junc_end=int(rec["junc_end"][0]))):
        want_idx = T.search_records(rec, **f)
        n_dev = eng.search_device(d_rec.data_ptr(), N, d_idx.data_ptr(), N, st, **f)
        assert n_dev == len(want_idx), f
        assert np.array_equal(d_idx.cpu().numpy()[:n_dev].astype(np.uint64), want_idx), f
    assert len(T.search_records(rec, all=True)) > 1000 and eng.search_device(d_rec.data_ptr(), N, 0, 0, st, all=True) == len(T.search_records(rec, all=True))
    # `treat norm` on the device (cmd/treat/storage.go:801-878): Norm of the standard reads as float64 and inside their blobs ==
    # the host's normalize_records == scale * ReadCount in IEEE double (0.0 for reads with a mutation, whose blobs stay as loaded)
    std_total = int(rec["read_count"][rec["has_mutation"] == 0].astype(np.uint64).sum())
    scale = T.norm_scale(std_total, 4321.5)
    d_norm = torch.zeros(N, dtype=torch.float64, device="cuda")
    eng.normalize_device(d_rec.data_ptr(), N, scale, d_norm.data_ptr(), d_blob.data_ptr(), d_boff.data_ptr(), st)
    torch.cuda.synchronize()
    want = T.normalize_records(rec, scale, hblob, hboff)
    assert (rec["has_mutation"] != 0).sum() > 50 and (rec["read_count"] > 1).sum() > 50
    assert np.array_equal(d_norm.cpu().numpy().view(np.uint64), want.view(np.uint64))
    assert np.array_equal(want, np.where(rec["has_mutation"] == 0, scale * rec["read_count"].astype(np.float64), 0.0))
    assert np.array_equal(d_blob.cpu().numpy(), hblob)
    i = int(np.flatnonzero(rec["has_mutation"] == 0)[7])
    nb = T.UnmarshalBinary(hblob[int(hboff[i]):int(hboff[i + 1])].tobytes())
    assert nb["Norm"] == want[i] and want[i] != 0.0


def test_other_edit_base_is_isomorphic(eng, examples, tmp_path):
    """Swapping the letters A and T everywhere and editing on 'A' must give the same records: exercises
    the edit-base-dependent alphabet (codes for C, G, T; A stripped) against the oracle and itself."""
    swap = bytes.maketrans(b"ATat", b"TAta")
    tsrc = P.read_fasta(examples["templates.fa"])
    p = tmp_path / "swapped.fa"
    p.write_text("".join(">%s\n%s\n" % (i, s.encode().translate(swap).decode()) for i, s in tsrc))
    reads = P.read_fasta(examples["clones.fa"])
    t_T = T.NewTemplateFromFasta(examples["templates.fa"], T.FORWARD, "T")
    t_A = T.NewTemplateFromFasta(str(p), T.FORWARD, "a")
    o_A = O.NewTemplateFromFasta(str(p), O.FORWARD, "a")
    a_T = eng.NewAlignments(reads, t_T, False)
    sw = [(i, s.encode().translate(swap).decode()) for i, s in reads]
    a_A = eng.NewAlignments(sw, t_A, False)
    for (rid, seq), x, y in zip(sw, a_T, a_A):
        fx, fy = x.fields(), y.fields()
        fx["junc_seq"] = fx["junc_seq"].encode().translate(swap).decode()
        assert fx == fy, rid
        assert x.ops() == y.ops()
        of = O.NewFragment(rid, seq, O.FORWARD, "a")
        oa = O.NewAlignment(of, o_A)
        assert fy == oa.fields(), rid
        assert y.WriteTo(T.NewFragment(rid, seq, T.FORWARD, "a"), t_A, 80) == oa.WriteTo(of, o_A, 80)


def test_reverse_orientation(eng, examples):
    """REVERSE fragments (fragment.go:95-97: plain reversal) through the Go-shaped API."""
    tmpl = T.NewTemplateFromFasta(examples["test-templates.fa"], T.FORWARD, "t")
    om = O.NewTemplateFromFasta(examples["test-templates.fa"], O.FORWARD, "t")
    for rid, seq in P.read_fasta(examples["test-sample.fa"]):
        fr = T.NewFragment(rid, seq[::-1], T.REVERSE, "t")
        a = eng.NewAlignment(fr, tmpl, False)
        of = O.NewFragment(rid, seq[::-1], O.REVERSE, "t")
        oa = O.NewAlignment(of, om)
        assert a.fields() == oa.fields(), rid
        assert a.WriteTo(fr, tmpl, 80) == oa.WriteTo(of, om, 80)


def test_treat_align_binary(examples):
    """The treat_align executable (cmd/treat/align.go driver on the GPU) prints the oracle's bytes."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "treat_b200", "bin", "treat_align")
    out = subprocess.run([exe, "-t", examples["templates.fa"], "-f", examples["clones.fa"], "-b", "T"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout == O.cli_align(template=examples["templates.fa"], fragment=examples["clones.fa"], base="T")
    # the same tool written in C against include/treat_b200.h alone (examples/treat_align_capi.c: the calls a cgo binding makes)
    capi = os.path.join(os.path.dirname(exe), "treat_align_capi")
    for t, f, base, off in (("templates.fa", "clones.fa", "T", 0), ("templates.fa", "sample-1.fa", "T", 10), ("test-templates.fa", "test-sample.fa", "t", 0)):
        out = subprocess.run([capi, "-t", examples[t], "-f", examples[f], "-b", base, "--offset", str(off)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        assert out.stdout == O.cli_align(template=examples[t], fragment=examples[f], base=base, offset=off)
    out = subprocess.run([exe, "-1", "ATCTGTATGT", "-2", "ATTCGATTG", "-b", "T"], capture_output=True, text=True, timeout=120)
    assert out.stdout == "A-TCTGTA-TGT\nATTC-G-ATTG-\n\n"
    bad = subprocess.run([exe, "-1", "ATCTGTATGT", "-b", "T"], capture_output=True, text=True, timeout=120)
    assert bad.returncode != 0 and "Please provide 2 fragments to align" in bad.stderr


def test_zz_kernel_bounds_flags_clean(eng):
    """Runs last: no kernel index computation left its bounds during this module (meaningful when the
    library was built with TREAT_B200_CHECKS=1; a normal build always reports zero)."""
    flags, compiled = eng.debug_flags()
    assert flags == 0, "kernel bounds-check flags: 0x%x (checks compiled: %s)" % (flags, compiled)


def test_zz_checked_build_bounds_clean():
    """compute-sanitizer is closed on the GPU pool.  Substitute: a second library built with -DTREAT_B200_CHECKS (bounds checks
    compiled into the kernels' data-dependent index computations: band window, scratch, ops array, site maps, pair area)
    runs every kernel variant against the oracle in its own process; no check may fire."""
    import subprocess
    import sys
    from treat_b200 import _build
    lib = _build.build_variant("checked", ["-DTREAT_B200_CHECKS"])  # up to date after build(): nothing to compile then
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TREAT_B200_LIB=lib, TREAT_B200_EXPECT_CHECKS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "sanitize_smoke.py")], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "checks_compiled 1" in out.stdout and "debug_flags 0x0" in out.stdout


def test_tie_orders_match_oracle(eng, examples, tmp_path):
    """treat_b200_set_tie_order: the three traceback orders the reference's goldens allow (Up > Left > Diagonal as nwalgo
    publishes it, Up > Diagonal > Left, Left > Up > Diagonal) against the oracle's switch, on reads with many ties
    (indels next to repeats) and on the synthetic sets; records and every ops row."""
    rng = random.Random(4711)
    srcs = [s for _, s in P.read_fasta(examples["templates.fa"])]
    recs = [("t-%d" % (i + 1), _mutate(rng, rng.choice(srcs))) for i in range(600)]
    # homopolymer / dinucleotide repeats make Left-vs-Diagonal and Up-vs-Left ties certain
    recs += [("rep-%d" % i, "AAAAGGGGAAAACCCC"[: 4 + i % 12] * (1 + i % 5) + "ACG" * (i % 7)) for i in range(120)]
    fb = T.FastaBatch.from_records(recs)
    tm = T.NewTemplateFromFasta(examples["templates.fa"], T.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    differ = 0
    base_ops = None
    try:
        for order in ("ULD", "UDL", "LUD"):
            O.set_tie_order(order)
            eng.set_tie_order(order)
            eng.SetTemplate(tm)
            for extra in (0, cabi.DP_EVERY_READ, cabi.NO_BAND, cabi.ONE_READ_PER_WARP | cabi.DP_EVERY_READ, cabi.FULL_TRACEBACK_STORE):
                grec, gops = eng.align_batch(fb.seqs, fb.offsets, fb.read_counts, cabi.WANT_OPS | extra)
                orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, ops_stride=gops.shape[1])
                assert compare_records(grec, orec) == [], (order, extra)
                assert np.array_equal(gops, oops), (order, extra)
            if base_ops is None:
                base_ops = gops.copy()
            else:
                differ += int((gops != base_ops).any(axis=1).sum())
            # the synthetic RPS12 and long-gene sets (paired kernel with two words per step, scratch fills)
            for t, n, tr in ((synth.rps12_template(), 6000, 0.1), (synth.long_template(300, 8), 1500, 0.3)):
                omt, tmt = _oracle_template(t, tmp_path)
                eng.SetTemplate(tmt)
                seqs, off, rc = synth.reads(t, n, seed=synth.SEED_BASE + 90, p_trunc=tr, p_sub_base=0.01, p_indel_read=0.3)
                grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
                orec, oops, _ = O.align_batch(omt, seqs, off, rc, ops_stride=gops.shape[1])
                assert compare_records(grec, orec) == [], order
                assert np.array_equal(gops, oops), order
        assert differ > 0  # the orders do give different alignments on this set: the switch is not a no-op
    finally:
        O.set_tie_order("ULD")
        eng.set_tie_order("ULD")


def test_band_and_pure_deletion_paths(eng, examples, tmp_path):
    """k_band (band_kernel.cu) in front of the full-matrix kernels: reads whose bases are a subsequence of the template's
    (truncated ends, deletions inside, repeats where only the LEFTMOST embedding is nwalgo's), reads a few edits away
    (banded fill + certificate) and reads the certificate must refuse -- records and every ops row against the oracle,
    for the three tie orders, with the path counters showing that all three routes were taken."""
    rng = random.Random(99)
    t = synth.rps12_template()
    om, tm = _oracle_template(t, tmp_path)
    eb = "T"
    fe = t.raw_sequences()[0]
    bases = t.bases.decode()

    def with_counts(bs):
        # a read with the given bases and random edit-base runs
        return "".join(eb * rng.choice((0, 0, 1, 2, 5)) + b for b in bs) + eb * rng.randint(0, 3)

    recs = []
    for i in range(300):  # pure deletions: drop random bases, cut the ends
        bs = list(bases)
        for _ in range(rng.randint(0, 6)):
            del bs[rng.randrange(len(bs))]
        lo, hi = rng.choice((0, 0, rng.randint(0, 60))), rng.choice((0, 0, rng.randint(0, 60)))
        bs = bs[lo:len(bs) - hi]
        recs.append(("del-%d" % i, with_counts(bs)))
    for i in range(300):  # a few substitutions / indels: certified by the band
        bs = list(bases)
        for _ in range(rng.randint(1, 8)):
            k, pos = rng.randrange(3), rng.randrange(len(bs))
            if k == 0:
                del bs[pos]
            elif k == 1:
                bs.insert(pos, rng.choice("ACG"))
            else:
                bs[pos] = rng.choice("ACGN")
        recs.append(("edit-%d" % i, with_counts(bs)))
    for i in range(200):  # far from the template: the certificate fails, or the length difference leaves the band
        bs = list(bases)
        for _ in range(rng.randint(15, 60)):
            k, pos = rng.randrange(3), rng.randrange(len(bs))
            if k == 0:
                del bs[pos]
            elif k == 1:
                bs.insert(pos, rng.choice("ACG"))
            else:
                bs[pos] = rng.choice("ACG")
        recs.append(("far-%d" % i, with_counts(bs)))
    recs += [("fe", fe), ("one", "A"), ("two", "CG"), ("first", bases[0]), ("last", bases[-1]), ("allT", eb * 7)]
    fb = T.FastaBatch.from_records(recs)
    try:
        for order in ("ULD", "UDL", "LUD"):
            O.set_tie_order(order)
            eng.set_tie_order(order)
            eng.SetTemplate(tm)
            orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, ops_stride=None)
            eng.dp_stats()
            for extra in (0, cabi.DP_EVERY_READ, cabi.NO_BAND):
                grec, gops = eng.align_batch(fb.seqs, fb.offsets, fb.read_counts, cabi.WANT_OPS | extra)
                w = min(gops.shape[1], oops.shape[1])
                assert compare_records(grec, orec) == [], (order, extra)
                assert np.array_equal(gops[:, :w], oops[:, :w]), (order, extra)
                sub, band, full = eng.dp_stats()
                if extra == 0:
                    assert sub > 200 and band > 200 and full > 50, (sub, band, full)
                elif extra == cabi.DP_EVERY_READ:
                    assert sub == 0 and band > 400 and full > 50, (sub, band, full)
                else:
                    assert sub == band == full == 0
        # a template with repeats: many embeddings of a truncated read exist, nwalgo's is the leftmost
        rep = "ACG" * 30 + "AACCGG" * 10 + "ACG" * 10
        tpath = tmp_path / "rep.fa"
        tpath.write_text(">FE\n%s\n>PE\n%s\n" % ("T".join(rep), rep))
        tm2 = T.NewTemplateFromFasta(str(tpath), T.FORWARD, "T")
        om2 = O.NewTemplateFromFasta(str(tpath), O.FORWARD, "T")
        recs2 = [("r-%d" % i, rep[rng.randint(0, 80):len(rep) - rng.randint(0, 80)]) for i in range(200)]
        recs2 += [("s-%d" % i, "".join(c for c in rep if rng.random() > 0.1)) for i in range(200)]
        fb2 = T.FastaBatch.from_records(recs2)
        for order in ("ULD", "UDL", "LUD"):
            O.set_tie_order(order)
            eng.set_tie_order(order)
            eng.SetTemplate(tm2)
            orec, oops, _ = O.align_batch(om2, fb2.seqs, fb2.offsets, fb2.read_counts, ops_stride=None)
            for extra in (0, cabi.DP_EVERY_READ, cabi.NO_BAND):
                grec, gops = eng.align_batch(fb2.seqs, fb2.offsets, fb2.read_counts, cabi.WANT_OPS | extra)
                w = min(gops.shape[1], oops.shape[1])
                assert compare_records(grec, orec) == [], (order, extra)
                assert np.array_equal(gops[:, :w], oops[:, :w]), (order, extra)
    finally:
        O.set_tie_order("ULD")
        eng.set_tie_order("ULD")


def test_band_min_reads_threshold(eng, tmp_path):
    """treat_b200_set_band_min_reads: below the threshold a launch sends its band candidates to the full-matrix kernel (the
    banded fill's fixed latency does not pay for a handful of reads) and keeps the pure-deletion route for truncated reads
    only; results are the same either way."""
    t = synth.long_template(300, 8)
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    seqs, off, rc = synth.reads(t, 6000, seed=synth.SEED_BASE + 61, p_trunc=0.3, p_sub_base=0.005, p_indel_read=0.1)
    orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=8, ops_stride=None)
    try:
        got = {}
        for thr in (0, 1 << 30):
            eng.set_band_min_reads(thr)
            eng.dp_stats()
            grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS | cabi.NO_SUMMARY)
            w = min(gops.shape[1], oops.shape[1])
            assert compare_records(grec, orec) == [] and np.array_equal(gops[:, :w], oops[:, :w]), thr
            got[thr] = eng.dp_stats()
        assert got[0][1] > 100 and got[0][0] > 500          # banded fill and pure deletions at any size
        assert got[1 << 30][1] == 0 and got[1 << 30][0] > 500 and got[1 << 30][2] > got[0][2]  # no band: more reads in the full matrix
    finally:
        eng.set_band_min_reads(0)


def _collapse_case(eng, t, tmp_path, base_n, total_n, seed, p_trunc=0.0, lower_frac=0.2):
    """A batch with many exact duplicates (sampled with replacement from base_n distinct reads, some of them in lower case)
    against the oracle-side collapse: a dict of upper-cased sequences in order of first occurrence, ReadCounts summed."""
    om, tm = _oracle_template(t, tmp_path)
    eng.SetTemplate(tm)
    bseqs, boff, _ = synth.reads(t, base_n, seed=seed, p_trunc=p_trunc)
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, base_n, size=total_n)
    rc = rng.integers(1, 50, size=total_n).astype(np.uint32)
    parts = []
    for i, q in enumerate(pick):
        s = bseqs[int(boff[q]):int(boff[q + 1])].tobytes()
        parts.append(s.lower() if rng.random() < lower_frac else s)
    seqs = np.frombuffer(b"".join(parts), dtype=np.uint8)
    off = np.zeros(total_n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(p_) for p_ in parts])
    # oracle-side collapse
    seen, first, member, sums = {}, [], np.zeros(total_n, dtype=np.uint32), []
    for i, s in enumerate(parts):
        k = s.upper()
        u = seen.get(k)
        if u is None:
            u = seen[k] = len(first)
            first.append(i)
            sums.append(0)
        member[i] = u
        sums[u] = (sums[u] + int(rc[i])) & 0xffffffff
    first = np.array(first, dtype=np.uint32)
    useqs = np.frombuffer(b"".join(parts[i] for i in first), dtype=np.uint8)
    uoff = np.zeros(len(first) + 1, dtype=np.uint64)
    uoff[1:] = np.cumsum([len(parts[i]) for i in first])
    urc = np.array(sums, dtype=np.uint32)
    eng.reset_summary()
    grec, gops, gfirst, gmember = eng.collapse_align_batch(seqs, off, rc, cabi.WANT_OPS)
    assert len(grec) == len(first) and len(first) <= base_n
    assert np.array_equal(gfirst, first) and np.array_equal(gmember, member)
    orec, oops, _ = O.align_batch(om, useqs, uoff, urc, nthreads=8, ops_stride=gops.shape[1])
    assert compare_records(grec, orec) == []
    assert np.array_equal(gops, oops)
    assert np.array_equal(eng.summary(), summary_from_records(orec, t.m, 0, om.EditStop))
    # the same records as aligning every read and picking the representatives (ReadCount aside)
    arec, _ = eng.align_batch(seqs, off, rc, cabi.NO_SUMMARY)
    for f in ("edit_stop", "junc_start", "junc_end", "junc_len", "score", "mismatches", "indel", "alt_editing", "junc_seq_len"):
        assert np.array_equal(arec[f], grec[f][member]), f
    return len(first) / total_n


def test_collapse_exact_duplicates(eng, tmp_path):
    """fastx_collapser on the device: >= 50 % duplicates, mixed case; uniques, members, summed ReadCounts, records,
    ops and summary equal the oracle's on the collapsed set."""
    frac = _collapse_case(eng, synth.rps12_template(), tmp_path, 9000, 40000, synth.SEED_BASE + 61, p_trunc=0.1)
    assert frac < 0.5
    _collapse_case(eng, synth.long_template(300, 8), tmp_path, 700, 3000, synth.SEED_BASE + 62, p_trunc=0.2)
    # every read distinct, one read, and the wide path (a template base outside the packed alphabet)
    assert _collapse_case(eng, synth.rps12_template(), tmp_path, 3000, 3000, synth.SEED_BASE + 63, lower_frac=0.0) <= 1.0
    _collapse_case(eng, synth.rps12_template(), tmp_path, 1, 5, synth.SEED_BASE + 64)
    p = tmp_path / "w.fa"
    p.write_text(">fe\nACNGTTAGC%s\n>pe\nACNGAGC%s\n" % ("AG" * 20, "AG" * 20))
    tm, om = T.NewTemplateFromFasta(str(p), T.FORWARD, "T"), O.NewTemplateFromFasta(str(p), O.FORWARD, "T")
    eng.SetTemplate(tm)
    reads = [b"ACNGTTAGC" + b"AG" * 20, b"acngttagc" + b"ag" * 20, b"ACNG" + b"T" * 300 + b"AGC" + b"AG" * 20, b"ACNGTTAGC" + b"AG" * 20]
    seqs = np.frombuffer(b"".join(reads), dtype=np.uint8)
    off = np.array([0] + list(np.cumsum([len(r) for r in reads])), dtype=np.uint64)
    grec, gops, gfirst, gmember = eng.collapse_align_batch(seqs, off, None, cabi.WANT_OPS)
    assert list(gfirst) == [0, 2] and list(gmember) == [0, 0, 1, 0] and list(grec["read_count"]) == [3, 1]
    useqs = np.frombuffer(reads[0] + reads[2], dtype=np.uint8)
    uoff = np.array([0, len(reads[0]), len(reads[0]) + len(reads[2])], dtype=np.uint64)
    orec, oops, _ = O.align_batch(om, useqs, uoff, np.array([3, 1], dtype=np.uint32), ops_stride=gops.shape[1])
    assert compare_records(grec, orec) == [] and np.array_equal(gops, oops)


def test_frontier_kernel_edges(eng, tmp_path):
    """k_frontier (reads finished from two string compares on the raw characters): FE / PE themselves, edit bases added
    or removed at either end, lower case, foreign characters, junctions of every length, reads at every alignment in
    memory and longer than one pass -- records, ops and summary equal to the oracle's and to the run without the kernel."""
    for t in (synth.rps12_template(), synth.long_template()):
        om, tm = _oracle_template(t, tmp_path)
        eng.SetTemplate(tm)
        raws = t.raw_sequences()
        F, P = raws[0], raws[1]
        rng = random.Random(20261020)
        cases = [F, P, F.lower(), "T" + F, "TT" + P, F + "T", P + "TTT", F[1:], P[:-1], "", "T", "TTTT", P[:50] + F[len(F) - 200:],
                 P[:120] + "TTTTT" + F[len(F) - 60:], F.replace("A", "N", 1), P[:30] + "\x00" + P[31:], P[:40] + " " + F[len(F) - 100:],
                 "T" * 300 + F, F + "T" * 700, P[:10] + "T" * 260 + F[len(F) - 30:], "t" * 5 + P.lower()] + raws[2:]
        es = t.edit_site
        # one substituted base (k_frontier steps over it once): in the prefix, the suffix, the junction, at the very ends, lower
        # case, two of them, one of them outside the alphabet, the edit base in place of a base
        swap = {"A": "C", "C": "G", "G": "A"}

        def sub(sq, k):
            idx = [i for i, ch in enumerate(sq) if ch != "T"][k]
            return sq[:idx] + swap[sq[idx]] + sq[idx + 1:]
        mid = P[:120] + "TT" + F[len(F) - 150:]
        for base in (F, P, mid, P[:60] + "TTTT" + F[len(F) - 220:]):
            nb = sum(ch != "T" for ch in base)
            cases += [sub(base, k) for k in (0, 1, 5, nb // 3, nb // 2, nb - 2, nb - 1)]
            cases += [sub(sub(base, 5), nb // 2), sub(base, 7).lower(), sub(base, 30).replace("A", "N", 1), sub(base, 3) + "T", "TT" + sub(base, nb - 4)]
            k = [i for i, ch in enumerate(base) if ch != "T"][20]
            cases += [base[:k] + "T" + base[k + 1:], base[:k] + "N" + base[k + 1:], base[:k] + "n" + base[k + 1:]]
        for _ in range(1500):
            # a read of the model: FE below a frontier, a junction with arbitrary counts, PE above it ...
            ess = rng.randint(-1, t.m)
            jl = min(rng.choice([0, 0, 1, 2, 3, 5, 8, 13, 30, 80]), t.m - ess)
            cnt = [int(es[0, i]) if (t.m - i) <= ess else (rng.randint(0, 9) if (t.m - i) <= ess + jl else int(es[1, i])) for i in range(t.m + 1)]
            bs = list(t.bases.decode())
            u = rng.random()
            if u < 0.08:    # ... sometimes with a substitution, a foreign character, a deleted or an inserted base
                bs[rng.randrange(t.m)] = rng.choice("ACGN")
            elif u < 0.12:
                k = rng.randrange(t.m)
                del bs[k]
                cnt[k + 1] += cnt[k]
                del cnt[k]
            elif u < 0.16:
                k = rng.randrange(t.m)
                bs.insert(k, rng.choice("ACG"))
                cnt.insert(k, 0)
            sq = "".join("T" * c + b for c, b in zip(cnt, bs)) + "T" * cnt[-1]
            if rng.random() < 0.1:
                sq = sq.lower()
            cases.append(sq)
        raw = "".join(cases).encode("latin-1")
        off = np.zeros(len(cases) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cases])
        seqs = np.frombuffer(raw, dtype=np.uint8).copy()
        rc = np.arange(1, len(cases) + 1, dtype=np.uint32)
        for flags, exclude in ((cabi.WANT_OPS | cabi.EXCLUDE_SNPS, True), (cabi.WANT_OPS, False)):
            eng.reset_summary()
            grec, gops = eng.align_batch(seqs, off, rc, flags)
            gsum = eng.summary()
            orec, oops, _ = O.align_batch(om, seqs, off, rc, exclude_snps=exclude, nthreads=8, ops_stride=gops.shape[1])
            assert compare_records(grec, orec) == []
            assert np.array_equal(gops, oops)
            assert np.array_equal(gsum, summary_from_records(orec, t.m, 0, om.EditStop))
            eng.reset_summary()
            nrec, nops = eng.align_batch(seqs, off, rc, flags | cabi.NO_FRONTIER)
            assert np.array_equal(nrec, grec) and np.array_equal(nops, gops) and np.array_equal(eng.summary(), gsum)
        # every start alignment of the first read inside its 64-byte block
        for shift in (1, 3, 15, 16, 17, 31, 33, 47, 63):
            s2 = np.concatenate([np.full(shift, ord("G"), dtype=np.uint8), seqs])
            o2 = np.concatenate([[0], off + shift]).astype(np.uint64)
            r2 = np.concatenate([[7], rc]).astype(np.uint32)
            srec, sops = eng.align_batch(s2, o2, r2, flags | cabi.NO_SUMMARY)
            assert np.array_equal(srec[1:], grec) and np.array_equal(sops[1:, :gops.shape[1]], gops[:, :sops.shape[1]])
