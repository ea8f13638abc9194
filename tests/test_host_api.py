"""Host-side mirror of the Go API (treat_b200/csrc/host) on CPU: the reference's own unit tests
(fragment_test.go, template_test.go) plus differential checks against the oracle."""
import ctypes as C
import random

import numpy as np
import pytest

import treat_b200 as T
from treat_b200 import cabi
from oracle import oracle as O
from oracle import pyref as P
from helpers import summary_from_records


def test_fragment(fx):
    """TestFragment, fragment_test.go:25-50."""
    for s in fx["fragment_test"]["seqs"]:
        frag = T.NewFragment("id", s, T.FORWARD, "t")
        assert frag.String() == s.upper()
        assert frag.ReadCount == 1
        frag = T.NewFragment("1-143", s, T.REVERSE, "t")
        assert frag.String() == s[::-1].upper()
        assert frag.ReadCount == 143


def test_parse_read_counts(fx):
    """TestParseReadCounts, fragment_test.go:52-76."""
    for rid, count in fx["parse_read_counts"]["table"].items():
        frag = T.NewFragment(rid, "CTGCTG", T.FORWARD, "t")
        assert frag.ReadCount == count, rid
    for rid in ("a-99999999999999999999", "a-4294967297", "x_y-12", "-", "a-", "a--3", "a b-4", "  a_7  z-9"):
        assert T.parseMergeCount(rid) == O.parseMergeCount(rid), rid


def test_template(fx, examples):
    """TestTemplate, template_test.go:26-70."""
    tt = fx["template_test"]
    tmpl = T.NewTemplateFromFasta(examples["templates.fa"], T.FORWARD, "t")
    assert tmpl.Size() == tt["rps12_size"]
    full = T.NewFragment("full", tt["bad"]["full"], T.FORWARD, "t")
    pre = T.NewFragment("pre", tt["bad"]["pre"], T.FORWARD, "t")
    with pytest.raises(ValueError, match="Full and Pre templates must have the same non-edit bases"):
        T.NewTemplate(full, pre, None, None)
    full = T.NewFragment("full", tt["good"]["full"], T.FORWARD, "t")
    pre = T.NewFragment("pre", tt["good"]["pre"], T.FORWARD, "t")
    tmpl = T.NewTemplate(full, pre, None, None)
    rebuilt = "".join(tmpl.EditBase * b + (tmpl.Bases[i] if i < len(tmpl.Bases) else "")
                      for i, b in enumerate(tmpl.EditSite[1]))
    assert rebuilt == pre.String()
    with pytest.raises(ValueError, match="Please specify the alt regions"):
        T.NewTemplate(full, pre, None, [None, None])


def test_template_matches_oracle(examples):
    for name, base in (("templates.fa", "T"), ("test-templates.fa", "t"), ("simple-templates.fa", "T")):
        for offset in (0, 10):
            a = T.NewTemplateFromFasta(examples[name], T.FORWARD, base)
            b = O.NewTemplateFromFasta(examples[name], O.FORWARD, base)
            a.SetOffset(offset)
            b.SetOffset(offset)
            assert (a.Size(), a.Len(), a.EditStop, a.EditOffset, a.EditBase) == (b.Size(), b.Len(), b.EditStop, b.EditOffset, b.EditBase)
            assert a.Bases == b.Bases and a.EditSite == b.EditSite and a.BaseIndex == b.BaseIndex
            assert a.AltRegion == b.AltRegion
            assert [a.Max(i) for i in range(a.Len())] == [b.Max(i) for i in range(b.Len())]


def test_template_errors(tmp_path):
    p = tmp_path / "one.fa"
    p.write_text(">only\nACGT\n")
    with pytest.raises(ValueError, match="Must provide at least 2 templates"):
        T.NewTemplateFromFasta(str(p), T.FORWARD, "T")
    with pytest.raises(ValueError, match="Invalid FASTA file"):
        T.NewTemplateFromFasta(str(tmp_path / "missing.fa"), T.FORWARD, "T")
    p3 = tmp_path / "noalt.fa"
    p3.write_text(">fe\nATTC\n>pe\nAC\n>alt without region\nATC\n")
    with pytest.raises(ValueError, match="Please specify the alt regions"):
        T.NewTemplateFromFasta(str(p3), T.FORWARD, "T")


def test_fragment_fuzz_vs_oracle():
    rng = random.Random(7)
    for _ in range(300):
        s = "".join(rng.choice("ACGTTTacgtNn-x") for _ in range(rng.randint(0, 60)))
        base = rng.choice("TtAc")
        o = rng.choice((T.FORWARD, T.REVERSE))
        a, b = T.NewFragment("r_%d" % rng.randint(0, 99), s, o, base), None
        b = O.NewFragment(a.Name, s, o, base)
        assert (a.Bases, a.EditSite, a.EditBase, a.ReadCount, a.String()) == (b.Bases, b.EditSite, b.EditBase, b.ReadCount, b.String())


def test_fasta_reader(examples, tmp_path):
    for name in ("clones.fa", "templates.fa", "test-sample.fa"):
        fb = T.read_fasta(examples[name])
        want = P.read_fasta(examples[name])
        assert fb.ids == [i for i, _ in want]
        assert [fb.seq(i).decode() for i in range(len(fb))] == [s for _, s in want]
        assert fb.read_counts.tolist() == [O.parseMergeCount(i) for i, _ in want]
    p = tmp_path / "crlf.fa"
    p.write_bytes(b"junk\n>a-3 x \r\nAC\r\nGT\r\n>b\nTT")
    fb = T.read_fasta(str(p))
    assert fb.ids == ["a-3 x", "b"] and [fb.seq(i) for i in range(2)] == [b"ACGT", b"TT"]
    assert fb.read_counts.tolist() == [3, 1]


def _oracle_record_as_product(orec):
    rec = np.zeros(len(orec), dtype=cabi.RECORD_DTYPE)
    for f in orec.dtype.names:
        if f in rec.dtype.names and f not in ("pad",):
            rec[f] = orec[f]
    return rec


def test_format_and_marshal_vs_oracle(examples):
    """WriteTo / MarshalBinary text assembly on the host, fed with the oracle's record + ops
    (so this host logic is covered without a GPU)."""
    for tname, rname, base, offset in (("templates.fa", "clones.fa", "T", 0), ("test-templates.fa", "test-sample.fa", "t", 0),
                                       ("templates.fa", "sample-1.fa", "T", 10)):
        tm = T.NewTemplateFromFasta(examples[tname], T.FORWARD, base)
        tm.SetOffset(offset)
        om = O.NewTemplateFromFasta(examples[tname], O.FORWARD, base)
        om.SetOffset(offset)
        fb = T.read_fasta(examples[rname])
        orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts)
        rec = _oracle_record_as_product(orec)
        for i in range(len(fb)):
            seq = fb.seq(i)
            name = fb.ids[i].encode()
            need = cabi.lib().treat_b200_format_alignment(tm._h, rec[i:i + 1].ctypes.data, oops[i].ctypes.data, name, seq, len(seq), 80, None, 0)
            buf = C.create_string_buffer(int(need))
            cabi.lib().treat_b200_format_alignment(tm._h, rec[i:i + 1].ctypes.data, oops[i].ctypes.data, name, seq, len(seq), 80, buf, need)
            of = O.NewFragment(fb.ids[i], seq, O.FORWARD, base)
            oa = O.NewAlignment(of, om, False)
            assert buf.raw[:need].decode("latin-1") == oa.WriteTo(of, om, 80), fb.ids[i]
            need = cabi.lib().treat_b200_marshal_record(rec[i:i + 1].ctypes.data, seq, len(seq), None, 0)
            mb = (C.c_uint8 * int(need))()
            cabi.lib().treat_b200_marshal_record(rec[i:i + 1].ctypes.data, seq, len(seq), mb, need)
            assert bytes(mb) == oa.MarshalBinary(), fb.ids[i]


def test_print_alignment():
    a, b = "A" * 170, "C" * 170
    assert T.PrintAlignment(a, b, 80) == P.PrintAlignment(a, b, 80)


def test_marshal_batch_vs_oracle(examples):
    """`treat load` record path: MarshalBinary for a batch (storage.go:758-770, align.go:316-335)."""
    for tname, rname, base in (("templates.fa", "clones.fa", "T"), ("test-templates.fa", "test-sample.fa", "t")):
        om = O.NewTemplateFromFasta(examples[tname], O.FORWARD, base)
        fb = T.read_fasta(examples[rname])
        orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, want_ops=False)
        rec = _oracle_record_as_product(orec)
        for threads in (1, 3):
            blob, off = T.marshal_batch(rec, fb.seqs, fb.offsets, threads)
            assert int(off[-1]) == len(blob)
            for i in range(len(fb)):
                oa = O.NewAlignment(O.NewFragment(fb.ids[i], fb.seq(i), O.FORWARD, base), om)
                assert blob[int(off[i]):int(off[i + 1])].tobytes() == oa.MarshalBinary(), fb.ids[i]
    blob, off = T.marshal_batch(rec[:0], fb.seqs, fb.offsets[:1])
    assert len(blob) == 0 and off.tolist() == [0]


def test_normalize_vs_oracle(examples):
    """`treat norm` (cmd/treat/norm.go:25-73, cmd/treat/storage.go:801-878): default norm, per-sample scale, Norm of the
    standard reads as float64 and inside their MarshalBinary blobs -- against the pure-Python restatement, bit for bit."""
    pt = P.template_from_fasta(examples["templates.fa"], P.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    files = ("clones.fa", "sample-1.fa", "sample-2.fa") if "sample-2.fa" in examples else ("clones.fa", "sample-1.fa")
    for exclude in (False, True):
        samples, batches = [], []
        for rname in files:
            recs = P.read_fasta(examples[rname])
            recs = recs + [(i + "x", sq) for i, sq in recs[3:9]]  # a few read counts above one
            fb = T.FastaBatch.from_records(recs)
            orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, exclude_snps=exclude, want_ops=False)
            samples.append([P.NewAlignment(P.Fragment(i, sq, P.FORWARD, "T"), pt, exclude) for i, sq in recs])
            batches.append((fb, _oracle_record_as_product(orec)))
        totals = [summary_from_records(rec, om.Len() - 1, 0, om.EditStop)[1] for _, rec in batches]  # "Std", weighted by ReadCount
        for alns, tot in zip(samples, totals):
            assert int(tot) == sum(int(a.ReadCount) for a in alns if a.HasMutation == 0)
        for norm in (0.0, 1000.0, 12345.678):
            gnorm = norm if norm else T.norm_default(totals)
            if not norm:
                assert gnorm == P.normalize_default(samples)
            for (fb, rec), alns, tot in zip(batches, samples, totals):
                scale = T.norm_scale(int(tot), gnorm)
                assert scale == P.normalize_sample(alns, gnorm)
                blob, off = T.marshal_batch(rec, fb.seqs, fb.offsets, 2)
                got = T.normalize_records(rec, scale, blob, off)
                for i, a in enumerate(alns):
                    assert got[i] == (a.Norm if a.HasMutation == 0 else 0.0)
                    assert blob[int(off[i]):int(off[i + 1])].tobytes() == a.MarshalBinary(), (i, exclude, norm)
                for a in alns:
                    a.Norm = 0.0
    assert T.norm_scale(0, 500.0) == 1.0  # storage.go:840-843: an empty sample keeps scale 1


def test_search_records_vs_oracle(examples):
    """The record filter of Storage.Search (cmd/treat/storage.go:162-188, 257-286) on a batch's records against the
    pure-Python restatement walking Alignment objects: every combination of the fields the handlers set."""
    pt = P.template_from_fasta(examples["templates.fa"], P.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    recs = P.read_fasta(examples["clones.fa"]) + P.read_fasta(examples["sample-1.fa"])
    recs = recs + [(i + "y", sq) for i, sq in recs]
    fb = T.FastaBatch.from_records(recs)
    for exclude in (False, True):
        orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, exclude_snps=exclude, want_ops=False)
        rec = _oracle_record_as_product(orec)
        alns = [P.NewAlignment(P.Fragment(i, sq, P.FORWARD, "T"), pt, exclude) for i, sq in recs]
        es = sorted(set(int(a.EditStop) for a in alns))[:3] + [-1]
        jl = sorted(set(int(a.JuncLen) for a in alns))[:2] + [-1]
        alt = sorted(set(int(a.AltEditing) for a in alns))
        n = 0
        for kw in (dict(), dict(all=True), dict(has_mutation=True), dict(has_alt=True, all=True), dict(all=True, alt_region=max(alt)),
                   dict(all=True, alt_region=300), dict(all=True, offset=3, limit=5), dict(offset=1000), dict(limit=1), dict(all=True, has_alt=True, limit=2, offset=1)):
            for e in es:
                for j in jl:
                    for je in (-1, int(alns[0].JuncEnd)):
                        f = dict(kw, edit_stop=e, junc_len=j, junc_end=je)
                        assert T.search_records(rec, **f).tolist() == P.search(alns, **f), f
                        n += 1
        assert n > 100 and len(T.search_records(rec, all=True, has_alt=False)) > 10
        if max(alt) > 0:
            assert len(T.search_records(rec, all=True, has_alt=True)) > 0


def test_search_text_vs_oracle_and_readme(examples):
    """`treat search` rows (cmd/treat/search.go:30-89): the README's CSV block (README.rst:175-178) byte for byte, and tab /
    comma output with norms, alt editing, mutants and names that need Go's csv quoting against the pure-Python restatement."""
    pt = P.template_from_fasta(examples["templates.fa"], P.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    recs = P.read_fasta(examples["sample-1.fa"])
    fb = T.FastaBatch.from_records(recs)
    orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, want_ops=False)
    rec = _oracle_record_as_product(orec)
    idx = T.search_records(rec, limit=2)  # (the README's sample held other reads behind these two; examples/sample-1.fa has one more)
    norm = T.normalize_records(rec, 1.0)
    readme = ("gene,sample,norm,read_count,alt_editing,has_mutation,edit_stop,junc_end,junc_len,junc_seq\n"
              "RPS12,sample-1,10.0000,10,0,0,137,143,6,ATATAATATTTTTG\n"
              "RPS12,sample-1,9.0000,9,0,0,95,123,28,TTCGGTATTTGTTTTATGTTATTATATGAGTCCGCGATTGCCCAGCTCTG\n")
    assert T.search_text("RPS12", "sample-1", rec, fb.seqs, fb.offsets, idx, norm, csv=True) == readme
    recs = P.read_fasta(examples["clones.fa"]) + recs
    fb = T.FastaBatch.from_records(recs)
    for exclude in (False, True):
        orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, exclude_snps=exclude, want_ops=False)
        rec = _oracle_record_as_product(orec)
        alns = [P.NewAlignment(P.Fragment(i, sq, P.FORWARD, "T"), pt, exclude) for i, sq in recs]
        P.normalize_sample(alns, 100000.0 if not exclude else 4321.0987)
        scale = T.norm_scale(sum(int(a.ReadCount) for a in alns if a.HasMutation == 0), 100000.0 if not exclude else 4321.0987)
        norm = T.normalize_records(rec, scale)
        for gene, sample in (("RPS12", "sample-1"), ("gene,with comma", " leading blank"), ('a"quote', "tab\there"), ("", "\\.")):
            for kw in (dict(all=True), dict(all=True, has_alt=True), dict(has_mutation=True, limit=3)):
                idx = T.search_records(rec, **kw)
                assert idx.tolist() == P.search(alns, **kw)
                for csv in (False, True):
                    for header in (True, False):
                        assert T.search_text(gene, sample, rec, fb.seqs, fb.offsets, idx, norm, csv=csv, header=header) == \
                            P.search_text(gene, sample, alns, idx.tolist(), csv=csv, header=header), (gene, sample, kw, csv)
    assert T.search_text("g", "s", rec, fb.seqs, fb.offsets, [], None, csv=True, header=False) == ""


def test_hist_norm_vs_oracle(examples):
    """The three histogram series of the web UI (highChartHist, cmd/treat/handlers.go:162-299) for one sample: Norm-weighted,
    filtered like a search, summed in record order -- bit for bit against the pure-Python restatement; and the device
    summary's ReadCount-weighted histograms are the unfiltered form of the same series."""
    pt = P.template_from_fasta(examples["templates.fa"], P.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    recs = P.read_fasta(examples["clones.fa"]) + P.read_fasta(examples["sample-1.fa"])
    recs = recs + [(i + "z", sq) for i, sq in recs[::2]]
    fb = T.FastaBatch.from_records(recs)
    orec, _, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts, want_ops=False)
    rec = _oracle_record_as_product(orec)
    alns = [P.NewAlignment(P.Fragment(i, sq, P.FORWARD, "T"), pt, False) for i, sq in recs]
    scale = P.normalize_sample(alns, 98765.4321)
    norm = T.normalize_records(rec, scale)
    for a, v in zip(alns, norm):  # mutants keep Norm 0 on both sides
        assert (a.Norm if a.HasMutation == 0 else 0.0) == v
    lo, hi = -1, om.Len() + 1
    for kw in (dict(), dict(all=True), dict(has_mutation=True), dict(all=True, has_alt=True), dict(edit_stop=int(alns[0].EditStop))):
        for which in (0, 1, 2):
            got = T.hist_norm(rec, norm, om.EditStop, which, lo, hi, **kw)
            want = P.hist_norm(alns, which, om.EditStop, **kw)
            assert {v: got[v - lo] for v in range(lo, hi + 1) if got[v - lo] != 0.0} == {v: x for v, x in want.items() if x != 0.0}, (kw, which)
    # every read, weighted by ReadCount: what the kernels accumulate (helpers.summary_from_records states it)
    rc_norm = rec["read_count"].astype(np.float64)
    summ = summary_from_records(rec, om.Len() - 1, 0, om.EditStop)
    B = om.Len() + 2
    for which in (0, 1, 2):
        # reads without and with alt editing together = every read; values -2 .. B - 3 <-> the summary's B bins
        tot = T.hist_norm(rec, rc_norm, om.EditStop, which, -2, B - 3, all=True, has_alt=False) + \
            T.hist_norm(rec, rc_norm, om.EditStop, which, -2, B - 3, all=True, has_alt=True)
        dev = summ[16 + which * B:16 + (which + 1) * B].astype(np.float64)
        assert tot.sum() > 0
        if which == 1:  # bin = JuncLen
            assert np.array_equal(tot[2:], dev[:B - 2]) and dev[B - 2:].sum() == 0
        else:           # bin = EditStop / JuncEnd - EditOffset + 2
            assert np.array_equal(tot, dev)


def test_mutant_report_vs_oracle(examples):
    """`treat mutant` (cmd/treat/mutant.go:55-110) from records + ops; ops here come from the oracle."""
    tm = T.NewTemplateFromFasta(examples["templates.fa"], T.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    pt = P.template_from_fasta(examples["templates.fa"], P.FORWARD, "T")
    recs = P.read_fasta(examples["clones.fa"])
    recs = recs + [(i + "x", s) for i, s in recs[7:13]] + recs[9:11]  # repeat some indel reads: counts > 1
    fb = T.FastaBatch.from_records(recs)
    orec, oops, _ = O.align_batch(om, fb.seqs, fb.offsets, fb.read_counts)
    rec = _oracle_record_as_product(orec)
    want_tm, want_fm = P.mutant(pt, [P.Fragment(i, s, P.FORWARD, "T") for i, s in recs])
    assert want_tm and want_fm
    for n in (0, 1, 10):
        assert T.mutant_report(tm, rec, oops, fb.seqs, fb.offsets, n) == P.mutant_report(want_tm, want_fm, n)


def _random_fasta(rng, n):
    parts = []
    if rng.random() < 0.5:
        parts.append(b"junk before\nthe first>mid_3 line record\n")
    for i in range(n):
        L = rng.choice([0, 1, 30, 61, 200, 330])
        seq = "".join(rng.choice("ACGT>") if rng.random() < 0.02 else rng.choice("ACGT") for _ in range(L)).encode()
        w = rng.choice([10000, 60, 7])
        body = b"\n".join(seq[j:j + w] for j in range(0, len(seq), w))
        eol = b"\r\n" if rng.random() < 0.2 else b"\n"
        parts.append(b">r%d_%d some text" % (i, rng.randrange(1, 99)) + eol + body + (eol if body or rng.random() < 0.5 else b""))
        if rng.random() < 0.1:
            parts.append(b"\n")
    return b"".join(parts)


def test_fasta_parse_bytes_and_split(tmp_path):
    """treat_b200_fasta_parse == treat_b200_fasta_read, and the slices of treat_b200_fasta_split start at record starts:
    parsing them one after the other gives the records of the whole file (how a file is shared between ranks and how
    treat_b200_fasta_stream cuts its pieces)."""
    import random
    rng = random.Random(99)
    for trial in range(12):
        text = _random_fasta(rng, rng.choice([0, 1, 5, 200]))
        path = tmp_path / ("f%d.fa" % trial)
        path.write_bytes(text)
        whole = T.read_fasta(str(path))
        mem = T.parse_fasta(text)
        assert mem.ids == whole.ids and np.array_equal(mem.offsets, whole.offsets)
        assert np.array_equal(mem.seqs, whole.seqs) and np.array_equal(mem.read_counts, whole.read_counts)
        ref = P.read_fasta(str(path))
        assert [r[0] for r in ref] == whole.ids and [r[1].encode() for r in ref] == [whole.seq(i) for i in range(len(whole))]
        for parts in (1, 2, 3, 8, 50):
            off = T.fasta_split(text, parts)
            assert off[0] == 0 and off[-1] == len(text) and (np.diff(off.astype(np.int64)) >= 0).all()
            ids, seqs = [], []
            for k in range(parts):
                lo, hi = int(off[k]), int(off[k + 1])
                if lo and lo < len(text):
                    assert text[lo:lo + 1] == b">" and text[lo - 1:lo] == b"\n"
                piece = T.parse_fasta(text[lo:hi])
                ids += piece.ids
                seqs += [piece.seq(i) for i in range(len(piece))]
            assert ids == whole.ids and seqs == [whole.seq(i) for i in range(len(whole))], (trial, parts)


def _go_stats_block(gene, tmpl, rows, unique, norm):
    """cmd/treat/stats.go:62-137 (ShowStats, one gene) restated with Python's printf verbs; rows = [(sample, Stats)] with
    Stats = (Total, Std, NonStd, Indels, Snps, SingleMismatch, DoubleMismatch) (stats.go:33-41, 140-187)."""
    def pct(x, y):
        return 0.0 if y == 0 else x / y * 100.0
    g = [sum(r[1][k] for r in rows) for k in range(7)]
    out = ["=" * 80, gene, "=" * 80, "%20s%11d" % ("Total Alignments:", g[0])]
    if not norm:
        out += ["%20s%11d" % ("Standard:", g[1]), "%20s%11d" % ("Non-Standard:", g[2]), "%20s%11d" % ("1-Mismatch:", g[5]),
                "%20s%11d" % ("2-Mismatch:", g[6]), "%20s%11d" % (">3-Mismatch:", g[4]), "%20s%11d" % ("Indels:", g[3])]
    out += ["%20s%11d" % ("Template Edit Stop:", tmpl.EditStop), "%20s%11s" % ("Edit Base:", tmpl.EditBase),
            "%20s%11d" % ("Alt Templates:", len(tmpl.AltRegion)), "-" * 80]
    out.append("%-15s%9s%9s%5s%9s%5s%9s%5s%9s%5s" % ("Sample", "Total", "Std", "%", "Non-Std", "%", "1MM", "%", "2MM", "%")
               if not norm else "%-15s%9s" % ("Sample", "Total"))
    out.append("-" * 80)
    for name, c in rows:
        if len(name) > 12:
            name = name[:12] + ".."
        out.append("%-15s%9d%9d%5.1f%9d%5.1f%9d%5.1f%9d%5.1f" % (name, c[0], c[1], pct(c[1], c[0]), c[2], pct(c[2], c[0]), c[5],
                                                                   pct(c[5], c[0]), c[6], pct(c[6], c[0]))
                   if not norm else "%-15s%9d" % (name, c[0]))
    return "\n".join(out) + "\n\n"


def test_stats_text(examples):
    """`treat stats` block (cmd/treat/stats.go:62-187) from summary counters: per-sample Stats from the oracle's records."""
    from helpers import summary_from_records
    tm = T.NewTemplateFromFasta(examples["templates.fa"], T.FORWARD, "T")
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    samples = []
    for name, f in (("clones", "clones.fa"), ("a-very-long-sample-name", "sample-1.fa"), ("empty", None)):
        if f is None:
            summ = np.zeros(16 + 3 * (tm.Len() + 2), dtype=np.uint64)
        else:
            reads = P.read_fasta(examples[f])
            seqs = np.frombuffer("".join(s for _, s in reads).encode(), dtype=np.uint8)
            off = np.cumsum([0] + [len(s) for _, s in reads]).astype(np.uint64)
            rc = np.array([O.parseMergeCount(i) for i, _ in reads], dtype=np.uint32)
            rec, _, _ = O.align_batch(om, seqs, off, rc)
            summ = summary_from_records(rec, len(tm.Bases), 0, om.EditStop)
        samples.append((name, summ))
    for unique in (False, True):
        for norm in (False, True):
            rows = [(n, [int(x) for x in s[7 if unique else 0:][:7]]) for n, s in samples]
            assert T.stats_text(tm, "RPS12", samples, unique, norm) == _go_stats_block("RPS12", tm, rows, unique, norm)
    assert T.stats_text(tm, "none", [], False, False) == _go_stats_block("none", tm, [], False, False)


def test_unmarshal_binary_round_trip(examples):
    """Alignment.UnmarshalBinary (align.go:295-314) inverts MarshalBinary: blobs made by the oracle for clones.fa."""
    om = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    for rid, seq in P.read_fasta(examples["clones.fa"]):
        a = O.NewAlignment(O.NewFragment(rid, seq), om)
        got = T.UnmarshalBinary(a.MarshalBinary())
        assert got == dict(EditStop=a.EditStop, JuncStart=a.JuncStart, JuncEnd=a.JuncEnd, JuncLen=a.JuncLen, ReadCount=a.ReadCount,
                           HasMutation=a.HasMutation, AltEditing=a.AltEditing, Mismatches=a.Mismatches, Indel=a.Indel, Norm=0.0,
                           JuncSeq=a.JuncSeq), rid
    # negative fields keep their sign (readInt64 is a 32-bit big-endian read, align.go:281-287); Norm is an IEEE double
    import struct
    blob = struct.pack(">iiiiIBBBBdI", -1, 0, -7, 3, 4000000000, 1, 2, 255, 1, 0.125, 3) + b"TTA"
    got = T.UnmarshalBinary(blob)
    assert (got["EditStop"], got["JuncEnd"], got["ReadCount"], got["Mismatches"], got["Norm"], got["JuncSeq"]) == (-1, -7, 4000000000, 255, 0.125, "TTA")
    with pytest.raises(T.TreatError):
        T.UnmarshalBinary(blob[:37])  # shorter than its header says: the reference panics
