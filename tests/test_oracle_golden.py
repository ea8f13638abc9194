"""Pin the CPU oracle (oracle/treat_oracle.c) and the independent pure-Python restatement
(oracle/pyref.py) against every golden the reference owns for this path."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P


@pytest.fixture(autouse=True)
def _uld():
    O.set_tie_order("ULD")
    yield
    O.set_tie_order("ULD")


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


def test_align_fragments_c(fx, examples):
    """TestAlignFragments, align_test.go:68-137."""
    ts = fx["test_sample"]
    tmpl = O.NewTemplateFromFasta(examples[ts["templates"]], O.FORWARD, ts["base"])
    assert tmpl.Size() == ts["size"]
    assert len(ts["reads"]) == 15
    for r in ts["reads"]:
        frag = O.NewFragment(r["id"], r["seq"], O.FORWARD, ts["base"])
        _check_attrs(O.NewAlignment(frag, tmpl, False), r["attr"], r["id"])


def test_align_fragments_pyref(fx, examples):
    ts = fx["test_sample"]
    tmpl = P.template_from_fasta(examples[ts["templates"]], P.FORWARD, ts["base"])
    for r in ts["reads"]:
        frag = P.Fragment(r["id"], r["seq"], P.FORWARD, ts["base"])
        _check_attrs(P.NewAlignment(frag, tmpl, False), r["attr"], r["id"])


def test_tie_order_pins(fx, examples):
    """ULD/UDL/LUD satisfy the reference's goldens, LDU/DUL/DLU fail read 9-4."""
    ts = fx["test_sample"]
    tmpl = O.NewTemplateFromFasta(examples[ts["templates"]], O.FORWARD, ts["base"])
    verdict = {}
    for order in O.TIE_ORDERS:
        O.set_tie_order(order)
        bad = []
        for r in ts["reads"]:
            a = O.NewAlignment(O.NewFragment(r["id"], r["seq"], O.FORWARD, ts["base"]), tmpl, False)
            try:
                _check_attrs(a, r["attr"], r["id"])
            except AssertionError:
                bad.append(r["id"].split()[0])
        verdict[order] = bad
    assert verdict["ULD"] == verdict["UDL"] == verdict["LUD"] == []
    assert verdict["LDU"] == verdict["DUL"] == verdict["DLU"] == ["9-4"]


def test_readme_block(fx, examples):
    rb = fx["readme_block"]
    out = O.cli_align(template=examples[rb["templates"]], fragment=examples[rb["reads"]], base=rb["base"])
    assert out == rb["stdout"]


def test_readme_pairwise(fx):
    rp = fx["readme_pairwise"]
    assert O.cli_align(s1=rp["s1"], s2=rp["s2"], base=rp["base"]) == rp["stdout"]
    f1, f2 = P.Fragment("1-1", rp["s1"]), P.Fragment("2-1", rp["s2"])
    assert P.PrintAlignment(*P.SimpleAlign(f1, f2)) == rp["stdout"]


def test_readme_csv_rows(fx, examples):
    rc = fx["readme_csv"]
    tmpl = O.NewTemplateFromFasta(examples[rc["templates"]], O.FORWARD, "T")
    tmpl.SetOffset(rc["offset"])
    reads = dict(P.read_fasta(examples[rc["reads"]]))
    for row in rc["rows"]:
        a = O.NewAlignment(O.NewFragment(row["id"], reads[row["id"]]), tmpl, False)
        for k in ("read_count", "alt_editing", "has_mutation", "edit_stop", "junc_end", "junc_len", "junc_seq"):
            assert a.fields()[k] == row[k], (row["id"], k)


def test_fragment(fx):
    """TestFragment, fragment_test.go:25-50."""
    for s in fx["fragment_test"]["seqs"]:
        for mod in (O, ):
            f = mod.NewFragment("id", s, O.FORWARD, "t")
            assert f.String() == s.upper() and f.ReadCount == 1
            f = mod.NewFragment("1-143", s, O.REVERSE, "t")
            assert f.String() == s[::-1].upper() and f.ReadCount == 143
        assert P.Fragment("id", s, P.FORWARD, "t").String() == s.upper()
        assert P.Fragment("1-143", s, P.REVERSE, "t").String() == s[::-1].upper()


def test_parse_read_counts(fx):
    """TestParseReadCounts, fragment_test.go:52-76."""
    for rid, cnt in fx["parse_read_counts"]["table"].items():
        assert O.parseMergeCount(rid) == cnt, rid
        assert P.parseMergeCount(rid) == cnt, rid
        assert O.NewFragment(rid, "CTGCTG", O.FORWARD, "t").ReadCount == cnt
    # Atoi range error and uint32 truncation (fragment.go:80-85)
    assert O.parseMergeCount("a-99999999999999999999") == 1 == P.parseMergeCount("a-99999999999999999999")
    assert O.parseMergeCount("a-4294967297") == 1 == P.parseMergeCount("a-4294967297")


def test_template(fx, examples):
    """TestTemplate, template_test.go:26-70."""
    tt = fx["template_test"]
    assert O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "t").Size() == tt["rps12_size"]
    with pytest.raises(ValueError):
        O.NewTemplate(O.NewFragment("full", tt["bad"]["full"], O.FORWARD, "t"),
                      O.NewFragment("pre", tt["bad"]["pre"], O.FORWARD, "t"))
    full = O.NewFragment("full", tt["good"]["full"], O.FORWARD, "t")
    pre = O.NewFragment("pre", tt["good"]["pre"], O.FORWARD, "t")
    tmpl = O.NewTemplate(full, pre)
    rebuilt = "".join("T" * b + (tmpl.Bases[i] if i < len(tmpl.Bases) else "") for i, b in enumerate(tmpl.EditSite[1]))
    assert rebuilt == pre.String()
    with pytest.raises(ValueError):
        O.NewTemplate(full, pre, None, [None, None])
    rps = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    assert (rps.Len(), rps.EditStop, rps.AltRegion) == (164, 6, [(27, 34), (112, 119), (116, 119)])


def test_align_smoke(fx):
    """TestAlign, align_test.go:32-52 (asserts nothing in the reference; C and pyref must agree)."""
    s = fx["align_smoke"]
    a, b, c = (O.NewFragment(n, s[k], O.FORWARD, s["base"]) for n, k in (("a", "fe"), ("b", "pe"), ("c", "read")))
    tmpl = O.NewTemplate(a, b)
    pa, pb, pc = (P.Fragment(n, s[k], P.FORWARD, s["base"]) for n, k in (("a", "fe"), ("b", "pe"), ("c", "read")))
    ptmpl = P.Template(pa, pb)
    assert O.NewAlignment(c, tmpl).WriteTo(c, tmpl, 80) == P.NewAlignment(pc, ptmpl).WriteTo(pc, ptmpl, 80)


def test_c_oracle_matches_derived_fixtures(fx, examples):
    """The frozen pyref outputs (tests/golden) must be reproduced by the C oracle byte for byte."""
    for s in fx["derived"]["sets"]:
        tmpl = O.NewTemplateFromFasta(examples[s["templates"]], O.FORWARD, s["base"])
        tmpl.SetOffset(s["offset"])
        reads = P.read_fasta(examples[s["reads"]])
        assert len(reads) == len(s["results"])
        for (rid, seq), want in zip(reads, s["results"]):
            f = O.NewFragment(rid, seq, O.FORWARD, s["base"])
            a = O.NewAlignment(f, tmpl, s["exclude_snps"])
            assert a.fields() == want["fields"], rid
            assert a.WriteTo(f, tmpl, 80) == want["text"], rid
            assert a.MarshalBinary().hex() == want["marshal_hex"], rid
            assert O.nw_align(tmpl.Bases, f.Bases)[:2] == (want["aln1"], want["aln2"])


def _rand_read(rng, tmpl_str, base="T"):
    s = list(tmpl_str.upper())
    for _ in range(rng.randint(0, 4)):
        k = rng.randint(0, 4)
        pos = rng.randrange(len(s) + 1)
        if k == 0 and s:
            del s[min(pos, len(s) - 1)]
        elif k == 1:
            s.insert(pos, rng.choice("ACGTTTN"))
        elif k == 2 and s:
            s[min(pos, len(s) - 1)] = rng.choice("ACGTTN")
        elif k == 3:
            s[pos:pos] = "T" * rng.randint(1, 6)
        else:
            lo = rng.randrange(len(s) + 1)
            del s[lo:lo + rng.randint(0, 8)]
    return "".join(s)


def test_c_vs_pyref_fuzz(fx, examples):
    """Differential fuzz of the two restatements on mutated template strings (all tie orders)."""
    rng = random.Random(1234)
    for tf, base in (("test-templates.fa", "t"), ("simple-templates.fa", "T")):
        ct = O.NewTemplateFromFasta(examples[tf], O.FORWARD, base)
        pt = P.template_from_fasta(examples[tf], P.FORWARD, base)
        srcs = [s for _, s in P.read_fasta(examples[tf])]
        for it in range(150):
            seq = _rand_read(rng, rng.choice(srcs))
            tie = O.TIE_ORDERS[it % 6] if it % 5 == 0 else "ULD"
            O.set_tie_order(tie)
            name = "r-%d" % (it + 1)
            cf, pf = O.NewFragment(name, seq, O.FORWARD, base), P.Fragment(name, seq, P.FORWARD, base)
            for ex in (False, True):
                ca, pa = O.NewAlignment(cf, ct, ex), P.NewAlignment(pf, pt, ex, tie)
                assert ca.fields() == pa.fields(), (seq, tie)
                assert ca.WriteTo(cf, ct, 80) == pa.WriteTo(pf, pt, 80, tie), (seq, tie)
                assert ca.MarshalBinary() == pa.MarshalBinary()
            assert O.SimpleAlign(cf, O.NewFragment("x-1", srcs[0], O.FORWARD, base)) == \
                P.SimpleAlign(pf, P.Fragment("x-1", srcs[0], P.FORWARD, base), tie)


def test_edge_cases(fx, examples):
    """SURVEY 8c edge behaviours: empty / all-edit-base reads, uint8 wrap, reference panic case."""
    tmpl = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    for seq in ("", "TTTTTT"):
        a = O.NewAlignment(O.NewFragment("e-1", seq), tmpl)
        assert (a.HasMutation, a.Indel, a.Mismatches) == (1, 1, 0)
    fe = O.NewFragment("fe", "AC" * 300)
    t2 = O.NewTemplate(fe, O.NewFragment("pe", "AC" * 300))
    a = O.NewAlignment(O.NewFragment("x-1", "CA" * 300), t2)
    assert a.Indel == 0 or a.HasMutation == 1
    mm = O.NewAlignment(O.NewFragment("x-1", "G" * 300), O.NewTemplate(O.NewFragment("fe", "A" * 300), O.NewFragment("pe", "A" * 300)))
    assert (mm.Mismatches, mm.HasMutation, mm.Indel) == (300 % 256, 1, 0)
    t3 = O.NewTemplate(O.NewFragment("fe", "AT"), O.NewFragment("pe", "A"))
    p = O.NewAlignment(O.NewFragment("r-1", "TA"), t3)
    assert p.WouldPanic and p.JuncSeq == ""
    pp = P.NewAlignment(P.Fragment("r-1", "TA"), P.Template(P.Fragment("fe", "AT"), P.Fragment("pe", "A")))
    assert pp.WouldPanic and pp.fields() == p.fields()


def test_batch_driver_matches_scalar(fx, examples):
    tmpl = O.NewTemplateFromFasta(examples["templates.fa"], O.FORWARD, "T")
    reads = P.read_fasta(examples["clones.fa"])
    seqs = np.frombuffer("".join(s for _, s in reads).encode(), dtype=np.uint8)
    off = np.cumsum([0] + [len(s) for _, s in reads]).astype(np.uint64)
    rc = np.array([O.parseMergeCount(i) for i, _ in reads], dtype=np.uint32)
    for nthreads in (1, 3):
        rec, ops, _ = O.align_batch(tmpl, seqs, off, rc, nthreads=nthreads)
        for k, (rid, seq) in enumerate(reads):
            f = O.NewFragment(rid, seq)
            a = O.NewAlignment(f, tmpl)
            r = rec[k]
            assert (r["edit_stop"], r["junc_start"], r["junc_end"], r["junc_len"]) == (a.EditStop, a.JuncStart, a.JuncEnd, a.JuncLen)
            assert (r["has_mutation"], r["alt_editing"], r["mismatches"], r["indel"]) == (a.HasMutation, a.AltEditing, a.Mismatches, a.Indel)
            assert r["read_count"] == a.ReadCount and r["score"] == a.Score
            js = seq.upper()[r["junc_seq_off"]:r["junc_seq_off"] + r["junc_seq_len"]]
            assert js == a.JuncSeq
            aln1, aln2, _ = O.nw_align(tmpl.Bases, f.Bases)
            got = [(ops[k, c // 4] >> (2 * (c % 4))) & 3 for c in range(r["aln_len"])]
            want = [2 if x == "-" else 1 if y == "-" else 0 for x, y in zip(aln1, aln2)]
            assert got == want


def test_text_batch_matches_cli(fx, examples, tmp_path):
    """orc_text_batch (the checker of the device-side WriteTo at scale) == orc_cli_align on the same reads written as a
    FASTA file with the names it generates, which in turn is pinned to README.rst:70-82 by test_readme_block."""
    from treat_b200 import synth
    t = synth.rps12_template()
    tp = t.write_fasta(str(tmp_path / "t.fa"))
    om = O.NewTemplateFromFasta(tp, O.FORWARD, "T")
    seqs, off, rc = synth.reads(t, 120, seed=9, p_trunc=0.2)
    text, toff = O.text_batch(om, seqs, off, rc, prefix="r", nthreads=3)
    fp = str(tmp_path / "r.fa")
    with open(fp, "wb") as fh:
        for i in range(120):
            fh.write(b">r%d-%d\n" % (i, rc[i]) + seqs[int(off[i]):int(off[i + 1])].tobytes() + b"\n")
    want = O.cli_align(template=tp, fragment=fp)
    assert text.tobytes().decode("latin-1") == want
    assert toff[0] == 0 and toff[-1] == len(text) and np.all(np.diff(toff.astype(np.int64)) > 0)
    one = O.NewFragment("r7-%d" % rc[7], seqs[int(off[7]):int(off[8])].tobytes().decode())
    assert text[int(toff[7]):int(toff[8])].tobytes().decode() == O.NewAlignment(one, om).WriteTo(one, om, 80)
