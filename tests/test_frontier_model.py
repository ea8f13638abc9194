"""The claim behind k_frontier (treat_b200/csrc/frontier_kernel.cu), checked on the CPU against the oracle.

For a read whose stripped bases equal the template's, NewAlignment (align.go:95-279) never needs the edit-site counts
one by one: with R the upper-cased raw read, F / P the fully edited / pre-edited template re-inflated to raw strings,
  * the highest site whose count differs from FE is  m - (bases inside the longest common suffix of R and F),
  * the lowest site whose count differs from PE is   (bases inside the longest common prefix of R and P),
  * T[0] is all ones iff R == F, T[1] all ones iff R == P,
  * the bases are the template's iff the characters between that prefix and that suffix, edit bases dropped, are the
    template's bases between them (and, when prefix and suffix overlap, iff the base counts add up to m),
  * JuncSeq is the window of R between the end of base (prefix bases - 1) and the end of the first base of the suffix.
The model below takes exactly the decisions the kernel takes (same tables, same "leave it to the general path"
conditions) and must reproduce the oracle's record for every read it accepts.

Next step, stated and checked here but NOT taken by the kernel (`one_sub=True`): a read of the same form in which ONE base
was replaced by another base of the alphabet (never the edit base) aligns along the diagonal as well (n = m and one
mismatch: unique optimum), and its edit-site counts are those of the unsubstituted read.  The compares find
it as a difference in which neither side holds the edit base -- a frontier between two editing states always has the
edit base on one side --, step over it once and go on; the junction check tolerates it once.  Mismatches = 1, score
m - 2, and with excludeSnps the read is a mutant (align.go:240-241).  Measured in the kernel (round 2): 94 % instead of
88 % of an RPS12 batch and 44 % instead of 16 % of the noisy set are finished by k_frontier, k_pack_finish2 drops from
101 to 78 us per 10^6 reads, but k_frontier itself grows from 168 to 204 us (30 more warp instructions per read): a net
loss on RPS12, so the kernel leaves these reads to k_pack_finish2 for now.
"""
import numpy as np
import pytest

from oracle import oracle as O
from treat_b200 import synth


class FrontierModel:
    def __init__(self, t: synth.TemplateArrays):
        eb = t.edit_base
        raws = t.raw_sequences()
        self.F, self.P = raws[0], raws[1]
        self.tm = t.m
        self.len = t.m + 1
        self.K = t.K
        self.eb = eb
        self.bases = t.bases.decode()
        self.alt_start = [int(x) for x in t.alt_start]
        P, F = self.P, self.F
        self.pcum = [0] * (len(P) + 1)
        for i, ch in enumerate(P):
            self.pcum[i + 1] = self.pcum[i] + (ch != eb)
        self.fsuf = [0] * (len(F) + 1)
        for i in range(len(F)):
            self.fsuf[i + 1] = self.fsuf[i] + (F[len(F) - 1 - i] != eb)
        # raw offset just behind base i-1 of P (0 for i = 0)
        self.ppos_after = [0] * (t.m + 1)
        k = 0
        for i, ch in enumerate(P):
            if ch != eb:
                k += 1
                self.ppos_after[k] = i + 1
        # characters of F behind base m - bs (bs >= 1)
        self.ftail_after = [0] * (t.m + 1)
        k = 0
        for i in range(len(F) - 1, -1, -1):
            if F[i] != eb:
                k += 1
                self.ftail_after[k] = len(F) - 1 - i
        fe = t.edit_site[0]
        pe = t.edit_site[1]
        self.mid_cap = 254 - int(fe.max()) - int(pe.max())
        self.alphabet = set(ch for ch in "ACGT" if ch != eb)  # the packed alphabet: the first three of ACGT that are not the edit base

    def record(self, raw: str, offset: int = 0, exclude_snps: bool = False, one_sub: bool = False):
        """(edit_stop, junc_start, junc_end, junc_len, js_off, js_len, mismatches) or None: left to the general path."""
        R = raw.upper()
        L, P, F, tm, ln, eb = len(R), self.P, self.F, self.tm, self.len, self.eb
        lp, lf = min(L, len(P)), min(L, len(F))

        def prefix(start):
            t = start
            while t < lp and R[t] == P[t]:
                t += 1
            return t

        def suffix(start):
            s = start
            while s < lf and R[L - 1 - s] == F[len(F) - 1 - s]:
                s += 1
            return s
        t, s = prefix(0), suffix(0)
        # a difference with no edit base on either side is a substituted base (of the alphabet): stepped over once
        sub_at = set()
        if one_sub and t < lp and R[t] != eb and P[t] != eb:
            if R[t] not in self.alphabet:
                return None
            sub_at.add(t)
            t = prefix(t + 1)
        if one_sub and s < lf and R[L - 1 - s] != eb and F[len(F) - 1 - s] != eb:
            if R[L - 1 - s] not in self.alphabet:
                return None
            sub_at.add(L - 1 - s)
            s = suffix(s + 1)
        if len(sub_at) > 1:
            return None
        mism = len(sub_at)
        bp, bs = self.pcum[t], self.fsuf[s]
        t0_all = s == L and L == len(F)
        pe_all = t == L and L == len(P)
        if t + s >= L:
            ident = bp + self.fsuf[L - t] == tm
        else:
            if L - s - t > self.mid_cap:
                return None
            rank, ident = bp, True
            for ch in R[t:L - s]:
                if ch != eb:
                    if rank >= tm:
                        ident = False
                        break
                    if self.bases[rank] != ch:
                        if not one_sub:
                            ident = False
                            break
                        if ch not in self.alphabet:
                            return None
                        mism += 1
                        if mism > 1:
                            ident = False
                            break
                    rank += 1
            ident = ident and rank + bs == tm
        if not ident or not (t0_all or bs >= 1):
            return None
        junc_start = ln - 1 if t0_all else bs
        if self.K > 2 and junc_start in self.alt_start:
            return None
        has_mutation = exclude_snps and mism > 0
        junc_end = -1 if pe_all else tm - bp
        edit_stop = junc_start - 1
        junc_len, js_off, js_len = 0, 0, 0
        if junc_end > edit_stop:
            junc_len = junc_end - edit_stop
            if not t0_all and not has_mutation:
                o0 = self.ppos_after[bp]
                o1 = L - self.ftail_after[bs]
                js_len = o1 - o0
                js_off = o0 if js_len else 0
        elif junc_end < edit_stop:
            junc_end = edit_stop
        if t0_all:
            junc_end = edit_stop = junc_start
            junc_len = 0
        return edit_stop + offset, junc_start + offset, junc_end + offset, junc_len, js_off, js_len, mism


def _oracle_template(t, tmp_path):
    path = t.write_fasta(str(tmp_path / "tmpl.fa"))
    return O.NewTemplateFromFasta(path, O.FORWARD, t.edit_base)


def _check(t, seqs, off, rc, tmp_path, min_accept, min_subs=0, one_sub=False):
    om = _oracle_template(t, tmp_path)
    model = FrontierModel(t)
    raw = seqs.tobytes().decode("latin-1")
    for exclude_snps in (False, True):
        orec, _, _ = O.align_batch(om, seqs, off, rc, exclude_snps=exclude_snps, nthreads=4, want_ops=False)
        accepted = subs = 0
        for r in range(len(off) - 1):
            got = model.record(raw[int(off[r]):int(off[r + 1])], exclude_snps=exclude_snps, one_sub=one_sub)
            if got is None:
                continue
            accepted += 1
            o = orec[r]
            want = (int(o["edit_stop"]), int(o["junc_start"]), int(o["junc_end"]), int(o["junc_len"]), int(o["junc_seq_off"]),
                    int(o["junc_seq_len"]), int(o["mismatches"]))
            assert got == want, (r, got, want)
            mism = got[6]
            subs += mism
            assert mism in ((0, 1) if one_sub else (0,)) and o["has_mutation"] == (1 if exclude_snps and mism else 0)
            assert o["indel"] == 0 and o["alt_editing"] == 0
            assert o["score"] == t.m - 2 * mism and o["n_bases"] == t.m and o["aln_len"] == t.m and o["status"] == 0
        assert accepted >= min_accept * (len(off) - 1), accepted
        assert subs >= min_subs, subs
    return accepted


def test_frontier_model_rps12(tmp_path):
    t = synth.rps12_template()
    seqs, off, rc = synth.reads(t, 6000, seed=synth.SEED_BASE + 3)
    _check(t, seqs, off, rc, tmp_path, 0.80)
    _check(t, seqs, off, rc, tmp_path, 0.85, min_subs=150, one_sub=True)


def test_frontier_model_truncated_and_noisy(tmp_path):
    t = synth.rps12_template()
    seqs, off, rc = synth.reads(t, 3000, seed=synth.SEED_BASE + 11, p_trunc=0.3, p_sub_base=0.002, p_indel_read=0.02)
    _check(t, seqs, off, rc, tmp_path, 0.3)
    _check(t, seqs, off, rc, tmp_path, 0.3, one_sub=True)


def test_frontier_model_long_template(tmp_path):
    t = synth.long_template()
    seqs, off, rc = synth.reads(t, 1500, seed=synth.SEED_BASE + 7, p_trunc=0.2)
    _check(t, seqs, off, rc, tmp_path, 0.5)


def test_frontier_model_edges(tmp_path):
    """FE, PE, edit bases added / removed at both ends, lower case, foreign characters, the empty read."""
    t = synth.rps12_template()
    F, P = t.raw_sequences()[:2]
    cases = [F, P, F.lower(), "T" + F, "TT" + P, F + "T", P + "TTT", F[1:], P[:-1], "", "T", "TTTT", P[:50] + F[len(F) - 200:],
             P[:120] + "TTTTT" + F[len(F) - 60:], F.replace("A", "N", 1), P[:30] + "\x00" + P[31:], P[:40] + " " + F[len(F) - 100:]]
    # one substituted base: in the prefix, in the suffix, in the junction, at the very ends, lower case, twice, by the edit base
    swap = {"A": "C", "C": "G", "G": "A"}

    def sub(sq, k):  # the k-th base of sq replaced
        idx = [i for i, ch in enumerate(sq) if ch != "T"][k]
        return sq[:idx] + swap[sq[idx]] + sq[idx + 1:]
    mid = P[:120] + "TT" + F[len(F) - 150:]
    cases += [sub(F, 0), sub(F, 5), sub(F, -1), sub(P, 0), sub(P, -1), sub(P, 40), sub(mid, 3), sub(mid, -3), sub(mid, 60), sub(mid, 61),
              sub(sub(F, 5), 50), sub(F, 7).lower(), F[:20] + "T" + F[21:] if F[20] != "T" else F, sub(F, 30).replace("A", "N", 1)]
    # every site edited a little further than FE / not at all
    es = t.edit_site
    for k in range(2, t.K):
        cases.append(t.raw_sequences()[k])
    raw = "".join(cases).encode("latin-1")
    off = np.zeros(len(cases) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(c) for c in cases])
    seqs = np.frombuffer(raw, dtype=np.uint8)
    _check(t, seqs, off, None, tmp_path, 0.1)
    _check(t, seqs, off, None, tmp_path, 0.3, min_subs=6, one_sub=True)
    assert es.shape[0] == t.K
