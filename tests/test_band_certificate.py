"""Exactness argument for a banded fill ("next"): if the score S_band of a Needleman-Wunsch fill restricted
to the diagonals |i - j| <= w beats the best score any path that leaves the band can reach, then (1) S_band is the global
score and (2) nwalgo's traceback (Up, then Left, then Diagonal) walks the same cells in the banded matrix as in the full one,
so the gapped strings are identical.  This test checks both claims numerically against the full matrix and against the
oracle, and that the certificate holds for the reads TREAT actually sees (few indels near the diagonal).

Bound: a path that leaves the band touches diagonal +(w+1) or -(w+1); starting on diagonal 0 and ending on diagonal m - n it
then holds at least g = (w+1) + |+-(w+1) - (m-n)| gap steps, and a path with g gap steps scores at most
(m + n - g) / 2 - g (every non-gap step is at best a match)."""
import random

import numpy as np

from oracle import oracle as O

NEG = -10**9


def fill(a: str, b: str, w=None):
    """F[i][j] for template a (rows consumed by Up) and read b, gap -1, match +1, mismatch -1; cells outside the band are NEG."""
    m, n = len(a), len(b)
    F = np.full((m + 1, n + 1), NEG, dtype=np.int64)
    for i in range(m + 1):
        for j in range(n + 1):
            if w is not None and abs(i - j) > w:
                continue
            if i == 0 and j == 0:
                F[i][j] = 0
                continue
            best = NEG
            if i > 0 and j > 0 and F[i - 1][j - 1] > NEG:
                best = max(best, F[i - 1][j - 1] + (1 if a[i - 1] == b[j - 1] else -1))
            if i > 0 and F[i - 1][j] > NEG:
                best = max(best, F[i - 1][j] - 1)
            if j > 0 and F[i][j - 1] > NEG:
                best = max(best, F[i][j - 1] - 1)
            F[i][j] = best
    return F


def traceback(F, a, b):
    """nwalgo's pointer order: Up if F == up + gap, else Left if F == left + gap, else Diagonal; first row / column forced."""
    i, j, ops = len(a), len(b), []
    while i > 0 or j > 0:
        if j == 0 or (i > 0 and F[i - 1][j] > NEG and F[i][j] == F[i - 1][j] - 1):
            ops.append(1)  # template base vs '-'
            i -= 1
        elif i == 0 or (F[i][j - 1] > NEG and F[i][j] == F[i][j - 1] - 1):
            ops.append(2)  # '-' vs read base
            j -= 1
        else:
            ops.append(0)
            i -= 1
            j -= 1
    return ops[::-1]


def outside_bound(m, n, w):
    g = min((w + 1) + abs((w + 1) - (m - n)), (w + 1) + abs(-(w + 1) - (m - n)))
    return (m + n - g) / 2 - g


def ops_of(aln1, aln2):
    return [2 if x == "-" else 1 if y == "-" else 0 for x, y in zip(aln1, aln2)]


def mutate(rng, s, n_edits):
    s = list(s)
    for _ in range(n_edits):
        k, pos = rng.randrange(3), rng.randrange(len(s) + 1)
        if k == 0 and s:
            del s[min(pos, len(s) - 1)]
        elif k == 1:
            s.insert(pos, rng.choice("ACG"))
        elif s:
            s[min(pos, len(s) - 1)] = rng.choice("ACG")
    return "".join(s)


def test_band_certificate_implies_identical_traceback():
    rng = random.Random(20260)
    O.set_tie_order("ULD")
    certified = total = 0
    for trial in range(140):
        m = rng.randrange(20, 70)
        a = "".join(rng.choice("ACG") for _ in range(m))
        b = mutate(rng, a, rng.randrange(0, 7))
        if trial % 10 == 0:
            b = b[rng.randrange(0, 12):]  # a truncated read: the certificate must refuse what it cannot prove
        n = len(b)
        Ff = fill(a, b)
        full_ops = traceback(Ff, a, b)
        aln1, aln2, score = O.nw_align(a, b)
        assert int(Ff[m][n]) == score and full_ops == ops_of(aln1, aln2)  # the restatement above is the oracle's nwalgo
        for w in (3, 6, 10):
            if abs(m - n) > w:
                continue  # (m, n) itself lies outside the band
            Fb = fill(a, b, w)
            total += 1
            assert Fb[m][n] <= Ff[m][n]
            if Fb[m][n] > outside_bound(m, n, w):  # the certificate
                certified += 1
                assert Fb[m][n] == Ff[m][n], (a, b, w)
                assert traceback(Fb, a, b) == full_ops, (a, b, w)
    assert certified > total // 2  # the certificate is not vacuous on reads a few edits away from the template


def test_certificate_covers_rps12_like_reads():
    """On RPS12-sized pairs (m = 163) a band of +-16 diagonals is certified for reads with a handful of edits: 33 x 163 cells
    instead of 163 x 163."""
    rng = random.Random(7)
    m, w = 163, 16
    a = "".join(rng.choice("ACG") for _ in range(m))
    ok = 0
    for _ in range(12):
        b = mutate(rng, a, rng.randrange(2, 6))
        Fb = fill(a, b, w)
        if abs(m - len(b)) <= w and Fb[m][len(b)] > outside_bound(m, len(b), w):
            ok += 1
            aln1, aln2, score = O.nw_align(a, b)
            assert int(Fb[m][len(b)]) == score and traceback(Fb, a, b) == ops_of(aln1, aln2)
    assert ok == 12


def test_pure_deletion_reads_align_to_their_leftmost_embedding():
    """The shortcut of band_kernel.cu: when the read's bases are a subsequence of the template's, nwalgo's alignment is the
    LEFTMOST embedding (every read base on the earliest template base it can take), score 2n - m, for each of the three tie
    orders -- F[i][k] <= 2k - i with equality iff read[0:k] embeds in template[0:i], so Up is optimal exactly while the
    prefix still embeds and Left never is.  Checked against the oracle on random and repetitive templates."""
    rng = random.Random(5)

    def leftmost(a, b):
        pos, out = 0, []
        for ch in b:
            q = a.find(ch, pos)
            if q < 0:
                return None
            out.append(q)
            pos = q + 1
        return out

    checked = 0
    try:
        for order in ("ULD", "UDL", "LUD"):
            O.set_tie_order(order)
            for trial in range(300):
                m = rng.randrange(5, 80)
                a = "".join(rng.choice("ACG") for _ in range(m)) if trial % 3 else ("ACG" * 30 + "AACCGG" * 5)[:m]
                b = "".join(c for c in a if rng.random() > 0.2)
                if trial % 2:
                    b = b[rng.randrange(0, len(b) // 2 + 1):len(b) - rng.randrange(0, len(b) // 2 + 1)]
                if not b:
                    continue
                where = set(leftmost(a, b))
                aln1, aln2, score = O.nw_align(a, b)
                assert (aln1, aln2, score) == (a, "".join(a[i] if i in where else "-" for i in range(m)), 2 * len(b) - m), (order, a, b)
                checked += 1
    finally:
        O.set_tie_order("ULD")
    assert checked > 800
