"""Pure-Python line-by-line restatement of the reference path (small inputs only).

TEST INFRASTRUCTURE ONLY.  Independent of treat_oracle.c on purpose: the two
restatements are checked against each other and against the reference's goldens
(tests/test_oracle_golden.py), so a slip in one shows up as a disagreement.

Follows fragment.go:75-147, template.go:96-196, align.go:88-279,337-520 and the
published algorithm of github.com/aebruno/nwalgo@4a232086e3ad (see treat_oracle.c).
"""
from __future__ import annotations

import re
import struct

FORWARD, REVERSE = 1, -1
_fastx = re.compile(r".+[_\-](\d+)$")


def parseMergeCount(rec_id: str) -> int:  # fragment.go:75-89
    parts = rec_id.strip(" \t\n\v\f\r").split(" ", 1)
    m = _fastx.search(parts[0])  # Go FindStringSubmatch is unanchored on the left
    if m:
        v = int(m.group(1))
        if v > 2**63 - 1:
            return 1
        return v & 0xFFFFFFFF
    return 1


class Fragment:
    def __init__(self, name, seq, orientation=FORWARD, base="T"):  # fragment.go:91-121
        base = base.upper()
        if orientation == REVERSE:
            seq = seq[::-1]
        seq = seq.upper()
        self.Name, self.EditBase, self.Norm = name, base, 0.0
        self.EditSite, bases, run = [], [], 0
        for ch in seq:
            if ch != base:
                self.EditSite.append(run)
                run = 0
                bases.append(ch)
            else:
                run += 1
        self.EditSite.append(run)
        self.Bases = "".join(bases)
        self.ReadCount = parseMergeCount(name)

    def String(self):  # fragment.go:136-147
        out = []
        for i, b in enumerate(self.EditSite):
            out.append(self.EditBase * b)
            if i < len(self.Bases):
                out.append(self.Bases[i])
        return "".join(out)


class Template:
    def __init__(self, full, pre, alt=(), altRegion=()):  # template.go:96-151
        if full.Bases != pre.Bases or full.EditBase != pre.EditBase:
            raise ValueError("Invalid template sequences. Full and Pre templates must have the same non-edit bases")
        for a in alt:
            if full.Bases != a.Bases or full.EditBase != a.EditBase:
                raise ValueError("Invalid alt template sequence. All templates must have the same non-edit bases")
        if len(alt) != len(altRegion):
            raise ValueError("Invalid alt templates. Please specify the alt regions")
        self.Bases, self.EditBase, self.EditOffset = full.Bases, full.EditBase, 0
        self.EditSite = [list(full.EditSite), list(pre.EditSite)] + [list(a.EditSite) for a in alt]
        self.AltRegion = [list(r) for r in altRegion]
        self.BaseIndex, index = [], 0
        n = len(self.EditSite[0])
        for i in range(n):
            index += max(self.EditSite[0][i], self.EditSite[1][i])
            if i > 0 and i != n - 1:
                index += 1
            self.BaseIndex.append(index)
        self.EditStop = n - 1
        for j in range(n - 1, -1, -1):
            if self.EditSite[0][j] != self.EditSite[1][j]:
                self.EditStop = (n - 1) - j
                break
        self.EditStop -= 1

    def SetOffset(self, offset):  # template.go:153-161
        self.EditOffset = offset
        self.EditStop += offset
        for r in self.AltRegion:
            r[0] -= offset
            r[1] -= offset

    def Size(self):
        return len(self.EditSite)

    def Len(self):
        return len(self.EditSite[0])

    def Max(self, i):
        return max(es[i] for es in self.EditSite)


def read_fasta(path):
    """gofasta.SimpleParser under the assumptions stated in treat_oracle.c."""
    recs, cur = [], None
    with open(path, "rb") as fh:
        data = fh.read().decode("latin-1")
    started = False
    for line in data.split("\n"):
        if line.endswith("\r"):
            line = line[:-1]
        if not started:
            k = line.find(">")
            if k < 0:
                continue
            line = line[k:]
            started = True
        if line.startswith(">"):
            cur = [line[1:].strip(" \t\n\v\f\r"), []]
            recs.append(cur)
        elif cur is not None:
            cur[1].append(line)
    return [(i, "".join(s)) for i, s in recs]


def template_from_fasta(path, orientation=FORWARD, base="T"):  # template.go:51-94
    frags, regions = [], []
    for rid, seq in read_fasta(path):
        frags.append(Fragment(rid, seq, orientation, base))
        if len(frags) > 2:
            ms, me = re.search(r"\s*alt_start=(\d+)\s*", rid), re.search(r"\s*alt_stop=(\d+)\s*", rid)
            start = int(ms.group(1)) if ms else -1
            end = int(me.group(1)) if me else -1
            if start > -1 and end > -1:
                regions.append((start, end))
    if len(frags) < 2:
        raise ValueError("Must provide at least 2 templates. Full and Pre edited")
    return Template(frags[0], frags[1], frags[2:], regions)


def nw_align(a, b, match=1, mismatch=-1, gap=-1, tie="ULD"):
    """nwalgo.Align: Up (consume a) beats Left (consume b) beats Diagonal on ties."""
    al, bl = len(a) + 1, len(b) + 1
    f = [[0] * bl for _ in range(al)]
    p = [[0] * bl for _ in range(al)]
    for i in range(1, al):
        f[i][0], p[i][0] = gap * i, "U"
    for j in range(1, bl):
        f[0][j], p[0][j] = gap * j, "L"
    for i in range(1, al):
        for j in range(1, bl):
            d = f[i - 1][j - 1] + (match if a[i - 1] == b[j - 1] else mismatch)
            h = f[i - 1][j] + gap
            v = f[i][j - 1] + gap
            mx = max(d, h, v)
            cand = {"U": mx == h, "L": mx == v, "D": mx == d}
            p[i][j] = next(c for c in tie if cand[c])
            f[i][j] = mx
    x, y = [], []
    i, j = al - 1, bl - 1
    while i > 0 or j > 0:
        c = p[i][j]
        if c == "D":
            x.append(a[i - 1]); y.append(b[j - 1]); i -= 1; j -= 1
        elif c == "U":
            x.append(a[i - 1]); y.append("-"); i -= 1
        else:
            x.append("-"); y.append(b[j - 1]); j -= 1
    return "".join(reversed(x)), "".join(reversed(y)), f[al - 1][bl - 1]


class Alignment:
    def __init__(self):
        self.EditStop = self.JuncStart = self.JuncEnd = self.JuncLen = 0
        self.ReadCount, self.Norm = 0, 0.0
        self.HasMutation = self.Mismatches = self.Indel = self.AltEditing = 0
        self.JuncSeq, self.WouldPanic, self.Score = "", False, 0

    def fields(self):
        return dict(
            edit_stop=self.EditStop, junc_start=self.JuncStart, junc_end=self.JuncEnd, junc_len=self.JuncLen,
            read_count=self.ReadCount, has_mutation=self.HasMutation, mismatches=self.Mismatches,
            indel=self.Indel, alt_editing=self.AltEditing, junc_seq=self.JuncSeq, score=self.Score,
            would_panic=self.WouldPanic,
        )

    def computeT(self, frag, tmpl, excludeSnps, tie):  # align.go:122-181
        size = tmpl.Len()
        T = [[False] * size for _ in range(tmpl.Size())]
        aln1, aln2, self.Score = nw_align(tmpl.Bases, frag.Bases, 1, -1, -1, tie)
        fi = ti = 0
        for ai in range(len(aln1)):
            if aln1[ai] == "-":
                fi += 1
                self.HasMutation = self.Indel = 1
                continue
            count = 0
            if aln2[ai] != "-":
                count = frag.EditSite[fi]
                if frag.Bases[fi] != tmpl.Bases[ti]:
                    self.Mismatches = (self.Mismatches + 1) & 0xFF
                    if excludeSnps or self.Mismatches > 2:
                        self.HasMutation = 1
            else:
                self.HasMutation = self.Indel = 1
            for k in range(tmpl.Size()):
                if tmpl.EditSite[k][ti] == count:
                    T[k][(size - 1) - ti] = True
            if aln2[ai] != "-":
                fi += 1
            ti += 1
        for k in range(tmpl.Size()):
            if tmpl.EditSite[k][ti] == frag.EditSite[fi]:
                T[k][(size - 1) - ti] = True
        return T

    def WriteTo(self, frag, tmpl, tw=80, tie="ULD"):  # align.go:388-520
        if tw <= 0:
            tw = 80
        aln1, aln2, _ = nw_align(tmpl.Bases, frag.Bases, 1, -1, -1, tie)
        K = tmpl.Size()
        buf = [[] for _ in range(K + 1)]
        wb = lambda base, count, mx: "-" * (mx - count) + base * count  # align.go:88-93
        fi = ti = 0
        for ai in range(len(aln1)):
            if aln1[ai] == "-":
                for k in range(K):
                    buf[k].append(wb(tmpl.EditBase, 0, frag.EditSite[fi]) + "-")
                buf[K].append(wb(frag.EditBase, frag.EditSite[fi], frag.EditSite[fi]) + frag.Bases[fi])
                fi += 1
            elif aln2[ai] == "-":
                mx = tmpl.Max(ti)
                for k in range(K):
                    buf[k].append(wb(tmpl.EditBase, tmpl.EditSite[k][ti], mx) + tmpl.Bases[ti])
                buf[K].append(wb("-", 0, mx) + "-")
                ti += 1
            else:
                mx = max(tmpl.Max(ti), frag.EditSite[fi])
                for k in range(K):
                    buf[k].append(wb(tmpl.EditBase, tmpl.EditSite[k][ti], mx) + tmpl.Bases[ti])
                buf[K].append(wb(frag.EditBase, frag.EditSite[fi], mx) + frag.Bases[fi])
                fi += 1
                ti += 1
        mx = max(tmpl.Max(ti), frag.EditSite[fi])
        for k in range(K):
            buf[k].append(wb(tmpl.EditBase, tmpl.EditSite[k][ti], mx))
        buf[K].append(wb(frag.EditBase, frag.EditSite[fi], mx))
        rows_txt = ["".join(b) for b in buf]
        cols = tw - 4
        rows = len(rows_txt[0]) // cols + (1 if len(rows_txt[0]) % cols else 0)
        labels = ["FE", "PE"] + ["A%d" % (i + 1) for i in range(len(tmpl.AltRegion))] + ["RD"]
        out = ["=" * tw + "\n", frag.Name + "\n", "=" * tw + "\n"]
        out.append("JSS: %d\nESS: %d\nJES: %d\nJunc Len: %d\n" % (self.JuncStart, self.EditStop, self.JuncEnd, self.JuncLen))
        if self.AltEditing > 0:
            out.append("Alt: %d\n" % self.AltEditing)
        out.append("=" * tw + "\n\n")
        for r in range(rows):
            for i, b in enumerate(rows_txt):
                out.append(labels[i] + ": " + b[r * cols:min(r * cols + cols, len(b))] + "\n")
            out.append("\n")
        return "".join(out)

    def MarshalBinary(self):  # align.go:316-335
        js = self.JuncSeq.encode("latin-1")
        pk = lambda v: struct.pack(">I", v & 0xFFFFFFFF)
        return (pk(self.EditStop) + pk(self.JuncStart) + pk(self.JuncEnd) + pk(self.JuncLen) + pk(self.ReadCount)
                + bytes([self.HasMutation, self.AltEditing, self.Mismatches, self.Indel])
                + struct.pack(">d", self.Norm) + pk(len(js)) + js)


def _test(bits, i):  # bitset.Test: out of range (incl. uint(-1)) -> false
    return 0 <= i < len(bits) and bits[i]


def findJES(b):  # align.go:95-108
    if all(b):
        return -1
    if not any(b):
        return len(b) - 1
    return max(i for i, v in enumerate(b) if not v)


def findJSS(b):  # align.go:110-120
    if all(b):
        return len(b) - 1
    if not any(b):
        return -1
    return min(i for i, v in enumerate(b) if not v)


def NewAlignment(frag, tmpl, excludeSnps=False, tie="ULD"):  # align.go:237-279
    a = Alignment()
    T = a.computeT(frag, tmpl, excludeSnps, tie)
    a.JuncStart = findJSS(T[0])
    # computeAltEditing, align.go:183-235
    shift, alt, hasAlt = a.JuncStart, 0, False
    for i, v in enumerate(T[2:]):
        if _test(v, shift):
            alt, hasAlt = i, True
            break
    if hasAlt and shift == tmpl.AltRegion[alt][0]:
        fullMatch = True
        for x in range(shift, tmpl.Len()):
            if _test(T[alt + 2], x):
                continue
            fullMatch, shift = False, x
            break
        if fullMatch or shift >= tmpl.AltRegion[alt][1]:
            a.AltEditing = alt + 1
            if fullMatch:
                a.JuncStart = tmpl.Len() - 1
            else:
                for j in range(shift, tmpl.Len()):
                    if not _test(T[0], j):
                        a.JuncStart = j
                        break
    a.JuncEnd = findJES(T[1])
    a.EditStop = a.JuncStart - 1
    if a.JuncEnd > a.EditStop:
        a.JuncLen = a.JuncEnd - a.EditStop
        if a.HasMutation == 0:
            frm, to = (tmpl.Len() - 1) - a.JuncEnd, (tmpl.Len() - 1) - a.EditStop
            for i in range(frm, to):
                if i >= len(frag.EditSite):
                    a.WouldPanic = True
                    break
                a.JuncSeq += frag.EditBase * frag.EditSite[i]
                if i < len(frag.Bases):
                    a.JuncSeq += frag.Bases[i]
    elif a.JuncEnd < a.EditStop:
        a.JuncEnd, a.JuncLen = a.EditStop, 0
    if all(T[0]):
        a.JuncEnd = a.EditStop = a.JuncStart
        a.JuncLen, a.JuncSeq = 0, ""
    if a.WouldPanic:
        a.JuncSeq = ""
    a.ReadCount, a.Norm = frag.ReadCount, frag.Norm
    a.EditStop += tmpl.EditOffset
    a.JuncStart += tmpl.EditOffset
    a.JuncEnd += tmpl.EditOffset
    return a


def SimpleAlign(f1, f2, tie="ULD"):  # align.go:337-386
    aln1, aln2, _ = nw_align(f1.Bases, f2.Bases, 1, -1, -1, tie)
    wb = lambda base, count, mx: "-" * (mx - count) + base * count
    b0, b1, fi, ti = [], [], 0, 0
    for ai in range(len(aln1)):
        if aln1[ai] == "-":
            b0.append(wb(f1.EditBase, 0, f2.EditSite[fi]) + "-")
            b1.append(wb(f2.EditBase, f2.EditSite[fi], f2.EditSite[fi]) + f2.Bases[fi])
            fi += 1
        elif aln2[ai] == "-":
            b0.append(wb(f1.EditBase, f1.EditSite[ti], f1.EditSite[ti]) + f1.Bases[ti])
            b1.append(wb("-", 0, f1.EditSite[ti]) + "-")
            ti += 1
        else:
            mx = max(f1.EditSite[ti], f2.EditSite[fi])
            b0.append(wb(f1.EditBase, f1.EditSite[ti], mx) + f1.Bases[ti])
            b1.append(wb(f2.EditBase, f2.EditSite[fi], mx) + f2.Bases[fi])
            fi += 1
            ti += 1
    mx = max(f1.EditSite[ti], f2.EditSite[fi])
    b0.append(wb(f1.EditBase, f1.EditSite[ti], mx))
    b1.append(wb(f2.EditBase, f2.EditSite[fi], mx))
    return "".join(b0), "".join(b1)


def PrintAlignment(a1, a2, tw=80):  # cmd/treat/align.go:39-57
    out = []
    for r in range(0, len(a1), tw):
        out += [a1[r:r + tw] + "\n", a2[r:r + tw] + "\n", "\n"]
    return "".join(out)


def mutant(tmpl, frags, tie="ULD"):
    """cmd/treat/mutant.go:70-89: counts of the distinct gapped template strings (reads with an insertion)
    and gapped read strings (reads with a deletion).  Returns (tm, fm) dicts."""
    tm, fm = {}, {}
    for f in frags:
        aln1, aln2, _ = nw_align(tmpl.Bases, f.Bases, 1, -1, -1, tie)
        if "-" in aln1:
            tm[aln1] = tm.get(aln1, 0) + 1
        if "-" in aln2:
            fm[aln2] = fm.get(aln2, 0) + 1
    return tm, fm


def mutant_report(tm, fm, n):
    """cmd/treat/mutant.go:91-109 with ties ordered by string (Go's sort.Sort leaves them unspecified)."""
    out = ["Template Indels\n"]
    for cnt, (k, v) in enumerate(sorted(tm.items(), key=lambda kv: (-kv[1], kv[0])), 1):
        out.append("%d: %s\n" % (v, k))
        if cnt > n:
            break
    out.append("\nFragment Indels\n")
    for cnt, (k, v) in enumerate(sorted(fm.items(), key=lambda kv: (-kv[1], kv[0])), 1):
        out.append("%d: %s\n" % (v, k))
        if cnt > n:
            break
    return "".join(out)


# ---- `treat norm` (cmd/treat/norm.go:25-73, cmd/treat/storage.go:801-878) ---------------------------------------------
def normalize_default(samples):
    """norm.go:48-63: the default norm is the number of standard reads over all samples divided by the number of samples.
    samples: one list of Alignment per sample."""
    total = 0
    for alns in samples:
        for a in alns:
            if a.HasMutation == 0:  # Search(... HasMutation: false ...)
                total += int(a.ReadCount)
    return float(total) / float(len(samples))


def normalize_sample(alns, norm):
    """storage.go:801-878 NormalizeSample: scale = norm / (ReadCount summed over the standard reads), 1.0 for an empty
    total; Norm = scale * float64(ReadCount) for the standard reads, the others keep theirs.  Returns the scale."""
    total = 0
    for a in alns:
        if a.HasMutation == 0:
            total += int(a.ReadCount)
    scale = 1.0
    if total > 0:
        scale = norm / float(total)
    for a in alns:
        if a.HasMutation == 0:
            a.Norm = scale * float(a.ReadCount)
    return scale


# ---- Storage.Search's record filter (cmd/treat/storage.go:48-64, 162-188, 257-286) --------------------------------------
def has_match(a, edit_stop=-1, junc_end=-1, junc_len=-1, alt_region=0, has_mutation=False, has_alt=False, all=False):  # :162-188
    if not all:
        if has_mutation and a.HasMutation == 0:
            return False
        elif not has_mutation and a.HasMutation == 1:
            return False
    if edit_stop >= 0 and edit_stop != a.EditStop:
        return False
    if junc_len >= 0 and junc_len != a.JuncLen:
        return False
    if junc_end >= 0 and junc_end != a.JuncEnd:
        return False
    if has_alt and a.AltEditing == 0:
        return False
    if alt_region > 0 and (alt_region & 0xFF) != a.AltEditing:
        return False
    return True


def search(alns, offset=0, limit=0, **fields):  # :257-286, one sample's alignments in order; returns their indices
    out, count, off = [], 0, 0
    for i, a in enumerate(alns):
        if not has_match(a, **fields):
            continue
        if not fields.get("has_alt", False) and a.AltEditing > 0:  # by default, don't include alt editing
            continue
        if offset > 0 and off < offset:
            off += 1
            continue
        if limit > 0 and count >= limit:
            break
        out.append(i)
        count += 1
        off += 1
    return out


# ---- `treat search` output (cmd/treat/search.go:30-89) through Go's encoding/csv writer ---------------------------------
def _csv_field(f, comma):  # encoding/csv: Writer.fieldNeedsQuotes + the quoted form (UseCRLF false)
    need = False
    if f != "":
        if f == "\\.":
            need = True
        elif any(c in f for c in ("\n", "\r", '"', comma)):
            need = True
        else:
            need = f[0].isspace()  # unicode.IsSpace
    if not need:
        return f
    return '"' + f.replace('"', '""') + '"'


def round_plus(f, places):  # storage.go:190-198
    import math
    shift = math.pow(10, float(places))
    return math.floor(f * shift + .5) / shift


def search_text(gene, sample, alns, idxs, csv=False, header=True):
    comma = "," if csv else "\t"
    rows = []
    if header:
        rows.append(["gene", "sample", "norm", "read_count", "alt_editing", "has_mutation", "edit_stop", "junc_end", "junc_len", "junc_seq"])
    for i in idxs:
        a = alns[i]
        alt = "%d" % a.AltEditing
        if a.AltEditing != 0:
            alt = "A%d" % a.AltEditing
        rows.append([gene, sample, "%.4f" % round_plus(a.Norm, 4), "%d" % a.ReadCount, alt, "%d" % a.HasMutation, "%d" % a.EditStop,
                     "%d" % a.JuncEnd, "%d" % a.JuncLen, a.JuncSeq])
    return "".join(comma.join(_csv_field(f, comma) for f in r) + "\n" for r in rows)


# ---- highChartHist (cmd/treat/handlers.go:162-299): one sample's series --------------------------------------------------
def hist_norm(alns, which, tmpl_edit_stop, **fields):
    """samples[key.Sample][f(a)] += a.Norm over Storage.Search's matches (Limit = Offset = 0), without the reads at the
    template's edit stop that have no junction; which: 0 EditStop, 1 JuncLen, 2 JuncEnd.  Returns {value: sum}."""
    fields = dict(fields, offset=0, limit=0)
    out = {}
    for i in search(alns, **fields):
        a = alns[i]
        if a.EditStop == tmpl_edit_stop and a.JuncLen == 0:
            continue
        v = (a.EditStop, a.JuncLen, a.JuncEnd)[which]
        out[v] = out.get(v, 0.0) + a.Norm
    return out
