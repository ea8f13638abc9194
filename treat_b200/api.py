"""Python mirror of the reference's Go package API over the C ABI (include/treat_b200.h).

Names, argument meaning and error behaviour follow package `treat`
(fragment.go / template.go / align.go) so tests read like the reference's own tests:

    tmpl = NewTemplateFromFasta("templates.fa", FORWARD, "t"); tmpl.SetOffset(0)
    eng = Aligner(device=0)
    frag = NewFragment(rec_id, rec_seq, FORWARD, "t")
    aln = eng.NewAlignment(frag, tmpl, False)      # one read
    alns = eng.NewAlignments(reads, tmpl, False)   # the reference's per-read loop as one batch
    text = aln.WriteTo(frag, tmpl, 80)

All alignment work happens in libtreat_b200.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import cabi
from .cabi import (EXCLUDE_SNPS, FORWARD, NO_SUMMARY, OPS_GAPPED_ONLY, RECORD_DTYPE, REVERSE, WANT_OPS, TreatError, check, lib)

__all__ = [
    "FORWARD", "REVERSE", "Fragment", "Template", "Alignment", "Aligner", "FastaBatch", "NewFragment",
    "NewTemplate", "NewTemplateFromFasta", "parseMergeCount", "PrintAlignment", "read_fasta", "TreatError",
    "marshal_batch", "mutant_report", "DeviceGroup", "Comm", "norm_default", "norm_scale", "normalize_records", "search_records", "search_text", "hist_norm",
]


def _b(s) -> bytes:
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("latin-1")


def parseMergeCount(rec_id) -> int:
    """fragment.go:75-89"""
    b = _b(rec_id)
    return lib().treat_b200_parse_merge_count(b, len(b))


class Fragment:
    """fragment.go:57-64"""

    def __init__(self, name: str, seq, orientation: int, base: str):
        raw = _b(seq)
        self.Name = name
        self.Seq = raw  # the raw record; the device re-derives Bases/EditSite from it
        self.Orientation = orientation
        bases = np.zeros(max(1, len(raw)), dtype=np.uint8)
        sites = np.zeros(len(raw) + 1, dtype=np.uint32)
        n = C.c_size_t()
        check(lib().treat_b200_new_fragment(raw, len(raw), orientation, ord(_b(base).upper()), bases.ctypes.data,
                                            sites.ctypes.data, C.byref(n)))
        self.Bases = bases[: n.value].tobytes().decode("latin-1")
        self.EditSite = sites[: n.value + 1].tolist()
        self.EditBase = _b(base).upper().decode("latin-1")
        self.ReadCount = parseMergeCount(name)
        self.Norm = 0.0

    def Len(self) -> int:
        return len(self.EditSite)

    def String(self) -> str:
        """fragment.go:136-147"""
        out = []
        for i, c in enumerate(self.EditSite):
            out.append(self.EditBase * c)
            if i < len(self.Bases):
                out.append(self.Bases[i])
        return "".join(out)

    def ToFasta(self) -> str:
        return ">" + self.Name + "\n" + self.String()

    def forward_seq(self) -> bytes:
        """The record in 5'->3' orientation, as NewFragment sees it after the optional reversal."""
        return self.Seq[::-1] if self.Orientation == REVERSE else self.Seq


def NewFragment(name, seq, orientation: int = FORWARD, base: str = "T") -> Fragment:
    """fragment.go:91-121"""
    return Fragment(name, seq, orientation, base)


class Template:
    """template.go:41-49 (handle to the C++ treat::Template)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().treat_b200_template_free(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def SetOffset(self, offset: int) -> None:
        lib().treat_b200_template_set_offset(self._h, offset)

    def Size(self) -> int:
        return lib().treat_b200_template_size(self._h)

    def Len(self) -> int:
        return lib().treat_b200_template_len(self._h)

    def Max(self, i: int) -> int:
        return lib().treat_b200_template_max(self._h, i)

    def IndexLabel(self, i: int) -> int:
        return i + self.EditOffset

    @property
    def EditStop(self) -> int:
        return lib().treat_b200_template_edit_stop(self._h)

    @property
    def EditOffset(self) -> int:
        return lib().treat_b200_template_edit_offset(self._h)

    @property
    def EditBase(self) -> str:
        return chr(lib().treat_b200_template_edit_base(self._h))

    @property
    def Bases(self) -> str:
        return C.string_at(lib().treat_b200_template_bases(self._h), self.Len() - 1).decode("latin-1")

    @property
    def EditSite(self) -> List[List[int]]:
        n = self.Len()
        return [list((C.c_uint32 * n).from_address(lib().treat_b200_template_edit_site(self._h, k)))
                for k in range(self.Size())]

    @property
    def BaseIndex(self) -> List[int]:
        return list((C.c_uint32 * self.Len()).from_address(lib().treat_b200_template_base_index(self._h)))

    @property
    def AltRegion(self) -> List[Tuple[int, int]]:
        L = lib()
        return [(L.treat_b200_template_alt_start(self._h, i), L.treat_b200_template_alt_end(self._h, i))
                for i in range(L.treat_b200_template_n_alt(self._h))]

    def String(self) -> str:
        es, b = self.EditSite[0], self.Bases
        return "".join(self.EditBase * c + (b[i] if i < len(b) else "") for i, c in enumerate(es))


def NewTemplateFromFasta(path: str, orientation: int = FORWARD, base: str = "T") -> Template:
    """template.go:51-94; raises ValueError with the reference's message on its error paths."""
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().treat_b200_template_from_fasta(_b(path), orientation, ord(_b(base)), 0, C.byref(h), err, 512)
    if rc != cabi.OK:
        raise ValueError(err.value.decode() or lib().treat_b200_strerror(rc).decode())
    return Template(h)


def NewTemplate(full: Fragment, pre: Fragment, alt: Optional[Sequence[Fragment]] = None,
                altRegion: Optional[Sequence] = None) -> Template:
    """template.go:96-151"""
    alt = list(alt or [])
    altRegion = list(altRegion or [])
    frs = [full, pre] + alt
    if any(f.EditBase != full.EditBase for f in frs):
        raise ValueError("Invalid template sequences. Full and Pre templates must have the same non-edit bases")
    arr = (C.c_char_p * len(frs))(*[f.forward_seq() for f in frs])
    st = np.array([r[0] if r else 0 for r in altRegion] or [0], dtype=np.int32)
    en = np.array([r[1] if r else 0 for r in altRegion] or [0], dtype=np.int32)
    h = C.c_void_p()
    err = C.create_string_buffer(512)
    rc = lib().treat_b200_template_new(arr, len(frs), st.ctypes.data, en.ctypes.data, len(altRegion), FORWARD,
                                       ord(_b(full.EditBase)), C.byref(h), err, 512)
    if rc != cabi.OK:
        raise ValueError(err.value.decode() or lib().treat_b200_strerror(rc).decode())
    return Template(h)


class FastaBatch:
    """Concatenated reads + offsets + ReadCounts: the batch layout of treat_b200_align_batch."""

    def __init__(self, ids: List[str], seqs: np.ndarray, offsets: np.ndarray, read_counts: np.ndarray):
        self.ids, self.seqs, self.offsets, self.read_counts = ids, seqs, offsets, read_counts

    def __len__(self):
        return len(self.offsets) - 1

    def seq(self, i: int) -> bytes:
        return self.seqs[int(self.offsets[i]): int(self.offsets[i + 1])].tobytes()

    @staticmethod
    def from_records(records: Iterable[Tuple[str, object]]) -> "FastaBatch":
        ids, chunks = [], []
        for rid, seq in records:
            ids.append(rid)
            chunks.append(_b(seq))
        off = np.zeros(len(ids) + 1, dtype=np.uint64)
        if ids:
            off[1:] = np.cumsum([len(c) for c in chunks], dtype=np.uint64)
        seqs = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if ids else np.zeros(0, np.uint8)
        rc = np.array([parseMergeCount(i) for i in ids], dtype=np.uint32)
        return FastaBatch(ids, seqs, off, rc)


def fasta_split(data, parts: int) -> np.ndarray:
    """treat_b200_fasta_split: offsets[parts + 1] of slices of a FASTA file's bytes that start at record starts."""
    if isinstance(data, (bytes, bytearray)):
        data = np.frombuffer(data, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    off = np.zeros(parts + 1, dtype=np.uint64)
    check(lib().treat_b200_fasta_split(data.ctypes.data if data.size else None, data.size, parts, off.ctypes.data))
    return off


def read_fasta(path: str) -> FastaBatch:
    """gofasta.SimpleParser record semantics (cmd/treat/align.go:114), whole file as one batch."""
    h = C.c_void_p()
    check(lib().treat_b200_fasta_read(_b(path), C.byref(h)))
    return _fasta_batch(h)


def parse_fasta(data) -> FastaBatch:
    """The same on bytes in memory (treat_b200_fasta_parse)."""
    if isinstance(data, (bytes, bytearray)):
        data = np.frombuffer(data, dtype=np.uint8)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    h = C.c_void_p()
    check(lib().treat_b200_fasta_parse(data.ctypes.data if data.size else None, data.size, C.byref(h)))
    return _fasta_batch(h)


def _fasta_batch(h) -> FastaBatch:
    try:
        L = lib()
        n = L.treat_b200_fasta_count(h)
        ids = [L.treat_b200_fasta_id(h, i).decode("latin-1") for i in range(n)]
        off = np.array((C.c_uint64 * (n + 1)).from_address(L.treat_b200_fasta_offsets(h)), dtype=np.uint64) if n else np.zeros(1, np.uint64)
        total = int(off[-1])
        seqs = np.array((C.c_uint8 * total).from_address(L.treat_b200_fasta_seqs(h)), dtype=np.uint8) if total else np.zeros(0, np.uint8)
        rc = np.array((C.c_uint32 * n).from_address(L.treat_b200_fasta_read_counts(h)), dtype=np.uint32) if n else np.zeros(0, np.uint32)
        return FastaBatch(ids, seqs, off, rc)
    finally:
        lib().treat_b200_fasta_free(h)


def marshal_batch(rec: np.ndarray, seqs: np.ndarray, seq_off: np.ndarray, threads: int = 4):
    """Alignment.MarshalBinary (align.go:316-335) for a whole batch: returns (blob bytes, offsets[N+1])."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    N = len(rec)
    dummy = np.zeros(1, np.uint8)
    sp = seqs.ctypes.data if seqs.size else dummy.ctypes.data
    need = lib().treat_b200_marshal_batch(rec.ctypes.data, sp, seq_off.ctypes.data, N, None, 0, None, 1)
    if need < 0:
        raise TreatError(int(need))
    out = np.zeros(max(int(need), 1), dtype=np.uint8)
    off = np.zeros(N + 1, dtype=np.uint64)
    got = lib().treat_b200_marshal_batch(rec.ctypes.data, sp, seq_off.ctypes.data, N, out.ctypes.data, int(need),
                                         off.ctypes.data, threads)
    if got < 0:
        raise TreatError(int(got))
    return out[: int(need)], off


def _search_fields(edit_stop=-1, junc_end=-1, junc_len=-1, offset=0, limit=0, alt_region=0, has_mutation=False, has_alt=False, all=False):
    from .cabi import Search
    return Search(int(edit_stop), int(junc_end), int(junc_len), int(offset), int(limit), int(alt_region), int(bool(has_mutation)),
                  int(bool(has_alt)), int(bool(all)), 0)


def search_records(rec: np.ndarray, **fields) -> np.ndarray:
    """Indices of the records Storage.Search would hand out (cmd/treat/storage.go:162-188, 257-286), in record order."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    f = _search_fields(**fields)
    n = C.c_uint64(0)
    idx = np.zeros(max(len(rec), 1), dtype=np.uint64)
    check(lib().treat_b200_search_records(C.byref(f), rec.ctypes.data, len(rec), idx.ctypes.data, len(idx), C.byref(n)))
    return idx[: int(n.value)]


def search_text(gene: str, sample: str, rec: np.ndarray, seqs: np.ndarray, seq_off: np.ndarray, idx, norm: np.ndarray = None,
                csv: bool = False, header: bool = True) -> str:
    """`treat search` output (cmd/treat/search.go:30-89) for the records idx of one gene / sample."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    idx = np.ascontiguousarray(idx, dtype=np.uint64)
    nv = None if norm is None else np.ascontiguousarray(norm, dtype=np.float64)
    dummy = np.zeros(1, np.uint8)
    out, ln = C.c_void_p(), C.c_size_t()
    check(lib().treat_b200_search_text(_b(gene), _b(sample), rec.ctypes.data, None if nv is None else nv.ctypes.data,
                                       (seqs if seqs.size else dummy).ctypes.data_as(C.POINTER(C.c_uint8)), seq_off.ctypes.data,
                                       idx.ctypes.data if idx.size else None, len(idx), int(csv), int(header), C.byref(out), C.byref(ln)))
    try:
        return C.string_at(out.value, ln.value).decode("latin-1")
    finally:
        lib().treat_b200_free(out)


def hist_norm(rec: np.ndarray, norm: np.ndarray, tmpl_edit_stop: int, which: int, lo: int, hi: int, **fields) -> np.ndarray:
    """highChartHist (cmd/treat/handlers.go:162-299): Norm-weighted histogram of EditStop (0) / JuncLen (1) / JuncEnd (2) over the
    matching records of one sample; out[v - lo] for v in [lo, hi]."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    norm = np.ascontiguousarray(norm, dtype=np.float64)
    f = _search_fields(**fields)
    out = np.zeros(hi - lo + 1, dtype=np.float64)
    check(lib().treat_b200_hist_norm(C.byref(f), rec.ctypes.data, norm.ctypes.data, len(rec), int(tmpl_edit_stop), int(which), int(lo), int(hi),
                                     out.ctypes.data))
    return out


def norm_default(std_reads_per_sample) -> float:
    """`treat norm` without --norm: the average of the samples' standard-read totals (cmd/treat/norm.go:48-63)."""
    v = np.ascontiguousarray(std_reads_per_sample, dtype=np.uint64)
    out = C.c_double(0.0)
    check(lib().treat_b200_norm_default(v.ctypes.data, len(v), C.byref(out)))
    return float(out.value)


def norm_scale(std_reads_of_sample: int, norm: float) -> float:
    """NormalizeSample's scaling factor (cmd/treat/storage.go:840-843)."""
    out = C.c_double(0.0)
    check(lib().treat_b200_norm_scale(int(std_reads_of_sample), float(norm), C.byref(out)))
    return float(out.value)


def normalize_records(rec: np.ndarray, scale: float, blob: np.ndarray = None, blob_off: np.ndarray = None) -> np.ndarray:
    """Norm = scale * ReadCount of the standard reads (cmd/treat/storage.go:845-870); patches `blob` (from marshal_batch) in
    place when given.  Returns the float64 Norm of every record (0.0 for reads with a mutation)."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    out = np.zeros(len(rec), dtype=np.float64)
    bp = op = None
    if blob is not None:
        assert blob.dtype == np.uint8 and blob.flags["C_CONTIGUOUS"] and blob_off is not None
        blob_off = np.ascontiguousarray(blob_off, dtype=np.uint64)
        bp, op = blob.ctypes.data, blob_off.ctypes.data
    check(lib().treat_b200_normalize_records(rec.ctypes.data, len(rec), float(scale), out.ctypes.data, bp, op))
    return out


def mutant_report(tmpl: "Template", rec: np.ndarray, ops: np.ndarray, seqs: np.ndarray, seq_off: np.ndarray, n: int = 10) -> str:
    """`treat mutant` text (cmd/treat/mutant.go:55-110) from the records + ops of an align call."""
    rec = np.ascontiguousarray(rec, dtype=RECORD_DTYPE)
    ops = np.ascontiguousarray(ops, dtype=np.uint8)
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    out, ln = C.c_void_p(), C.c_size_t()
    dummy = np.zeros(1, np.uint8)
    check(lib().treat_b200_mutant_report(tmpl._h, rec.ctypes.data, ops.ctypes.data, ops.shape[1] if ops.ndim == 2 else 0,
                                         seqs.ctypes.data if seqs.size else dummy.ctypes.data, seq_off.ctypes.data,
                                         len(rec), n, C.byref(out), C.byref(ln)))
    try:
        return C.string_at(out.value, ln.value).decode("latin-1")
    finally:
        lib().treat_b200_free(out)


def UnmarshalBinary(blob: bytes) -> dict:
    """Alignment.UnmarshalBinary (align.go:295-314): the stored fields of a `treat load` blob."""
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    norm, js = C.c_double(), C.c_void_p()
    buf = C.create_string_buffer(bytes(blob), len(blob))
    check(lib().treat_b200_unmarshal_record(C.cast(buf, C.c_void_p), len(blob), rec.ctypes.data, C.byref(norm), C.byref(js)))
    r = rec[0]
    return dict(EditStop=int(r["edit_stop"]), JuncStart=int(r["junc_start"]), JuncEnd=int(r["junc_end"]), JuncLen=int(r["junc_len"]),
                ReadCount=int(r["read_count"]), HasMutation=int(r["has_mutation"]), AltEditing=int(r["alt_editing"]),
                Mismatches=int(r["mismatches"]), Indel=int(r["indel"]), Norm=norm.value,
                JuncSeq=C.string_at(js.value, int(r["junc_seq_len"])).decode("latin-1"))


def stats_text(tmpl: "Template", gene: str, samples: Sequence[Tuple[str, np.ndarray]], unique: bool = False, norm: bool = False) -> str:
    """`treat stats` block of one gene (cmd/treat/stats.go:62-137) from (sample name, summary counters) pairs."""
    n = len(samples)
    names = (C.c_char_p * max(n, 1))(*[_b(s[0]) for s in samples])
    arrs = [np.ascontiguousarray(s[1], dtype=np.uint64) for s in samples]
    ptrs = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    out, ln = C.c_void_p(), C.c_size_t()
    check(lib().treat_b200_stats_text(tmpl._h, _b(gene), names, ptrs, n, int(unique), int(norm), C.byref(out), C.byref(ln)))
    try:
        return C.string_at(out.value, ln.value).decode("latin-1")
    finally:
        lib().treat_b200_free(out)


def PrintAlignment(a1: str, a2: str, tw: int = 80) -> str:
    """cmd/treat/align.go:39-57"""
    out = []
    for r in range(0, len(a1), tw):
        out += [a1[r:r + tw], "\n", a2[r:r + tw], "\n", "\n"]
    return "".join(out)


class Alignment:
    """align.go:41-55, filled from one treat_b200_record."""

    def __init__(self, rec: np.void, ops: Optional[np.ndarray], seq: bytes, engine: "Aligner"):
        self._rec = np.array([rec], dtype=RECORD_DTYPE)
        self._ops = None if ops is None else np.ascontiguousarray(ops, dtype=np.uint8)
        self._seq = seq
        self._eng = engine
        r = self._rec[0]
        self.EditStop, self.JuncStart, self.JuncEnd, self.JuncLen = (int(r["edit_stop"]), int(r["junc_start"]),
                                                                    int(r["junc_end"]), int(r["junc_len"]))
        self.ReadCount, self.Norm = int(r["read_count"]), 0.0
        self.HasMutation, self.Mismatches, self.Indel, self.AltEditing = (int(r["has_mutation"]), int(r["mismatches"]),
                                                                          int(r["indel"]), int(r["alt_editing"]))
        self.Score, self.Status = int(r["score"]), int(r["status"])
        o, n = int(r["junc_seq_off"]), int(r["junc_seq_len"])
        self.JuncSeq = seq[o:o + n].decode("latin-1").upper()
        self.WouldPanic = bool(self.Status & cabi.ST_REF_PANIC)

    def fields(self) -> dict:
        return dict(
            edit_stop=self.EditStop, junc_start=self.JuncStart, junc_end=self.JuncEnd, junc_len=self.JuncLen,
            read_count=self.ReadCount, has_mutation=self.HasMutation, mismatches=self.Mismatches,
            indel=self.Indel, alt_editing=self.AltEditing, junc_seq=self.JuncSeq, score=self.Score,
            would_panic=self.WouldPanic,
        )

    def ops(self) -> List[int]:
        n = int(self._rec[0]["aln_len"])
        if not int(self._rec[0]["indel"]):
            return [0] * n  # base over base; with OPS_GAPPED_ONLY the row is not even written
        return [(int(self._ops[c >> 2]) >> (2 * (c & 3))) & 3 for c in range(n)]

    def gapped(self, tmpl_bases: str, frag_bases: str) -> Tuple[str, str]:
        """nwalgo.Align's two strings rebuilt from the device ops."""
        a, b, i, j = [], [], 0, 0
        for op in self.ops():
            if op == cabi.OP_INS:
                a.append("-"); b.append(frag_bases[j]); j += 1
            elif op == cabi.OP_DEL:
                a.append(tmpl_bases[i]); b.append("-"); i += 1
            else:
                a.append(tmpl_bases[i]); b.append(frag_bases[j]); i += 1; j += 1
        return "".join(a), "".join(b)

    def WriteTo(self, frag: Fragment, tmpl: Template, tw: int = 80) -> str:
        """align.go:388-520 (text assembled on the host from the device's record + ops)."""
        if self._ops is None:
            raise ValueError("alignment was computed without ops")
        seq = frag.forward_seq()
        name = _b(frag.Name)
        need = lib().treat_b200_format_alignment(tmpl._h, self._rec.ctypes.data, self._ops.ctypes.data, name, seq,
                                                 len(seq), tw, None, 0)
        if need < 0:
            raise TreatError(int(need))
        buf = C.create_string_buffer(int(need) + 1)
        lib().treat_b200_format_alignment(tmpl._h, self._rec.ctypes.data, self._ops.ctypes.data, name, seq, len(seq),
                                          tw, buf, int(need))
        return buf.raw[: int(need)].decode("latin-1")

    def MarshalBinary(self) -> bytes:
        """align.go:316-335"""
        need = lib().treat_b200_marshal_record(self._rec.ctypes.data, self._seq, len(self._seq), None, 0)
        buf = (C.c_uint8 * int(need))()
        lib().treat_b200_marshal_record(self._rec.ctypes.data, self._seq, len(self._seq), buf, int(need))
        return bytes(buf)

    def SimpleAlign(self, f1: Fragment, f2: Fragment) -> Tuple[str, str]:
        return self._eng.SimpleAlign(f1, f2)


class Aligner:
    """One GPU context (one process per GPU).  Wraps treat_b200_create / set_template / align_batch."""

    def __init__(self, device: int = 0):
        self._ctx = C.c_void_p()
        check(lib().treat_b200_create(device, C.byref(self._ctx)))
        self.device = device
        self._tmpl = None

    def close(self):
        if getattr(self, "_ctx", None):
            lib().treat_b200_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    # ---- template --------------------------------------------------------------------------
    def SetTemplate(self, tmpl: Template) -> None:
        check(lib().treat_b200_use_template(self._ctx, tmpl._h), self._ctx)
        # the context keeps a device copy; remember the identity + offset state it was made from
        self._tmpl = (tmpl, tmpl.EditOffset)

    def _ensure_template(self, tmpl: Template) -> None:
        if self._tmpl is None or self._tmpl[0] is not tmpl or self._tmpl[1] != tmpl.EditOffset:
            self.SetTemplate(tmpl)

    # ---- raw batch entry points ---------------------------------------------------------------
    def ops_stride(self, max_len: int) -> int:
        return lib().treat_b200_ops_stride(self._ctx, max_len)

    def align_batch(self, seqs: np.ndarray, seq_off: np.ndarray, read_count: Optional[np.ndarray] = None,
                    flags: int = WANT_OPS, rec: Optional[np.ndarray] = None, ops: Optional[np.ndarray] = None):
        """treat_b200_align_batch on host buffers; returns (records, ops or None)."""
        if not (isinstance(seqs, np.ndarray) and seqs.dtype == np.uint8 and seqs.flags.c_contiguous):
            seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        if not (isinstance(seq_off, np.ndarray) and seq_off.dtype == np.uint64 and seq_off.flags.c_contiguous):
            seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        N = len(seq_off) - 1
        if rec is None:
            rec = np.zeros(N, dtype=RECORD_DTYPE)
        stride = 0
        if flags & WANT_OPS:
            if ops is None:
                max_len = int((seq_off[1:] - seq_off[:-1]).max()) if N else 0
                stride = self.ops_stride(max_len)
                ops = np.zeros((N, stride), dtype=np.uint8)
            else:
                stride = ops.shape[1]
        rc = None if read_count is None else np.ascontiguousarray(read_count, dtype=np.uint32)
        dummy = np.zeros(1, np.uint8)
        check(lib().treat_b200_align_batch(
            self._ctx, seqs.ctypes.data if seqs.size else dummy.ctypes.data, seq_off.ctypes.data,
            rc.ctypes.data if rc is not None else None, N, flags, rec.ctypes.data,
            ops.ctypes.data if (flags & WANT_OPS) else None, stride), self._ctx)
        return rec, (ops if flags & WANT_OPS else None)

    def collapse_align_batch(self, seqs: np.ndarray, seq_off: np.ndarray, read_count: Optional[np.ndarray] = None,
                             flags: int = WANT_OPS, want_members: bool = True):
        """treat_b200_collapse_align_batch: exact duplicates (equal upper-cased sequences) are collapsed on the device, only the
        distinct sequences are aligned, with the members' ReadCounts summed (fastx_collapser upstream of `treat load`).
        Returns (records[n_unique], ops[n_unique] or None, first[n_unique], member[N] or None)."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        N = len(seq_off) - 1
        rec = np.zeros(N, dtype=RECORD_DTYPE)
        stride, ops = 0, None
        if flags & WANT_OPS:
            max_len = int((seq_off[1:] - seq_off[:-1]).max()) if N else 0
            stride = self.ops_stride(max_len)
            ops = np.zeros((N, stride), dtype=np.uint8)
        first = np.zeros(N, dtype=np.uint32)
        member = np.zeros(N, dtype=np.uint32) if want_members else None
        rc = None if read_count is None else np.ascontiguousarray(read_count, dtype=np.uint32)
        nu = C.c_uint64()
        dummy = np.zeros(1, np.uint8)
        check(lib().treat_b200_collapse_align_batch(
            self._ctx, seqs.ctypes.data if seqs.size else dummy.ctypes.data, seq_off.ctypes.data, rc.ctypes.data if rc is not None else None,
            N, flags, rec.ctypes.data, ops.ctypes.data if ops is not None else None, stride, first.ctypes.data,
            member.ctypes.data if member is not None else None, C.byref(nu)), self._ctx)
        k = nu.value
        return rec[:k], (ops[:k] if ops is not None else None), first[:k], member

    def align_batch_device(self, d_seqs: int, d_seq_off: int, d_read_count: int, N: int, max_len: int, flags: int,
                           d_rec: int, d_ops: int, ops_stride: int, stream: int = 0) -> None:
        """treat_b200_align_batch_device on raw device pointers (e.g. torch tensors' data_ptr())."""
        check(lib().treat_b200_align_batch_device(self._ctx, d_seqs, d_seq_off, d_read_count or None, N, max_len, flags,
                                                  d_rec, d_ops or None, ops_stride, stream or None), self._ctx)

    def upload_fasta(self, data) -> Tuple[int, int]:
        """treat_b200_fasta_upload: FASTA bytes (bytes, or a uint8 array, e.g. pinned) -> (records, longest sequence).
        The parse runs on the GPU; the file stays resident for align_fasta()."""
        if isinstance(data, (bytes, bytearray)):
            data = np.frombuffer(data, dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        n, ml = C.c_uint64(), C.c_uint32()
        check(lib().treat_b200_fasta_upload(self._ctx, data.ctypes.data, data.size, C.byref(n), C.byref(ml)), self._ctx)
        self._fasta = (data, n.value, ml.value)
        return n.value, ml.value

    def align_fasta(self, flags: int = WANT_OPS, rec: Optional[np.ndarray] = None, ops: Optional[np.ndarray] = None,
                    want_ids: bool = True):
        """treat_b200_fasta_align on the uploaded file: (records, ops or None, ids or None, read counts)."""
        data, N, ml = self._fasta
        if rec is None:
            rec = np.zeros(N, dtype=RECORD_DTYPE)
        stride = 0
        if flags & WANT_OPS:
            if ops is None:
                stride = self.ops_stride(ml)
                ops = np.zeros((N, stride), dtype=np.uint8)
            else:
                stride = ops.shape[1]
        id_off = np.zeros(N, dtype=np.uint64)
        id_len = np.zeros(N, dtype=np.uint32)
        rcs = np.zeros(N, dtype=np.uint32)
        check(lib().treat_b200_fasta_align(self._ctx, flags, rec.ctypes.data, ops.ctypes.data if (flags & WANT_OPS) else None, stride,
                                           id_off.ctypes.data, id_len.ctypes.data, rcs.ctypes.data, None, None), self._ctx)
        ids = None
        if want_ids:
            raw = data.tobytes()
            ids = [raw[int(o):int(o) + int(n)].decode("latin-1") for o, n in zip(id_off, id_len)]
        return rec, (ops if flags & WANT_OPS else None), ids, rcs

    def stream_fasta(self, data, flags: int = 0, ops_stride: int = 0):
        """treat_b200_fasta_stream: yields one dict of numpy arrays (copies) per piece, in file order."""
        from .cabi import FASTA_SINK
        if isinstance(data, (bytes, bytearray)):
            data = np.frombuffer(data, dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        pieces = []

        def arr(ptr, n, dt):
            if not ptr or not n:
                return np.zeros(0, dtype=dt)
            nbytes = int(n) * np.dtype(dt).itemsize
            return np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=dt).copy()

        def sink(_user, pp):
            q = pp.contents
            n = q.n
            d = dict(first_record=q.first_record, n=n, rec=arr(q.rec, n, RECORD_DTYPE), id_off=arr(q.id_off, n, np.uint64),
                     seq_off=arr(q.seq_off, n, np.uint64), id_len=arr(q.id_len, n, np.uint32), read_count=arr(q.read_count, n, np.uint32),
                     seq_len=arr(q.seq_len, n, np.uint32), ops_stride=q.ops_stride, seq_in_place=bool(q.seq_in_place), ops=None, ops_index=None)
            if q.ops and q.n_ops_rows:
                d["ops"] = arr(q.ops, q.n_ops_rows * q.ops_stride, np.uint8).reshape(q.n_ops_rows, q.ops_stride)
            elif q.ops_stride:
                d["ops"] = np.zeros((0, q.ops_stride), dtype=np.uint8)
            if q.ops_index:
                d["ops_index"] = arr(q.ops_index, q.n_ops_rows, np.uint32)
            d["blob_off"] = d["blob"] = None
            if q.blob_off:
                d["blob_off"] = arr(q.blob_off, n + 1, np.uint64)
                d["blob"] = arr(q.blob, int(d["blob_off"][-1]), np.uint8)
            pieces.append(d)
            return 0
        cb = FASTA_SINK(sink)
        total = C.c_uint64()
        check(lib().treat_b200_fasta_stream(self._ctx, data.ctypes.data, data.size, flags, ops_stride, cb, None, C.byref(total)), self._ctx)
        return pieces, total.value

    def text_device(self, d_seqs: int, d_seq_off: int, d_names: int, d_name_off: int, d_rec: int, d_ops: int, ops_stride: int,
                    N: int, max_len: int, tw: int, d_text_off: int, d_text: int = 0, text_cap: int = 0, stream: int = 0) -> int:
        """treat_b200_text_device on raw device pointers: Alignment.WriteTo for a batch; returns the text size in bytes
        (d_text = 0: sizing call)."""
        total = C.c_uint64()
        check(lib().treat_b200_text_device(self._ctx, d_seqs, d_seq_off, d_names, d_name_off, d_rec, d_ops, ops_stride, N, max_len, tw,
                                           d_text_off, d_text or None, text_cap, C.byref(total), stream or None), self._ctx)
        return total.value

    def search_device(self, d_rec: int, N: int, d_idx: int, cap: int, stream: int = 0, **fields) -> int:
        """Storage.Search's record filter for records in HBM: the matching indices go to d_idx (uint64), returns their number."""
        f = _search_fields(**fields)
        n = C.c_uint64(0)
        check(lib().treat_b200_search_device(self._ctx, C.byref(f), d_rec, N, d_idx or None, cap, C.byref(n), stream or None), self._ctx)
        return int(n.value)

    def normalize_device(self, d_rec: int, N: int, scale: float, d_norm: int = 0, d_blob: int = 0, d_blob_off: int = 0, stream: int = 0):
        """`treat norm` for records (and marshal_device's blobs) in HBM: Norm = scale * ReadCount of the standard reads."""
        check(lib().treat_b200_normalize_device(self._ctx, d_rec, N, float(scale), d_norm or None, d_blob or None, d_blob_off or None,
                                                stream or None), self._ctx)

    def marshal_device(self, d_seqs: int, d_seq_off: int, d_rec: int, N: int, d_blob_off: int, d_blob: int = 0, blob_cap: int = 0,
                       stream: int = 0) -> int:
        """treat_b200_marshal_device on raw device pointers: Alignment.MarshalBinary for a batch; returns the size in bytes
        (d_blob = 0: sizing call)."""
        total = C.c_uint64()
        check(lib().treat_b200_marshal_device(self._ctx, d_seqs, d_seq_off, d_rec, N, d_blob_off, d_blob or None, blob_cap,
                                              C.byref(total), stream or None), self._ctx)
        return total.value

    def text_batch(self, seqs: np.ndarray, seq_off: np.ndarray, names: Sequence, flags: int = 0, tw: int = 80, out=None) -> bytes:
        """treat_b200_text_batch: NewFragment + NewAlignment + WriteTo for reads the caller holds (ids in `names`); returns the
        text `treat align` would print for them (or writes the chunks to `out` as they arrive)."""
        from .cabi import TEXT_SINK
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        enc = [_b(x) for x in names]
        nbuf = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
        noff = np.zeros(len(enc) + 1, dtype=np.uint64)
        noff[1:] = np.cumsum([len(x) for x in enc])
        parts = []

        def sink(_user, ptr, n, _first, _nrec):
            buf = C.string_at(ptr, n) if n else b""
            if out is not None:
                out.write(buf)
            else:
                parts.append(buf)
            return 0
        cb = TEXT_SINK(sink)
        dummy = np.zeros(1, dtype=np.uint8)
        check(lib().treat_b200_text_batch(self._ctx, seqs.ctypes.data if seqs.size else dummy.ctypes.data, seq_off.ctypes.data,
                                          nbuf.ctypes.data, noff.ctypes.data, len(enc), flags, tw, cb, None), self._ctx)
        return b"".join(parts)

    def fasta_text(self, data, flags: int = 0, tw: int = 80, out=None) -> Tuple[bytes, int]:
        """treat_b200_fasta_text: FASTA bytes -> the text `treat align -t ... -f ...` prints (parse, alignment and
        WriteTo on the GPU).  Returns (text, records); with `out` (a binary file object) the pieces are written there
        as they arrive and the returned text is empty."""
        from .cabi import TEXT_SINK
        if isinstance(data, (bytes, bytearray)):
            data = np.frombuffer(data, dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        parts = []

        def sink(_user, ptr, n, _first, _nrec):
            buf = C.string_at(ptr, n) if n else b""
            if out is not None:
                out.write(buf)
            else:
                parts.append(buf)
            return 0
        cb = TEXT_SINK(sink)
        total = C.c_uint64()
        check(lib().treat_b200_fasta_text(self._ctx, data.ctypes.data, data.size, flags, tw, cb, None, C.byref(total)), self._ctx)
        return b"".join(parts), total.value

    def synchronize(self) -> None:
        check(lib().treat_b200_synchronize(self._ctx), self._ctx)

    def set_tie_order(self, order: str = "ULD") -> None:
        """Traceback tie order of the DP kernels: "ULD" (nwalgo as published), "UDL" or "LUD" (treat_b200_set_tie_order)."""
        check(lib().treat_b200_set_tie_order(self._ctx, cabi.TIE_ORDERS.index(order)), self._ctx)

    def debug_flags(self) -> Tuple[int, bool]:
        """(flag word, checks compiled in): non-zero flags mean a kernel index left its bounds."""
        f, c = C.c_uint32(), C.c_int32()
        check(lib().treat_b200_debug_flags(self._ctx, C.byref(f), C.byref(c)), self._ctx)
        return f.value, bool(c.value)

    def set_band_min_reads(self, reads: int) -> None:
        """The banded fill runs only for launches with at least this many reads to align (treat_b200_set_band_min_reads)."""
        check(lib().treat_b200_set_band_min_reads(self._ctx, int(reads)), self._ctx)

    def dp_stats(self, reset: bool = True) -> Tuple[int, int, int]:
        """Reads that needed a DP since the last reset: (pure deletions finished without one, banded fill, full matrix)."""
        out = (C.c_uint64 * 3)()
        check(lib().treat_b200_dp_stats(self._ctx, out, 1 if reset else 0), self._ctx)
        return int(out[0]), int(out[1]), int(out[2])

    def launch_count(self) -> int:
        return lib().treat_b200_launch_count(self._ctx)

    # ---- summaries -------------------------------------------------------------------------------
    def summary(self) -> np.ndarray:
        n = lib().treat_b200_summary_len(self._ctx)
        out = np.zeros(n, dtype=np.uint64)
        check(lib().treat_b200_get_summary(self._ctx, out.ctypes.data, n), self._ctx)
        return out

    def reset_summary(self) -> None:
        check(lib().treat_b200_reset_summary(self._ctx), self._ctx)

    def heatmap(self) -> np.ndarray:
        """ESS x junction-length heat map [dim, dim] accumulated by align calls with the HEATMAP flag (handlers.go:578-589)."""
        d = lib().treat_b200_heatmap_dim(self._ctx)
        out = np.zeros((d, d), dtype=np.uint64)
        check(lib().treat_b200_get_heatmap(self._ctx, out.ctypes.data, d * d), self._ctx)
        return out

    def heatmap_device_ptr(self) -> Tuple[int, int]:
        p = C.c_void_p()
        check(lib().treat_b200_heatmap_device_ptr(self._ctx, C.byref(p)), self._ctx)
        d = lib().treat_b200_heatmap_dim(self._ctx)
        return p.value, d * d

    def summary_device_ptr(self) -> Tuple[int, int]:
        p = C.c_void_p()
        check(lib().treat_b200_summary_device_ptr(self._ctx, C.byref(p)), self._ctx)
        return p.value, lib().treat_b200_summary_len(self._ctx)

    def measure_int32_peak(self) -> Tuple[float, float]:
        a, b = C.c_double(), C.c_double()
        check(lib().treat_b200_measure_int32_peak(self._ctx, C.byref(a), C.byref(b)), self._ctx)
        return a.value, b.value

    # ---- Go-shaped API -----------------------------------------------------------------------------
    def NewAlignments(self, reads, tmpl: Template, excludeSnps: bool = False, want_ops: bool = True) -> List[Alignment]:
        """The reference's per-read loop (cmd/treat/align.go:114-120, storage.go:738-785) as one batch.
        `reads` is a FastaBatch or an iterable of (id, seq)."""
        batch = reads if isinstance(reads, FastaBatch) else FastaBatch.from_records(reads)
        self._ensure_template(tmpl)
        flags = (EXCLUDE_SNPS if excludeSnps else 0) | (WANT_OPS | OPS_GAPPED_ONLY if want_ops else 0)
        rec, ops = self.align_batch(batch.seqs, batch.offsets, batch.read_counts, flags)
        return [Alignment(rec[i], None if ops is None else ops[i], batch.seq(i), self) for i in range(len(batch))]

    def NewAlignment(self, frag: Fragment, tmpl: Template, excludeSnps: bool = False) -> Alignment:
        """align.go:237-279 for one fragment (a batch of one)."""
        batch = FastaBatch.from_records([(frag.Name, frag.forward_seq())])
        batch.read_counts[0] = frag.ReadCount
        return self.NewAlignments(batch, tmpl, excludeSnps)[0]

    def SimpleAlign(self, f1: Fragment, f2: Fragment) -> Tuple[str, str]:
        """align.go:337-386"""
        s1, s2 = f1.forward_seq(), f2.forward_seq()
        cap = 2 * (len(s1) + len(s2)) + 64
        o1, o2 = C.create_string_buffer(cap), C.create_string_buffer(cap)
        n = lib().treat_b200_simple_align(self._ctx, s1, len(s1), s2, len(s2), ord(_b(f1.EditBase)), o1, o2, cap)
        self._tmpl = None  # SimpleAlign installs f1 as the context's template
        if n < 0:
            raise TreatError(int(n), lib().treat_b200_last_error(self._ctx).decode())
        return o1.raw[: int(n)].decode("latin-1"), o2.raw[: int(n)].decode("latin-1")

    def cli_align(self, template: Optional[str] = None, fragment: Optional[str] = None, s1: Optional[str] = None,
                  s2: Optional[str] = None, base: str = "T", offset: int = 0) -> str:
        """`treat align` stdout (cmd/treat/align.go:59-122)."""
        out, n = C.c_void_p(), C.c_size_t()
        err = C.create_string_buffer(512)
        rc = lib().treat_b200_cli_align(self._ctx, _b(template) if template else None, _b(fragment) if fragment else None,
                                        _b(s1) if s1 else None, _b(s2) if s2 else None, ord(_b(base)), offset,
                                        C.byref(out), C.byref(n), err, 512)
        self._tmpl = None
        if rc != cabi.OK:
            raise ValueError(err.value.decode() or lib().treat_b200_last_error(self._ctx).decode())
        try:
            return C.string_at(out.value, n.value).decode("latin-1")
        finally:
            lib().treat_b200_free(out)


class DeviceGroup:
    """Several GPUs of one process behind one handle (treat_b200_group_*): one context and one host worker thread per
    device, the batch cut into contiguous read ranges, the summary counters summed by ncclAllReduce inside the library."""

    def __init__(self, devices: Optional[Sequence[int]] = None, n: Optional[int] = None):
        devs = list(devices) if devices is not None else list(range(n or 1))
        arr = (C.c_int * len(devs))(*devs)
        self._g = C.c_void_p()
        check(lib().treat_b200_group_create(arr, len(devs), C.byref(self._g)))
        self.devices = devs

    def _check(self, code: int) -> None:
        if code != cabi.OK:
            raise TreatError(code, lib().treat_b200_group_last_error(self._g).decode())

    def close(self):
        if getattr(self, "_g", None):
            lib().treat_b200_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().treat_b200_group_size(self._g)

    def SetTemplate(self, tmpl: Template) -> None:
        self._check(lib().treat_b200_group_use_template(self._g, tmpl._h))

    def ops_stride(self, max_len: int) -> int:
        return lib().treat_b200_group_ops_stride(self._g, max_len)

    def align_batch(self, seqs: np.ndarray, seq_off: np.ndarray, read_count: Optional[np.ndarray] = None,
                    flags: int = WANT_OPS, rec: Optional[np.ndarray] = None, ops: Optional[np.ndarray] = None):
        """treat_b200_group_align_batch: same arguments and results as Aligner.align_batch, reads shared between the GPUs."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        N = len(seq_off) - 1
        if rec is None:
            rec = np.zeros(N, dtype=RECORD_DTYPE)
        stride = 0
        if flags & WANT_OPS:
            if ops is None:
                max_len = int((seq_off[1:] - seq_off[:-1]).max()) if N else 0
                stride = self.ops_stride(max_len)
                ops = np.zeros((N, stride), dtype=np.uint8)
            else:
                stride = ops.shape[1]
        rc = None if read_count is None else np.ascontiguousarray(read_count, dtype=np.uint32)
        dummy = np.zeros(1, np.uint8)
        self._check(lib().treat_b200_group_align_batch(
            self._g, seqs.ctypes.data if seqs.size else dummy.ctypes.data, seq_off.ctypes.data,
            rc.ctypes.data if rc is not None else None, N, flags, rec.ctypes.data,
            ops.ctypes.data if (flags & WANT_OPS) else None, stride))
        return rec, (ops if flags & WANT_OPS else None)

    def summary(self, from_rank: int = 0) -> np.ndarray:
        """All-GPU summary counters (one grouped ncclAllReduce), read from device `from_rank`."""
        n = lib().treat_b200_summary_len(lib().treat_b200_group_ctx(self._g, 0))
        out = np.zeros(n, dtype=np.uint64)
        self._check(lib().treat_b200_group_summary(self._g, from_rank, out.ctypes.data, n))
        return out

    def heatmap(self, from_rank: int = 0) -> np.ndarray:
        d = lib().treat_b200_heatmap_dim(lib().treat_b200_group_ctx(self._g, 0))
        out = np.zeros((d, d), dtype=np.uint64)
        self._check(lib().treat_b200_group_heatmap(self._g, from_rank, out.ctypes.data, d * d))
        return out

    def reset_summary(self) -> None:
        self._check(lib().treat_b200_group_reset_summary(self._g))

    def launch_count(self) -> int:
        return lib().treat_b200_group_launch_count(self._g)


class Comm:
    """One process per GPU: the library's own NCCL communicator over an Aligner's context (treat_b200_comm_*).  Rank 0
    makes the id (Comm.unique_id()) and the host program hands it to the other ranks."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * cabi.COMM_ID_BYTES)()
        check(lib().treat_b200_comm_unique_id(buf, cabi.COMM_ID_BYTES))
        return bytes(buf)

    def __init__(self, engine: Aligner, nranks: int, rank: int, uid: Optional[bytes] = None):
        self._c = C.c_void_p()
        self.engine, self.nranks, self.rank = engine, nranks, rank
        buf = (C.c_uint8 * cabi.COMM_ID_BYTES).from_buffer_copy(uid) if uid else None
        check(lib().treat_b200_comm_create(engine._ctx, nranks, rank, buf, cabi.COMM_ID_BYTES if uid else 0, C.byref(self._c)))

    def _check(self, code: int) -> None:
        if code != cabi.OK:
            raise TreatError(code, lib().treat_b200_comm_last_error(self._c).decode())

    def close(self):
        if getattr(self, "_c", None):
            lib().treat_b200_comm_destroy(self._c)
            self._c = None

    def reduce_summary(self) -> np.ndarray:
        n = lib().treat_b200_summary_len(self.engine._ctx)
        out = np.zeros(n, dtype=np.uint64)
        self._check(lib().treat_b200_comm_reduce_summary(self._c, out.ctypes.data, n))
        return out

    def reduce_heatmap(self) -> np.ndarray:
        d = lib().treat_b200_heatmap_dim(self.engine._ctx)
        out = np.zeros((d, d), dtype=np.uint64)
        self._check(lib().treat_b200_comm_reduce_heatmap(self._c, out.ctypes.data, d * d))
        return out

    def reduce_summary_device(self, d_out: int, stream: int = 0) -> None:
        """Snapshot on `stream`, ncclAllReduce on the communicator's stream (overlaps the caller's next step); see join()."""
        self._check(lib().treat_b200_comm_reduce_summary_device(self._c, d_out, stream or None))

    def join(self, stream: int = 0) -> None:
        """Work queued on `stream` after this call sees the results of every reduce_summary_device issued so far."""
        self._check(lib().treat_b200_comm_join(self._c, stream or None))
