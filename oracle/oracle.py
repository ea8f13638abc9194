"""ctypes binding of oracle/_build/liboracle.so (the C restatement in treat_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/treat_oracle.h.  Function names follow the
reference's Go API (fragment.go, template.go, align.go) so tests read like the
reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

FORWARD = 1
REVERSE = -1
TIE_ORDERS = ["ULD", "UDL", "LUD", "LDU", "DUL", "DLU"]


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("treat_oracle.c", "treat_oracle.h", "oracle_cli.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Fragment(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("read_count", C.c_uint32),
        ("norm", C.c_double),
        ("bases", C.POINTER(C.c_char)),
        ("n", C.c_size_t),
        ("edit_base", C.c_char),
        ("edit_site", C.POINTER(C.c_uint32)),
    ]


class _Template(C.Structure):
    _fields_ = [
        ("bases", C.POINTER(C.c_char)),
        ("m", C.c_size_t),
        ("edit_offset", C.c_uint32),
        ("edit_stop", C.c_int),
        ("edit_base", C.c_char),
        ("K", C.c_size_t),
        ("edit_site", C.POINTER(C.POINTER(C.c_uint32))),
        ("base_index", C.POINTER(C.c_uint32)),
        ("n_alt", C.c_size_t),
        ("alt_start", C.POINTER(C.c_int)),
        ("alt_end", C.POINTER(C.c_int)),
    ]


class _Alignment(C.Structure):
    _fields_ = [
        ("edit_stop", C.c_int),
        ("junc_start", C.c_int),
        ("junc_end", C.c_int),
        ("junc_len", C.c_int),
        ("read_count", C.c_uint32),
        ("norm", C.c_double),
        ("has_mutation", C.c_uint8),
        ("mismatches", C.c_uint8),
        ("indel", C.c_uint8),
        ("alt_editing", C.c_uint8),
        ("junc_seq", C.POINTER(C.c_char)),
        ("junc_seq_len", C.c_size_t),
        ("would_panic", C.c_int),
        ("score", C.c_int),
    ]


RECORD_DTYPE = np.dtype(
    [
        ("edit_stop", "<i4"),
        ("junc_start", "<i4"),
        ("junc_end", "<i4"),
        ("junc_len", "<i4"),
        ("read_count", "<u4"),
        ("has_mutation", "u1"),
        ("alt_editing", "u1"),
        ("mismatches", "u1"),
        ("indel", "u1"),
        ("score", "<i4"),
        ("junc_seq_off", "<u4"),
        ("junc_seq_len", "<u4"),
        ("aln_len", "<u2"),
        ("n_bases", "<u2"),
        ("status", "u1"),
        ("pad", "u1", (3,)),
    ]
)
assert RECORD_DTYPE.itemsize == 44

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_parse_merge_count.restype = C.c_uint32
        L.orc_parse_merge_count.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_new_fragment.restype = C.POINTER(_Fragment)
        L.orc_new_fragment.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_char]
        L.orc_fragment_string.restype = C.c_void_p
        L.orc_fragment_string.argtypes = [C.POINTER(_Fragment)]
        L.orc_fragment_free.argtypes = [C.POINTER(_Fragment)]
        L.orc_new_template.restype = C.POINTER(_Template)
        L.orc_new_template.argtypes = [
            C.POINTER(_Fragment), C.POINTER(_Fragment), C.POINTER(C.POINTER(_Fragment)), C.c_size_t,
            C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_size_t, C.c_char_p, C.c_size_t,
        ]
        L.orc_template_from_fasta.restype = C.POINTER(_Template)
        L.orc_template_from_fasta.argtypes = [C.c_char_p, C.c_int, C.c_char, C.c_char_p, C.c_size_t]
        L.orc_template_set_offset.argtypes = [C.POINTER(_Template), C.c_int]
        L.orc_template_max.restype = C.c_uint32
        L.orc_template_max.argtypes = [C.POINTER(_Template), C.c_size_t]
        L.orc_template_free.argtypes = [C.POINTER(_Template)]
        L.orc_nw_align.restype = C.c_size_t
        L.orc_nw_align.argtypes = [
            C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
        ]
        L.orc_new_alignment.restype = C.POINTER(_Alignment)
        L.orc_new_alignment.argtypes = [C.POINTER(_Fragment), C.POINTER(_Template), C.c_int]
        L.orc_alignment_free.argtypes = [C.POINTER(_Alignment)]
        L.orc_write_to.restype = C.c_void_p
        L.orc_write_to.argtypes = [C.POINTER(_Alignment), C.POINTER(_Fragment), C.POINTER(_Template), C.c_int,
                                   C.POINTER(C.c_size_t)]
        L.orc_simple_align.argtypes = [C.POINTER(_Fragment), C.POINTER(_Fragment), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p)]
        L.orc_marshal_binary.restype = C.c_void_p
        L.orc_marshal_binary.argtypes = [C.POINTER(_Alignment), C.POINTER(C.c_size_t)]
        L.orc_cli_align.restype = C.c_void_p
        L.orc_cli_align.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char, C.c_int,
                                    C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        L.orc_align_batch.restype = C.c_uint64
        L.orc_align_batch.argtypes = [C.POINTER(_Template), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_text_batch.restype = C.c_void_p
        L.orc_text_batch.argtypes = [C.POINTER(_Template), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p, C.c_int,
                                     C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_set_tie_order.argtypes = [C.c_int]
        L.orc_get_tie_order.restype = C.c_int
        _lib = L
    return _lib


def _b(s) -> bytes:
    return s if isinstance(s, bytes) else s.encode("latin-1")


def _take(ptr, n=None) -> bytes:
    s = C.string_at(ptr) if n is None else C.string_at(ptr, n)
    lib().orc_free(ptr)
    return s


def set_tie_order(name: str) -> None:
    lib().orc_set_tie_order(TIE_ORDERS.index(name))


def parseMergeCount(rec_id) -> int:
    b = _b(rec_id)
    return lib().orc_parse_merge_count(b, len(b))


class Fragment:
    """fragment.go:57-64"""

    def __init__(self, ptr):
        self._p = ptr

    def __del__(self):
        try:
            if self._p:
                lib().orc_fragment_free(self._p)
                self._p = None
        except Exception:  # interpreter shutdown
            pass

    Name = property(lambda s: s._p.contents.name.decode("latin-1"))
    ReadCount = property(lambda s: s._p.contents.read_count)
    Norm = property(lambda s: s._p.contents.norm)
    Bases = property(lambda s: C.string_at(s._p.contents.bases, s._p.contents.n).decode("latin-1"))
    EditBase = property(lambda s: s._p.contents.edit_base.decode("latin-1"))
    EditSite = property(lambda s: [s._p.contents.edit_site[i] for i in range(s._p.contents.n + 1)])

    def Len(self):
        return self._p.contents.n + 1

    def String(self):
        return _take(lib().orc_fragment_string(self._p)).decode("latin-1")


def NewFragment(name, seq, orientation=FORWARD, base="T") -> Fragment:
    s = _b(seq)
    return Fragment(lib().orc_new_fragment(_b(name), s, len(s), orientation, _b(base)))


class Template:
    """template.go:41-49"""

    def __init__(self, ptr):
        self._p = ptr

    def __del__(self):
        try:
            if self._p:
                lib().orc_template_free(self._p)
                self._p = None
        except Exception:  # interpreter shutdown
            pass

    Bases = property(lambda s: C.string_at(s._p.contents.bases, s._p.contents.m).decode("latin-1"))
    EditOffset = property(lambda s: s._p.contents.edit_offset)
    EditStop = property(lambda s: s._p.contents.edit_stop)
    EditBase = property(lambda s: s._p.contents.edit_base.decode("latin-1"))
    EditSite = property(
        lambda s: [[s._p.contents.edit_site[k][i] for i in range(s._p.contents.m + 1)] for k in range(s._p.contents.K)]
    )
    BaseIndex = property(lambda s: [s._p.contents.base_index[i] for i in range(s._p.contents.m + 1)])
    AltRegion = property(
        lambda s: [(s._p.contents.alt_start[i], s._p.contents.alt_end[i]) for i in range(s._p.contents.n_alt)]
    )

    def Size(self):
        return self._p.contents.K

    def Len(self):
        return self._p.contents.m + 1

    def Max(self, i):
        return lib().orc_template_max(self._p, i)

    def SetOffset(self, offset):
        lib().orc_template_set_offset(self._p, offset)


def NewTemplateFromFasta(path, orientation=FORWARD, base="T") -> Template:
    err = C.create_string_buffer(512)
    p = lib().orc_template_from_fasta(_b(path), orientation, _b(base), err, 512)
    if not p:
        raise ValueError(err.value.decode())
    return Template(p)


def NewTemplate(full: Fragment, pre: Fragment, alt=None, altRegion=None) -> Template:
    alt = alt or []
    altRegion = altRegion or []
    arr = (C.POINTER(_Fragment) * max(1, len(alt)))(*[a._p for a in alt])
    st = (C.c_int * max(1, len(altRegion)))(*[r[0] if r else 0 for r in altRegion])
    en = (C.c_int * max(1, len(altRegion)))(*[r[1] if r else 0 for r in altRegion])
    err = C.create_string_buffer(512)
    p = lib().orc_new_template(full._p, pre._p, arr, len(alt), st, en, len(altRegion), err, 512)
    if not p:
        raise ValueError(err.value.decode())
    return Template(p)


def nw_align(a, b, match=1, mismatch=-1, gap=-1):
    """nwalgo.Align"""
    a, b = _b(a), _b(b)
    p1, p2, sc = C.c_void_p(), C.c_void_p(), C.c_int()
    n = lib().orc_nw_align(a, len(a), b, len(b), match, mismatch, gap, C.byref(p1), C.byref(p2), C.byref(sc))
    return _take(p1.value, n).decode("latin-1"), _take(p2.value, n).decode("latin-1"), sc.value


class Alignment:
    """align.go:41-55"""

    def __init__(self, ptr):
        self._p = ptr

    def __del__(self):
        try:
            if self._p:
                lib().orc_alignment_free(self._p)
                self._p = None
        except Exception:  # interpreter shutdown
            pass

    EditStop = property(lambda s: s._p.contents.edit_stop)
    JuncStart = property(lambda s: s._p.contents.junc_start)
    JuncEnd = property(lambda s: s._p.contents.junc_end)
    JuncLen = property(lambda s: s._p.contents.junc_len)
    ReadCount = property(lambda s: s._p.contents.read_count)
    Norm = property(lambda s: s._p.contents.norm)
    HasMutation = property(lambda s: s._p.contents.has_mutation)
    Mismatches = property(lambda s: s._p.contents.mismatches)
    Indel = property(lambda s: s._p.contents.indel)
    AltEditing = property(lambda s: s._p.contents.alt_editing)
    JuncSeq = property(lambda s: C.string_at(s._p.contents.junc_seq, s._p.contents.junc_seq_len).decode("latin-1"))
    WouldPanic = property(lambda s: bool(s._p.contents.would_panic))
    Score = property(lambda s: s._p.contents.score)

    def WriteTo(self, frag: Fragment, tmpl: Template, tw: int = 80) -> str:
        n = C.c_size_t()
        return _take(lib().orc_write_to(self._p, frag._p, tmpl._p, tw, C.byref(n)), n.value).decode("latin-1")

    def MarshalBinary(self) -> bytes:
        n = C.c_size_t()
        return _take(lib().orc_marshal_binary(self._p, C.byref(n)), n.value)

    def fields(self) -> dict:
        return dict(
            edit_stop=self.EditStop, junc_start=self.JuncStart, junc_end=self.JuncEnd, junc_len=self.JuncLen,
            read_count=self.ReadCount, has_mutation=self.HasMutation, mismatches=self.Mismatches,
            indel=self.Indel, alt_editing=self.AltEditing, junc_seq=self.JuncSeq, score=self.Score,
            would_panic=self.WouldPanic,
        )


def NewAlignment(frag: Fragment, tmpl: Template, excludeSnps: bool = False) -> Alignment:
    return Alignment(lib().orc_new_alignment(frag._p, tmpl._p, int(excludeSnps)))


def SimpleAlign(f1: Fragment, f2: Fragment):
    p1, p2 = C.c_void_p(), C.c_void_p()
    lib().orc_simple_align(f1._p, f2._p, C.byref(p1), C.byref(p2))
    return _take(p1.value).decode("latin-1"), _take(p2.value).decode("latin-1")


def cli_align(template=None, fragment=None, s1=None, s2=None, base="T", offset=0) -> str:
    """`treat align` stdout (cmd/treat/align.go:59-122)."""
    n = C.c_size_t()
    err = C.create_string_buffer(512)
    p = lib().orc_cli_align(
        _b(template) if template else None, _b(fragment) if fragment else None,
        _b(s1) if s1 else None, _b(s2) if s2 else None, _b(base), offset, C.byref(n), err, 512,
    )
    if not p:
        raise ValueError(err.value.decode())
    return _take(p, n.value).decode("latin-1")


def ops_stride_for(m: int, nmax: int) -> int:
    return ((m + nmax + 3) // 4 + 15) // 16 * 16


def align_batch(tmpl: Template, seqs: np.ndarray, seq_off: np.ndarray, read_count=None, exclude_snps=False,
                mode=0, nthreads=1, want_ops=True, ops_stride=None):
    """Batch driver: returns (records[N] of RECORD_DTYPE, ops[N, stride] uint8 or None, text_bytes)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    N = len(seq_off) - 1
    rec = np.zeros(N, dtype=RECORD_DTYPE)
    ops = None
    if want_ops:
        if ops_stride is None:
            lens = np.diff(seq_off.astype(np.int64))
            ops_stride = ops_stride_for(tmpl.Len() - 1, int(lens.max()) if N else 0)
        ops = np.zeros((N, ops_stride), dtype=np.uint8)
    rc = None
    if read_count is not None:
        rc = np.ascontiguousarray(read_count, dtype=np.uint32)
    tb = lib().orc_align_batch(
        tmpl._p, seqs.ctypes.data, seq_off.ctypes.data, rc.ctypes.data if rc is not None else None, N,
        int(exclude_snps), mode, nthreads, rec.ctypes.data, ops.ctypes.data if ops is not None else None,
        ops_stride or 0,
    )
    return rec, ops, tb


def text_batch(tmpl: Template, seqs: np.ndarray, seq_off: np.ndarray, read_count=None, prefix: str = "", nthreads: int = 1):
    """`treat align -t -f` stdout for reads named "<prefix><index>-<read_count>" (cmd/treat/align.go:113-121):
    returns (text uint8 array, text_off uint64[N+1])."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    N = len(seq_off) - 1
    rc = None if read_count is None else np.ascontiguousarray(read_count, dtype=np.uint32)
    toff = np.zeros(N + 1, dtype=np.uint64)
    total = C.c_uint64()
    p = lib().orc_text_batch(tmpl._p, seqs.ctypes.data, seq_off.ctypes.data, rc.ctypes.data if rc is not None else None, N,
                             prefix.encode(), nthreads, toff.ctypes.data, C.byref(total))
    try:
        text = np.frombuffer((C.c_uint8 * total.value).from_address(p), dtype=np.uint8).copy()
    finally:
        lib().orc_free(p)
    return text, toff
