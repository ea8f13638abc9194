"""Synthetic RPS12-shaped / long-gene read sets -- bench and test tooling.

The generator itself is tools/synth_reads.c (host C, deterministic, shardable); this module
builds it on demand and exposes numpy-friendly wrappers.  The RPS12 templates are the
reference's examples/templates.fa as frozen in tests/golden/fixtures.json.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SRC = os.path.join(_ROOT, "tools", "synth_reads.c")
_LIB = os.path.join(_ROOT, "tools", "_build", "libsynth.so")
SEED_BASE = 0x5EED0000

_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-std=c11", "-fPIC", "-shared", "-o", _LIB, _SRC])
    return _LIB


def _l():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.synth_max_len.restype = C.c_uint64
        L.synth_max_len.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.synth_reads.restype = C.c_uint64
        L.synth_reads.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint8,
                                  C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_void_p, C.c_uint64, C.c_void_p,
                                  C.c_void_p]
        L.synth_reads_noisy.restype = C.c_uint64
        L.synth_reads_noisy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint8,
                                        C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p]
        L.synth_long_template.restype = None
        L.synth_long_template.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class TemplateArrays:
    """Plain arrays of a template: what treat_b200_set_template takes."""

    def __init__(self, bases: bytes, edit_site: np.ndarray, alt_start, alt_end, edit_base: str = "T"):
        self.bases = bases
        self.edit_site = np.ascontiguousarray(edit_site, dtype=np.uint32)  # [K][m+1]
        self.alt_start = np.ascontiguousarray(alt_start, dtype=np.int32)
        self.alt_end = np.ascontiguousarray(alt_end, dtype=np.int32)
        self.edit_base = edit_base
        self.m = len(bases)
        self.K = self.edit_site.shape[0]

    def raw_sequences(self):
        """Re-inflated template sequences (FE, PE, alts) as FASTA-ready strings."""
        out = []
        eb = self.edit_base
        for k in range(self.K):
            s = []
            for i in range(self.m + 1):
                s.append(eb * int(self.edit_site[k, i]))
                if i < self.m:
                    s.append(chr(self.bases[i]))
            out.append("".join(s))
        return out

    def write_fasta(self, path: str) -> str:
        names = ["Fully Edited", "Pre-Edited"] + [
            "Alt %d alt_start=%d alt_stop=%d" % (i + 1, self.alt_start[i], self.alt_end[i]) for i in range(self.K - 2)]
        with open(path, "w") as fh:
            for nme, s in zip(names, self.raw_sequences()):
                fh.write(">%s\n%s\n" % (nme, s))
        return path


def _strip(seq: str, eb: str):
    bases, sites, run = [], [], 0
    for ch in seq.upper():
        if ch == eb:
            run += 1
        else:
            sites.append(run)
            run = 0
            bases.append(ch)
    sites.append(run)
    return "".join(bases), sites


def rps12_template() -> TemplateArrays:
    """examples/templates.fa (FE, PE, A1..A3) from the frozen fixtures."""
    import re
    with open(os.path.join(_ROOT, "tests", "golden", "fixtures.json")) as fh:
        text = json.load(fh)["fasta"]["templates.fa"]
    recs, cur = [], None
    for line in text.split("\n"):
        if line.startswith(">"):
            cur = [line[1:].strip(), []]
            recs.append(cur)
        elif cur is not None:
            cur[1].append(line.strip())
    rows, st, en, bases = [], [], [], None
    for i, (rid, parts) in enumerate(recs):
        b, s = _strip("".join(parts), "T")
        bases = bases or b
        assert b == bases
        rows.append(s)
        if i >= 2:
            st.append(int(re.search(r"alt_start=(\d+)", rid).group(1)))
            en.append(int(re.search(r"alt_stop=(\d+)", rid).group(1)))
    return TemplateArrays(bases.encode(), np.array(rows, dtype=np.uint32), st, en, "T")


def long_template(m: int = 300, n_alt: int = 8, seed: int = SEED_BASE + 5) -> TemplateArrays:
    bases = np.zeros(m, dtype=np.uint8)
    es = np.zeros((2 + n_alt, m + 1), dtype=np.uint32)
    st = np.zeros(n_alt, dtype=np.int32)
    en = np.zeros(n_alt, dtype=np.int32)
    _l().synth_long_template(m, n_alt, seed, bases.ctypes.data, es.ctypes.data, st.ctypes.data, en.ctypes.data)
    return TemplateArrays(bases.tobytes(), es, st, en, "T")


def reads(t: TemplateArrays, N: int, seed: int = SEED_BASE + 3, first: int = 0, p_trunc: float = 0.0,
          pinned_alloc=None, p_sub_base: float = 0.0, p_indel_read: float = 0.0):
    """Generate reads [first, first+N): returns (seqs uint8, seq_off uint64[N+1], read_count uint32[N]).
    p_sub_base / p_indel_read: sequencing noise on top of the model (per-base substitutions, one more indel per read)."""
    L = _l()
    mx = int(L.synth_max_len(t.m, t.edit_site.ctypes.data, t.K)) + 2
    cap = mx * max(N, 1)
    seqs = pinned_alloc(cap) if pinned_alloc else np.empty(cap, dtype=np.uint8)
    off = np.zeros(N + 1, dtype=np.uint64)
    rc = np.zeros(N, dtype=np.uint32)
    used = L.synth_reads_noisy(t.bases, t.m, t.edit_site.ctypes.data, t.K, t.alt_start.ctypes.data, t.alt_end.ctypes.data,
                               ord(t.edit_base), seed, first, N, p_trunc, p_sub_base, p_indel_read, seqs.ctypes.data, cap,
                               off.ctypes.data, rc.ctypes.data)
    assert used > 0 or N == 0
    return seqs[:used], off, rc
