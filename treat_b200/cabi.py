"""ctypes binding of libtreat_b200.so: one Python function per symbol of include/treat_b200.h.

This is the same binding a cgo caller would write; nothing here computes.
Loading fails loudly when the library is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TREAT_B200_LIB") or os.path.join(_PKG, "libtreat_b200.so")  # the override: a checked / experimental build

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_TOO_LONG, ERR_NO_TEMPLATE, ERR_NO_DEVICE, ERR_TEMPLATE, ERR_IO = range(-1, -9, -1)
FORWARD, REVERSE = 1, -1
EXCLUDE_SNPS, WANT_OPS, NO_SUMMARY, ONE_READ_PER_WARP, FULL_TRACEBACK_STORE = 0x1, 0x2, 0x4, 0x8, 0x10
DP_EVERY_READ, USE_PACK_MARKS, OPS_GAPPED_ONLY, HEATMAP, WANT_BLOBS, SKIP_DP = 0x20, 0x40, 0x80, 0x100, 0x200, 0x400
NO_BAND = 0x800
NO_FRONTIER = 0x1000
FRONTIER_ONLY = 0x2000
ST_REF_PANIC, ST_WIDE_BASE, ST_SATURATED = 0x1, 0x2, 0x4
OP_MATCH, OP_DEL, OP_INS = 0, 1, 2
SUMMARY_HEADER = 16
COMM_ID_BYTES = 128
TIE_ORDERS = ("ULD", "UDL", "LUD")

RECORD_DTYPE = np.dtype(
    [
        ("edit_stop", "<i4"), ("junc_start", "<i4"), ("junc_end", "<i4"), ("junc_len", "<i4"),
        ("read_count", "<u4"),
        ("has_mutation", "u1"), ("alt_editing", "u1"), ("mismatches", "u1"), ("indel", "u1"),
        ("score", "<i4"), ("junc_seq_off", "<u4"), ("junc_seq_len", "<u4"),
        ("aln_len", "<u2"), ("n_bases", "<u2"),
        ("status", "u1"), ("reserved", "u1", (3,)),
        ("raw_len", "<u4"),
    ]
)
assert RECORD_DTYPE.itemsize == 48


class Search(C.Structure):
    """treat_b200_search: the fields of SearchFields that HasMatch / Storage.Search read (cmd/treat/storage.go:48-64, 162-188, 257-286)."""
    _fields_ = [("edit_stop", C.c_int32), ("junc_end", C.c_int32), ("junc_len", C.c_int32), ("offset", C.c_int32), ("limit", C.c_int32),
                ("alt_region", C.c_int32), ("has_mutation", C.c_uint8), ("has_alt", C.c_uint8), ("all", C.c_uint8), ("reserved", C.c_uint8)]


class Packed(C.Structure):
    """treat_b200_packed"""
    _fields_ = [
        ("N", C.c_uint64), ("max_len", C.c_uint32), ("wide", C.c_uint32),
        ("bases_stride", C.c_uint32), ("counts_stride", C.c_uint32),
        ("d_n", C.c_void_p), ("d_flags", C.c_void_p), ("d_bases", C.c_void_p), ("d_counts", C.c_void_p),
        ("d_raw_len", C.c_void_p),
    ]


class AlignGeometry(C.Structure):
    """treat_b200_geometry"""
    _fields_ = [(n, C.c_uint32) for n in ("cpl", "strips", "reads_per_warp", "warps_per_cta", "smem_cta_bytes", "smem_template_bytes",
                                          "smem_warp_bytes", "band_rows", "band_diagonals", "scratch_bytes_per_warp", "regs_per_thread",
                                          "occupancy_warps_per_sm")]


class FastaPiece(C.Structure):
    """treat_b200_fasta_piece"""
    _fields_ = [
        ("first_record", C.c_uint64), ("n", C.c_uint64), ("rec", C.c_void_p), ("id_off", C.c_void_p), ("seq_off", C.c_void_p),
        ("id_len", C.c_void_p), ("read_count", C.c_void_p), ("seq_len", C.c_void_p), ("ops", C.c_void_p), ("ops_index", C.c_void_p),
        ("n_ops_rows", C.c_uint64), ("ops_stride", C.c_uint32), ("seq_in_place", C.c_uint32),
        ("blob", C.c_void_p), ("blob_off", C.c_void_p),
    ]


FASTA_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FastaPiece))
TEXT_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64)

# name -> (restype, argtypes); must list every function include/treat_b200.h declares
_P, _U8P, _SZ = C.c_void_p, C.c_void_p, C.c_size_t
SIGNATURES = {
    "treat_b200_abi_version": (C.c_int, []),
    "treat_b200_strerror": (C.c_char_p, [C.c_int]),
    "treat_b200_last_error": (C.c_char_p, [_P]),
    "treat_b200_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "treat_b200_destroy": (None, [_P]),
    "treat_b200_host_alloc": (C.c_int, [C.POINTER(_P), _SZ]),
    "treat_b200_host_register": (C.c_int, [_P, _SZ]),
    "treat_b200_host_unregister": (C.c_int, [_P]),
    "treat_b200_host_free": (None, [_P]),
    "treat_b200_set_template": (C.c_int, [_P, _U8P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, C.c_uint32, C.c_int32, C.c_uint8]),
    "treat_b200_use_template": (C.c_int, [_P, _P]),
    "treat_b200_align_batch": (C.c_int, [_P, _U8P, _P, _P, C.c_uint64, C.c_uint32, _P, _U8P, C.c_uint32]),
    "treat_b200_align_batch_device": (C.c_int, [_P, _U8P, _P, _P, C.c_uint64, C.c_uint32, C.c_uint32, _P, _U8P, C.c_uint32, _P]),
    "treat_b200_collapse_align_batch": (C.c_int, [_P, _U8P, _P, _P, C.c_uint64, C.c_uint32, _P, _U8P, C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    "treat_b200_ops_stride": (C.c_uint32, [_P, C.c_uint32]),
    "treat_b200_synchronize": (C.c_int, [_P]),
    "treat_b200_heatmap_dim": (C.c_uint32, [_P]),
    "treat_b200_get_heatmap": (C.c_int, [_P, _P, C.c_uint64]),
    "treat_b200_heatmap_device_ptr": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "treat_b200_fasta_parse": (C.c_int, [_U8P, C.c_uint64, C.POINTER(C.c_void_p)]),
    "treat_b200_fasta_split": (C.c_int, [_U8P, C.c_uint64, C.c_uint32, _P]),
    "treat_b200_fasta_upload": (C.c_int, [_P, _U8P, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "treat_b200_fasta_align": (C.c_int, [_P, C.c_uint32, _P, _U8P, C.c_uint32, _P, _P, _P, _P, _P]),
    "treat_b200_fasta_stream": (C.c_int, [_P, _U8P, C.c_uint64, C.c_uint32, C.c_uint32, _P, _P, C.POINTER(C.c_uint64)]),
    "treat_b200_text_device": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int32, _P, _P, C.c_uint64,
                                        C.POINTER(C.c_uint64), _P]),
    "treat_b200_text_batch": (C.c_int, [_P, _U8P, _P, _U8P, _P, C.c_uint64, C.c_uint32, C.c_int32, _P, _P]),
    "treat_b200_fasta_text": (C.c_int, [_P, _U8P, C.c_uint64, C.c_uint32, C.c_int32, _P, _P, C.POINTER(C.c_uint64)]),
    "treat_b200_marshal_device": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64), _P]),
    "treat_b200_unmarshal_record": (C.c_int, [_P, _SZ, _P, C.POINTER(C.c_double), C.POINTER(_P)]),
    "treat_b200_search_records": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "treat_b200_hist_norm": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "treat_b200_search_text": (C.c_int, [C.c_char_p, C.c_char_p, _P, _P, _U8P, _P, _P, C.c_uint64, C.c_int32, C.c_int32, C.POINTER(_P), C.POINTER(_SZ)]),
    "treat_b200_search_device": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64), _P]),
    "treat_b200_norm_default": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double)]),
    "treat_b200_norm_scale": (C.c_int, [C.c_uint64, C.c_double, C.POINTER(C.c_double)]),
    "treat_b200_normalize_records": (C.c_int, [_P, C.c_uint64, C.c_double, _P, _P, _P]),
    "treat_b200_normalize_device": (C.c_int, [_P, _P, C.c_uint64, C.c_double, _P, _P, _P, _P]),
    "treat_b200_stats_text": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(_P), C.c_int32, C.c_int32, C.c_int32,
                                       C.POINTER(_P), C.POINTER(_SZ)]),
    "treat_b200_packed_layout": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(Packed)]),
    "treat_b200_pack_device": (C.c_int, [_P, _U8P, _P, C.POINTER(Packed), _P]),
    "treat_b200_align_packed_device": (C.c_int, [_P, C.POINTER(Packed), _P, C.c_uint32, _P, _U8P, C.c_uint32, _P]),
    "treat_b200_launch_count": (C.c_uint64, [_P]),
    "treat_b200_set_band_rows": (C.c_int, [_P, C.c_uint32]),
    "treat_b200_set_band_min_reads": (C.c_int, [_P, C.c_uint32]),
    "treat_b200_set_tie_order": (C.c_int, [_P, C.c_int32]),
    "treat_b200_align_geometry": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.POINTER(AlignGeometry)]),
    "treat_b200_debug_flags": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "treat_b200_dp_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.c_int32]),
    "treat_b200_summary_bins": (C.c_uint32, [_P]),
    "treat_b200_summary_len": (C.c_uint64, [_P]),
    "treat_b200_get_summary": (C.c_int, [_P, _P, C.c_uint64]),
    "treat_b200_reset_summary": (C.c_int, [_P]),
    "treat_b200_summary_device_ptr": (C.c_int, [_P, C.POINTER(_P)]),
    "treat_b200_measure_int32_peak": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "treat_b200_measure_h2d": (C.c_int, [_P, _P, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "treat_b200_measure_transfer": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    "treat_b200_group_create": (C.c_int, [_P, C.c_int32, C.POINTER(_P)]),
    "treat_b200_group_destroy": (None, [_P]),
    "treat_b200_group_size": (C.c_int32, [_P]),
    "treat_b200_group_ctx": (_P, [_P, C.c_int32]),
    "treat_b200_group_last_error": (C.c_char_p, [_P]),
    "treat_b200_group_set_template": (C.c_int, [_P, _U8P, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32, C.c_uint32, C.c_int32, C.c_uint8]),
    "treat_b200_group_use_template": (C.c_int, [_P, _P]),
    "treat_b200_group_align_batch": (C.c_int, [_P, _U8P, _P, _P, C.c_uint64, C.c_uint32, _P, _U8P, C.c_uint32]),
    "treat_b200_group_ops_stride": (C.c_uint32, [_P, C.c_uint32]),
    "treat_b200_group_summary": (C.c_int, [_P, C.c_int32, _P, C.c_uint64]),
    "treat_b200_group_heatmap": (C.c_int, [_P, C.c_int32, _P, C.c_uint64]),
    "treat_b200_group_reset_summary": (C.c_int, [_P]),
    "treat_b200_group_launch_count": (C.c_uint64, [_P]),
    "treat_b200_comm_unique_id": (C.c_int, [_U8P, _SZ]),
    "treat_b200_comm_create": (C.c_int, [_P, C.c_int32, C.c_int32, _U8P, _SZ, C.POINTER(_P)]),
    "treat_b200_comm_destroy": (None, [_P]),
    "treat_b200_comm_reduce_summary": (C.c_int, [_P, _P, C.c_uint64]),
    "treat_b200_comm_reduce_heatmap": (C.c_int, [_P, _P, C.c_uint64]),
    "treat_b200_comm_reduce_summary_device": (C.c_int, [_P, _P, _P]),
    "treat_b200_comm_join": (C.c_int, [_P, _P]),
    "treat_b200_comm_last_error": (C.c_char_p, [_P]),
    "treat_b200_nccl_version": (C.c_int32, []),
    "treat_b200_parse_merge_count": (C.c_uint32, [C.c_char_p, _SZ]),
    "treat_b200_new_fragment": (C.c_int, [_U8P, _SZ, C.c_int, C.c_uint8, _U8P, _P, C.POINTER(_SZ)]),
    "treat_b200_template_from_fasta": (C.c_int, [C.c_char_p, C.c_int, C.c_uint8, C.c_int32, C.POINTER(_P), C.c_char_p, _SZ]),
    "treat_b200_template_new": (C.c_int, [C.POINTER(C.c_char_p), C.c_int32, _P, _P, C.c_int32, C.c_int, C.c_uint8, C.POINTER(_P), C.c_char_p, _SZ]),
    "treat_b200_template_set_offset": (None, [_P, C.c_int32]),
    "treat_b200_template_free": (None, [_P]),
    "treat_b200_template_size": (C.c_int32, [_P]),
    "treat_b200_template_len": (C.c_int32, [_P]),
    "treat_b200_template_edit_stop": (C.c_int32, [_P]),
    "treat_b200_template_edit_offset": (C.c_uint32, [_P]),
    "treat_b200_template_edit_base": (C.c_uint8, [_P]),
    "treat_b200_template_bases": (_P, [_P]),
    "treat_b200_template_edit_site": (_P, [_P, C.c_int32]),
    "treat_b200_template_base_index": (_P, [_P]),
    "treat_b200_template_n_alt": (C.c_int32, [_P]),
    "treat_b200_template_alt_start": (C.c_int32, [_P, C.c_int32]),
    "treat_b200_template_alt_end": (C.c_int32, [_P, C.c_int32]),
    "treat_b200_template_max": (C.c_uint32, [_P, C.c_int32]),
    "treat_b200_fasta_read": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "treat_b200_fasta_free": (None, [_P]),
    "treat_b200_fasta_count": (C.c_uint64, [_P]),
    "treat_b200_fasta_id": (C.c_char_p, [_P, C.c_uint64]),
    "treat_b200_fasta_seqs": (_P, [_P]),
    "treat_b200_fasta_offsets": (_P, [_P]),
    "treat_b200_fasta_read_counts": (_P, [_P]),
    "treat_b200_format_alignment": (C.c_int64, [_P, _P, _U8P, C.c_char_p, _U8P, _SZ, C.c_int32, _P, _SZ]),
    "treat_b200_simple_align": (C.c_int64, [_P, _U8P, _SZ, _U8P, _SZ, C.c_uint8, _P, _P, _SZ]),
    "treat_b200_marshal_record": (C.c_int64, [_P, _U8P, _SZ, _U8P, _SZ]),
    "treat_b200_marshal_batch": (C.c_int64, [_P, _U8P, _P, C.c_uint64, _U8P, C.c_uint64, _P, C.c_int32]),
    "treat_b200_mutant_report": (C.c_int, [_P, _P, _U8P, C.c_uint32, _U8P, _P, C.c_uint64, C.c_int32, C.POINTER(_P),
                                           C.POINTER(_SZ)]),
    "treat_b200_cli_align": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint8, C.c_int32,
                                       C.POINTER(_P), C.POINTER(_SZ), C.c_char_p, _SZ]),
    "treat_b200_free": (None, [_P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libtreat_b200.so (built in-tree by treat_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "treat_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class TreatError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        msg = lib().treat_b200_strerror(code).decode()
        super().__init__(f"treat_b200 error {code}: {msg}" + (f" ({detail})" if detail else ""))


def check(code: int, ctx=None):
    if code != OK:
        detail = lib().treat_b200_last_error(ctx).decode() if ctx else ""
        raise TreatError(code, detail)
