"""treat_b200 -- B200 (sm_100a) implementation of ubccr/treat's read-vs-template alignment path.

Host API mirrors the reference's Go package `treat`; all alignment work runs in
libtreat_b200.so (hand-written CUDA) behind the C ABI of include/treat_b200.h.
"""
from .cabi import (EXCLUDE_SNPS, FORWARD, NO_SUMMARY, OP_DEL, OP_INS, OP_MATCH, RECORD_DTYPE, REVERSE, ST_REF_PANIC,
                   ST_SATURATED, ST_WIDE_BASE, SUMMARY_HEADER, WANT_OPS, TreatError)
from .api import (Aligner, Alignment, Comm, DeviceGroup, FastaBatch, Fragment, NewFragment, NewTemplate, NewTemplateFromFasta,
                  PrintAlignment, Template, fasta_split, marshal_batch, mutant_report, norm_default, norm_scale, normalize_records, search_records, search_text, hist_norm,
                  parseMergeCount, parse_fasta,
                  read_fasta, stats_text, UnmarshalBinary)

__version__ = "0.1.0"
