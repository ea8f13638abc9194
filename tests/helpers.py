"""Shared test helpers (test infrastructure: may use the oracle)."""
import numpy as np

SUMMARY_HEADER = 16


def summary_from_records(rec, m, edit_offset, tmpl_edit_stop, n_bases=None):
    """Reference definition of the summary counters (cmd/treat/stats.go:140-187,
    cmd/treat/handlers.go:203-215) computed with numpy from per-read records."""
    bins = m + 1 + 2
    out = np.zeros(SUMMARY_HEADER + 3 * bins, dtype=np.uint64)
    rc = rec["read_count"].astype(np.uint64)
    one = np.ones(len(rec), dtype=np.uint64)
    mut = rec["has_mutation"] != 0
    indel = rec["indel"] == 1
    mm = rec["mismatches"]
    cls = {1: ~mut, 2: mut, 3: indel, 5: ~indel & (mm == 1), 6: ~indel & (mm == 2), 4: ~indel & (mm > 2)}
    for base, w in ((0, rc), (7, one)):
        out[base + 0] = w.sum()
        for k, mask in cls.items():
            out[base + k] = w[mask].sum()
    nb = rec["n_bases"] if n_bases is None else n_bases
    out[14] = len(rec)
    out[15] = (nb.astype(np.uint64) * np.uint64(m)).sum()
    keep = ~((rec["edit_stop"] == tmpl_edit_stop) & (rec["junc_len"] == 0))
    es = rec["edit_stop"].astype(np.int64) - edit_offset + 2
    jl = rec["junc_len"].astype(np.int64)
    je = rec["junc_end"].astype(np.int64) - edit_offset + 2
    for idx, vals in enumerate((es, jl, je)):
        v = vals[keep]
        w = rc[keep]
        ok = (v >= 0) & (v < bins)
        np.add.at(out, SUMMARY_HEADER + idx * bins + v[ok], w[ok])
    return out


def compare_records(got, want, seqs=None, seq_off=None):
    """Field-by-field comparison of treat_b200 records with oracle records; returns list of mismatches."""
    bad = []
    for f in ("edit_stop", "junc_start", "junc_end", "junc_len", "read_count", "has_mutation", "alt_editing",
              "mismatches", "indel", "score", "junc_seq_len", "aln_len", "n_bases"):
        d = np.nonzero(got[f] != want[f])[0]
        if len(d):
            bad.append((f, d[:5].tolist(), got[f][d[:5]].tolist(), want[f][d[:5]].tolist()))
    has = want["junc_seq_len"] > 0
    d = np.nonzero(has & (got["junc_seq_off"] != want["junc_seq_off"]))[0]
    if len(d):
        bad.append(("junc_seq_off", d[:5].tolist(), got["junc_seq_off"][d[:5]].tolist(), want["junc_seq_off"][d[:5]].tolist()))
    d = np.nonzero((got["status"] & 1) != (want["status"] & 1))[0]
    if len(d):
        bad.append(("status.panic", d[:5].tolist()))
    return bad


def heatmap_from_records(rec: np.ndarray, len_: int, offset: int) -> np.ndarray:
    """cmd/treat/handlers.go:578-589 with ReadCount for Norm: heat[EditStop - offset][JuncLen] += ReadCount, EditStop >= offset."""
    heat = np.zeros((len_, len_), dtype=np.uint64)
    es = rec["edit_stop"].astype(np.int64) - offset
    jl = rec["junc_len"].astype(np.int64)
    ok = (es >= 0) & (es < len_) & (jl < len_)
    np.add.at(heat, (es[ok], jl[ok]), rec["read_count"][ok].astype(np.uint64))
    return heat
