"""Full-size parity run (BASELINE configs[3]: 10 M synthetic RPS12 reads): every record and every traceback ops row of
the GPU path against the CPU oracle on ALL reads, and the `treat align` text (Alignment.WriteTo on the device, through
treat_b200_fasta_text) byte for byte on the first --text-reads of them.  Test infrastructure: uses oracle/.

    python tests/parity_full.py --reads 10000000 --text-reads 1000000 --out gpurun_out/parity_full.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # helpers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--text-reads", type=int, default=1_000_000)
    ap.add_argument("--piece", type=int, default=1_000_000)
    ap.add_argument("--workload", default="rps12", choices=["rps12", "long"])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import tempfile
    import treat_b200 as T
    from treat_b200 import cabi, synth
    from oracle import oracle as O
    from helpers import compare_records, summary_from_records

    threads = os.cpu_count() or 1
    t = synth.rps12_template() if a.workload == "rps12" else synth.long_template(300, 8)
    p_trunc = 0.0 if a.workload == "rps12" else 0.2
    with tempfile.TemporaryDirectory() as d:
        path = t.write_fasta(os.path.join(d, "t.fa"))
        tm = T.NewTemplateFromFasta(path, T.FORWARD, "T")
        om = O.NewTemplateFromFasta(path, O.FORWARD, "T")
    eng = T.Aligner(0)
    eng.SetTemplate(tm)
    res = {"workload": a.workload, "m": t.m, "K": t.K, "reads": a.reads, "host_threads": threads, "pieces": [],
           "record_mismatches": 0, "ops_rows_differing": 0, "gapped_reads": 0, "text_reads": 0, "text_bytes": 0,
           "text_bytes_differing": 0}
    t_gpu = t_cpu = 0.0
    text_left = a.text_reads
    first = 0
    summ_want = None
    while first < a.reads:
        n = min(a.piece, a.reads - first)
        seqs, off, rc = synth.reads(t, n, seed=synth.SEED_BASE + 3, first=first, p_trunc=p_trunc)
        t0 = time.time()
        grec, gops = eng.align_batch(seqs, off, rc, cabi.WANT_OPS)
        t_gpu += time.time() - t0
        t0 = time.time()
        orec, oops, _ = O.align_batch(om, seqs, off, rc, nthreads=threads, ops_stride=gops.shape[1])
        t_cpu += time.time() - t0
        bad = compare_records(grec, orec)
        res["record_mismatches"] += len(bad)
        diff_rows = int((gops != oops).any(axis=1).sum())
        res["ops_rows_differing"] += diff_rows
        res["gapped_reads"] += int((orec["indel"] != 0).sum())
        s = summary_from_records(orec, t.m, 0, om.EditStop)
        summ_want = s if summ_want is None else summ_want + s
        piece = {"first": first, "n": n, "record_fields_differing": [b[0] for b in bad], "ops_rows_differing": diff_rows}
        if text_left > 0:
            k = min(text_left, n)
            text_left -= k
            raw = seqs.tobytes()
            tb = td = 0
            dt_gpu = dt_cpu = 0.0
            for k0 in range(0, k, 200_000):  # bounded host memory: ~0.6 GB of text per side and step
                k1 = min(k, k0 + 200_000)
                # the FASTA file `treat align -f` would read: ids "r<index>-<read_count>" (the oracle builds the same names)
                lines = []
                for i in range(k0, k1):
                    lines.append(b">r%d-%d\n" % (i - k0, rc[i]))
                    lines.append(raw[int(off[i]):int(off[i + 1])])
                    lines.append(b"\n")
                fasta = b"".join(lines)
                t0 = time.time()
                got, nrec = eng.fasta_text(fasta, cabi.NO_SUMMARY)
                dt_gpu += time.time() - t0
                assert nrec == k1 - k0
                t0 = time.time()
                want, _ = O.text_batch(om, seqs[int(off[k0]):int(off[k1])], off[k0:k1 + 1] - off[k0], rc[k0:k1], prefix="r", nthreads=threads)
                dt_cpu += time.time() - t0
                g = np.frombuffer(got, dtype=np.uint8)
                c = min(len(g), len(want))
                td += abs(len(g) - len(want)) + int((g[:c] != want[:c]).sum())
                tb += int(len(want))
                del got, want, g, fasta, lines
            res["text_reads"] += k
            res["text_bytes"] += tb
            res["text_bytes_differing"] += td
            piece.update({"text_reads": k, "text_bytes": tb, "text_bytes_differing": td,
                          "text_seconds_gpu_path": round(dt_gpu, 3), "text_seconds_oracle": round(dt_cpu, 3)})
        res["pieces"].append(piece)
        print("[parity] reads %d..%d: record fields differing %s, ops rows differing %d%s" % (
            first, first + n, piece["record_fields_differing"], diff_rows,
            (", text bytes differing %d of %d" % (piece["text_bytes_differing"], piece["text_bytes"])) if "text_bytes" in piece else ""),
            file=sys.stderr, flush=True)
        first += n
    res["summary_equal"] = bool(np.array_equal(eng.summary(), summ_want))
    res["seconds_gpu_path"] = round(t_gpu, 2)
    res["seconds_oracle"] = round(t_cpu, 2)
    res["bit_exact"] = (res["record_mismatches"] == 0 and res["ops_rows_differing"] == 0 and res["text_bytes_differing"] == 0
                        and res["summary_equal"])
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(line + "\n")
    eng.close()
    return 0 if res["bit_exact"] else 1


if __name__ == "__main__":
    sys.exit(main())
