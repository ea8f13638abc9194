"""The two `treat align` binaries on one reads file: treat_b200/bin/treat_align (GPU: FASTA parse, alignment and WriteTo on the
device, text streamed to stdout) against oracle/_build/oracle_align (the CPU port's CLI, one thread like the Go tool), compared
by the MD5 of their stdout.  Test infrastructure: runs the oracle.

    python tests/cli_full.py --reads 300000 --out gpurun_out/cli_full.json
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_md5(cmd):
    t0 = time.time()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE)
    h, n = hashlib.md5(), 0
    while True:
        b = p.stdout.read(1 << 22)
        if not b:
            break
        h.update(b)
        n += len(b)
    rc = p.wait()
    return h.hexdigest(), n, time.time() - t0, rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=300000)
    ap.add_argument("--oracle-reads", type=int, default=0, help="reads the one-thread oracle CLI gets (0 = all)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from treat_b200 import synth, _build
    from oracle import oracle as O
    O.build()
    _build.build()
    t = synth.rps12_template()
    d = tempfile.mkdtemp(prefix="treat_cli_")
    tp = t.write_fasta(os.path.join(d, "templates.fa"))
    seqs, off, rc = synth.reads(t, a.reads, seed=synth.SEED_BASE + 3)
    raw = seqs.tobytes()
    fp = os.path.join(d, "reads.fa")
    with open(fp, "wb") as fh:
        for i in range(a.reads):
            fh.write(b">%d-%d\n" % (i + 1, rc[i]))  # fastx_collapser-style ids (fragment.go:43-55)
            fh.write(raw[int(off[i]):int(off[i + 1])])
            fh.write(b"\n")
    gpu = run_md5([os.path.join(ROOT, "treat_b200", "bin", "treat_align"), "-t", tp, "-f", fp])
    cpu = run_md5([os.path.join(ROOT, "oracle", "_build", "oracle_align"), "-t", tp, "-f", fp])
    res = {"reads": a.reads, "file_bytes": os.path.getsize(fp),
           "treat_align_gpu": {"md5": gpu[0], "stdout_bytes": gpu[1], "seconds": round(gpu[2], 2), "rc": gpu[3]},
           "oracle_align_cpu_1_thread": {"md5": cpu[0], "stdout_bytes": cpu[1], "seconds": round(cpu[2], 2), "rc": cpu[3]},
           "identical": gpu[0] == cpu[0] and gpu[1] == cpu[1] and gpu[3] == 0 and cpu[3] == 0}
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(line + "\n")
    return 0 if res["identical"] else 1


if __name__ == "__main__":
    sys.exit(main())
