#!/usr/bin/env python
"""bench.py -- reads aligned/s and GCUPS of the TREAT alignment hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (NewFragment + NewAlignment for every read: 48-byte records
+ 2-bit traceback ops) over one batch of synthetic RPS12-shaped reads per GPU (SURVEY.md 8d,
BASELINE.json configs[2]; weak scaling: every rank aligns its own --reads reads, no data-path
collective: one all-reduce of the summary counters per job, issued behind the last timed step inside the timed region;
`nccl.every_step` times the stricter form with an all-reduce behind every step).

One JSON line on rank 0:
  value      whole-job reads/s with the raw reads already resident in HBM (device-timed, max over ranks)
  e2e        the same through treat_b200_align_batch with pinned HOST buffers (H2D + kernels + D2H)
  roofline   the kernel that takes the largest share of a step: k_frontier (finishes the reads that are the template
             edited up to a frontier from their raw characters) against the measured HBM copy bandwidth, timed alone
  roofline_fused  k_frontier + k_pack_finish2 (every read that needs no DP finished, the others packed and listed)
  roofline_dp   the DP stage (k_band_* / k_align2, run on every read) against the INT32 ALU peak measured live on this GPU
  kernels_ms    CUDA-event time of each stage alone; reads whose stripped bases equal the template's, or differ
             in exactly one base, are finished without a DP -- exact, see DESIGN.md 4
  cpu_baseline  the CPU oracle (a C port of the Go path; the Go reference cannot be built here)

--impl reference times that CPU port with all host threads on bounded samples of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_aligned_per_sec"
UNIT = "reads/s"
INT_OPS_PER_CELL = 8  # 1 compare/select, 3 adds, 2 max, 2 equality predicates


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000, help="reads per GPU per step")
    ap.add_argument("--workload", default="rps12", choices=["rps12", "long", "xlong"],
                    help="rps12: BASELINE configs[2]; long: configs[4] (m = 300, K = 10); xlong: m = 700 (column strips)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fasta", action="store_true", help="skip the FASTA-ingest measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the strong-scaling parity leg (digests vs tests/golden)")
    ap.add_argument("--no-side", action="store_true", help="skip the long-gene and noisy workload points")
    return ap.parse_args()


def make_template(workload):
    from treat_b200 import synth
    if workload == "rps12":
        return synth.rps12_template(), 0.0, synth.SEED_BASE + 3
    if workload == "xlong":
        return synth.long_template(700, 4), 0.2, synth.SEED_BASE + 7
    return synth.long_template(300, 8), 0.2, synth.SEED_BASE + 5


def workload_name(args, t):
    return "synthetic-%s m=%d K=%d reads/gpu=%d (NewFragment+NewAlignment: records + traceback ops)" % (
        args.workload, t.m, t.K, args.reads)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, pci_bus_id=None):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.nvml, self.h, self.stop_flag, self.thread, self.samples = None, None, False, None, []
        try:  # NVML in-process: a sample costs microseconds, so even a 10 ms timed region gets several
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM), n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM),
                                     int(reasons(self.h))))
            except Exception:
                pass
            time.sleep(0.0005)

    def sample_now(self):
        """One sample from the calling thread (used right after the last step is enqueued, while the GPU still runs)."""
        if not self.nvml:
            return
        n = self.nvml
        reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            self.samples.append((n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM), n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM),
                                 int(reasons(self.h))))
        except Exception:
            pass

    def start(self):
        if self.nvml:
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=2)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            seen = sorted(k for k, b in bits.items() if any(smp[2] & b for smp in self.samples))
            sm = [smp[0] for smp in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smp[1] for smp in self.samples)) if sm else None,
                    "reasons": seen, "samples": len(sm), "source": "NVML, sampled during the timed region"}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- the only place outside tests/ and smoke() that executes oracle/
# ---------------------------------------------------------------------------------------------
def oracle_template(t):
    from oracle import oracle as O
    d = tempfile.mkdtemp(prefix="treat_bench_")
    path = t.write_fasta(os.path.join(d, "tmpl.fa"))
    return O, O.NewTemplateFromFasta(path, O.FORWARD, t.edit_base)


def cpu_sample_rate(t, p_trunc, seed, n_reads, threads, repeats=1, first=0):
    """reads/s of the CPU port (lean mode: one DP per read, record + ops) on n_reads reads."""
    from treat_b200 import synth
    O, om = oracle_template(t)
    seqs, off, rc = synth.reads(t, n_reads, seed=seed, first=first, p_trunc=p_trunc)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rec, _, _ = O.align_batch(om, seqs, off, rc, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cells = float((rec["n_bases"].astype(np.int64) * t.m).sum())
    return n_reads / best, cells / best / 1e9, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, p_trunc, seed = make_template(args.workload)
    threads = len(os.sched_getaffinity(0))
    sample = args.cpu_sample or 4096 * threads
    times, cells = [], 0.0
    from treat_b200 import synth
    O, om = oracle_template(t)
    for s in range(args.warmup + args.steps):
        seqs, off, rc = synth.reads(t, sample, seed=seed, first=s * sample, p_trunc=p_trunc)
        t0 = time.perf_counter()
        rec, _, _ = O.align_batch(om, seqs, off, rc, nthreads=threads)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            cells += float((rec["n_bases"].astype(np.int64) * t.m).sum())
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "gcups": cells / total / 1e9,
        "config": {"workload": workload_name(args, t), "sample_reads_per_step": sample,
                   "note": "CPU port of the Go path (oracle/treat_oracle.c, full int64 score matrix + byte pointer "
                           "matrix per read like nwalgo); the Go reference itself cannot be built here (no Go toolchain)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d reads per step x %d steps, lean mode (one DP per read, record + ops)" % (sample, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def pinned_array(nbytes):
    from treat_b200 import cabi
    p = C.c_void_p()
    cabi.check(cabi.lib().treat_b200_host_alloc(C.byref(p), max(nbytes, 1)))
    arr = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p.value))
    return arr, p


def pin_to_gpu_numa_node(gpu_index):
    """Run this rank (and allocate its pinned buffers) on the CPUs next to its GPU: with 8 ranks on a
    two-socket host the H2D/D2H copies otherwise cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as e:  # best effort
        log("cpu affinity not set:", e)



# ---------------------------------------------------------------------------------------------
# legs added around the main measurement
# ---------------------------------------------------------------------------------------------
PARITY_READS = 1_000_000


def parity_leg(T, cabi, parallel, synth, eng, comm, args, t, seed, p_trunc, rank, world, barrier, max_over_ranks, alloc):
    """Strong-scaling parity: the FIRST 1,000,000 reads of the workload shared between the N GPUs (contiguous ranges, host
    buffers, every ops row), then SHA-256 over the ranks' canonical records, over their traceback ops and over the
    all-reduced summary -- the same digests at every N, and for rps12 equal to the CPU port's committed in
    tests/golden/parity_digests.json (tests/golden/make_digests.py)."""
    import hashlib
    P = PARITY_READS
    lo, hi = parallel.shard_range(P, rank, world)
    seqs, off, rc = synth.reads(t, hi - lo, seed=seed, first=lo, p_trunc=p_trunc, pinned_alloc=alloc)
    max_len = int(np.diff(off.astype(np.int64)).max()) if hi > lo else 0
    stride = eng.ops_stride(max_len)
    rec = alloc((hi - lo) * 48).view(cabi.RECORD_DTYPE)
    ops = alloc((hi - lo) * stride).reshape(hi - lo, stride)
    h_off, h_rc = alloc((hi - lo + 1) * 8).view(np.uint64), alloc((hi - lo) * 4).view(np.uint32)
    h_off[:], h_rc[:] = off, rc
    eng.align_batch(seqs, h_off, h_rc, cabi.WANT_OPS | cabi.NO_SUMMARY, rec=rec, ops=ops)  # warm-up (buffers, clocks)
    eng.reset_summary()
    ops[:] = 0
    barrier()
    t0 = time.perf_counter()
    eng.align_batch(seqs, h_off, h_rc, cabi.WANT_OPS, rec=rec, ops=ops)
    red = comm.reduce_summary()
    ms = 1e3 * max_over_ranks(time.perf_counter() - t0)
    eng.reset_summary()
    tag = "/dev/shm/treat_b200_parity_%s_%%d_%%s.bin" % os.environ.get("MASTER_PORT", str(os.getppid()))
    parallel.canonical_records(rec).tofile(tag % (rank, "rec"))
    parallel.compact_ops(rec, ops).tofile(tag % (rank, "ops"))
    np.asarray(red).tofile(tag % (rank, "sum"))
    np.concatenate([seqs, np.frombuffer(off[1:].tobytes() if rank else off.tobytes(), dtype=np.uint8)]).tofile(tag % (rank, "in"))
    barrier()
    out = None
    if rank == 0:
        def cat(kind):
            h = hashlib.sha256()
            for r in range(world):
                with open(tag % (r, kind), "rb") as fh:
                    while True:
                        b = fh.read(1 << 24)
                        if not b:
                            break
                        h.update(b)
            return h.hexdigest()
        sums = [np.fromfile(tag % (r, "sum"), dtype=np.uint64) for r in range(world)]
        same_on_all = all(np.array_equal(x, sums[0]) for x in sums)
        out = {"reads": P, "n_gpus": world, "records_sha256": cat("rec"), "ops_sha256": cat("ops"),
               "summary_sha256": parallel.sha256_hex(sums[0]), "summary_identical_on_all_ranks": bool(same_on_all),
               "summary_reads": int(sums[0][14]), "strong_scaling_e2e_ms": ms, "strong_scaling_e2e_reads_per_s": P / (ms * 1e-3),
               "api": "treat_b200_align_batch per rank on its contiguous range (host buffers, every ops row) + "
                      "treat_b200_comm_reduce_summary (the library's ncclAllReduce)"}
        want = None
        gp = os.path.join(ROOT, "tests", "golden", "parity_digests.json")
        if os.path.exists(gp):
            want = json.load(open(gp)).get(args.workload)
        if want and want["reads"] == P and want["seed"] == seed and want["p_trunc"] == p_trunc:
            out["golden"] = "tests/golden/parity_digests.json (CPU port of the Go path on the same reads)"
            out["matches_golden"] = bool(out["records_sha256"] == want["records_sha256"] and out["ops_sha256"] == want["ops_sha256"]
                                         and out["summary_sha256"] == want["summary_sha256"] and same_on_all)
        else:
            out["golden"] = None
            out["matches_golden"] = None
        if out["matches_golden"] is False or not same_on_all or out["summary_reads"] != P:
            raise AssertionError("parity leg: digests differ from the oracle's: %s vs %s" % (out, want))
    barrier()
    for kind in ("rec", "ops", "sum", "in"):
        try:
            os.remove(tag % (rank, kind))
        except OSError:
            pass
    return out


def h2d_floor(cabi, eng, host_arr, nbytes, barrier, max_over_ranks, world):
    """What the host's PCIe / memory path gives: one plain pinned cudaMemcpyAsync of a step's input bytes, all ranks at once."""
    ms = C.c_double()
    barrier()
    t0 = time.perf_counter()
    cabi.check(cabi.lib().treat_b200_measure_h2d(eng._ctx, host_arr.ctypes.data, nbytes, 5, C.byref(ms)), eng._ctx)
    wall = time.perf_counter() - t0
    worst = max_over_ranks(ms.value)
    return {"h2d_floor_ms": worst, "h2d_floor_gb_per_s_per_gpu": nbytes / (worst * 1e-3) / 1e9,
            "h2d_floor_gb_per_s_all_gpus": world * nbytes / (worst * 1e-3) / 1e9, "h2d_floor_this_rank_ms": ms.value,
            "h2d_floor_what": "cudaMemcpyAsync H2D of the step's %d input bytes from pinned memory, 5 copies back to back, "
                              "CUDA events, every rank at once, max over ranks (%.0f ms wall incl. warm-up)" % (nbytes, 1e3 * wall)}


def side_workload(T, cabi, synth, eng, name, N, steps, alloc):
    """A second workload point on this GPU (N = 1 only): device-resident and end-to-end reads/s, DP share, geometry."""
    import torch
    if name == "long":
        t, p_trunc, seed, noise = synth.long_template(300, 8), 0.2, synth.SEED_BASE + 5, {}
        what = "BASELINE configs[4]: long-gene set m=300 K=10, variable-length reads (20%% truncated by up to 100 bases)"
    else:
        t, p_trunc, seed, noise = synth.rps12_template(), 0.0, synth.SEED_BASE + 3, {"p_sub_base": 0.01, "p_indel_read": 0.05}
        what = "RPS12 reads with sequencing noise: 1% substitutions per base and one more single-base indel in 5% of the reads"
    with tempfile.TemporaryDirectory() as d:
        tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "tmpl.fa")), T.FORWARD, t.edit_base)
    eng.SetTemplate(tm)
    seqs, off, rc = synth.reads(t, N, seed=seed, p_trunc=p_trunc, pinned_alloc=alloc, **noise)
    max_len = int(np.diff(off.astype(np.int64)).max())
    stride = eng.ops_stride(max_len)
    h_off, h_rc = alloc((N + 1) * 8).view(np.uint64), alloc(N * 4).view(np.uint32)
    h_off[:], h_rc[:] = off, rc
    h_rec = alloc(N * 48).view(cabi.RECORD_DTYPE)
    h_ops = alloc(N * stride).reshape(N, stride)
    flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
    dev = torch.device("cuda", eng.device)
    d_seqs = torch.from_numpy(np.ascontiguousarray(seqs)).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_rc = torch.from_numpy(rc.astype(np.int32)).to(dev)
    d_rec = torch.empty(N * 48, dtype=torch.uint8, device=dev)
    d_ops = torch.empty(N * stride, dtype=torch.uint8, device=dev)
    sraw = torch.cuda.current_stream().cuda_stream

    def step_dev(fl):
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, fl,
                               d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw)

    def timed(fl):
        for _ in range(2):
            step_dev(fl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        for _ in range(steps):
            step_dev(fl)
        b.record(torch.cuda.current_stream())
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps
    dev_ms = timed(flags)
    torch.cuda.synchronize()
    eng.dp_stats()
    step_dev(flags)
    torch.cuda.synchronize()  # the step ran on this process's stream: the counters are final only behind it
    paths = eng.dp_stats()  # (pure deletions, banded fill, full matrix) of one step, counted by the kernels
    dp_ms = timed(flags | cabi.DP_EVERY_READ)
    full_ms = timed(flags | cabi.DP_EVERY_READ | cabi.NO_BAND)
    step_dev(flags)
    torch.cuda.synchronize()
    rec = d_rec.cpu().numpy().view(cabi.RECORD_DTYPE)
    no_dp = (rec["indel"] == 0) & (rec["mismatches"] <= 1) & (rec["n_bases"] == t.m)
    cells = rec["n_bases"].astype(np.int64) * t.m
    for _ in range(2):
        eng.align_batch(seqs, h_off, h_rc, flags, rec=h_rec, ops=h_ops)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.align_batch(seqs, h_off, h_rc, flags, rec=h_rec, ops=h_ops)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / steps
    assert np.array_equal(h_rec, rec), "side workload %s: host-buffer and device-buffer paths disagree" % name
    out = {"workload": what, "reads_per_step": N, "steps": steps, "reads_total": N * steps, "m": t.m, "K": t.K,
           "raw_bytes_per_read": float(len(seqs)) / N, "max_raw_len": max_len,
           "value": N / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms,
           "value_dp_every_read": N / (dp_ms * 1e-3), "ms_per_step_dp_every_read": dp_ms,
           "e2e": {"value": N / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(len(seqs) + (N + 1) * 8 + N * 4)},
           "reads_finished_without_dp": float(no_dp.mean()), "reads_with_gap": float((rec["indel"] != 0).mean()),
           "dp_paths": {"pure_deletion_no_dp": paths[0] / N, "banded_fill": paths[1] / N, "full_matrix": paths[2] / N,
                        "what": "fraction of the reads by the kernel that finished them after the fused kernel (k_band_classify / "
                                "k_band_fill + k_band_finish / k_align*), counted on the device"},
           "gcups_performed": performed_cells(rec, t.m)[0] / (dev_ms * 1e-3) / 1e9,
           "gcups_effective_dp_every_read": float(cells.sum()) / (dp_ms * 1e-3) / 1e9,
           "value_full_matrix_dp_every_read": N / (full_ms * 1e-3), "ms_per_step_full_matrix_dp_every_read": full_ms,
           "gcups_full_matrix_dp_every_read": float(cells.sum()) / (full_ms * 1e-3) / 1e9}
    geom = cabi.AlignGeometry()
    if cabi.lib().treat_b200_align_geometry(eng._ctx, int(rec["n_bases"].max()), 0, C.byref(geom)) == 0:
        out["geometry"] = {f: getattr(geom, f) for f, _ in cabi.AlignGeometry._fields_}
    del d_seqs, d_off, d_rc, d_rec, d_ops
    return out


def performed_cells(rec, m):
    """DP cells the kernels fill for these records (what gcups_performed counts), from the records alone: reads the fused
    kernel finishes (no gap, at most one mismatch, n = m) and pure deletions (gap-only reads with score 2n - m: leftmost
    embedding) need none; reads within 15 bases of the template's length take the banded fill (32 diagonals x max(m, n)
    rows; the few whose certificate fails are counted as banded too); the rest fill the whole m x n matrix."""
    n = rec["n_bases"].astype(np.int64)
    no_dp = (rec["indel"] == 0) & (rec["mismatches"] <= 1) & (n == m)
    subseq = (rec["indel"] != 0) & (rec["mismatches"] == 0) & (n <= m) & (rec["score"].astype(np.int64) == 2 * n - m)
    band = ~no_dp & ~subseq & (np.abs(n - m) < 16)
    full = ~no_dp & ~subseq & ~band
    cells = np.zeros(len(rec), dtype=np.int64)
    cells[band] = 32 * np.maximum(n[band], m)
    cells[full] = n[full] * m
    return float(cells.sum()), {"no_dp": float(no_dp.mean()), "pure_deletion": float(subseq.mean()), "banded": float(band.mean()),
                                "full_matrix": float(full.mean())}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import treat_b200 as T
    from treat_b200 import cabi, parallel, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; treat_b200 has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL_DEBUG (if the caller set it) stays: its lines go to stderr so that stdout keeps the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    t, p_trunc, seed = make_template(args.workload)
    with tempfile.TemporaryDirectory() as d:
        tm = T.NewTemplateFromFasta(t.write_fasta(os.path.join(d, "tmpl.fa")), T.FORWARD, t.edit_base)
    eng = T.Aligner(local)
    eng.SetTemplate(tm)
    # The one collective of the path (the summary counters) is the library's own ncclAllReduce (treat_b200_comm_*): rank 0
    # makes the NCCL id, torch.distributed only carries its 128 bytes to the other ranks (plus barriers / max-over-ranks).
    uid_t = torch.zeros(cabi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    if world > 1:
        if rank == 0:
            uid_t.copy_(torch.frombuffer(bytearray(T.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid_t, 0)
    comm = T.Comm(eng, world, rank, bytes(uid_t.cpu().numpy().tobytes()) if world > 1 else None)

    N = args.reads
    t0 = time.perf_counter()
    hold = []

    def alloc(nbytes):
        arr, p = pinned_array(nbytes)
        hold.append(p)
        return arr
    seqs, off, rc = synth.reads(t, N, seed=seed, first=rank * N, p_trunc=p_trunc, pinned_alloc=alloc)
    lens = np.diff(off.astype(np.int64))
    max_len = int(lens.max())
    stride = eng.ops_stride(max_len)
    h_rec_raw = alloc(N * 48)
    h_ops_raw = alloc(N * stride)
    h_rec = h_rec_raw.view(cabi.RECORD_DTYPE)
    h_ops = h_ops_raw.reshape(N, stride)
    h_off, h_rc = alloc((N + 1) * 8).view(np.uint64), alloc(N * 4).view(np.uint32)
    h_off[:] = off
    h_rc[:] = rc
    log("[rank %d] generated %d reads (%.1f MB raw, max_len %d) in %.1fs" % (rank, N, len(seqs) / 1e6, max_len, time.perf_counter() - t0))

    int_peak, _ = eng.measure_int32_peak()

    # ---- device-resident inputs --------------------------------------------------------------
    d_seqs = torch.from_numpy(np.ascontiguousarray(seqs)).to(dev)
    d_off = torch.from_numpy(off.astype(np.int64)).to(dev)
    d_rc = torch.from_numpy(rc.astype(np.int32)).to(dev)
    d_rec = torch.empty(N * 48, dtype=torch.uint8, device=dev)
    d_ops = torch.empty(N * stride, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)  # explicit stream: the C ABI launches on it, the events time it
    torch.cuda.set_stream(stream)
    sraw = stream.cuda_stream
    assert sraw != 0
    # records + the traceback ops of every read that has a gap: a read without a gap aligns base over base (its row
    # would be all zero), so its alignment strings follow from the record alone -- what NewAlignments asks for
    flags = cabi.WANT_OPS | cabi.OPS_GAPPED_ONLY
    d_red = torch.zeros(parallel.summary_tensor(eng).numel(), dtype=torch.int64, device=dev)

    def step_device(reduce=False):
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, flags,
                               d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw)
        # the only collective on the path: ~4 KB of uint64 counters, ncclAllReduce issued by the library.  A job reduces its
        # summary once, behind its last batch (the counters accumulate on the device; the reference sums its per-read records
        # when the load is over: cmd/treat/stats.go:140-187) -- `reduce` marks that batch; `nccl.every_step` times the
        # stricter form with an all-reduce behind every step
        if reduce:
            comm.reduce_summary_device(d_red.data_ptr(), sraw)
        return d_red

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            tt = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return x

    for _ in range(args.warmup):
        step_device(reduce=True)
    barrier()
    try:
        pr = torch.cuda.get_device_properties(dev)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        bus = None
    sampler = ClockSampler(local, bus)
    sampler.start()
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step_device(reduce=(k == args.steps - 1))
    comm.join(sraw)  # the timed region ends when the job's all-reduce has
    e1.record(stream)
    sampler.sample_now()  # the queue is still draining: a sample inside the timed region even when it lasts milliseconds
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - l0
    # the one collective of the path: the all-reduced "reads aligned" counter must hold every rank's reads of every step so far
    torch.cuda.synchronize()
    reduced_reads = int(d_red[14].item())
    assert reduced_reads == world * N * (args.warmup + args.steps), (reduced_reads, world, N)
    # the stricter form: an all-reduce behind every step (it runs on the communicator's stream next to the following step)
    every_ms = None
    if world > 1:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_device(reduce=True)
        comm.join(sraw)
        e1.record(stream)
        barrier()
        every_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        assert int(d_red[14].item()) == world * N * (args.warmup + 2 * args.steps)
    clocks = sampler.stop()
    rec_dev = d_rec.cpu().numpy().view(cabi.RECORD_DTYPE)
    cells_per_step = float((rec_dev["n_bases"].astype(np.int64) * t.m).sum())
    # reads finished without a DP: no gap, at most one mismatch, as long as the template (k_pack's marks)
    no_dp = (rec_dev["indel"] == 0) & (rec_dev["mismatches"] <= 1) & (rec_dev["n_bases"] == t.m)
    cells_performed, path_frac = performed_cells(rec_dev, t.m)
    value = world * N * args.steps / (dev_ms * 1e-3)

    # ---- the dominant kernel alone (k_align on already-packed reads), CUDA events ----------------
    pk = cabi.Packed()
    cabi.check(cabi.lib().treat_b200_packed_layout(eng._ctx, N, max_len, 0, C.byref(pk)), eng._ctx)
    d_n = torch.empty(N, dtype=torch.int16, device=dev)
    d_fl = torch.empty(N, dtype=torch.uint8, device=dev)
    d_b = torch.empty(N * pk.bases_stride, dtype=torch.uint8, device=dev)
    d_c = torch.empty(N * pk.counts_stride, dtype=torch.uint8, device=dev)
    d_rl = torch.empty(N, dtype=torch.int32, device=dev)
    pk.d_n, pk.d_flags, pk.d_bases, pk.d_counts, pk.d_raw_len = (d_n.data_ptr(), d_fl.data_ptr(), d_b.data_ptr(),
                                                                 d_c.data_ptr(), d_rl.data_ptr())
    cabi.check(cabi.lib().treat_b200_pack_device(eng._ctx, d_seqs.data_ptr(), d_off.data_ptr(), C.byref(pk), sraw), eng._ctx)
    torch.cuda.synchronize()
    max_n = int(d_n.max().item())
    pk.max_len = max_n  # size the align launch for the longest stripped read, as the batch entry point does

    def kernel_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    pk_full = cabi.Packed()
    C.memmove(C.byref(pk_full), C.byref(pk), C.sizeof(pk))
    pk_full.max_len = max_len
    # the DP stage on every read (what roofline_dp is about): k_band_classify / k_band_fill / k_band_finish, then k_align2 for
    # the reads whose certificate fails ...
    align_ms = kernel_ms(lambda: cabi.check(cabi.lib().treat_b200_align_packed_device(
        eng._ctx, C.byref(pk), d_rc.data_ptr(), flags, d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw), eng._ctx), args.steps)
    eng.dp_stats()
    cabi.check(cabi.lib().treat_b200_align_packed_device(eng._ctx, C.byref(pk), d_rc.data_ptr(), flags, d_rec.data_ptr(), d_ops.data_ptr(),
                                                         stride, sraw), eng._ctx)
    torch.cuda.synchronize()
    dp_paths_every = eng.dp_stats()
    # ... and the full m x n matrix of every read (k_align2 alone, TREAT_B200_NO_BAND): the kernel round 1 reported
    full_ms = kernel_ms(lambda: cabi.check(cabi.lib().treat_b200_align_packed_device(
        eng._ctx, C.byref(pk), d_rc.data_ptr(), flags | cabi.NO_BAND, d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw), eng._ctx),
        max(2, args.steps // 2))
    # ... and what the batch call launches: k_same_bases finishes the reads whose bases equal the template's without a
    # DP and lists the others for the DP kernel
    marks_ms = kernel_ms(lambda: cabi.check(cabi.lib().treat_b200_align_packed_device(
        eng._ctx, C.byref(pk), d_rc.data_ptr(), flags | cabi.USE_PACK_MARKS, d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw),
        eng._ctx), args.steps)
    same_frac = float((d_fl.cpu().numpy() & 0x80).astype(bool).mean())
    pack_ms = kernel_ms(lambda: cabi.check(cabi.lib().treat_b200_pack_device(
        eng._ctx, d_seqs.data_ptr(), d_off.data_ptr(), C.byref(pk_full), sraw), eng._ctx), args.steps)
    n_arr = d_n.cpu().numpy().astype(np.int64) & 0xffff
    # the fused kernel alone (k_pack_finish2 + the multi-pass form for reads longer than one pass): what the batch call
    # launches before the DP kernel.  TREAT_B200_SKIP_DP leaves the DP launch out (the 8-byte read-back of max n stays).
    fused_ms = kernel_ms(lambda: eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len,
                                                        flags | cabi.SKIP_DP | cabi.NO_SUMMARY, d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw),
                         args.steps)
    # k_frontier alone (TREAT_B200_FRONTIER_ONLY: nothing runs behind it): the kernel that takes the largest share of a step.
    # Which reads it finishes is read off the records: the buffer is filled with 0xFF first, a finished read has its length in
    # the last word.
    d_rec.fill_(0xff)
    frontier_args = (d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, flags | cabi.SKIP_DP | cabi.FRONTIER_ONLY | cabi.NO_SUMMARY,
                     d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw)
    eng.align_batch_device(*frontier_args)
    torch.cuda.synchronize()
    by_frontier = d_rec.cpu().numpy().view(cabi.RECORD_DTYPE)["raw_len"] != 0xffffffff
    frontier_ms = kernel_ms(lambda: eng.align_batch_device(*frontier_args), args.steps) if by_frontier.any() else None
    eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len, flags | cabi.NO_SUMMARY, d_rec.data_ptr(),
                           d_ops.data_ptr(), stride, sraw)  # the records of the whole batch again
    # the whole step with the DP forced on every read (TREAT_B200_DP_EVERY_READ: k_pack16 + k_align2 on all reads)
    dp_every_ms = kernel_ms(lambda: eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), N, max_len,
                                                           flags | cabi.DP_EVERY_READ | cabi.NO_SUMMARY, d_rec.data_ptr(), d_ops.data_ptr(), stride,
                                                           sraw), max(2, args.steps // 2))
    # ---- Alignment.WriteTo for the batch, everything resident in HBM (k_text: sizes, scan, write) ----------------
    text_dev = None
    if rank == 0 and world == 1 and not args.no_fasta:
        tn = min(N, 250_000)  # 0.7 GB of text
        names = np.frombuffer(b"".join(b"r%d-%d" % (i, rc[i]) for i in range(tn)), dtype=np.uint8)
        noff = np.zeros(tn + 1, dtype=np.int64)
        noff[1:] = np.cumsum([len(b"r%d-%d" % (i, rc[i])) for i in range(tn)])
        d_names, d_noff = torch.from_numpy(names.copy()).to(dev), torch.from_numpy(noff).to(dev)
        d_toff = torch.empty(tn + 1, dtype=torch.int64, device=dev)
        # every read's ops row stays on the device for the formatter
        eng.align_batch_device(d_seqs.data_ptr(), d_off.data_ptr(), d_rc.data_ptr(), tn, max_len, cabi.WANT_OPS | cabi.NO_SUMMARY,
                               d_rec.data_ptr(), d_ops.data_ptr(), stride, sraw)
        targs = (d_seqs.data_ptr(), d_off.data_ptr(), d_names.data_ptr(), d_noff.data_ptr(), d_rec.data_ptr(), d_ops.data_ptr(), stride,
                 tn, max_len, 80, d_toff.data_ptr())
        tbytes = eng.text_device(*targs, 0, 0, sraw)
        d_text = torch.empty(tbytes, dtype=torch.uint8, device=dev)
        text_ms = kernel_ms(lambda: eng.text_device(*targs, d_text.data_ptr(), tbytes, sraw), 5)
        raw_bytes = int(off[tn])
        text_dev = {"value": tn / (text_ms * 1e-3), "unit": UNIT, "ms_per_launch_pair": text_ms, "reads": tn, "text_bytes": int(tbytes),
                    "text_gb_per_s": tbytes / (text_ms * 1e-3) / 1e9,
                    # algorithmic HBM bytes: raw reads twice (both passes), records + ops rows twice, the text once
                    "hbm_frac": (2 * raw_bytes + 2 * tn * (48 + stride) + tbytes) / (text_ms * 1e-3) / 1e9 / (
                        float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0),
                    "api": "treat_b200_text_device: k_text sizes + scan + k_text write for reads, names, records and ops rows in HBM "
                           "(includes the one host synchronisation that returns the text size)"}
        del d_text, d_names, d_noff, d_toff
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        hbm_peak, hbm_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"

    def ncu_traffic(kernel):
        # DRAM bytes per launch from the committed ncu capture (per read x reads of this launch)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r04_traffic.json")))[kernel]
            return (tr["dram_read_bytes_per_read"] + tr["dram_write_bytes_per_read"]) * N if args.workload == "rps12" else None
        except Exception:
            return None
    def ncu_facts(kernel):
        # issue-slot / pipe utilisation and instruction counts of the committed ncu capture (profiles/r04_*_ncu_full.md)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r04_traffic.json")))[kernel]
            return {k: v for k, v in tr.items() if not k.startswith("dram_")}
        except Exception:
            return None
    # k_pack_finish2 (the fused kernel, largest share of a step): raw characters + offsets + read counts in, the 48-byte
    # record out; for the reads left for the DP also the packed read (2-bit bases, counts, n / flags / raw_len) and a list entry
    listed = ~no_dp
    packed_bytes = float(((n_arr + 3) // 4 + (n_arr + 1))[listed].sum() + int(listed.sum()) * (2 + 1 + 4 + 4))
    fused_bytes = float(len(seqs) + N * 8 + N * 4 + int(no_dp.sum()) * 48) + packed_bytes
    roofline_fused = {"bound": "hbm", "kernel": "k_frontier + k_pack_finish2 on the reads it leaves (NewFragment + editing-site assembly of "
                                                "every read that needs no DP, the others packed and listed)",
                      "achieved": fused_bytes / (fused_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": fused_bytes / (fused_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                      "bytes_per_launch": fused_bytes, "ms_per_launch": fused_ms, "bytes_per_read": fused_bytes / N,
                      "note": "ms_per_launch covers both launches (plus the multi-pass form for reads longer than one pass) and one 8-byte read-back"}
    if frontier_ms is not None:
        # k_frontier: raw characters + offsets + read counts in; the 48-byte record of every read it finishes and a 4-byte list
        # entry for every other read out
        n_fr = int(by_frontier.sum())
        fr_bytes = float(len(seqs) + N * 8 + N * 4 + n_fr * 48 + (N - n_fr) * 4)
        roofline = {"bound": "hbm", "kernel": "k_frontier (NewFragment + NewAlignment of the reads that are the template edited up to a frontier, "
                                              "from two string compares on the raw characters; largest share of a step)",
                    "achieved": fr_bytes / (frontier_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": fr_bytes / (frontier_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                    "bytes_per_launch": fr_bytes, "ms_per_launch": frontier_ms, "traffic": ncu_traffic("k_frontier"), "ncu": ncu_facts("k_frontier"),
                    "bytes_per_read": fr_bytes / N, "reads_finished": n_fr / N,
                    "note": "byte work on the raw characters (XOR against the two template strings, first / last difference, junction "
                            "check): issue slots, not DRAM, are the nearer bound -- see `issue` and profiles/"}
    else:  # the template does not take k_frontier (wide alphabet, very long strings): the fused kernel is the dominant one
        roofline = dict(roofline_fused)
    # the same kernel against the bound that actually limits it: warp instructions issued per second (from the committed ncu
    # capture's instruction count per read) against 4 issue slots per SM and clock
    facts = ncu_facts("k_frontier")
    if frontier_ms is not None and facts and facts.get("warp_instructions_per_read") and args.workload == "rps12":
        props = torch.cuda.get_device_properties(dev)
        mhz = float(clocks.get("sm_mhz") or 1965.0) if isinstance(clocks, dict) else 1965.0
        issue_peak = props.multi_processor_count * 4 * mhz * 1e6
        issued = facts["warp_instructions_per_read"] * N / (frontier_ms * 1e-3)
        roofline["issue"] = {"bound": "warp instructions issued", "achieved": issued / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-instr/s",
                             "frac": issued / issue_peak, "warp_instructions_per_read": facts["warp_instructions_per_read"],
                             "note": "instruction count per read from profiles/r04_traffic.json (ncu), time measured here"}
    # the unfused NewFragment kernel (treat_b200_pack_device, used by the stage calls and for long / wide templates)
    pack_bytes = float(len(seqs) + N * 8 + ((n_arr + 3) // 4 + (n_arr + 1)).sum() + N * (2 + 1 + 4))
    roofline_pack = {"bound": "hbm", "kernel": "k_pack16 (stage call treat_b200_pack_device)", "achieved": pack_bytes / (pack_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s", "frac": pack_bytes / (pack_ms * 1e-3) / 1e9 / hbm_peak, "ms_per_launch": pack_ms,
                     "bytes_per_launch": pack_bytes}
    # the DP stage on every read.  Algorithmic work: m x n cells x 8 scalar integer ops; the banded fill
    # performs 32 x max(m, n) cells for a read it certifies, so two figures are reported: `frac` on the cells the kernels
    # actually fill, `frac_algorithmic` on the full matrices they replace (the banded kernel makes that one exceed 1).
    alg_bytes = float(((n_arr + 3) // 4 + (n_arr + 1)).sum() + N * (2 + 1 + 4 + 4) + N * 48 + ((t.m + n_arr + 3) // 4).sum())
    int_ops = cells_per_step * INT_OPS_PER_CELL
    band_reads, full_reads = dp_paths_every[1], dp_paths_every[2]
    near = np.abs(n_arr - t.m) < 16
    cells_filled = float((32 * np.maximum(n_arr, t.m))[near].sum() + (n_arr * t.m)[~near].sum())
    roofline_dp = {"bound": "int32", "kernel": "DP stage on every read: k_band_fill + k_band_finish (banded, certified), k_align2 for the rest",
                   "achieved": cells_filled * INT_OPS_PER_CELL / (align_ms * 1e-3) / 1e12, "peak": int_peak / 1e12,
                   "unit": "Tiop/s", "frac": cells_filled * INT_OPS_PER_CELL / (align_ms * 1e-3) / int_peak,
                   "frac_algorithmic": int_ops / (align_ms * 1e-3) / int_peak,
                   "peak_source": "measured live: k_int32_peak (dependent-free add+max, all SMs)",
                   "ops_per_cell": INT_OPS_PER_CELL, "cells_per_launch": cells_per_step, "cells_filled_per_launch": cells_filled,
                   "ms_per_launch": align_ms,
                   "gcups": cells_per_step / (align_ms * 1e-3) / 1e9, "gcups_performed": cells_filled / (align_ms * 1e-3) / 1e9,
                   "reads_banded": int(band_reads), "reads_full_matrix": int(full_reads),
                   "full_matrix_every_read": {"kernel": "k_align2 (TREAT_B200_NO_BAND)", "ms_per_launch": full_ms,
                                              "gcups": cells_per_step / (full_ms * 1e-3) / 1e9,
                                              "frac": int_ops / (full_ms * 1e-3) / int_peak},
                   "traffic": ncu_traffic("k_band_fill"), "ncu": ncu_facts("k_band_fill"),
                   "hbm_frac": alg_bytes / (align_ms * 1e-3) / 1e9 / hbm_peak,
                   "note": "frac counts the cells the kernels fill (a 32-diagonal band for certified reads) at 8 scalar ops per "
                           "cell against the measured INT32 peak; the kernels need ~4 instructions per cell PAIR (two reads per "
                           "register: PRMT score select, IMAD add, VIMNMX3.S16x2, IMAD direction digits), so frac can exceed 1"}
    del d_b, d_c, d_n, d_fl, d_rl

    # ---- end to end through the host-buffer C ABI call -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        def step_host(fl=flags):
            eng.align_batch(seqs, h_off, h_rc, fl, rec=h_rec, ops=h_ops)
            return int(comm.reduce_summary()[0])  # all-GPU counters of the step, read back to the host
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        dev_stride = ((t.m + max_n + 3) // 4 + 15) // 16 * 16
        n_gapped = int((rec_dev["indel"] != 0).sum())
        e2e = {"value": world * N * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(len(seqs) + (N + 1) * 8 + N * 4),
               # the library compacts the gapped reads' rows only while they are rare (< 1/32 of a chunk), else every row
               "d2h_bytes_per_step": int(N * 48 + (n_gapped * (dev_stride + 4) if n_gapped * 32 <= N else N * dev_stride)
                                         + 32 * ((N + 65535) // 65536) + 8),
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "api": "treat_b200_align_batch (pinned host buffers), records + ops rows of the %d reads with a gap "
                      "(TREAT_B200_OPS_GAPPED_ONLY)" % n_gapped}
        # the same call returning the ops row of every read
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host(cabi.WANT_OPS)
        barrier()
        full_s = max_over_ranks(time.perf_counter() - t0)
        e2e["every_ops_row"] = {"value": world * N * args.steps / full_s, "ms_per_step": 1e3 * full_s / args.steps,
                                "d2h_bytes_per_step": int(N * 48 + N * dev_stride + 8)}
        # the floor of this call: the same input bytes as one plain pinned H2D copy per rank, all ranks at once
        e2e.update(h2d_floor(cabi, eng, seqs, int(len(seqs)), barrier, max_over_ranks, world))
        e2e["frac_of_h2d_floor"] = e2e["h2d_floor_ms"] / e2e["ms_per_step"]
        # ... and with everything the step moves: the inputs (reads, offsets, read counts) in and the results out at once
        tms = C.c_double()
        barrier()
        cabi.check(cabi.lib().treat_b200_measure_transfer(eng._ctx, seqs.ctypes.data, min(int(e2e["h2d_bytes_per_step"]), seqs.base.nbytes if seqs.base is not None else seqs.nbytes),
                                                          h_ops_raw.ctypes.data, min(int(e2e["d2h_bytes_per_step"]), h_ops_raw.nbytes), 5, C.byref(tms)), eng._ctx)  # (scratch: the ops buffer)
        e2e["transfer_floor_ms"] = max_over_ranks(tms.value)
        e2e["frac_of_transfer_floor"] = e2e["transfer_floor_ms"] / e2e["ms_per_step"]
        e2e["transfer_floor_what"] = ("per step one pinned H2D copy of the step's input bytes and, on a second stream, one D2H copy of its "
                                      "result bytes, every rank at once, max over ranks: what the box's PCIe / memory path allows")
        if rank == 0 and world == 1:
            # the same call from pageable buffers (what a cgo caller passing Go slices has): staged through pinned buffers
            # by the library's host copy threads
            pg_seqs, pg_off, pg_rc = np.array(seqs), np.array(h_off), np.array(h_rc)
            pg_rec, pg_ops = np.zeros(N, dtype=cabi.RECORD_DTYPE), np.zeros((N, stride), dtype=np.uint8)
            eng.align_batch(pg_seqs, pg_off, pg_rc, flags, rec=pg_rec, ops=pg_ops)
            t0 = time.perf_counter()
            for _ in range(3):
                eng.align_batch(pg_seqs, pg_off, pg_rc, flags, rec=pg_rec, ops=pg_ops)
            pg_s = (time.perf_counter() - t0) / 3
            assert np.array_equal(pg_rec, rec_dev)
            e2e["pageable_buffers"] = {"value": N / pg_s, "ms_per_step": 1e3 * pg_s,
                                       "what": "the same call from pageable caller buffers (Go slices, numpy arrays): the library "
                                               "stages them through pinned buffers with host copy threads; callers that want the "
                                               "pinned rate allocate with treat_b200_host_alloc (INTEGRATION.md)"}
            del pg_seqs, pg_rec, pg_ops
        if not np.array_equal(h_rec, rec_dev):
            badi = np.nonzero(h_rec != rec_dev)[0]
            log("MISMATCH host vs device at", len(badi), "reads, first:", badi[:8])
            for i in badi[:4]:
                log(int(i), h_rec[i], rec_dev[i])
            raise AssertionError("host-buffer and device-buffer paths disagree")

    # ---- from the bytes of a FASTA file: parse on the GPU vs the host reader -------------------
    fasta = None
    if rank == 0 and world == 1 and not args.no_fasta and not args.no_e2e:
        parts = []
        for i in range(N):
            parts.append(b">r%d_%d\n" % (i, rc[i]))
            parts.append(seqs[int(off[i]):int(off[i + 1])].tobytes())
            parts.append(b"\n")
        text = b"".join(parts)
        del parts
        h_txt = alloc(len(text))
        h_txt[:] = np.frombuffer(text, dtype=np.uint8)
        id_off, id_len, rcs = alloc(N * 8).view(np.uint64), alloc(N * 4).view(np.uint32), alloc(N * 4).view(np.uint32)

        def step_fasta():
            n_rec, ml = C.c_uint64(), C.c_uint32()
            cabi.check(cabi.lib().treat_b200_fasta_upload(eng._ctx, h_txt.ctypes.data, h_txt.size, C.byref(n_rec), C.byref(ml)), eng._ctx)
            cabi.check(cabi.lib().treat_b200_fasta_align(eng._ctx, flags, h_rec.ctypes.data, h_ops.ctypes.data, stride,
                                                         id_off.ctypes.data, id_len.ctypes.data, rcs.ctypes.data, None, None), eng._ctx)
            return n_rec.value
        h_rec[:] = np.zeros(1, dtype=cabi.RECORD_DTYPE)[0]
        assert step_fasta() == N
        if not (np.array_equal(h_rec, rec_dev) and np.array_equal(rcs, rc)):
            raise AssertionError("FASTA-ingest path disagrees with the device-buffer path")
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            step_fasta()
        fa_s = (time.perf_counter() - t0) / reps
        # the host parser on the same bytes in memory, one thread like gofasta.SimpleParser
        hh = C.c_void_p()
        t0 = time.perf_counter()
        cabi.check(cabi.lib().treat_b200_fasta_parse(h_txt.ctypes.data, h_txt.size, C.byref(hh)))
        host_s = time.perf_counter() - t0
        assert cabi.lib().treat_b200_fasta_count(hh) == N
        cabi.lib().treat_b200_fasta_free(hh)
        # streaming form: pieces of the file in flight, results delivered to a sink piece by piece
        seen = [0]

        def sink_count(_u, pp):
            seen[0] += pp.contents.n
            return 0

        def sink_copy(_u, pp):  # a consumer that keeps the records: one memmove per piece
            q = pp.contents
            C.memmove(h_rec.ctypes.data + q.first_record * 48, q.rec, q.n * 48)
            seen[0] += q.n
            return 0
        stream_ms = {}
        for name, fn in (("count", sink_count), ("copy_records", sink_copy), ("blobs", sink_count)):
            cb = cabi.FASTA_SINK(fn)
            tot = C.c_uint64()
            sflags = flags | cabi.WANT_BLOBS if name == "blobs" else flags  # `treat load`: MarshalBinary blobs built on the device

            def step_stream():
                cabi.check(cabi.lib().treat_b200_fasta_stream(eng._ctx, h_txt.ctypes.data, h_txt.size, sflags, 0, cb, None, C.byref(tot)), eng._ctx)
            seen[0] = 0
            step_stream()
            assert tot.value == N and seen[0] == N
            t0 = time.perf_counter()
            for _ in range(reps):
                step_stream()
            stream_ms[name] = 1e3 * (time.perf_counter() - t0) / reps
        if not np.array_equal(h_rec, rec_dev):
            raise AssertionError("FASTA streaming path disagrees with the device-buffer path")
        # `treat align -t -f` text for the same file: parse, alignment and Alignment.WriteTo on the GPU, text to a sink
        tbytes = [0]

        def sink_text(_u, _ptr, n, _first, _nrec):
            tbytes[0] += n
            return 0
        tcb = cabi.TEXT_SINK(sink_text)
        ttot = C.c_uint64()

        def step_text():
            tbytes[0] = 0
            cabi.check(cabi.lib().treat_b200_fasta_text(eng._ctx, h_txt.ctypes.data, h_txt.size, cabi.NO_SUMMARY, 80, tcb, None, C.byref(ttot)), eng._ctx)
        step_text()
        assert ttot.value == N
        t0 = time.perf_counter()
        for _ in range(2):
            step_text()
        text_s = (time.perf_counter() - t0) / 2
        write_text = {"value": N / text_s, "unit": UNIT, "ms_per_step": 1e3 * text_s, "text_bytes": int(tbytes[0]),
                      "d2h_gb_per_s": tbytes[0] / text_s / 1e9,
                      "api": "treat_b200_fasta_text: file bytes -> the text `treat align -t -f` prints (k_text: WriteTo on the GPU), "
                             "delivered piece by piece to a sink that only counts the bytes"}
        fasta = {"value": N / fa_s, "unit": UNIT, "ms_per_step": 1e3 * fa_s, "file_bytes": len(text), "write_text": write_text,
                 "stream": {"value": N / (stream_ms["count"] * 1e-3), "unit": UNIT, "ms_per_step": stream_ms["count"],
                            "ms_per_step_sink_copies_records": stream_ms["copy_records"],
                            "ms_per_step_with_marshal_blobs": stream_ms["blobs"],
                            "api": "treat_b200_fasta_stream: ~32 MB pieces cut at record starts, copy / index / align / deliver overlapped"},
                 "api": "treat_b200_fasta_upload + treat_b200_fasta_align: file bytes in pinned host memory -> records, "
                        "gapped ops rows, id index, read counts (parse, NewFragment, NewAlignment on the GPU)",
                 "host_parser": {"value": N / host_s, "unit": UNIT, "seconds": host_s,
                                 "what": "treat_b200_fasta_parse (one host thread) on the same bytes in memory, parse only"}}
        del text

    # ---- strong-scaling parity leg: the first 1M reads shared between the N GPUs, digests vs the oracle's ------------
    parity = None
    if not args.no_parity:
        parity = parity_leg(T, cabi, parallel, synth, eng, comm, args, t, seed, p_trunc, rank, world, barrier, max_over_ranks, alloc)

    # ---- other workload points on one GPU: BASELINE configs[4] (long genes) and a noisy RPS12 set ---------------------
    side = {}
    if rank == 0 and world == 1 and not args.no_side and args.workload == "rps12":
        side["long"] = side_workload(T, cabi, synth, eng, "long", 500_000, 10, alloc)   # 5 M variable-length reads in all
        side["noisy"] = side_workload(T, cabi, synth, eng, "noisy", 500_000, 6, alloc)
        eng.SetTemplate(tm)

    # ---- exact-duplicate collapse on the device (fastx_collapser upstream of `treat load`) -----------------------------
    collapse = None
    if rank == 0 and world == 1 and not args.no_side and args.workload == "rps12":
        cn, distinct = min(N, 500_000), 100_000
        rng = np.random.default_rng(7)
        pick = rng.integers(0, distinct, size=cn)
        lens_c = lens[pick]
        coff = np.zeros(cn + 1, dtype=np.uint64)
        coff[1:] = np.cumsum(lens_c)
        cseqs = alloc(int(coff[-1]))
        src_off = off.astype(np.int64)
        idx = np.repeat(src_off[pick] - coff[:-1].astype(np.int64), lens_c) + np.arange(int(coff[-1]), dtype=np.int64)
        cseqs[:] = seqs[idx]
        crc = rc[pick].copy()
        t0 = time.perf_counter()
        crec, _, cfirst, _ = eng.collapse_align_batch(cseqs, coff, crc, cabi.NO_SUMMARY, want_members=False)
        for _ in range(3):
            t0 = time.perf_counter()
            crec, _, cfirst, _ = eng.collapse_align_batch(cseqs, coff, crc, cabi.NO_SUMMARY, want_members=False)
        c_ms = 1e3 * (time.perf_counter() - t0)
        arec = np.zeros(cn, dtype=cabi.RECORD_DTYPE)
        for _ in range(2):
            t0 = time.perf_counter()
            eng.align_batch(cseqs, coff, crc, cabi.NO_SUMMARY, rec=arec)
        a_ms = 1e3 * (time.perf_counter() - t0)
        same = all(np.array_equal(arec[f][cfirst], crec[f]) for f in ("edit_stop", "junc_end", "junc_len", "score", "mismatches", "indel"))
        collapse = {"reads": cn, "unique": int(len(crec)), "unique_fraction": len(crec) / cn, "ms_collapse_align": c_ms, "ms_align_every_read": a_ms,
                    "records_returned": int(len(crec)), "agrees_with_uncollapsed": bool(same),
                    "api": "treat_b200_collapse_align_batch vs treat_b200_align_batch on %d reads drawn from %d distinct ones "
                           "(host buffers): duplicates collapse on the device, only the distinct sequences are aligned and returned" % (cn, distinct)}
        if not same:
            raise AssertionError("collapsed and uncollapsed records disagree")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU port gets every host core
        threads = len(all_cpus)
        sample = args.cpu_sample or min(N, 40000 * threads)
        v, g, secs = cpu_sample_rate(t, p_trunc, seed, sample, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "gcups": g,
               "sample": "first %d reads of the workload, %d threads, %.1f s; CPU port of the Go path (oracle/treat_oracle.c), "
                         "lean mode: one DP per read, record + ops" % (sample, threads, secs)}
        v1, g1, s1 = cpu_sample_rate(t, p_trunc, seed, min(sample, 20000), 1)
        cpu["single_thread"] = {"value": v1, "unit": UNIT, "gcups": g1, "sample": "first %d reads, 1 thread, %.1f s, lean mode" % (min(sample, 20000), s1)}
        # what `treat align` itself does per read: NewAlignment + WriteTo (a second DP and the text)
        from treat_b200 import synth as _synth
        O, om = oracle_template(t)
        fs = min(sample, 10000 * threads)
        fseqs, foff, frc = _synth.reads(t, fs, seed=seed, p_trunc=p_trunc)
        t0 = time.perf_counter()
        ftext, _ = O.text_batch(om, fseqs, foff, frc, prefix="r", nthreads=threads)
        fsecs = time.perf_counter() - t0
        cpu["faithful"] = {"value": fs / fsecs, "unit": UNIT, "text_bytes": int(len(ftext)),
                           "sample": "first %d reads, %d threads, %.1f s; NewFragment + NewAlignment + WriteTo per read "
                                     "(two DPs and the text, as cmd/treat/align.go:113-121)" % (fs, threads, fsecs)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            # cells of the reads' full DP matrices per second -- NOT performed work: reads whose stripped bases equal the
            # template's (or differ in one base) are finished without a DP (exact) ...
            "gcups_effective": world * cells_per_step * args.steps / (dev_ms * 1e-3) / 1e9,
            # ... and the cells the DP kernels actually filled per second in the same timed region
            "gcups_performed": world * cells_performed * args.steps / (dev_ms * 1e-3) / 1e9,
            "config": {"workload": workload_name(args, t), "l2": "inputs larger than L2 (%.0f MB raw reads per GPU per step)" % (len(seqs) / 1e6),
                       "parallelism": "reads sharded over %d GPU(s), summary all-reduce only" % world,
                       "outputs": "48-byte record per read + %d-byte 2-bit ops row per read with a gap" % stride},
            "kernels_ms": {"k_frontier": frontier_ms,
                           "k_frontier + k_pack_finish2 (NewFragment + assembly of every read without a DP)": fused_ms,
                           "k_band_* + k_align2 (reads that need a DP) = step - fused": max(dev_ms / args.steps - fused_ms, 0.0),
                           "stage calls: k_pack16": pack_ms, "stage calls: k_same_bases+k_align(listed reads)": marks_ms,
                           "DP stage on every read (k_band_* + k_align2)": align_ms,
                           "k_align2 on every read (full matrix, NO_BAND)": full_ms},
            "dp_paths": dict(path_frac, what="fraction of the reads by how they are finished: fused kernel without a DP, pure "
                                             "deletion (leftmost embedding, no DP), banded fill, full matrix (from the records)"),
            "reads_equal_to_template_bases": same_frac,
            # the same step with the DP forced on every read (k_pack16 + k_align2 on all reads), from the stage timings
            "value_dp_every_read": world * N / (dp_every_ms * 1e-3),
            "roofline": roofline, "roofline_fused": roofline_fused, "roofline_dp": roofline_dp, "roofline_pack": roofline_pack, "cpu_baseline": cpu, "e2e": e2e, "fasta_ingest": fasta,
            "text_device": text_dev, "parity": parity, "long": side.get("long"), "noisy": side.get("noisy"), "collapse": collapse,
            "nccl": {"version": cabi.lib().treat_b200_nccl_version() if world > 1 else None,
                     "collective": "ncclAllReduce(uint64, sum) of the summary counters, issued by libtreat_b200.so "
                                   "(treat_b200_comm_reduce_summary_device) once per job: behind the last of the timed steps, inside "
                                   "the timed region" if world > 1 else "none (1 GPU)",
                     "every_step": {"ms_per_step": every_ms, "value": world * N / (every_ms * 1e-3),
                                    "what": "the same steps with an all-reduce behind every one of them"} if every_ms else None},
            "gpu_launches": int(launches), "clocks": clocks, "summary_reads_all_ranks": reduced_reads,
        }
        emit_line(line)
    comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


_RESULT_FD = None


def emit_line(line):
    """The ONE JSON line of this run, on the process's real stdout (see main)."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # stdout carries exactly one JSON line.  Libraries print there too (NCCL writes its "NCCL version ..." line to stdout when
    # NCCL_DEBUG=VERSION -- NCCL_DEBUG_FILE is only honoured from WARN up --, torch.distributed has its banners): keep the real
    # stdout for the result and point file descriptor 1 at stderr for everything else.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "treat_b200", "libtreat_b200.so")
    if not os.path.exists(lib) or int(os.environ.get("LOCAL_RANK", "0")) == 0 and os.environ.get("TREAT_B200_REBUILD"):
        g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
