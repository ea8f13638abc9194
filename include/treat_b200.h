/*
 * treat_b200.h -- C ABI of libtreat_b200.so: the B200 (sm_100a) implementation of
 * ubccr/treat's read-versus-template alignment hot path.
 *
 * The reference has no FFI; the path sits behind the Go package API of package `treat`
 * (SURVEY.md 8b).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference repository).  A cgo / ctypes caller binds exactly
 * these symbols (INTEGRATION.md shows both stubs).
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns
 * TREAT_B200_OK (0) or a negative TREAT_B200_ERR_* code and never throws; the callee does
 * not retain caller pointers after return (cgo pointer rule); one context = one GPU =
 * one caller thread at a time and one call in flight (no internal locking; the device-buffer
 * entry points say what "in flight" means for them).  There is NO CPU fallback: without a usable CUDA device
 * treat_b200_create fails with TREAT_B200_ERR_NO_DEVICE.
 */
#ifndef TREAT_B200_H
#define TREAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TREAT_B200_ABI_VERSION 1

/* ---- error codes ------------------------------------------------------------------- */
#define TREAT_B200_OK 0
#define TREAT_B200_ERR_INVALID (-1)     /* bad argument */
#define TREAT_B200_ERR_CUDA (-2)        /* CUDA runtime error, see treat_b200_last_error */
#define TREAT_B200_ERR_NOMEM (-3)
#define TREAT_B200_ERR_TOO_LONG (-4)    /* template or read exceeds the kernel's shared-memory geometry */
#define TREAT_B200_ERR_NO_TEMPLATE (-5) /* align called before set_template */
#define TREAT_B200_ERR_NO_DEVICE (-6)   /* no CUDA device / not sm_100: there is no CPU path */
#define TREAT_B200_ERR_TEMPLATE (-7)    /* template validation failed (template.go:97-113) */
#define TREAT_B200_ERR_IO (-8)

/* const.go:20-23 */
#define TREAT_B200_FORWARD 1
#define TREAT_B200_REVERSE (-1)

/* ---- flags of the align calls ------------------------------------------------------- */
#define TREAT_B200_EXCLUDE_SNPS 0x1u /* NewAlignment(..., excludeSnps=true), storage.go:756 */
#define TREAT_B200_WANT_OPS 0x2u     /* also return the traceback (gapped strings in 2-bit form) */
#define TREAT_B200_NO_SUMMARY 0x4u   /* do not accumulate the per-context summary histograms */
#define TREAT_B200_ONE_READ_PER_WARP 0x8u /* use the one-read-per-warp kernel even where the paired kernel applies */
#define TREAT_B200_FULL_TRACEBACK_STORE 0x10u /* keep every direction word (global scratch) instead of the band */
#define TREAT_B200_DP_EVERY_READ 0x20u /* run the DP even for reads whose stripped bases equal the template's (results are the same) */
#define TREAT_B200_OPS_GAPPED_ONLY 0x80u /* with WANT_OPS: write the ops row only of reads whose alignment has a gap
                                            (record.indel != 0); the row of any other read would be all zero (m = n
                                            match columns) and is unspecified (left untouched, or zeroed), so D2H traffic
                                            can shrink to the gapped reads */
#define TREAT_B200_WANT_BLOBS 0x200u /* treat_b200_fasta_stream only: also hand the sink every record's MarshalBinary blob (`treat load`) */
#define TREAT_B200_HEATMAP 0x100u /* also accumulate the ESS x junction-length heat map (treat_b200_get_heatmap) */
#define TREAT_B200_SKIP_DP 0x400u /* measurement only (bench.py times the fused kernel alone with it): the DP launch for the reads
                                     that need one is skipped, so their records are NOT written */
#define TREAT_B200_NO_BAND 0x800u /* do not run the banded / pure-deletion kernel in front of the full-matrix DP (results are the
                                     same; tests compare both ways) */
#define TREAT_B200_NO_FRONTIER 0x1000u /* do not run k_frontier (reads that are the template edited up to a frontier, finished from
                                          two string compares on the raw characters) in front of the fused kernels (results are
                                          the same; tests compare both ways) */
#define TREAT_B200_FRONTIER_ONLY 0x2000u /* measurement only (bench.py times k_frontier alone with it): nothing runs behind k_frontier,
                                            so the records of the reads it leaves to the other kernels are NOT written */
#define TREAT_B200_USE_PACK_MARKS 0x40u /* treat_b200_align_packed_device only: the reads were packed by treat_b200_pack_device
                                           under the template that is current now, so reads it marked as equal to the template
                                           skip the DP; without this flag the stage call runs the DP for every read */

/* ---- per-read status bits ----------------------------------------------------------- */
#define TREAT_B200_ST_REF_PANIC 0x1u  /* reference indexes frag.EditSite out of range here (align.go:250-257) */
#define TREAT_B200_ST_WIDE_BASE 0x2u  /* read holds a base outside the 3-letter packed alphabet */
#define TREAT_B200_ST_SATURATED 0x4u  /* an edit-base run >= 255: batch was re-run on the wide path */

/* ---- traceback ops: 2 bits per alignment column, 4 columns per byte, low bits first -- */
#define TREAT_B200_OP_MATCH 0u /* template base over read base       (nwalgo pointer NW)   */
#define TREAT_B200_OP_DEL 1u   /* template base over '-'   ("Up",   consumes template)     */
#define TREAT_B200_OP_INS 2u   /* '-' over read base       ("Left", consumes read)         */

/*
 * Per-read result.  Bytes 0..23 are the fields of Alignment.MarshalBinary
 * (align.go:316-335) in the same order, host endianness; Norm is always 0 on this path
 * (fragment.go:120 leaves Fragment.Norm zero).  JuncSeq is returned as a window into the
 * upper-cased raw read: JuncSeq == upper(seq[junc_seq_off : junc_seq_off + junc_seq_len]).
 * EditStop/JuncStart/JuncEnd already include Template.EditOffset (align.go:274-276).
 */
typedef struct treat_b200_record {
    int32_t edit_stop;
    int32_t junc_start;
    int32_t junc_end;
    int32_t junc_len;
    uint32_t read_count;
    uint8_t has_mutation;
    uint8_t alt_editing;
    uint8_t mismatches;
    uint8_t indel;
    int32_t score;         /* nwalgo.Align's third return value */
    uint32_t junc_seq_off;
    uint32_t junc_seq_len;
    uint16_t aln_len;      /* alignment columns = len(aln1) */
    uint16_t n_bases;      /* len(Fragment.Bases) */
    uint8_t status;        /* TREAT_B200_ST_* */
    uint8_t reserved[3];
    uint32_t raw_len;      /* len(seq) */
} treat_b200_record; /* 48 bytes */

/*
 * Summary counters accumulated on the device over every align call since the last reset
 * (the only data that crosses GPUs: SURVEY.md 8e).  Layout of the uint64 array returned by
 * treat_b200_get_summary, with B = treat_b200_summary_bins(ctx) = Template.Len() + 2:
 *   [0..6]   Total, Std, NonStd, Indels, Snps, SingleMismatch, DoubleMismatch weighted by
 *            ReadCount (cmd/treat/stats.go:140-187, COUNT_FRAG)
 *   [7..13]  the same seven counters counting each read once (COUNT_UNIQUE)
 *   [14]     reads aligned, [15] DP cells (sum of m*n)
 *   [16 .. 16+B)        histogram of EditStop  (bin = EditStop - EditOffset + 2)
 *   [16+B .. 16+2B)     histogram of JuncLen   (bin = JuncLen)
 *   [16+2B .. 16+3B)    histogram of JuncEnd   (bin = JuncEnd - EditOffset + 2)
 * Histograms are weighted by ReadCount and skip reads with EditStop == Template.EditStop
 * && JuncLen == 0, as cmd/treat/handlers.go:203-215 does.
 */
#define TREAT_B200_SUMMARY_HEADER 16

typedef struct treat_b200_ctx treat_b200_ctx;
typedef struct treat_b200_template treat_b200_template;

/* ---- context ---------------------------------------------------------------------- */
int treat_b200_abi_version(void);
const char *treat_b200_strerror(int code);
const char *treat_b200_last_error(const treat_b200_ctx *ctx);
/* device = CUDA ordinal (one process per GPU: pass LOCAL_RANK) */
int treat_b200_create(int device, treat_b200_ctx **out);
void treat_b200_destroy(treat_b200_ctx *ctx);
/* Pinned host memory for the batch buffers.  Any host pointer works: treat_b200_align_batch moves pageable buffers (Go
 * slices, numpy arrays) through pinned staging buffers of its own with a few host copy threads (TREAT_B200_COPY_THREADS,
 * default 5), because the driver would stage such copies itself and block the caller for their whole duration.  Measured
 * on one B200 for 1 M RPS12 reads: 6.2 ms from / to pinned buffers, 6.4 ms with pageable outputs, 13.7 ms with pageable
 * inputs and outputs (host memcpy bound; 35.7 ms without the staging).  For the full rate allocate the input buffers once
 * with treat_b200_host_alloc and reuse them, or pin long-lived buffers of your own in place with treat_b200_host_register
 * (page-locking costs milliseconds per 100 MB: do it once, not per call). */
int treat_b200_host_alloc(void **out, size_t bytes);
int treat_b200_host_register(void *p, size_t bytes);
int treat_b200_host_unregister(void *p);
void treat_b200_host_free(void *p);

/* ---- template: replaces treat.NewTemplate / Template.SetOffset state (template.go:96-161) ----
 * bases: m non-edit bases (upper case); edit_site: K rows of m+1 counts, row 0 = FE,
 * row 1 = PE, rows 2.. = alternative templates; alt_start/alt_end: n_alt = K-2 regions
 * AFTER SetOffset (template.go:157-160); tmpl_edit_stop: Template.EditStop after SetOffset.
 * Validation of the FASTA-level invariants stays with the caller's NewTemplate. */
int treat_b200_set_template(treat_b200_ctx *ctx, const uint8_t *bases, int32_t m, const uint32_t *edit_site,
                            int32_t K, const int32_t *alt_start, const int32_t *alt_end, int32_t n_alt,
                            uint32_t edit_offset, int32_t tmpl_edit_stop, uint8_t edit_base);
int treat_b200_use_template(treat_b200_ctx *ctx, const treat_b200_template *tmpl);

/* ---- the hot path: replaces the per-read loop
 *        frag := treat.NewFragment(rec.Id, rec.Seq, FORWARD, base)   (fragment.go:91)
 *        aln  := treat.NewAlignment(frag, tmpl, excludeSnps)         (align.go:237)
 *      of cmd/treat/align.go:114-120 and cmd/treat/storage.go:738-785 for a whole batch.
 * seqs/seq_off: concatenated raw reads (any case), read i = seqs[seq_off[i] .. seq_off[i+1]);
 * read_count: Fragment.ReadCount per read (treat_b200_parse_merge_count) or NULL for 1;
 * rec: N records, caller owned; ops (may be NULL): ops_stride bytes per read, see
 * treat_b200_ops_stride; results are in input order.  Of every ops row only the leading
 * align16(ceil((m + max n of the chunk) / 4)) bytes are written (zero past the alignment); pass a
 * zeroed buffer if the tail of the rows is to be compared. */
int treat_b200_align_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off,
                           const uint32_t *read_count, uint64_t N, uint32_t flags, treat_b200_record *rec,
                           uint8_t *ops, uint32_t ops_stride);
/* same with every buffer already resident in device memory of ctx's GPU; max_len = max_i len(read i).  The kernels are
 * launched on `stream` (a cudaStream_t, NULL = ctx stream) and the results are ready when the work queued on that stream
 * is.  The call is NOT fully asynchronous: it waits once for `stream` (an 8-byte read-back of the longest stripped read,
 * which sizes the DP launch).  Contract: ONE call in flight per context.  The packed reads, the DP list and the scratch of
 * a call live in the context, so a second call on the same context -- from another stream or thread, including
 * treat_b200_pack_device / treat_b200_align_packed_device -- must not start before the work of the first has finished on
 * its stream.  For concurrent streams use one context per stream. */
int treat_b200_align_batch_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off,
                                  const uint32_t *d_read_count, uint64_t N, uint32_t max_len, uint32_t flags,
                                  treat_b200_record *d_rec, uint8_t *d_ops, uint32_t ops_stride, void *stream);
/* The same with exact duplicates collapsed on the device first -- the step the TREAT workflow runs upstream of `treat load`
 * (fastx_collapser: identical sequences become one record whose id carries the count that parseMergeCount reads back,
 * fragment.go:43-55,75-89).  Reads whose upper-cased sequences are equal give the same Fragment (fragment.go:91-121), hence
 * the same Alignment: only the first of them is aligned, with ReadCount = the sum of the members' counts (uint32 like
 * Fragment.ReadCount).  Outputs, for the *n_unique distinct sequences in the order of their first occurrence:
 *   rec[u], ops row u        the alignment of unique u (buffers sized for N: the caller cannot know n_unique before)
 *   first[u]                 index in the input of the first read with that sequence (junc_seq_off refers to it)
 *   member[r] (may be NULL)  for every input read the number u of its sequence
 * The summary counters count every distinct sequence once ("unique", cmd/treat/stats.go --unique) and weigh it with the summed
 * ReadCount -- what a load of the collapsed file gives.  Exact: candidates are found by a 64-bit hash and then compared byte
 * for byte; a hash collision makes the library repeat the collapse with another seed. */
int treat_b200_collapse_align_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off,
                                    const uint32_t *read_count, uint64_t N, uint32_t flags, treat_b200_record *rec,
                                    uint8_t *ops, uint32_t ops_stride, uint32_t *first, uint32_t *member, uint64_t *n_unique);
/* bytes per read needed for the ops of reads up to max_len against the current template */
uint32_t treat_b200_ops_stride(const treat_b200_ctx *ctx, uint32_t max_len);
int treat_b200_synchronize(treat_b200_ctx *ctx);

/* ---- stage-level entry points (device buffers): K1 = NewFragment for a batch, K2 = align --- */
typedef struct treat_b200_packed {
    uint64_t N;
    uint32_t max_len;        /* longest raw read */
    uint32_t wide;           /* 0: 2-bit bases + u8 counts, 1: u8 bases + u32 counts */
    uint32_t bases_stride;   /* bytes per read, multiple of 16 */
    uint32_t counts_stride;  /* bytes per read, multiple of 16 */
    uint16_t *d_n;           /* [N] len(Fragment.Bases) */
    uint8_t *d_flags;        /* [N] TREAT_B200_ST_WIDE_BASE | TREAT_B200_ST_SATURATED | 0x80 (internal mark: bases equal the template) */
    uint8_t *d_bases;        /* [N][bases_stride] */
    uint8_t *d_counts;       /* [N][counts_stride]  Fragment.EditSite, n+1 entries */
    uint32_t *d_raw_len;     /* [N] */
} treat_b200_packed;
/* sizes the strides for (N, max_len); the caller allocates the five arrays */
int treat_b200_packed_layout(const treat_b200_ctx *ctx, uint64_t N, uint32_t max_len, uint32_t wide,
                             treat_b200_packed *out);
int treat_b200_pack_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off,
                           const treat_b200_packed *dst, void *stream);
int treat_b200_align_packed_device(treat_b200_ctx *ctx, const treat_b200_packed *src,
                                   const uint32_t *d_read_count, uint32_t flags, treat_b200_record *d_rec,
                                   uint8_t *d_ops, uint32_t ops_stride, void *stream);
/* Traceback tie order: which pointer wins when several of Up (template base over '-'), Left ('-' over read base) and
 * Diagonal are optimal in a cell.  The DP itself lives in the un-vendored module github.com/aebruno/nwalgo (align.go:129,
 * 338,393); its published source takes Up, then Left, then Diagonal (the default here).  The reference's own goldens pin
 * only "Up beats Diagonal" (SURVEY.md 8c), which these three orders satisfy; should nwalgo turn out to differ, this call is
 * the whole fix.  Scores, and every read whose optimal alignment is unique, do not depend on it. */
#define TREAT_B200_TIE_ULD 0 /* Up > Left > Diagonal (nwalgo as published; default) */
#define TREAT_B200_TIE_UDL 1 /* Up > Diagonal > Left */
#define TREAT_B200_TIE_LUD 2 /* Left > Up > Diagonal */
int treat_b200_set_tie_order(treat_b200_ctx *ctx, int32_t order);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
uint64_t treat_b200_launch_count(const treat_b200_ctx *ctx);
/* tuning: the banded fill in front of the full-matrix DP runs only when a launch has at least `reads` reads to align
 * (default 16,384; environment TREAT_B200_BAND_MIN; 0 = always).  Below that its fixed latency exceeds what it saves, so
 * only truncated reads (pure deletions, 16 or more bases short) take a shortcut.  Results do not depend on it. */
int treat_b200_set_band_min_reads(treat_b200_ctx *ctx, uint32_t reads);
/* tuning: rows of traceback words kept per lane in shared memory (0 = default 64); any value is
 * exact, reads whose optimal path could leave the band are re-filled into global scratch */
int treat_b200_set_band_rows(treat_b200_ctx *ctx, uint32_t rows);
/* Geometry of the DP launch for reads of up to max_n stripped bases against the current template (what SURVEY.md 8d asks to
 * be reported for the long-gene configuration: shared memory per CTA, resident warps, traceback storage). */
typedef struct treat_b200_geometry {
    uint32_t cpl;                    /* template columns per lane of the wavefront */
    uint32_t strips;                 /* column strips (1 unless the template has more than 480 bases) */
    uint32_t reads_per_warp;         /* 2: k_align2 (16-bit halves), 1: k_align / k_align_strips */
    uint32_t warps_per_cta;          /* one persistent CTA per SM */
    uint32_t smem_cta_bytes;         /* dynamic shared memory of the CTA */
    uint32_t smem_template_bytes;    /* of which: template, edit-site counts, scoring profile, summary bins (CTA-wide) */
    uint32_t smem_warp_bytes;        /* of which per warp: TMA stage slots, read bases, band window, ops, site map */
    uint32_t band_rows;              /* direction-word rows kept per lane in shared memory */
    uint32_t band_diagonals;         /* B: a read is certified when (m - S)/2 <= B and (n - S)/2 <= B */
    uint32_t scratch_bytes_per_warp; /* global (L2-resident) scratch a full fill writes */
    uint32_t regs_per_thread;        /* of the kernel that would be launched */
    uint32_t occupancy_warps_per_sm; /* resident warps = warps_per_cta (one CTA per SM) */
} treat_b200_geometry;
int treat_b200_align_geometry(treat_b200_ctx *ctx, uint32_t max_n, uint32_t flags, treat_b200_geometry *out);
/* which kernel finished the reads that needed a DP since the last reset: out3[0] pure deletions (leftmost embedding, no
 * DP), out3[1] banded fill with an exactness certificate, out3[2] left for the full-matrix kernels (bench / tests).
 * Synchronises the context's own streams; work queued on a caller's stream must have finished before the call. */
int treat_b200_dp_stats(treat_b200_ctx *ctx, uint64_t *out3, int32_t reset);
/* bounds-check flags of the kernels' index computations: always 0 unless the library was built with
 * -DTREAT_B200_CHECKS (*checks_compiled says which); a substitute for compute-sanitizer on pools without it */
int treat_b200_debug_flags(treat_b200_ctx *ctx, uint32_t *flags_out, int32_t *checks_compiled);

/* ---- summaries: device-resident uint64 counters --------------------- */
uint32_t treat_b200_summary_bins(const treat_b200_ctx *ctx);
uint64_t treat_b200_summary_len(const treat_b200_ctx *ctx); /* TREAT_B200_SUMMARY_HEADER + 3*bins */
int treat_b200_get_summary(treat_b200_ctx *ctx, uint64_t *out, uint64_t len);
int treat_b200_reset_summary(treat_b200_ctx *ctx);
/* device pointer of the counters, for an NCCL all-reduce issued by the caller */
int treat_b200_summary_device_ptr(treat_b200_ctx *ctx, uint64_t **d_ptr);
/* ESS x junction-length heat map of the web UI (cmd/treat/handlers.go:578-589), accumulated on the device by align calls
 * that carry TREAT_B200_HEATMAP: heat[(EditStop - EditOffset) * dim + JuncLen] += ReadCount for records with
 * EditStop >= EditOffset, dim = Template.Len().  (The reference adds Norm, the ReadCount-derived weight of `treat norm`;
 * it blanks the cell (Template.EditStop, 0) when it renders -- left to the consumer.)  Cleared by treat_b200_reset_summary;
 * across GPUs: all-reduce dim*dim uint64 from treat_b200_heatmap_device_ptr. */
uint32_t treat_b200_heatmap_dim(const treat_b200_ctx *ctx);
int treat_b200_get_heatmap(treat_b200_ctx *ctx, uint64_t *out, uint64_t len);
int treat_b200_heatmap_device_ptr(treat_b200_ctx *ctx, uint64_t **d_ptr);

/* ---- roofline helper: dependent-free INT32 add/max throughput of this GPU ---------- */
int treat_b200_measure_int32_peak(treat_b200_ctx *ctx, double *lane_ops_per_s, double *sm_mhz_effective);
/* Floor of the host-buffer calls: `reps` plain cudaMemcpyAsync H2D copies of host[0, bytes) (pinned memory for the DMA rate)
 * into a device buffer of the context, timed with CUDA events on the context's stream: *ms_per_copy.  Run it on all ranks
 * at once to see what the host's PCIe / memory path gives N GPUs together (bench.py: e2e.h2d_floor_ms). */
int treat_b200_measure_h2d(treat_b200_ctx *ctx, const void *host, uint64_t bytes, int32_t reps, double *ms_per_copy);
/* The same with the results going the other way at the same time: per repetition one H2D copy of in_bytes and, on a second
 * stream, one D2H copy of out_bytes into host_out; *ms_per_step = time until both are done.  PCIe is full duplex but the
 * host's memory / IO path is shared by both directions. */
int treat_b200_measure_transfer(treat_b200_ctx *ctx, const void *host_in, uint64_t in_bytes, void *host_out, uint64_t out_bytes,
                                int32_t reps, double *ms_per_step);

/* =====================================================================================
 * Several GPUs.  Reads shard over the GPUs with no data-path exchange; the only collective is one
 * ncclAllReduce(ncclUint64, ncclSum) over the summary counters (and the heat map) -- what `treat stats`
 * (cmd/treat/stats.go:140-187) and the histogram handlers (cmd/treat/handlers.go:203-215) later read.  NCCL is loaded
 * at run time (dlopen of libnccl.so.2) the first time a communicator is made, so a single-GPU caller needs no NCCL.
 *
 * Two forms:
 *  (1) one process, n GPUs: treat_b200_group_* -- the library owns one context + one host worker thread per device and the
 *      ncclCommInitAll communicators; a batch is cut into n contiguous read ranges (range i = [N*i/n, N*(i+1)/n)), so the
 *      records, ops rows and their order are those of the one-GPU call.  This is the form a Go host uses.
 *  (2) one process per GPU: treat_b200_comm_* on top of a plain context; the caller hands the 128-byte NCCL id of rank 0
 *      to every rank by whatever means it has (a file, MPI, torch.distributed, a socket).
 * ===================================================================================== */
typedef struct treat_b200_group treat_b200_group;
/* devices: n CUDA ordinals (NULL: 0..n-1).  n = 1 works without NCCL. */
int treat_b200_group_create(const int *devices, int32_t n, treat_b200_group **out);
void treat_b200_group_destroy(treat_b200_group *g);
int32_t treat_b200_group_size(const treat_b200_group *g);
treat_b200_ctx *treat_b200_group_ctx(treat_b200_group *g, int32_t i); /* device i's context (owned by the group) */
const char *treat_b200_group_last_error(const treat_b200_group *g);
/* treat_b200_set_template / treat_b200_use_template on every device of the group */
int treat_b200_group_set_template(treat_b200_group *g, const uint8_t *bases, int32_t m, const uint32_t *edit_site, int32_t K,
                                  const int32_t *alt_start, const int32_t *alt_end, int32_t n_alt, uint32_t edit_offset,
                                  int32_t tmpl_edit_stop, uint8_t edit_base);
int treat_b200_group_use_template(treat_b200_group *g, const treat_b200_template *tmpl);
/* treat_b200_align_batch (same arguments, same results in the same order) with the reads shared between the group's GPUs:
 * the per-read loop of cmd/treat/align.go:114-120 / cmd/treat/storage.go:738-785 for a whole batch on n devices. */
int treat_b200_group_align_batch(treat_b200_group *g, const uint8_t *seqs, const uint64_t *seq_off, const uint32_t *read_count,
                                 uint64_t N, uint32_t flags, treat_b200_record *rec, uint8_t *ops, uint32_t ops_stride);
uint32_t treat_b200_group_ops_stride(const treat_b200_group *g, uint32_t max_len);
/* All-GPU summary counters (layout of treat_b200_get_summary): one grouped ncclAllReduce over the devices' counters, result
 * read from device `from_rank` (every device holds the same reduced array; pass 0).  The devices' own counters stay as
 * they are.  treat_b200_group_heatmap: the same for the ESS x junction-length heat map (len = dim * dim). */
int treat_b200_group_summary(treat_b200_group *g, int32_t from_rank, uint64_t *out, uint64_t len);
int treat_b200_group_heatmap(treat_b200_group *g, int32_t from_rank, uint64_t *out, uint64_t len);
int treat_b200_group_reset_summary(treat_b200_group *g);
uint64_t treat_b200_group_launch_count(const treat_b200_group *g); /* kernels launched on all devices */

typedef struct treat_b200_comm treat_b200_comm;
#define TREAT_B200_COMM_ID_BYTES 128
int treat_b200_comm_unique_id(uint8_t *id, size_t len); /* rank 0: ncclGetUniqueId; len >= TREAT_B200_COMM_ID_BYTES */
int treat_b200_comm_create(treat_b200_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t *id, size_t len, treat_b200_comm **out);
void treat_b200_comm_destroy(treat_b200_comm *c);
/* Sum of every rank's summary counters / heat map: blocking, result in host memory; the context's own counters stay as they
 * are.  Every rank must call it (it is a collective). */
int treat_b200_comm_reduce_summary(treat_b200_comm *c, uint64_t *out, uint64_t len);
int treat_b200_comm_reduce_heatmap(treat_b200_comm *c, uint64_t *out, uint64_t len);
/* The same without leaving the device: d_out[treat_b200_summary_len] = sum over ranks of the counters as they are once the
 * work queued on `stream` (a cudaStream_t; NULL = the context's stream) so far has run.  A snapshot of the counters is
 * taken on `stream`; the ncclAllReduce itself runs on the communicator's own stream, so the caller's next align call does
 * not wait for the other ranks.  d_out is complete for work queued on a stream after treat_b200_comm_join(c, that stream)
 * (or after treat_b200_comm_reduce_summary / cudaDeviceSynchronize). */
int treat_b200_comm_reduce_summary_device(treat_b200_comm *c, uint64_t *d_out, void *stream);
int treat_b200_comm_join(treat_b200_comm *c, void *stream);
const char *treat_b200_comm_last_error(const treat_b200_comm *c);
/* version of the NCCL library in use (e.g. 22809), 0 when none could be loaded */
int32_t treat_b200_nccl_version(void);

/* =====================================================================================
 * Host-side mirror of the reference's Go API (no GPU work).  Implemented in C++ under
 * treat_b200/csrc/host and exported here so ctypes/cgo callers can use it.
 * ===================================================================================== */

/* fragment.go:75-89 parseMergeCount */
uint32_t treat_b200_parse_merge_count(const char *id, size_t len);
/* fragment.go:91-121 NewFragment on the host (templates, SimpleAlign).  bases_out needs
 * len bytes, edit_site_out len+1 entries. */
int treat_b200_new_fragment(const uint8_t *seq, size_t len, int orientation, uint8_t base, uint8_t *bases_out,
                            uint32_t *edit_site_out, size_t *n_out);

/* template.go:51-94 NewTemplateFromFasta (+ SetOffset, template.go:153-161 when offset != 0) */
int treat_b200_template_from_fasta(const char *path, int orientation, uint8_t base, int32_t offset,
                                   treat_b200_template **out, char *err, size_t errlen);
/* template.go:96-151 NewTemplate from raw template sequences (seqs[0]=FE, seqs[1]=PE, rest alt) */
int treat_b200_template_new(const char *const *seqs, int32_t n_seqs, const int32_t *alt_start,
                            const int32_t *alt_end, int32_t n_regions, int orientation, uint8_t base,
                            treat_b200_template **out, char *err, size_t errlen);
void treat_b200_template_set_offset(treat_b200_template *t, int32_t offset); /* template.go:153-161 */
void treat_b200_template_free(treat_b200_template *t);
int32_t treat_b200_template_size(const treat_b200_template *t);      /* Size()  template.go:163 */
int32_t treat_b200_template_len(const treat_b200_template *t);       /* Len()   template.go:167 */
int32_t treat_b200_template_edit_stop(const treat_b200_template *t);
uint32_t treat_b200_template_edit_offset(const treat_b200_template *t);
uint8_t treat_b200_template_edit_base(const treat_b200_template *t);
const uint8_t *treat_b200_template_bases(const treat_b200_template *t);
const uint32_t *treat_b200_template_edit_site(const treat_b200_template *t, int32_t k);
const uint32_t *treat_b200_template_base_index(const treat_b200_template *t);
int32_t treat_b200_template_n_alt(const treat_b200_template *t);
int32_t treat_b200_template_alt_start(const treat_b200_template *t, int32_t i);
int32_t treat_b200_template_alt_end(const treat_b200_template *t, int32_t i);
uint32_t treat_b200_template_max(const treat_b200_template *t, int32_t i); /* Max(i) template.go:186 */

/* FASTA reader with gofasta.SimpleParser's record semantics (cmd/treat/align.go:114).
 * UNPINNED (like the traceback tie order): gofasta (github.com/aebruno/gofasta@e776ef625791, go.mod:6) is not vendored in the
 * reference, and none of its tests says how header / sequence lines are trimmed.  Assumed here, by the host reader, the
 * device parser and the test oracle alike: Id = the header line behind '>' with strings.TrimSpace applied; sequence lines
 * are concatenated with only the line break ('\n', or '\r\n') removed -- a trailing blank or tab inside a sequence line
 * stays in the sequence and counts as a base outside the alphabet; a last line without '\n' is kept; bytes before the
 * first '>' are ignored.  If gofasta trims sequence lines, such files would differ. */
typedef struct treat_b200_fasta treat_b200_fasta;
int treat_b200_fasta_read(const char *path, treat_b200_fasta **out);
int treat_b200_fasta_parse(const uint8_t *bytes, uint64_t len, treat_b200_fasta **out); /* the same on bytes in memory */
/* Cut the bytes of a FASTA file into `parts` slices of about equal size that start at record starts ('>' opening a line):
 * offsets[parts + 1], offsets[0] = 0, offsets[parts] = len.  Parsing the slices one after the other gives the records of
 * the whole file in order -- this is how a file is shared between GPUs (one slice per rank) and how
 * treat_b200_fasta_stream cuts its pieces. */
int treat_b200_fasta_split(const uint8_t *fasta, uint64_t len, uint32_t parts, uint64_t *offsets);
void treat_b200_fasta_free(treat_b200_fasta *f);
uint64_t treat_b200_fasta_count(const treat_b200_fasta *f);
const char *treat_b200_fasta_id(const treat_b200_fasta *f, uint64_t i);
const uint8_t *treat_b200_fasta_seqs(const treat_b200_fasta *f);     /* concatenated */
const uint64_t *treat_b200_fasta_offsets(const treat_b200_fasta *f); /* count+1 */
const uint32_t *treat_b200_fasta_read_counts(const treat_b200_fasta *f);

/* FASTA ingest on the device: the record loop of gofasta.SimpleParser (cmd/treat/align.go:114, storage.go:738) with
 * parseMergeCount (fragment.go:75-89), NewFragment and NewAlignment, straight from the bytes of the file.
 * treat_b200_fasta_upload copies the file to the GPU (fasta may be pageable; pinned is faster), finds its records
 * and returns their number and the longest sequence (size the ops rows with treat_b200_ops_stride(ctx, max_seq_len)).
 * The file stays resident until the next upload, so it can be aligned against several templates.
 * treat_b200_fasta_align aligns every record against the current template: rec[n_records] in file order; optional
 * outputs: ops rows, each record's Id as (offset, length) into the file bytes (strings.TrimSpace applied), ReadCount.
 * Record semantics are those of treat_b200_fasta_read (which runs on the host). */
int treat_b200_fasta_upload(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint64_t *n_records, uint32_t *max_seq_len);
int treat_b200_fasta_align(treat_b200_ctx *ctx, uint32_t flags, treat_b200_record *rec, uint8_t *ops, uint32_t ops_stride,
                           uint64_t *id_off, uint32_t *id_len, uint32_t *read_count, uint64_t *seq_off, uint32_t *seq_len);
/* (seq_off, seq_len: where each record's sequence starts in the file and how many bytes it has without line breaks;
 * record.junc_seq_off counts from there when the record has a single sequence line.)
 *
 * Streaming form: one call, nothing to size in advance.  The bytes are cut at record starts into pieces of ~32 MB; copy,
 * index, alignment and delivery of up to six pieces overlap; every finished piece is handed to `sink` in file order
 * (pointers are valid during the call only; a non-zero return stops the run; the sink must not call align functions of the
 * same context).  This is the shape of the reference's loop (storage.go:738-785): the sink is where MarshalBinary + Put go. */
typedef struct treat_b200_fasta_piece {
    uint64_t first_record;         /* number of the piece's first record in the file */
    uint64_t n;                    /* records in this piece */
    const treat_b200_record *rec;  /* [n] */
    const uint64_t *id_off;        /* [n] Id = file[id_off, id_off + id_len) */
    const uint64_t *seq_off;       /* [n] first sequence byte in the file */
    const uint32_t *id_len, *read_count, *seq_len; /* [n] */
    const uint8_t *ops;            /* WANT_OPS: n_ops_rows rows of ops_stride bytes, else NULL */
    const uint32_t *ops_index;     /* OPS_GAPPED_ONLY: row q belongs to record ops_index[q] of the piece; NULL: row q = record q */
    uint64_t n_ops_rows;
    uint32_t ops_stride;
    uint32_t seq_in_place;         /* 1: every sequence of the piece is file[seq_off, seq_off + seq_len) (no line break inside) */
    /* TREAT_B200_WANT_BLOBS: Alignment.MarshalBinary of every record (align.go:316-335), built on the device: blob i =
     * blob[blob_off[i], blob_off[i+1]) -- what cmd/treat/storage.go:758-770 puts into the alignment bucket; else NULL */
    const uint8_t *blob;
    const uint64_t *blob_off;      /* [n + 1] */
} treat_b200_fasta_piece;
typedef int (*treat_b200_fasta_sink)(void *user, const treat_b200_fasta_piece *piece);
/* ops_stride: bytes per ops row when every record gets one (0: chosen per piece from its longest read, see piece->ops_stride) */
int treat_b200_fasta_stream(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint32_t flags, uint32_t ops_stride,
                            treat_b200_fasta_sink sink, void *user, uint64_t *n_records);

/* Alignment.WriteTo (align.go:388-520) for a batch, on the device: the text block `treat align` prints per read (header
 * with name / JSS / ESS / JES / Junc Len / Alt, then the gapped FE, PE, A*, RD rows wrapped at tw - 4 columns; tw <= 0
 * means 80), built from the raw reads, the records and the ops rows of an align call made with TREAT_B200_WANT_OPS
 * (one row per read: not OPS_GAPPED_ONLY).  All pointers are device pointers; names are concatenated like the reads
 * (name r = d_names[d_name_off[r], d_name_off[r+1])).  The call computes d_text_off[N + 1] (text of read r =
 * d_text[d_text_off[r], d_text_off[r+1])), returns the total in *text_bytes after one stream synchronisation, and
 * writes the text when d_text is not NULL and text_cap is large enough (d_text = NULL: sizing call). */
int treat_b200_text_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off, const uint8_t *d_names,
                           const uint64_t *d_name_off, const treat_b200_record *d_rec, const uint8_t *d_ops,
                           uint32_t ops_stride, uint64_t N, uint32_t max_len, int32_t tw, uint64_t *d_text_off,
                           uint8_t *d_text, uint64_t text_cap, uint64_t *text_bytes, void *stream);
/* The loop of cmd/treat/align.go:114-120 for records the caller already holds (host buffers; names concatenated like the
 * reads, name r = names[name_off[r], name_off[r+1])): NewFragment, NewAlignment and WriteTo for reads [0, N) on the device,
 * the text handed to `sink` chunk by chunk (32,768 reads) in read order; first_record counts reads.  The typedef of the
 * sink is below.  flags: TREAT_B200_EXCLUDE_SNPS, TREAT_B200_NO_SUMMARY, TREAT_B200_HEATMAP. */
typedef int (*treat_b200_text_sink)(void *user, const uint8_t *text, uint64_t len, uint64_t first_record, uint64_t n_records);
int treat_b200_text_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off, const uint8_t *names,
                          const uint64_t *name_off, uint64_t N, uint32_t flags, int32_t tw, treat_b200_text_sink sink,
                          void *user);
/* `treat align -t templates.fa -f reads.fa` (cmd/treat/align.go:96-121) in one streaming call: FASTA bytes in, the exact
 * stdout text out.  The file is cut at record starts into pieces (~8 MB); every piece is indexed, aligned and formatted on
 * the device and its text handed to `sink` in file order while the next piece is in flight (text is valid during the
 * callback only; a non-zero return stops the run).  flags: TREAT_B200_EXCLUDE_SNPS, TREAT_B200_NO_SUMMARY, TREAT_B200_HEATMAP. */
int treat_b200_fasta_text(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint32_t flags, int32_t tw,
                          treat_b200_text_sink sink, void *user, uint64_t *n_records);

/* align.go:388-520 Alignment.WriteTo from a device result: text for one read.  Returns the
 * number of bytes needed; writes at most cap bytes to out (call twice or size generously). */
int64_t treat_b200_format_alignment(const treat_b200_template *t, const treat_b200_record *rec,
                                    const uint8_t *ops, const char *name, const uint8_t *seq, size_t seq_len,
                                    int32_t tw, char *out, size_t cap);
/* align.go:337-386 Alignment.SimpleAlign: runs f2 against f1 on the GPU (f1 as a one-row
 * template) and assembles the two gapped strings.  out1/out2 need cap bytes each. */
int64_t treat_b200_simple_align(treat_b200_ctx *ctx, const uint8_t *s1, size_t l1, const uint8_t *s2, size_t l2,
                                uint8_t base, char *out1, char *out2, size_t cap);
/* align.go:316-335 Alignment.MarshalBinary: 36-byte big-endian header + JuncSeq */
int64_t treat_b200_marshal_record(const treat_b200_record *rec, const uint8_t *seq, size_t seq_len, uint8_t *out,
                                  size_t cap);
/* The same on the device (`treat load`, cmd/treat/storage.go:758-770, with reads and records resident in HBM, e.g. after
 * treat_b200_align_batch_device): d_blob_off[N + 1] and the blobs back to back in d_blob; *blob_bytes after one stream
 * synchronisation; d_blob = NULL: sizing call. */
int treat_b200_marshal_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off, const treat_b200_record *d_rec,
                              uint64_t N, uint64_t *d_blob_off, uint8_t *d_blob, uint64_t blob_cap, uint64_t *blob_bytes,
                              void *stream);
/* align.go:295-314 Alignment.UnmarshalBinary: the fields MarshalBinary stores, back into a record (edit_stop ..
 * indel, read_count, junc_seq_len; the other fields are zero), Norm, and a pointer to the JuncSeq bytes inside the
 * blob.  ERR_INVALID when the blob is shorter than its header says (the reference panics there). */
int treat_b200_unmarshal_record(const uint8_t *blob, size_t len, treat_b200_record *rec, double *norm, const uint8_t **junc_seq);
/* cmd/treat/storage.go:758-770 (`treat load`): Alignment.MarshalBinary for a whole batch; blob i is
 * out[out_off[i] .. out_off[i+1]).  Returns the total size in bytes; nothing is written when out is NULL
 * or out_cap is too small (call once to size).  threads = host threads for the fill. */
int64_t treat_b200_marshal_batch(const treat_b200_record *rec, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N,
                                 uint8_t *out, uint64_t out_cap, uint64_t *out_off /* N+1 */, int32_t threads);
/* cmd/treat/mutant.go:55-110 (`treat mutant`): census of the distinct gapped template strings (reads with an
 * insertion) and gapped read strings (reads with a deletion), from the records + ops of an align call made
 * with TREAT_B200_WANT_OPS.  *out is the text `treat mutant -n n` prints (release with treat_b200_free);
 * entries with equal counts are ordered by string (Go's sort.Sort leaves that order unspecified). */
int treat_b200_mutant_report(const treat_b200_template *t, const treat_b200_record *rec, const uint8_t *ops,
                             uint32_t ops_stride, const uint8_t *seqs, const uint64_t *seq_off, uint64_t N, int32_t n,
                             char **out, size_t *out_len);
/* `treat norm` (cmd/treat/norm.go:25-73, cmd/treat/storage.go:801-878).  NormalizeSample sums ReadCount over a sample's
 * standard reads (HasMutation == 0) -- that sum is counter [1] ("Std", weighted by ReadCount) of the summary accumulated while
 * the sample's reads were aligned (treat_b200_get_summary; same flags as the records) --, takes scale = norm / total (1.0 when
 * the total is 0) and stores Norm = scale * float64(ReadCount) in every standard read's record (bytes 24..32 of its
 * MarshalBinary blob, big-endian float64); reads with a mutation keep their Norm.
 * treat_b200_norm_default: Normalize's default norm, the average of the samples' totals (norm.go:48-63: total / len(samples)
 * in float64).  treat_b200_norm_scale: NormalizeSample's scale for one sample. */
int treat_b200_norm_default(const uint64_t *std_reads_per_sample, int32_t n_samples, double *norm);
int treat_b200_norm_scale(uint64_t std_reads_of_sample, double norm, double *scale);
/* The per-record half of NormalizeSample for a batch on the host: norm_out[i] = scale * ReadCount for a standard read, 0.0 (the
 * value `treat load` stored) for a read with a mutation; with blob != NULL the value is also written into blob i (as
 * treat_b200_marshal_batch laid them out: blob_off[N + 1]) -- for standard reads only, like the reference. */
int treat_b200_normalize_records(const treat_b200_record *rec, uint64_t N, double scale, double *norm_out, uint8_t *blob,
                                 const uint64_t *blob_off);
/* The same for records (and blobs from treat_b200_marshal_device) resident in HBM; asynchronous on `stream`. */
int treat_b200_normalize_device(treat_b200_ctx *ctx, const treat_b200_record *d_rec, uint64_t N, double scale, double *d_norm,
                                uint8_t *d_blob, const uint64_t *d_blob_off, void *stream);
/* The record filter of the web handlers and of `treat search` (cmd/treat/storage.go:48-64 SearchFields, :162-188 HasMatch,
 * :257-286 the loop of Storage.Search over one sample's alignments): which records of a batch match, in record order, after
 * `offset` matches were skipped and up to `limit` of them (0: no limit).  Gene / sample / knock-down / replicate are key
 * matches (HasKeyMatch) and stay with the caller's storage.  edit_stop / junc_end / junc_len < 0: any value. */
typedef struct treat_b200_search {
    int32_t edit_stop;    /* SearchFields.EditStop */
    int32_t junc_end;     /* SearchFields.JuncEnd */
    int32_t junc_len;     /* SearchFields.JuncLen */
    int32_t offset;       /* SearchFields.Offset */
    int32_t limit;        /* SearchFields.Limit */
    int32_t alt_region;   /* SearchFields.AltRegion: > 0: AltEditing must equal uint8(alt_region) */
    uint8_t has_mutation; /* SearchFields.HasMutation */
    uint8_t has_alt;      /* SearchFields.HasAlt: only reads with alt editing (without it they are left out) */
    uint8_t all;          /* SearchFields.All: standard reads and mutants alike */
    uint8_t reserved;
} treat_b200_search;
/* idx[0 .. min(*count, cap)) = the matching records' indices; *count = how many match behind the offset and inside the limit */
int treat_b200_search_records(const treat_b200_search *f, const treat_b200_record *rec, uint64_t N, uint64_t *idx, uint64_t cap,
                              uint64_t *count);
/* The same for records resident in HBM (d_idx on the device); one stream synchronisation returns *count. */
int treat_b200_search_device(treat_b200_ctx *ctx, const treat_b200_search *f, const treat_b200_record *d_rec, uint64_t N,
                             uint64_t *d_idx, uint64_t cap, uint64_t *count, void *stream);
/* highChartHist (cmd/treat/handlers.go:162-299; the /data/es-hist, /data/jl-hist and /data/je-hist series of one sample): the
 * Norm-weighted histogram of EditStop (which = 0), JuncLen (1) or JuncEnd (2) over the records the search fields match
 * (limit and offset are ignored, :178-179), without the reads that sit at the template's edit stop with no junction
 * (EditStop == tmpl_edit_stop && JuncLen == 0, :203-205).  out[v - lo] += Norm for v in [lo, hi], added in record order in
 * float64 -- the reference's summation order within a sample, so the sums are its sums bit for bit.  (The device summary's
 * histograms are the ReadCount-weighted, unfiltered form of the same three series: exact integers that can be reduced
 * across GPUs; with Norm = scale * ReadCount they differ from these only in the rounding of the float sum.) */
int treat_b200_hist_norm(const treat_b200_search *f, const treat_b200_record *rec, const double *norm, uint64_t N,
                         int32_t tmpl_edit_stop, int32_t which, int32_t lo, int32_t hi, double *out);
/* `treat search` output (cmd/treat/search.go:30-89) for the records idx[0 .. n) of one gene / sample (e.g. from
 * treat_b200_search_records): the header line unless header == 0, then per record gene, sample, norm ("%.4f" of
 * RoundPlus(Norm, 4); norm == NULL: 0), read_count, alt_editing ("0" or "A<k>"), has_mutation, edit_stop, junc_end,
 * junc_len, junc_seq (the upper-cased window of the raw read), written by Go's encoding/csv rules (a field is quoted when
 * it holds the separator, a quote, CR / LF or starts with a blank; quotes are doubled) with ',' (csv != 0) or a tab between
 * the fields.  The --fasta form needs the stored fragments and stays with the caller's storage.  *out is malloc'ed
 * (treat_b200_free). */
int treat_b200_search_text(const char *gene, const char *sample, const treat_b200_record *rec, const double *norm, const uint8_t *seqs,
                           const uint64_t *seq_off, const uint64_t *idx, uint64_t n, int32_t csv, int32_t header, char **out,
                           size_t *out_len);
/* cmd/treat/stats.go:62-137 (`treat stats`): the block ShowStats prints for one gene -- rule, name, totals, template
 * facts, the per-sample table -- from summary counters (treat_b200_get_summary, or their sum over GPUs / samples).
 * samples[i] / summaries[i]: one row per sample (each summary holds at least 14 counters); the gene totals are their
 * sums.  unique = --unique (count each read once), norm = --norm (totals only).  *out is malloc'ed (treat_b200_free). */
int treat_b200_stats_text(const treat_b200_template *t, const char *gene, const char *const *samples,
                          const uint64_t *const *summaries, int32_t n_samples, int32_t unique, int32_t norm, char **out,
                          size_t *out_len);
/* cmd/treat/align.go:59-122 `treat align`: writes the exact stdout text to a malloc'ed
 * buffer (*out, release with treat_b200_free) */
int treat_b200_cli_align(treat_b200_ctx *ctx, const char *template_path, const char *fragment_path,
                         const char *s1, const char *s2, uint8_t base, int32_t offset, char **out,
                         size_t *out_len, char *err, size_t errlen);
void treat_b200_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* TREAT_B200_H */
