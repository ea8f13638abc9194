// device_format.h -- parameter blocks shared by the kernels (*_kernel.cu) and the C ABI (capi.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/treat_b200.h"

namespace treat_b200 {

// One pass of the wavefront handles 32 lanes x 15 columns per lane = 480 template bases; longer templates run in
// column strips (k_align_strips).  The site rows of the editing-site assembly hold 32 x 32 sites: m + 1 <= 1024.
constexpr int kOnePassTemplateBases = 480;
constexpr int kMaxTemplateBases = 1023;
// Raw reads longer than this are refused (ERR_TOO_LONG); keeps n in uint16 and bounds shared memory.
constexpr uint32_t kMaxReadLen = 16384;
constexpr int kMaxTemplates = 32;  // K = FE + PE + alternatives

// Device copy of the Template (template.go:41-49) in the form the kernels read.
struct TemplateDev {
    const uint8_t *tb;      // [m] base codes (2-bit alphabet codes, or raw bytes on the wide path)
    const void *tcount;     // [K][len_pad] EditSite, uint8 (packed path) or uint32 (wide path)
    const int32_t *alt;     // [2*n_alt] alt_start..., alt_end...  (after SetOffset)
    const uint8_t *lut;     // [256] upper-cased byte -> alphabet code (0..2, 3 = other)
    int32_t m, len, len_pad, K, n_alt;
    int32_t edit_offset, tmpl_edit_stop;
    uint8_t edit_base;
};

struct PackParams {
    const uint8_t *seqs;    // raw reads; read r = seqs[off[r]-off_bias .. off[r+1]-off_bias)
    const uint64_t *off;
    uint64_t off_bias;
    uint64_t N;
    uint16_t *n;
    uint8_t *flags;
    uint8_t *bases;
    uint8_t *counts;
    uint32_t *raw_len;
    uint32_t bases_stride, counts_stride;
    uint32_t stage_cap;     // per-warp staging capacity in bases (>= max_len + 1)
    const uint8_t *lut;
    uint8_t edit_base;
    uint32_t *scalars;      // [0] max n (atomicMax), [1] any saturated edit-base run
    const uint32_t *tpk;    // packed path: the template's bases in the packed-read format (may be null: no marks)
    uint32_t tm;            // template bases
    // reads that are not back to back (k_pack16 only; FASTA records read in place): read r = seqs[begin[r], begin[r] + len[r])
    // inside the buffer seqs[0, span).  begin == nullptr: the off[] form above.
    const uint64_t *begin;
    const uint32_t *len;
    uint64_t span;
    uint32_t padded;        // seqs is 32-byte aligned and readable 32 bytes past its last byte (library-owned buffers)
    // k_pack_finish2: reads too long for one of its passes are appended here ...
    uint32_t *long_list;
    uint32_t *long_count;
    // ... and k_pack_finish, given such a list, handles only reads sel[0 .. *sel_count)
    const uint32_t *sel;
    const uint32_t *sel_count;
};

// internal bit of the per-read flag byte written by k_pack: the read's stripped bases equal the template's
// (n == m, every code equal).  Never reaches treat_b200_record.status.
constexpr uint32_t kFlagSameBases = 0x80u;
// same length as the template and exactly one base differs: the diagonal (score m - 2) is the unique optimum, because
// an alignment of two equal-length sequences with g >= 1 gap pairs scores at most m - 3g.  Also internal.
constexpr uint32_t kFlagOneMismatch = 0x40u;
constexpr uint32_t kFlagsInternal = kFlagSameBases | kFlagOneMismatch;

struct AlignParams {
    // packed reads
    const uint16_t *n;
    const uint8_t *rflags;
    const uint8_t *bases;
    const uint8_t *counts;
    const uint32_t *raw_len;
    const uint32_t *read_count;  // may be null -> 1
    uint32_t bases_stride, counts_stride;
    uint64_t N;
    uint32_t ncap;               // largest n this launch is sized for
    // template
    TemplateDev t;
    // outputs
    treat_b200_record *rec;
    uint8_t *ops;                // may be null
    uint32_t ops_stride;
    uint64_t *summary;           // may be null
    uint32_t *skipped;           // may be null: counts reads (pairs) a speculative launch could not hold
    uint32_t *dbg;               // may be null: bounds-check flag word (TREAT_B200_CHECKS builds)
    uint32_t bins;
    uint32_t flags;
    // shared-memory carve-up (bytes, all multiples of 16)
    uint32_t sm_tmpl_bytes;      // CTA-wide template section
    uint32_t sm_prof_bytes;      // scoring profile inside it
    uint32_t sm_warp_bytes;      // per-warp section
    uint32_t sm_stage_b, sm_stage_c;  // one read inside a stage slot (bases, counts)
    uint32_t sm_stage_total;     // both slots, all reads of a warp
    uint32_t sm_rb, sm_dir, sm_ops, sm_t2f;
    uint32_t pair;               // 1: k_align2 (two reads per warp), 0: k_align
    uint32_t force_single;       // TREAT_B200_ONE_READ_PER_WARP
    uint32_t force_full;         // TREAT_B200_FULL_TRACEBACK_STORE: always take the global-scratch path
    uint32_t band_rows;          // direction-word rows kept per lane in shared memory = (2 band_chunks + 1)(CPL + 1)
    uint32_t band_chunks;        // C: lane l keeps fill chunk u iff |u - l| <= C
    void *scratch;               // global scratch for full fills: [CTAs x warps][scratch_stride] words
    uint32_t scratch_stride;
    // read selection.  k_same_bases: reads whose flag byte lacks kFlagSameBases are appended to list/list_count.
    // k_align / k_align2: when list is set, the reads to align are list[0 .. *list_count) instead of 0 .. N-1.
    uint32_t *list;
    uint32_t *list_count;
    // ops output.  ops_gapped_only: rows of reads without a gap (all-zero rows) are not written.  ops_idx set:
    // the written rows are compacted: row k = atomicAdd(ops_count) holds the ops of read ops_idx[k].
    uint32_t ops_gapped_only;
    uint32_t *ops_idx;
    uint32_t *ops_count;
    uint32_t cpl;                // template columns per lane (sizes the template section)
    uint32_t strips;             // column strips of 32 x cpl bases (1 unless the template is longer than 480 bases)
    uint32_t sm_cnt_stage;       // k_same_bases: bytes per warp for the staged edit-site counts
    uint32_t tie;                // TREAT_B200_TIE_*: which pointer nwalgo's traceback takes when several are optimal
};

// Alignment.WriteTo for a batch (text_kernel.cu)
struct TextParams {
    // raw reads: read r = seqs[off[r] - off_bias, off[r+1] - off_bias), or seqs[begin[r], begin[r] + len_[r]) when begin is set
    const uint8_t *seqs;
    const uint64_t *off;
    uint64_t off_bias;
    const uint64_t *begin;
    const uint32_t *len_;
    // names: name r = names[name_off[r], name_off[r+1]), or names[name_off[r], name_off[r] + name_len[r]) when name_len is set
    const uint8_t *names;
    const uint64_t *name_off;
    const uint32_t *name_len;
    const treat_b200_record *rec;
    const uint8_t *ops;          // one row per read (gapless reads: all zero)
    uint32_t ops_stride;
    uint64_t N;
    uint32_t max_len;            // longest raw read
    int32_t tw;                  // terminal width (>= 5)
    // template: bases as bytes, EditSite rows as uint32 [K][len_pad], Max(i) over the rows [len]
    const uint8_t *tb;
    const uint32_t *tc32;
    const uint32_t *tmax;
    int32_t m, len, len_pad, K;
    uint8_t edit_base;
    uint32_t *tlen;              // [N] pass 1: text bytes of read r
    uint32_t *gtot;              // [N] pass 1 -> pass 2: length of the gapped rows
    const uint64_t *text_off;    // [N+1] pass 2: exclusive scan of tlen
    uint8_t *text;
    uint32_t *gmax;              // pass 1 (may be null): atomicMax of the row lengths
    uint32_t g_longest;          // pass 2: that maximum (0: unknown -> column-wise writer)
    uint32_t stage_cap, tmpl_in_smem, warp_bytes, col_cap;  // set by launch_text
};
int launch_text(TextParams &p, bool write, int num_sms, int max_smem_optin, void *stream);
// Alignment.MarshalBinary for a batch (text_kernel.cu)
struct MarshalParams {
    const uint8_t *seqs;         // raw reads, addressed like TextParams
    const uint64_t *off;
    uint64_t off_bias;
    const uint64_t *begin;
    const treat_b200_record *rec;
    uint64_t N;
    const uint64_t *blob_off;    // [N+1] exclusive scan of 36 + junc_seq_len
    uint8_t *blob;
};
int launch_marshal_sizes(const treat_b200_record *d_rec, uint64_t N, uint32_t *d_blen, int num_sms, void *stream);
int launch_marshal_write(const MarshalParams &p, int num_sms, void *stream);
int configure_text_kernels(int max_smem_optin);

// launchers (pack_kernel.cu, align_kernel.cu)
int launch_pack(const PackParams &p, bool wide, int num_sms, int max_smem_optin, void *stream);
int launch_align(const AlignParams &p, bool wide, int num_sms, int max_smem_optin, void *stream);
// K1f (fused_kernel.cu): NewFragment for the batch + the editing-site assembly of the reads that need no DP in one pass;
// the other reads are packed to HBM and listed in q.list for launch_align
bool pack_finish_applies(const PackParams &p, const AlignParams &q);
int launch_pack_finish(const PackParams &p, const AlignParams &q, int num_sms, int max_smem_optin, void *stream);
bool pack_finish2_applies(const PackParams &p, const AlignParams &q);
int launch_pack_finish2(const PackParams &p, const AlignParams &q, int num_sms, int max_smem_optin, void *stream);
int configure_fused_kernels(int max_smem_optin);
// K0 (frontier_kernel.cu): reads that are the template edited up to a frontier -- a prefix of the pre-edited string, a
// junction, a suffix of the fully edited string -- are finished from two string compares on the raw characters; all other
// reads are appended to f.rest for the fused kernels above (which then run with pack.sel = f.rest).
struct FrontierGeom {
    uint32_t image_bytes;  // shared-memory image of the template (multiple of 16)
    uint32_t off_keep, off_p, off_f, off_pcum, off_fsuf, off_ppos, off_ftail, off_bases, off_alt;
    uint32_t sp, sf;       // bytes per shifted copy of P / F
    uint32_t Lp, Lf, tm;   // raw lengths of the pre-edited / fully edited strings, template bases
    uint32_t mid_cap;      // longest junction (raw characters) that cannot hold a saturating edit-base run
    uint32_t G, ws;        // lanes per read (4 / 8 / 16) and bytes between two reads' windows in a warp's staging area
};
struct FrontierParams {
    const uint8_t *image;
    FrontierGeom g;
    uint32_t *rest;        // reads left for the general path ...
    uint32_t *rest_count;  // ... and their number (zero on entry)
};
// host: builds the image for a template (false: the kernel does not apply to it)
bool build_frontier_image(const uint8_t *bases, int m, const uint32_t *fe, const uint32_t *pe, const int32_t *alt_start, int n_alt,
                          uint8_t edit_base, std::vector<uint8_t> &image, FrontierGeom &g);
bool frontier_applies(const PackParams &p, const AlignParams &q, const FrontierParams &f);
int launch_frontier(const PackParams &p, const AlignParams &q, const FrontierParams &f, int num_sms, int max_smem_optin, void *stream);
int configure_frontier_kernels(int max_smem_optin);
// K2b (band_kernel.cu): of the reads launch_align would get (p.list, or all of them), finishes the pure deletions and the
// ones a banded fill certifies; the others are appended to list2.  band_scratch_bytes: global scratch it needs (0: n/a).
// list2_count points at three zeroed words: reads left for launch_align, pure deletions, band reads.
uint64_t band_scratch_bytes(const AlignParams &p, int num_sms, int max_smem);
// min_reads: with fewer reads to align than this the banded fill is not worth its fixed latency; band candidates then go
// to list2 and only reads at least 16 bases shorter than the template are tested for the pure-deletion route.
int launch_band(const AlignParams &p, uint32_t *list2, uint32_t *list2_count, void *scratch, bool subseq, uint32_t min_reads,
                unsigned long long *stats, int num_sms, int max_smem, void *stream, int *kernels_launched = nullptr);
int configure_band_kernels(int max_smem_optin);
// K2s: finishes the reads marked kFlagSameBases without a DP and lists the others (p is the plan_align'ed block)
int launch_same_bases(const AlignParams &p, int num_sms, void *stream);
// fills the shared-memory geometry of p for (m, K, ncap, wide); returns columns-per-lane or -1
int plan_align(AlignParams &p, bool wide);
int align_warps(const AlignParams &p, int max_smem);
int align_kernel_regs(const AlignParams &p, bool wide);  // registers per thread of the kernel launch_align would pick
uint64_t align_scratch_bytes(const AlignParams &p, int num_sms, int max_smem);
int launch_merge_summary(uint64_t *dst, uint64_t *src, uint32_t len, void *stream);
int launch_heatmap(const treat_b200_record *d_rec, uint64_t N, int edit_offset, uint32_t n, uint64_t *d_heat, int num_sms, void *stream);
// FASTA ingest (fasta_kernel.cu)
int launch_fasta_mark(const uint8_t *file, uint64_t len, uint64_t tile0, uint64_t ntiles, uint32_t *tile_count,
                      uint64_t *first_gt, int num_sms, void *stream);
int launch_scan_u32(const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *sums, void *stream);
int launch_fasta_starts(const uint8_t *file, uint64_t len, uint64_t ntiles, const uint64_t *tile_off, const uint64_t *first_gt,
                        uint64_t *start, int num_sms, void *stream);
int launch_fasta_records(const uint8_t *file, uint64_t len, const uint64_t *start, uint64_t N, uint64_t *id_off, uint32_t *id_len,
                         uint32_t *read_count, uint64_t *seq_begin, uint32_t *seq_len, uint32_t *scalars, int num_sms, void *stream);
int launch_fasta_compact(const uint8_t *file, uint64_t len, const uint64_t *start, uint64_t N, const uint64_t *seq_begin,
                         const uint64_t *seq_off, uint8_t *seqs, int num_sms, void *stream);
constexpr uint32_t kFastaTileBytes = 4096;  // bytes per tile of the record-start count
// up to 64 bytes from device memory into pinned host memory by a one-warp kernel (no copy-engine queue)
int launch_peek(void *h_dst, const void *d_src, uint32_t bytes, void *stream);
int launch_int32_peak(int num_sms, int iters, int *d_out, void *stream);
int configure_pack_kernels(int max_smem_optin);
int configure_align_kernels(int max_smem_optin);

// exact-duplicate collapse (collapse_kernel.cu)
int launch_collapse_hash(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias, uint32_t N, uint32_t seed, uint64_t *key,
                         int num_sms, void *stream);
int launch_collapse_table(const uint64_t *key, uint32_t N, uint64_t *table, uint32_t cap, uint32_t *slot, uint32_t *head, int num_sms,
                          void *stream);
int launch_collapse_verify(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias, const uint32_t *read_count, uint32_t N,
                           const uint32_t *slot, const uint32_t *head, uint32_t *rep, uint32_t *is_rep, uint32_t *sum, uint32_t *collisions,
                           uint32_t *max_len, int num_sms, void *stream);
int launch_collapse_compact(const uint64_t *off, uint64_t off_bias, uint32_t N, const uint32_t *rep, const uint32_t *is_rep,
                            const uint64_t *uidx, const uint32_t *sum, uint32_t *first, uint64_t *begin, uint32_t *len, uint32_t *count,
                            uint32_t *member, int num_sms, void *stream);

int launch_collapse_gather(const uint8_t *seqs, const uint64_t *begin, const uint32_t *len, const uint64_t *dst_off, uint32_t N, uint8_t *dst,
                           int num_sms, void *stream);

#ifdef __CUDACC__
// resident CTAs per SM of a launch shape, asked of the runtime once per shape: the query costs far more than a launch, and
// treat_b200_align_batch launches every chunk's kernels through the same few shapes
int occupancy_cached(const void *kern, int threads, size_t smem);
#endif

// context internals other translation units need (group.cu)
int ctx_device(const treat_b200_ctx *ctx);
#ifdef __CUDACC__
cudaStream_t ctx_stream(treat_b200_ctx *ctx);
#endif

}  // namespace treat_b200
