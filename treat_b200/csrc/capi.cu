// capi.cu -- device side of the C ABI declared in include/treat_b200.h: the context, the
// template upload, the batch entry points (host buffers and device buffers) and summaries.
// There is no CPU path: every align call ends in launch_pack / launch_align.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in nsys / ncu timelines, a few ns when no tool is attached

#include "device_format.h"
#include "host/treat_host.hpp"

using namespace treat_b200;

namespace {

constexpr uint32_t a16(uint32_t x) { return (x + 15u) & ~15u; }
constexpr uint64_t kChunkReads = 1u << 16;  // reads per pipelined chunk of treat_b200_align_batch
constexpr int kSlots = 4;                   // most chunks in flight: H2D, kernels and D2H of different chunks overlap
constexpr size_t kScalarBytes = 64;         // per-slot device scalars (16 words; [6..8]: k_band's three list counters)
constexpr int kFastaSlots = 6;              // pieces in flight of treat_b200_fasta_stream: two being copied, index, align, deliver

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// Host threads that copy between the caller's pageable buffers and pinned staging buffers (see stage_in / unstage in
// treat_b200_align_batch): one memcpy runs at 10 GB/s or so, a handful of them keep up with the PCIe link.
// A large host copy whose destination is not read again soon (a pinned staging buffer the DMA engine reads, or the caller's
// result buffer): non-temporal stores skip the read-for-ownership of the destination lines, a third of the memory traffic of
// a copy that does not fit the caches.  (glibc's memcpy switches to them only for copies of tens of megabytes.)
static void stream_copy(void *dst_, const void *src_, size_t n)
{
#if defined(__SSE2__)
    if (n >= ((size_t)256 << 10)) {
        uint8_t *d = static_cast<uint8_t *>(dst_);
        const uint8_t *s = static_cast<const uint8_t *>(src_);
        const size_t head = (16u - (size_t)(reinterpret_cast<uintptr_t>(d) & 15u)) & 15u;
        memcpy(d, s, head);
        d += head, s += head, n -= head;
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; i++, d += 64, s += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
        }
        _mm_sfence();
        memcpy(d, s, n - blocks * 64);
        return;
    }
#endif
    if (n) memcpy(dst_, src_, n);
}

struct CopyPool {
    struct Job {
        uint8_t *dst = nullptr;
        const uint8_t *src = nullptr;
        size_t n = 0;
        bool nt = false;
    };
    std::vector<std::thread> th;
    std::vector<Job> job;
    std::mutex m;
    std::condition_variable cv, done;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    void start(int n)
    {
        if (!th.empty() || n < 1) return;
        job.resize((size_t)n);
        for (int i = 0; i < n; i++)
            th.emplace_back([this, i] {
                uint64_t seen = 0;
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(m);
                        cv.wait(lk, [&] { return stop || gen != seen; });
                        if (stop) return;
                        seen = gen;
                        j = job[(size_t)i];
                    }
                    if (j.nt) stream_copy(j.dst, j.src, j.n);
                    else if (j.n) memcpy(j.dst, j.src, j.n);
                    {
                        std::lock_guard<std::mutex> lk(m);
                        if (--pending == 0) done.notify_one();
                    }
                }
            });
    }
    // nt: write the destination with non-temporal stores (staging buffers the copy engine reads next; measured slower for
    // results landing in the caller's memory)
    void copy(void *dst, const void *src, size_t n, bool nt = false)
    {
        if (th.empty() || n < (1u << 20)) {
            if (n) memcpy(dst, src, n);
            return;
        }
        const size_t parts = th.size() + 1, each = ((n + parts - 1) / parts + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(m);
            for (size_t i = 0; i < th.size(); i++) {
                const size_t lo = std::min(n, (i + 1) * each), hi = std::min(n, (i + 2) * each);
                job[i] = Job{(uint8_t *)dst + lo, (const uint8_t *)src + lo, hi - lo, nt};
            }
            pending = (int)th.size();
            gen++;
        }
        cv.notify_all();
        if (nt) stream_copy(dst, src, std::min(n, each));
        else memcpy(dst, src, std::min(n, each));
        std::unique_lock<std::mutex> lk(m);
        done.wait(lk, [&] { return pending == 0; });
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m);
            stop = true;
        }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
};

// NVTX range around a host-side phase (copy in / pack + align launches / copy out / text / deliver)
struct Nvtx {
    explicit Nvtx(const char *name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
    Nvtx(const Nvtx &) = delete;
    Nvtx &operator=(const Nvtx &) = delete;
};

// device buffers of one in-flight chunk
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t h2d_done = nullptr;  // recorded behind the chunk's input copies: the next chunk's copies wait for it (see h2d_in_order)
    DevBuf seqs, off, rc, n, flags, bases, counts, raw_len, rec, ops, summary;
    DevBuf list;     // reads that need the DP (written by k_same_bases / k_pack_finish*, read by the align kernels)
    DevBuf long_list;  // reads too long for one pass of k_pack_finish2
    DevBuf rest_list;  // reads k_frontier leaves for k_pack_finish2
    DevBuf list2;      // reads k_band leaves for the full-matrix kernels
    DevBuf ops_idx;  // OPS_GAPPED_ONLY: read index of every compacted ops row
    // OPS_GAPPED_ONLY: pinned landing buffers of the compacted rows and their read indices
    uint8_t *h_gap_rows = nullptr;
    uint32_t *h_gap_idx = nullptr;
    size_t gap_rows_cap = 0, gap_idx_cap = 0;
    uint32_t gap_head = 0;     // rows copied ahead of knowing how many there are
    bool scatter = false;      // the chunk's gapped-read count (and compacted rows) still have to be consumed on the host
    bool sparse = false;       // this chunk returns only the gapped reads' ops rows, compacted
    uint8_t *land = nullptr;   // pinned landing buffer of a piece's / chunk's results (fasta_stream; pageable outputs of align_batch)
    size_t land_cap = 0;
    uint8_t *h_in = nullptr;   // pinned staging buffer of a chunk's inputs when the caller's buffers are pageable
    size_t h_in_cap = 0;
    bool unstage = false;      // the chunk's results sit in `land` and still have to be copied to the caller's buffers
    const uint64_t *doff = nullptr;  // the chunk's slice of the call's device copies of seq_off / read_count
    const uint32_t *drc = nullptr;
    DevBuf scratch;  // global scratch of this slot's align launches (launches of different slots overlap on the GPU)
    DevBuf text, text_off, tlen, gtot, tsums;  // Alignment.WriteTo on the device (treat_b200_text_device / _fasta_text)
    DevBuf names, name_off;                    // treat_b200_text_batch: the chunk's names
    uint8_t *tland = nullptr;                  // pinned landing buffer of a piece's text
    size_t tland_cap = 0;
    uint64_t *h_text_total = nullptr;          // pinned [1]
    uint32_t *scalars = nullptr;    // device [8]: max n, any saturated run, reads skipped by a speculative launch, reads listed
                                    // for the DP, compacted ops rows
    uint32_t *h_scalars = nullptr;  // pinned [8]
    // chunk whose speculative launch still has to be validated (treat_b200_align_batch)
    bool pending = false;
    uint64_t c0 = 0, cn = 0, b0 = 0;
    uint32_t max_len = 0, dev_stride = 0;
    void release()
    {
        for (DevBuf *b : {&seqs, &off, &rc, &n, &flags, &bases, &counts, &raw_len, &rec, &ops, &summary, &scratch, &list, &long_list, &rest_list, &list2, &ops_idx, &text,
                          &text_off, &tlen, &gtot, &tsums, &names, &name_off})
            b->release();
        if (tland) cudaFreeHost(tland);
        if (h_text_total) cudaFreeHost(h_text_total);
        if (scalars) cudaFree(scalars);
        if (h_scalars) cudaFreeHost(h_scalars);
        if (h_gap_rows) cudaFreeHost(h_gap_rows);
        if (h_gap_idx) cudaFreeHost(h_gap_idx);
        if (land) cudaFreeHost(land);
        if (h_in) cudaFreeHost(h_in);
        if (h2d_done) cudaEventDestroy(h2d_done);
        if (stream) cudaStreamDestroy(stream);
    }
};

// reads that are not back to back in their buffer (FASTA records aligned in place)
struct ReadSpan {
    const uint64_t *begin;
    const uint32_t *len;
    uint64_t span;
};

// a FASTA file resident on the device with its record index (treat_b200_fasta_upload)
struct FastaDev {
    DevBuf file, tile_count, tile_off, sums, start, id_off, id_len, rc, seq_begin, seq_len, seq_off, seqs, scal;
    uint64_t len = 0, N = 0, seq_bytes = 0;
    uint32_t max_len = 0;
    uint64_t *h = nullptr;   // pinned [8]: line-start '>' count, first '>', (longest sequence | multi-line flag << 32), sequence bytes
    uint8_t *h_stage = nullptr;  // pinned staging buffer (two 32 MB halves) for source bytes that are pageable
    size_t h_stage_cap = 0;
    cudaEvent_t stage_done[2] = {nullptr, nullptr};  // the H2D copy out of each half
    bool multi = false;      // some record has more than one sequence line: reads are compacted before packing
    bool compacted = false;  // seqs / seq_off hold the compacted reads
    bool open = false;
    void release()
    {
        for (DevBuf *b : {&file, &tile_count, &tile_off, &sums, &start, &id_off, &id_len, &rc, &seq_begin, &seq_len, &seq_off, &seqs, &scal})
            b->release();
        if (h) cudaFreeHost(h);
        h = nullptr;
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        h_stage_cap = 0;
        for (cudaEvent_t &ev : stage_done) {
            if (ev) cudaEventDestroy(ev);
            ev = nullptr;
        }
    }
};

}  // namespace

struct treat_b200_ctx {
    int device = 0;
    int num_sms = 0;
    int max_smem = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    // template
    bool have_template = false;
    bool tmpl_wide = false;  // template needs the wide path (non-alphabet base or count >= 255)
    TemplateDev td{};
    DevBuf d_tb, d_tb_wide, d_tc8, d_tc32, d_alt, d_lut;
    DevBuf d_tpk;  // template bases in the packed-read format (16 codes per word): k_pack marks reads that equal it
    DevBuf d_frontier;              // k_frontier: shared-memory image of the template (shifted copies of the FE / PE strings, tables)
    FrontierGeom fr_geom{};         //   its layout; image_bytes == 0: the kernel does not apply to this template
    DevBuf d_tmax; // Template.Max(i) (template.go:186) for every edit site: column widths of WriteTo
    uint32_t band_rows = 0;  // 0 = default window
    uint32_t band_min_reads = 16384;  // k_band's banded fill only pays for itself when a launch has at least this many reads to align
    uint32_t tie = TREAT_B200_TIE_ULD;  // traceback tie order (treat_b200_set_tie_order)
    uint32_t ncap_hint = 0;  // largest n seen with this template: sizes speculative align launches
    uint32_t gap_hint = 0;   // most gapped reads seen in a chunk: sizes the ahead-of-time copy of compacted ops rows
    uint32_t *d_dbg = nullptr;  // bounds-check flags (only written by TREAT_B200_CHECKS builds)
    unsigned long long *d_dp_stats = nullptr;  // [4] reads k_band finished as pure deletions / by the banded fill / left for the full DP
    std::vector<uint8_t> lut;
    uint64_t *d_summary = nullptr;
    uint32_t bins = 0;
    uint64_t *d_heat = nullptr;  // [len][len] ESS x junction-length heat map (TREAT_B200_HEATMAP)
    Slot slot[kSlots];
    // scratch for the device-resident entry point
    Slot dev_slot;
    CopyPool copy_pool;           // host copy threads, started on the first call that brings pageable buffers
    DevBuf d_off_all, d_rc_all;   // treat_b200_align_batch: the call's seq_off / read_count, copied once for all chunks
    uint8_t *h_meta = nullptr;    //   their pinned staging buffer when the caller's arrays are pageable
    size_t h_meta_cap = 0;
    cudaEvent_t meta_done = nullptr;
    // treat_b200_collapse_align_batch
    struct {
        DevBuf seqs, key, table, slot, head, rep, is_rep, sum, uidx, sums, first, begin, len, count, member, scal, rec, ops, gather, goff;
        uint64_t *h = nullptr;  // pinned [4]: collisions | longest unique read << 32, n_unique
    } col;
    FastaDev fasta;               // treat_b200_fasta_upload / treat_b200_fasta_align
    FastaDev fasta_slot[kFastaSlots];  // pieces in flight of treat_b200_fasta_stream
    Slot fslot[kFastaSlots];
};

#define CK(ctx, expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);            \
            return TREAT_B200_ERR_CUDA;                                                 \
        }                                                                               \
    } while (0)

static int fail(treat_b200_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// Small device -> host read-backs the host waits on right away (sizes, counts): written by a one-warp kernel straight
// into pinned host memory instead of a cudaMemcpyAsync, which would queue on the D2H copy engine behind the bulk result
// copies of the other chunks in flight and stall the pipeline for their whole duration.
static int peek(treat_b200_ctx *ctx, void *h_dst, const void *d_src, uint32_t bytes, cudaStream_t st)
{
    int e = launch_peek(h_dst, d_src, bytes, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_peek launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}
#define PEEK(ctx, dst, src, bytes, st)                     \
    do {                                                   \
        int _r = peek((ctx), (dst), (src), (bytes), (st)); \
        if (_r) return _r;                                 \
    } while (0)

// Input copies of different chunks sit on different streams, and the copy engines serve concurrent H2D copies side by
// side: each then takes as long as all of them together, so every chunk's kernels start late and the last chunks all
// finish copying at the very end, with their kernels and result copies still to come.  Chaining the copies with events
// keeps the bus just as busy but delivers the chunks one after the other.
static int h2d_in_order_begin(treat_b200_ctx *ctx, Slot &s, const Slot *prev)
{
    if (!s.h2d_done) CK(ctx, cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    if (prev && prev->h2d_done) CK(ctx, cudaStreamWaitEvent(s.stream, prev->h2d_done, 0));
    return TREAT_B200_OK;
}
static int h2d_in_order_end(treat_b200_ctx *ctx, Slot &s)
{
    CK(ctx, cudaEventRecord(s.h2d_done, s.stream));
    return TREAT_B200_OK;
}

extern "C" {

int treat_b200_abi_version(void) { return TREAT_B200_ABI_VERSION; }

const char *treat_b200_strerror(int code)
{
    switch (code) {
    case TREAT_B200_OK: return "ok";
    case TREAT_B200_ERR_INVALID: return "invalid argument";
    case TREAT_B200_ERR_CUDA: return "CUDA error";
    case TREAT_B200_ERR_NOMEM: return "out of memory";
    case TREAT_B200_ERR_TOO_LONG: return "template or read too long for the kernel geometry";
    case TREAT_B200_ERR_NO_TEMPLATE: return "no template set";
    case TREAT_B200_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU path)";
    case TREAT_B200_ERR_TEMPLATE: return "invalid template";
    case TREAT_B200_ERR_IO: return "I/O error";
    default: return "unknown error";
    }
}

const char *treat_b200_last_error(const treat_b200_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }

int treat_b200_create(int device, treat_b200_ctx **out)
{
    if (!out) return TREAT_B200_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return TREAT_B200_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TREAT_B200_ERR_NO_DEVICE;
    if (prop.major != 10) return TREAT_B200_ERR_NO_DEVICE;  // the fatbin holds sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return TREAT_B200_ERR_NO_DEVICE;
    treat_b200_ctx *ctx = new treat_b200_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return TREAT_B200_ERR_CUDA;
    }
    if (cudaMalloc((void **)&ctx->d_dbg, 4) != cudaSuccess || cudaMemset(ctx->d_dbg, 0, 4) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TREAT_B200_ERR_CUDA;
    }
    if (cudaMalloc((void **)&ctx->d_dp_stats, 32) != cudaSuccess || cudaMemset(ctx->d_dp_stats, 0, 32) != cudaSuccess) {
        cudaFree(ctx->d_dbg);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TREAT_B200_ERR_CUDA;
    }
    if (const char *bm = getenv("TREAT_B200_BAND_MIN")) ctx->band_min_reads = (uint32_t)strtoul(bm, nullptr, 10);
    int e = configure_pack_kernels(ctx->max_smem);
    if (!e) e = configure_align_kernels(ctx->max_smem);
    if (!e) e = configure_text_kernels(ctx->max_smem);
    if (!e) e = configure_fused_kernels(ctx->max_smem);
    if (!e) e = configure_frontier_kernels(ctx->max_smem);
    if (!e) e = configure_band_kernels(ctx->max_smem);
    if (e) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return TREAT_B200_ERR_CUDA;
    }
    *out = ctx;
    return TREAT_B200_OK;
}

void treat_b200_destroy(treat_b200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf *b : {&ctx->d_tb, &ctx->d_tb_wide, &ctx->d_tc8, &ctx->d_tc32, &ctx->d_alt, &ctx->d_lut, &ctx->d_tpk, &ctx->d_tmax, &ctx->d_frontier}) b->release();
    if (ctx->d_summary) cudaFree(ctx->d_summary);
    if (ctx->d_heat) cudaFree(ctx->d_heat);
    if (ctx->d_dbg) cudaFree(ctx->d_dbg);
    if (ctx->d_dp_stats) cudaFree(ctx->d_dp_stats);
    for (Slot &sl : ctx->slot) sl.release();
    ctx->d_off_all.release();
    ctx->d_rc_all.release();
    if (ctx->h_meta) cudaFreeHost(ctx->h_meta);
    if (ctx->meta_done) cudaEventDestroy(ctx->meta_done);
    ctx->dev_slot.release();
    for (DevBuf *b : {&ctx->col.seqs, &ctx->col.key, &ctx->col.table, &ctx->col.slot, &ctx->col.head, &ctx->col.rep, &ctx->col.is_rep, &ctx->col.sum,
                      &ctx->col.uidx, &ctx->col.sums, &ctx->col.first, &ctx->col.begin, &ctx->col.len, &ctx->col.count, &ctx->col.member,
                      &ctx->col.scal, &ctx->col.rec, &ctx->col.ops, &ctx->col.gather, &ctx->col.goff})
        b->release();
    if (ctx->col.h) cudaFreeHost(ctx->col.h);
    ctx->fasta.release();
    for (FastaDev &f : ctx->fasta_slot) f.release();
    for (Slot &sl : ctx->fslot) sl.release();
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int treat_b200_host_alloc(void **out, size_t bytes)
{
    if (!out) return TREAT_B200_ERR_INVALID;
    if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return TREAT_B200_ERR_NOMEM;
    }
    return TREAT_B200_OK;
}
void treat_b200_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}
int treat_b200_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return TREAT_B200_ERR_INVALID;
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();
        return TREAT_B200_ERR_CUDA;
    }
    return TREAT_B200_OK;
}
int treat_b200_host_unregister(void *p)
{
    if (!p) return TREAT_B200_ERR_INVALID;
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return TREAT_B200_ERR_CUDA;
    }
    return TREAT_B200_OK;
}

int treat_b200_set_template(treat_b200_ctx *ctx, const uint8_t *bases, int32_t m, const uint32_t *edit_site, int32_t K,
                            const int32_t *alt_start, const int32_t *alt_end, int32_t n_alt, uint32_t edit_offset,
                            int32_t tmpl_edit_stop, uint8_t edit_base)
{
    if (!ctx || !bases || !edit_site || m < 1 || K < 2 || n_alt != K - 2 || (n_alt > 0 && (!alt_start || !alt_end)))
        return fail(ctx, TREAT_B200_ERR_INVALID, "set_template: bad arguments (need m >= 1, K >= 2, n_alt == K-2)");
    if (m > kMaxTemplateBases || K > kMaxTemplates)
        return fail(ctx, TREAT_B200_ERR_TOO_LONG, "set_template: template longer than 1023 bases or more than 32 rows");
    CK(ctx, cudaSetDevice(ctx->device));
    if (edit_base >= 'a' && edit_base <= 'z') edit_base -= 32;
    const int len = m + 1, len_pad = (len + 31) / 32 * 32;

    // packed alphabet: the first three of "ACGT" that are not the edit base; everything else is code 3
    std::vector<uint8_t> lut(256, 3);
    {
        int code = 0;
        for (const char *c = "ACGT"; *c && code < 3; c++)
            if ((uint8_t)*c != edit_base) lut[(uint8_t)*c] = (uint8_t)code++;
    }
    bool wide = false;
    std::vector<uint8_t> tb(m), tbw(m);
    for (int i = 0; i < m; i++) {
        tbw[i] = bases[i];
        tb[i] = lut[bases[i]];
        if (tb[i] == 3) wide = true;
    }
    std::vector<uint8_t> tc8((size_t)K * len_pad, 0);
    std::vector<uint32_t> tc32((size_t)K * len_pad, 0);
    for (int k = 0; k < K; k++)
        for (int i = 0; i < len; i++) {
            const uint32_t c = edit_site[(size_t)k * len + i];
            if (c >= 255u) wide = true;
            tc8[(size_t)k * len_pad + i] = (uint8_t)std::min<uint32_t>(c, 255u);
            tc32[(size_t)k * len_pad + i] = c;
        }
    std::vector<int32_t> alt(2 * std::max(n_alt, 1), 0);
    for (int i = 0; i < n_alt; i++) {
        alt[i] = alt_start[i];
        alt[n_alt + i] = alt_end[i];
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, ctx->d_tb.reserve(m));
    CK(ctx, ctx->d_tb_wide.reserve(m));
    CK(ctx, ctx->d_tc8.reserve(tc8.size()));
    CK(ctx, ctx->d_tc32.reserve(tc32.size() * 4));
    CK(ctx, ctx->d_alt.reserve(alt.size() * 4));
    CK(ctx, ctx->d_lut.reserve(256));
    CK(ctx, cudaMemcpy(ctx->d_tb.p, tb.data(), m, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(ctx->d_tb_wide.p, tbw.data(), m, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(ctx->d_tc8.p, tc8.data(), tc8.size(), cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(ctx->d_tc32.p, tc32.data(), tc32.size() * 4, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(ctx->d_alt.p, alt.data(), alt.size() * 4, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(ctx->d_lut.p, lut.data(), 256, cudaMemcpyHostToDevice));
    {
        std::vector<uint32_t> tpk((size_t)(m + 15) / 16, 0u);
        if (!wide)
            for (int i = 0; i < m; i++) tpk[i >> 4] |= (uint32_t)tb[i] << (2 * (i & 15));
        CK(ctx, ctx->d_tpk.reserve(tpk.size() * 4));
        CK(ctx, cudaMemcpy(ctx->d_tpk.p, tpk.data(), tpk.size() * 4, cudaMemcpyHostToDevice));
    }
    {
        // k_frontier: the FE / PE rows re-inflated to raw strings, as a shared-memory image
        std::vector<uint8_t> image;
        FrontierGeom fg{};
        std::vector<uint8_t> up(bases, bases + m);
        for (uint8_t &c : up)
            if (c >= 'a' && c <= 'z') c -= 32;
        ctx->fr_geom = FrontierGeom{};
        if (!wide && build_frontier_image(up.data(), m, edit_site, edit_site + len, alt_start, n_alt, edit_base, image, fg)) {
            CK(ctx, ctx->d_frontier.reserve(image.size()));
            CK(ctx, cudaMemcpy(ctx->d_frontier.p, image.data(), image.size(), cudaMemcpyHostToDevice));
            ctx->fr_geom = fg;
        }
    }
    {
        std::vector<uint32_t> tmax((size_t)len_pad, 0u);
        for (int i = 0; i < len; i++)
            for (int k = 0; k < K; k++) tmax[i] = std::max(tmax[i], edit_site[(size_t)k * len + i]);
        CK(ctx, ctx->d_tmax.reserve(tmax.size() * 4));
        CK(ctx, cudaMemcpy(ctx->d_tmax.p, tmax.data(), tmax.size() * 4, cudaMemcpyHostToDevice));
    }
    ctx->lut = lut;
    ctx->tmpl_wide = wide;
    ctx->td = TemplateDev{};
    ctx->td.alt = (const int32_t *)ctx->d_alt.p;
    ctx->td.lut = (const uint8_t *)ctx->d_lut.p;
    ctx->td.m = m;
    ctx->td.len = len;
    ctx->td.len_pad = len_pad;
    ctx->td.K = K;
    ctx->td.n_alt = n_alt;
    ctx->td.edit_offset = (int32_t)edit_offset;
    ctx->td.tmpl_edit_stop = tmpl_edit_stop;
    ctx->td.edit_base = edit_base;
    // summary counters are sized by the template
    const uint32_t bins = (uint32_t)len + 2;
    if (ctx->d_summary) cudaFree(ctx->d_summary);
    ctx->d_summary = nullptr;
    ctx->bins = bins;
    const size_t sl = (TREAT_B200_SUMMARY_HEADER + 3 * (size_t)bins) * 8;
    CK(ctx, cudaMalloc((void **)&ctx->d_summary, sl));
    CK(ctx, cudaMemset(ctx->d_summary, 0, sl));
    if (ctx->d_heat) cudaFree(ctx->d_heat);
    ctx->d_heat = nullptr;
    CK(ctx, cudaMalloc((void **)&ctx->d_heat, (size_t)len * len * 8));
    CK(ctx, cudaMemset(ctx->d_heat, 0, (size_t)len * len * 8));
    ctx->have_template = true;
    ctx->ncap_hint = 0;
    return TREAT_B200_OK;
}

int treat_b200_use_template(treat_b200_ctx *ctx, const treat_b200_template *tmpl)
{
    if (!ctx || !tmpl) return TREAT_B200_ERR_INVALID;
    const treat::Template &t = *reinterpret_cast<const treat::Template *>(tmpl);
    const int m = (int)t.Bases.size(), K = t.Size(), len = m + 1;
    std::vector<uint32_t> es((size_t)K * len);
    for (int k = 0; k < K; k++) std::copy(t.EditSite[k].begin(), t.EditSite[k].end(), es.begin() + (size_t)k * len);
    std::vector<int32_t> as, ae;
    for (const auto &r : t.AltRegion) {
        as.push_back(r.Start);
        ae.push_back(r.End);
    }
    return treat_b200_set_template(ctx, (const uint8_t *)t.Bases.data(), m, es.data(), K, as.data(), ae.data(),
                                   (int32_t)as.size(), t.EditOffset, t.EditStop, (uint8_t)t.EditBase);
}

uint32_t treat_b200_ops_stride(const treat_b200_ctx *ctx, uint32_t max_len)
{
    if (!ctx || !ctx->have_template) return 0;
    return a16(((uint32_t)ctx->td.m + max_len + 3) / 4);
}

int treat_b200_synchronize(treat_b200_ctx *ctx)
{
    if (!ctx) return TREAT_B200_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (Slot &sl : ctx->slot)
        if (sl.stream) CK(ctx, cudaStreamSynchronize(sl.stream));
    if (ctx->dev_slot.stream) CK(ctx, cudaStreamSynchronize(ctx->dev_slot.stream));
    for (Slot &sl : ctx->fslot)
        if (sl.stream) CK(ctx, cudaStreamSynchronize(sl.stream));
    return TREAT_B200_OK;
}

uint64_t treat_b200_launch_count(const treat_b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int treat_b200_dp_stats(treat_b200_ctx *ctx, uint64_t *out3, int32_t reset)
{
    if (!ctx || !out3) return TREAT_B200_ERR_INVALID;
    int rc = treat_b200_synchronize(ctx);
    if (rc) return rc;
    CK(ctx, cudaMemcpy(out3, ctx->d_dp_stats, 24, cudaMemcpyDeviceToHost));
    if (reset) CK(ctx, cudaMemset(ctx->d_dp_stats, 0, 32));
    return TREAT_B200_OK;
}

int treat_b200_debug_flags(treat_b200_ctx *ctx, uint32_t *flags_out, int32_t *checks_compiled)
{
    if (!ctx || !flags_out) return TREAT_B200_ERR_INVALID;
    int rc = treat_b200_synchronize(ctx);
    if (rc) return rc;
    CK(ctx, cudaMemcpy(flags_out, ctx->d_dbg, 4, cudaMemcpyDeviceToHost));
#ifdef TREAT_B200_CHECKS
    if (checks_compiled) *checks_compiled = 1;
#else
    if (checks_compiled) *checks_compiled = 0;
#endif
    return TREAT_B200_OK;
}

int treat_b200_set_tie_order(treat_b200_ctx *ctx, int32_t order)
{
    if (!ctx) return TREAT_B200_ERR_INVALID;
    if (order != TREAT_B200_TIE_ULD && order != TREAT_B200_TIE_UDL && order != TREAT_B200_TIE_LUD)
        return fail(ctx, TREAT_B200_ERR_INVALID, "set_tie_order: ULD (0), UDL (1) or LUD (2)");
    ctx->tie = (uint32_t)order;
    return TREAT_B200_OK;
}

int treat_b200_set_band_min_reads(treat_b200_ctx *ctx, uint32_t reads)
{
    if (!ctx) return TREAT_B200_ERR_INVALID;
    ctx->band_min_reads = reads;
    return TREAT_B200_OK;
}

int treat_b200_set_band_rows(treat_b200_ctx *ctx, uint32_t rows)
{
    if (!ctx) return TREAT_B200_ERR_INVALID;
    ctx->band_rows = rows;
    return TREAT_B200_OK;
}

int treat_b200_align_geometry(treat_b200_ctx *ctx, uint32_t max_n, uint32_t flags, treat_b200_geometry *out)
{
    if (!ctx || !out) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "geometry: no template");
    memset(out, 0, sizeof *out);
    AlignParams ap{};
    ap.t = ctx->td;
    ap.ncap = max_n;
    ap.bins = ctx->bins;
    ap.summary = (flags & TREAT_B200_NO_SUMMARY) ? nullptr : ctx->d_summary;
    ap.force_single = (flags & TREAT_B200_ONE_READ_PER_WARP) ? 1u : 0u;
    ap.band_rows = ctx->band_rows;
    if (plan_align(ap, ctx->tmpl_wide) < 0) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "geometry: template too long");
    const int warps = align_warps(ap, ctx->max_smem);
    if (warps < 1) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "geometry: read/template geometry exceeds shared memory");
    out->cpl = ap.cpl;
    out->strips = ap.strips;
    out->reads_per_warp = ap.pair ? 2u : 1u;
    out->warps_per_cta = (uint32_t)warps;
    out->smem_template_bytes = ap.sm_tmpl_bytes;
    out->smem_warp_bytes = ap.sm_warp_bytes;
    out->smem_cta_bytes = ap.sm_tmpl_bytes + (uint32_t)warps * ap.sm_warp_bytes;
    out->band_rows = ap.band_rows;
    out->band_diagonals = ap.band_chunks ? ap.band_chunks * (ap.cpl + 1) - 1 : 0;
    out->scratch_bytes_per_warp = (uint32_t)(align_scratch_bytes(ap, 1, ctx->max_smem) / (uint64_t)warps);
    out->regs_per_thread = (uint32_t)align_kernel_regs(ap, ctx->tmpl_wide);
    out->occupancy_warps_per_sm = (uint32_t)warps;
    return TREAT_B200_OK;
}

int treat_b200_packed_layout(const treat_b200_ctx *ctx, uint64_t N, uint32_t max_len, uint32_t wide,
                             treat_b200_packed *out)
{
    if (!ctx || !out) return TREAT_B200_ERR_INVALID;
    if (max_len > kMaxReadLen) return TREAT_B200_ERR_TOO_LONG;
    memset(out, 0, sizeof *out);
    out->N = N;
    out->max_len = max_len;
    out->wide = wide ? 1 : 0;
    out->bases_stride = std::max(16u, wide ? a16(max_len) : a16((max_len + 3) / 4));
    out->counts_stride = wide ? a16(4 * (max_len + 1)) : a16(max_len + 1);
    return TREAT_B200_OK;
}

// heat map of a finished, validated range of records (TREAT_B200_HEATMAP)
static int add_heatmap(treat_b200_ctx *ctx, uint32_t flags, const treat_b200_record *d_rec, uint64_t N, cudaStream_t st)
{
    if (!(flags & TREAT_B200_HEATMAP) || !N) return TREAT_B200_OK;
    int e = launch_heatmap(d_rec, N, ctx->td.edit_offset, (uint32_t)ctx->td.len, ctx->d_heat, ctx->num_sms, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_heatmap launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}

static int ensure_scalars(treat_b200_ctx *ctx, Slot &s)
{
    if (!s.scalars) CK(ctx, cudaMalloc((void **)&s.scalars, kScalarBytes));
    if (!s.h_scalars) CK(ctx, cudaMallocHost((void **)&s.h_scalars, kScalarBytes));
    return TREAT_B200_OK;
}

static int fill_pack_params(treat_b200_ctx *ctx, PackParams &pp, const uint8_t *d_seqs, const uint64_t *d_off, uint64_t off_bias,
                            const treat_b200_packed *dst, uint32_t *d_scalars, const ReadSpan *span, bool padded)
{
    pp = PackParams{};
    pp.seqs = d_seqs;
    pp.off = d_off;
    pp.off_bias = off_bias;
    pp.N = dst->N;
    pp.n = dst->d_n;
    pp.flags = dst->d_flags;
    pp.bases = dst->d_bases;
    pp.counts = dst->d_counts;
    pp.raw_len = dst->d_raw_len;
    pp.bases_stride = dst->bases_stride;
    pp.counts_stride = dst->counts_stride;
    pp.stage_cap = a16(dst->max_len + 1 + 48);  // room for the zeroed tails the word-wise output reads
    pp.lut = ctx->td.lut;
    pp.edit_base = ctx->td.edit_base;
    pp.scalars = d_scalars;
    pp.tpk = (!dst->wide && !ctx->tmpl_wide) ? (const uint32_t *)ctx->d_tpk.p : nullptr;
    pp.tm = (uint32_t)ctx->td.m;
    pp.padded = padded && ((uintptr_t)d_seqs & 31u) == 0 ? 1u : 0u;
    if (span && !dst->wide) {  // k_pack16 only
        pp.begin = span->begin;
        pp.len = span->len;
        pp.span = span->span;
    } else if (span) {
        return fail(ctx, TREAT_B200_ERR_INVALID, "pack: in-place reads are not supported on the wide path");
    }
    return TREAT_B200_OK;
}

static int do_pack(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_off, uint64_t off_bias,
                   const treat_b200_packed *dst, uint32_t *d_scalars, cudaStream_t st, const ReadSpan *span = nullptr,
                   bool padded = false)
{
    PackParams pp;
    int rc = fill_pack_params(ctx, pp, d_seqs, d_off, off_bias, dst, d_scalars, span, padded);
    if (rc) return rc;
    int e = launch_pack(pp, dst->wide != 0, ctx->num_sms, ctx->max_smem, st);
    if (e == -1000) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "pack: read too long for shared-memory staging");
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_pack launch: ") + cudaGetErrorString((cudaError_t)e));
    if (dst->N) ctx->launches++;
    return TREAT_B200_OK;
}

// optional device-side plumbing of one align launch
struct AlignOpts {
    uint64_t *summary = nullptr;      // accumulate here instead of the context's summary
    uint32_t *skipped = nullptr;      // speculative launch: counts the reads it could not hold
    DevBuf *list = nullptr;           // with list_count: finish reads marked by k_pack without a DP, align only the listed rest
    uint32_t *list_count = nullptr;   //   (must be zero on entry)
    uint32_t *ops_idx = nullptr;      // with ops_count: compact the ops rows (OPS_GAPPED_ONLY on the host-buffer path)
    uint32_t *ops_count = nullptr;    //   (must be zero on entry)
    bool listed = false;              // the list is already filled (k_pack_finish ran): no k_same_bases launch
    DevBuf *long_list = nullptr;      // k_pack_finish2: reads too long for one of its passes (taken up by k_pack_finish)
    uint32_t *long_count = nullptr;   //   (must be zero on entry)
    DevBuf *rest_list = nullptr;      // k_frontier: reads it leaves for k_pack_finish2
    uint32_t *rest_count = nullptr;   //   (must be zero on entry)
    DevBuf *list2 = nullptr;          // with list2_count: k_band runs first (pure deletions, banded fill) and leaves these reads
    uint32_t *list2_count = nullptr;  //   for the full-matrix kernels
    bool list2_zeroed = false;        //   the caller cleared the counter on this stream already
};

// the parameter block of an align launch (everything but the shared-memory plan, the scratch and the list)
static int fill_align_params(treat_b200_ctx *ctx, AlignParams &ap, const treat_b200_packed *src, uint32_t ncap,
                             const uint32_t *d_read_count, uint32_t flags, treat_b200_record *d_rec, uint8_t *d_ops,
                             uint32_t ops_stride, const AlignOpts &o)
{
    uint64_t *summary_override = o.summary;
    uint32_t *d_skipped = o.skipped;
    if (src->N > 0x7fffffffull) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "align: more than 2^31 reads in one launch");
    const bool wide = src->wide != 0;
    if (!wide && ctx->tmpl_wide)
        return fail(ctx, TREAT_B200_ERR_INVALID, "align: this template needs wide-packed reads (wide=1)");
    ap = AlignParams{};
    ap.n = src->d_n;
    ap.rflags = src->d_flags;
    ap.bases = src->d_bases;
    ap.counts = src->d_counts;
    ap.raw_len = src->d_raw_len;
    ap.read_count = d_read_count;
    ap.bases_stride = src->bases_stride;
    ap.counts_stride = src->counts_stride;
    ap.N = src->N;
    ap.ncap = ncap;
    ap.t = ctx->td;
    ap.t.tb = (const uint8_t *)(wide ? ctx->d_tb_wide.p : ctx->d_tb.p);
    ap.t.tcount = wide ? ctx->d_tc32.p : ctx->d_tc8.p;
    ap.rec = d_rec;
    ap.ops = (flags & TREAT_B200_WANT_OPS) ? d_ops : nullptr;
    ap.ops_stride = ops_stride;
    ap.ops_gapped_only = (flags & TREAT_B200_OPS_GAPPED_ONLY) ? 1u : 0u;
    if (ap.ops && o.ops_count) {
        ap.ops_count = o.ops_count;                        // gapped reads are counted ...
        if (ap.ops_gapped_only) ap.ops_idx = o.ops_idx;    // ... and their rows compacted when only they are wanted
    }
    ap.summary = (flags & TREAT_B200_NO_SUMMARY) ? nullptr : (summary_override ? summary_override : ctx->d_summary);
    ap.skipped = d_skipped;
    ap.dbg = ctx->d_dbg;
    ap.bins = ctx->bins;
    ap.flags = flags;
    ap.force_single = (flags & TREAT_B200_ONE_READ_PER_WARP) ? 1u : 0u;
    ap.force_full = (flags & TREAT_B200_FULL_TRACEBACK_STORE) ? 1u : 0u;
    ap.band_rows = ctx->band_rows;
    ap.tie = ctx->tie;
    if (ap.ops) {
        if (ops_stride % 16 || ops_stride < a16(((uint32_t)ctx->td.m + ncap + 3) / 4))
            return fail(ctx, TREAT_B200_ERR_INVALID, "align: ops_stride too small or not a multiple of 16");
    }
    return TREAT_B200_OK;
}

static int do_align(treat_b200_ctx *ctx, const treat_b200_packed *src, uint32_t ncap, const uint32_t *d_read_count,
                    uint32_t flags, treat_b200_record *d_rec, uint8_t *d_ops, uint32_t ops_stride, cudaStream_t st,
                    DevBuf &scratch, const AlignOpts &o = AlignOpts())
{
    uint32_t *d_list_count = o.list_count;
    DevBuf *list = o.list;
    if (src->N == 0) return TREAT_B200_OK;
    Nvtx nv("launch k_align*: DP + traceback + assembly");
    const bool wide = src->wide != 0;
    AlignParams ap;
    int frc = fill_align_params(ctx, ap, src, ncap, d_read_count, flags, d_rec, d_ops, ops_stride, o);
    if (frc) return frc;
    // The reads k_same_bases leaves for the DP are the hard ones (truncated reads fill into the global scratch).  With more
    // than 7 columns per lane the two-reads-per-warp kernel's scratch (8 bytes per lane and step) no longer fits L2 for
    // them, so listed reads of long templates take the one-read kernel; the paired one serves whole batches -- and the few
    // reads k_band leaves (1 % of the long-gene set: two to a warp they are one round of the persistent grid instead of two).
    const bool use_list = list && d_list_count && !wide && !(flags & TREAT_B200_DP_EVERY_READ);
    static const bool no_band_env = getenv("TREAT_B200_NO_BAND") != nullptr;
    const bool want_band = !no_band_env && !wide && o.list2 && o.list2_count &&
                           !(flags & (TREAT_B200_NO_BAND | TREAT_B200_FULL_TRACEBACK_STORE | TREAT_B200_ONE_READ_PER_WARP));
    static const bool long_single = getenv("TREAT_B200_LONG_SINGLE") != nullptr;  // (measurement) the one-read kernel behind k_band too
    const bool listed_long = (use_list || (o.list2 && !(flags & TREAT_B200_DP_EVERY_READ))) && ctx->td.m > 7 * 32;
    if (listed_long && (!want_band || long_single)) ap.force_single = 1u;
    if (plan_align(ap, wide) < 0) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "align: template too long");
    uint64_t band_bytes = want_band ? band_scratch_bytes(ap, ctx->num_sms, ctx->max_smem) : 0;  // 0: k_band does not apply
    if (listed_long && want_band && band_bytes == 0 && !ap.force_single) {  // no k_band after all: every listed read is for k_align
        ap.force_single = 1u;
        if (plan_align(ap, wide) < 0) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "align: template too long");
    }
    {
        // global scratch for reads whose path may leave the band (one region per resident warp); k_band's direction words
        // use the same buffer (the two launches follow each other on this stream)
        const uint64_t need = std::max<uint64_t>(align_scratch_bytes(ap, ctx->num_sms, ctx->max_smem), band_bytes);
        if (need > scratch.cap) {
            CK(ctx, cudaStreamSynchronize(st));  // the only launches that use this buffer run on this stream
            CK(ctx, scratch.reserve(need));
        }
        ap.scratch = scratch.p;
    }
    // reads marked by k_pack as equal to the template are finished without a DP; the align kernel gets the rest.
    // *d_list_count must be zero on entry (the callers clear the slot's scalars before the pack).
    if (use_list) {
        if (!o.listed) CK(ctx, list->reserve(src->N * 4));
        ap.list = (uint32_t *)list->p;
        ap.list_count = d_list_count;
        if (!o.listed) {
            int e0 = launch_same_bases(ap, ctx->num_sms, st);
            if (e0) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_same_bases launch: ") + cudaGetErrorString((cudaError_t)e0));
            ctx->launches++;
        }
    }
    // K2b: pure deletions and reads a few edits away from the template (banded fill with an exactness certificate) are
    // finished by k_band; what it cannot certify goes to list2 for the full-matrix kernels below
    if (band_bytes) {
        CK(ctx, o.list2->reserve(src->N * 4));
        if (!o.list2_zeroed) CK(ctx, cudaMemsetAsync(o.list2_count, 0, 12, st));  // + the pure-deletion and band list counters
        const bool subseq = !(flags & TREAT_B200_DP_EVERY_READ);
        int n_band = 0;
        int eb = launch_band(ap, (uint32_t *)o.list2->p, o.list2_count, scratch.p, subseq, ctx->band_min_reads, ctx->d_dp_stats, ctx->num_sms,
                             ctx->max_smem, st, &n_band);
        if (eb) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_band launch: ") + cudaGetErrorString((cudaError_t)eb));
        ctx->launches += (uint64_t)n_band;  // k_band_classify, k_band_fill and k_band_finish (twice when the pure deletions fork)
        ap.list = (uint32_t *)o.list2->p;
        ap.list_count = o.list2_count;
    }
    int e = launch_align(ap, wide, ctx->num_sms, ctx->max_smem, st);
    if (e == -1000 || e == -1001)
        return fail(ctx, TREAT_B200_ERR_TOO_LONG, "align: read/template geometry exceeds shared memory");
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_align launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}

static bool misaligned(const void *p) { return ((uintptr_t)p & 15u) != 0; }

int treat_b200_pack_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off,
                           const treat_b200_packed *dst, void *stream)
{
    if (!ctx || !dst || (dst->N && (!d_seqs || !d_seq_off))) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "pack: set a template first (edit base)");
    if (dst->max_len > kMaxReadLen) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "pack: read longer than 16384");
    if (misaligned(dst->d_bases) || misaligned(dst->d_counts) || dst->bases_stride % 16 || dst->counts_stride % 16)
        return fail(ctx, TREAT_B200_ERR_INVALID, "pack: packed buffers must be 16-byte aligned");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = ensure_scalars(ctx, ctx->dev_slot);
    if (rc) return rc;
    CK(ctx, cudaMemsetAsync(ctx->dev_slot.scalars, 0, kScalarBytes, st));
    return do_pack(ctx, d_seqs, d_seq_off, 0, dst, ctx->dev_slot.scalars, st);
}

int treat_b200_align_packed_device(treat_b200_ctx *ctx, const treat_b200_packed *src, const uint32_t *d_read_count,
                                   uint32_t flags, treat_b200_record *d_rec, uint8_t *d_ops, uint32_t ops_stride,
                                   void *stream)
{
    if (!ctx || !src || !d_rec) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    if (misaligned(src->d_bases) || misaligned(src->d_counts) || misaligned(d_rec) || misaligned(d_ops))
        return fail(ctx, TREAT_B200_ERR_INVALID, "align: buffers must be 16-byte aligned");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    // n <= raw length: size the launch for max_len (callers that know max n can pass it as max_len)
    if (flags & TREAT_B200_USE_PACK_MARKS) {
        // the caller vouches that src was packed under the current template: finish the marked reads without a DP
        int rc = ensure_scalars(ctx, ctx->dev_slot);
        if (rc) return rc;
        CK(ctx, cudaMemsetAsync(ctx->dev_slot.scalars + 3, 0, 4, st));
        AlignOpts o;
        o.list = &ctx->dev_slot.list;
        o.list_count = ctx->dev_slot.scalars + 3;
        o.list2 = &ctx->dev_slot.list2;
        o.list2_count = ctx->dev_slot.scalars + 6;
        return do_align(ctx, src, src->max_len, d_read_count, flags, d_rec, d_ops, ops_stride, st, ctx->dev_slot.scratch, o);
    }
    {
        int rc = ensure_scalars(ctx, ctx->dev_slot);
        if (rc) return rc;
        AlignOpts o;
        o.list2 = &ctx->dev_slot.list2;
        o.list2_count = ctx->dev_slot.scalars + 6;
        return do_align(ctx, src, src->max_len, d_read_count, flags, d_rec, d_ops, ops_stride, st, ctx->dev_slot.scratch, o);
    }
}

// K1f: pack + finish the reads that need no DP in one launch (k_pack_finish); the rest is listed for do_align(listed = true)
static int do_pack_finish(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_off, uint64_t off_bias,
                          const treat_b200_packed *dst, uint32_t *d_scalars, cudaStream_t st, const ReadSpan *span, bool padded,
                          const uint32_t *d_read_count, uint32_t flags, treat_b200_record *d_rec, uint8_t *d_ops, uint32_t ops_stride,
                          const AlignOpts &o)
{
    if (dst->N == 0) return TREAT_B200_OK;
    Nvtx nv("launch k_pack_finish*: NewFragment + assembly of reads without DP");
    PackParams pp;
    int rc = fill_pack_params(ctx, pp, d_seqs, d_off, off_bias, dst, d_scalars, span, padded);
    if (rc) return rc;
    AlignParams ap;
    // sized for the template's own length: the reads finished here have n = m; the ops stride is checked by do_align
    if ((rc = fill_align_params(ctx, ap, dst, (uint32_t)ctx->td.m, d_read_count, flags, d_rec, nullptr, ops_stride, o))) return rc;
    ap.ops = (flags & TREAT_B200_WANT_OPS) ? d_ops : nullptr;
    CK(ctx, o.list->reserve(dst->N * 4));
    ap.list = (uint32_t *)o.list->p;
    ap.list_count = o.list_count;
    ap.ops_idx = nullptr;  // reads finished here have no gap: with compacted rows they own none
    ap.ops_count = nullptr;
    ap.skipped = nullptr;
    static const bool no_fuse2 = getenv("TREAT_B200_NO_FUSE2") != nullptr;
    if (!no_fuse2 && o.long_list && o.long_count) {
        CK(ctx, o.long_list->reserve(dst->N * 4));
        pp.long_list = (uint32_t *)o.long_list->p;
        pp.long_count = o.long_count;
    }
    int e;
    // K0: reads that are the template edited up to a frontier are finished from their raw characters; the fused kernels
    // below then see only what it lists
    static const bool no_frontier = getenv("TREAT_B200_NO_FRONTIER") != nullptr;
    if (!no_frontier && !(flags & TREAT_B200_NO_FRONTIER) && !no_fuse2 && o.rest_list && o.rest_count && ctx->fr_geom.image_bytes && pack_finish2_applies(pp, ap)) {
        FrontierParams fp{};
        fp.image = (const uint8_t *)ctx->d_frontier.p;
        fp.g = ctx->fr_geom;
        CK(ctx, o.rest_list->reserve(dst->N * 4));
        fp.rest = (uint32_t *)o.rest_list->p;
        fp.rest_count = o.rest_count;
        if (frontier_applies(pp, ap, fp)) {
            e = launch_frontier(pp, ap, fp, ctx->num_sms, ctx->max_smem, st);
            if (e == -1000) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "pack: template too large for k_frontier's shared memory");
            if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_frontier launch: ") + cudaGetErrorString((cudaError_t)e));
            ctx->launches++;
            pp.sel = fp.rest;
            pp.sel_count = fp.rest_count;
            if (flags & TREAT_B200_FRONTIER_ONLY) return TREAT_B200_OK;  // measurement: the rest stays unaligned
        }
    }
    if (!no_fuse2 && pack_finish2_applies(pp, ap)) {
        // one pass per read, two reads per warp for templates of up to 255 bases ...
        e = launch_pack_finish2(pp, ap, ctx->num_sms, ctx->max_smem, st);
        if (e == -1000) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "pack: template too large for the fused kernel's shared memory");
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_pack_finish2 launch: ") + cudaGetErrorString((cudaError_t)e));
        ctx->launches++;
        // ... and the multi-pass form for the reads it listed as too long, if the chunk can hold any
        const uint32_t one_pass = (ctx->td.len <= 256 ? 16u : 32u) * 32u - 31u;
        if (dst->max_len <= one_pass) return TREAT_B200_OK;
        pp.sel = pp.long_list;
        pp.sel_count = pp.long_count;
    }
    e = launch_pack_finish(pp, ap, ctx->num_sms, ctx->max_smem, st);
    if (e == -1000) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "pack: read too long for shared-memory staging");
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_pack_finish launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}

// pack + (tiny sync for max n) + align on one slot's buffers
// ops_stride == 0: the ops go to the slot's own buffer with the tightest legal stride (returned in
// *ops_stride_used), so that the host copy moves only what the batch needs.
// spec_ncap != 0: speculative, sync-free variant: the align kernel is sized for spec_ncap, accumulates
// into the slot's private summary and counts what it could not hold; the caller validates later.
static int run_chunk(treat_b200_ctx *ctx, Slot &s, const uint8_t *d_seqs, const uint64_t *d_off, uint64_t off_bias,
                     const uint32_t *d_rc, uint64_t N, uint32_t max_len, uint32_t flags, treat_b200_record *d_rec,
                     uint8_t *d_ops, uint32_t ops_stride, cudaStream_t st, bool force_wide, bool *saturated,
                     uint32_t *ops_stride_used = nullptr, uint32_t spec_ncap = 0, const ReadSpan *span = nullptr,
                     bool padded = false)
{
    const bool wide = ctx->tmpl_wide || force_wide;
    // ops into the slot's own buffer (host-buffer path): gapped-only rows are compacted there
    const bool slot_own_ops = (flags & TREAT_B200_WANT_OPS) && d_ops == nullptr;
    if (slot_own_ops && (flags & TREAT_B200_OPS_GAPPED_ONLY)) CK(ctx, s.ops_idx.reserve(N * 4));
    auto opts = [&](bool own) {
        AlignOpts o;
        o.list = &s.list;
        o.list_count = s.scalars + 3;
        o.list2 = &s.list2;
        o.list2_count = s.scalars + 6;
        o.list2_zeroed = true;  // run_chunk clears the slot's scalars first
        if (own) {
            o.ops_count = s.scalars + 4;
            if (flags & TREAT_B200_OPS_GAPPED_ONLY) o.ops_idx = (uint32_t *)s.ops_idx.p;
        }
        return o;
    };
    treat_b200_packed pk;
    int rc = treat_b200_packed_layout(ctx, N, max_len, wide, &pk);
    if (rc) return fail(ctx, rc, "align: read longer than 16384");
    CK(ctx, s.n.reserve(N * 2));
    CK(ctx, s.flags.reserve(N));
    CK(ctx, s.bases.reserve(N * (size_t)pk.bases_stride));
    CK(ctx, s.counts.reserve(N * (size_t)pk.counts_stride));
    CK(ctx, s.raw_len.reserve(N * 4));
    pk.d_n = (uint16_t *)s.n.p;
    pk.d_flags = (uint8_t *)s.flags.p;
    pk.d_bases = (uint8_t *)s.bases.p;
    pk.d_counts = (uint8_t *)s.counts.p;
    pk.d_raw_len = (uint32_t *)s.raw_len.p;
    rc = ensure_scalars(ctx, s);
    if (rc) return rc;
    CK(ctx, cudaMemsetAsync(s.scalars, 0, kScalarBytes, st));
    // K1f: reads whose bases equal the template's (or differ in one base) are finished by the kernel that strips them
    static const bool no_fuse = getenv("TREAT_B200_NO_FUSE") != nullptr;
    const bool fuse = !no_fuse && !wide && ctx->d_tpk.p && ctx->td.len <= 512 && !(flags & TREAT_B200_DP_EVERY_READ);
    if (fuse) {
        if ((flags & TREAT_B200_WANT_OPS) && d_ops == nullptr) {
            // slot-owned rows: the fused kernel zeroes the rows of the reads it finishes, so they exist before it runs
            // (sized for the launch's largest possible n: the speculative bound, or the longest raw read)
            if (ops_stride == 0) ops_stride = a16(((uint32_t)ctx->td.m + (spec_ncap ? std::min(spec_ncap, max_len) : max_len) + 3) / 4);
            CK(ctx, s.ops.reserve(N * (size_t)ops_stride));
            d_ops = (uint8_t *)s.ops.p;
        }
        // its counters go to the slot's private summary: they join the context's once the chunk is known to be final (a
        // saturated edit-base run sends the chunk to the wide path, a speculative launch is validated later)
        uint64_t *slot_sum = nullptr;
        if (!(flags & TREAT_B200_NO_SUMMARY)) {
            const size_t sl = treat_b200_summary_len(ctx) * 8;
            if (s.summary.cap < sl) {
                CK(ctx, s.summary.reserve(sl));
                CK(ctx, cudaMemsetAsync(s.summary.p, 0, s.summary.cap, st));  // all of it: a later, larger template uses more of the buffer
            }
            slot_sum = (uint64_t *)s.summary.p;
        }
        AlignOpts o = opts(false);
        o.summary = slot_sum;
        o.long_list = &s.long_list;
        o.long_count = s.scalars + 5;
        o.rest_list = &s.rest_list;
        o.rest_count = s.scalars + 9;
        rc = do_pack_finish(ctx, d_seqs, d_off, off_bias, &pk, s.scalars, st, span, padded, d_rc, flags, d_rec, d_ops, ops_stride, o);
    } else {
        rc = do_pack(ctx, d_seqs, d_off, off_bias, &pk, s.scalars, st, span, padded);
    }
    if (rc) return rc;
    if (spec_ncap) {
        const uint32_t ncap = std::min(spec_ncap, max_len);
        if ((flags & TREAT_B200_WANT_OPS) && d_ops == nullptr) {
            if (ops_stride == 0) ops_stride = a16(((uint32_t)ctx->td.m + ncap + 3) / 4);
            CK(ctx, s.ops.reserve(N * (size_t)ops_stride));
            d_ops = (uint8_t *)s.ops.p;
        }
        if (ops_stride_used) *ops_stride_used = ops_stride;
        uint64_t *slot_sum = nullptr;
        if (!(flags & TREAT_B200_NO_SUMMARY)) {
            const size_t sl = treat_b200_summary_len(ctx) * 8;
            if (s.summary.cap < sl) {
                CK(ctx, s.summary.reserve(sl));
                CK(ctx, cudaMemsetAsync(s.summary.p, 0, s.summary.cap, st));  // all of it: a later, larger template uses more of the buffer
            }
            slot_sum = (uint64_t *)s.summary.p;
        }
        AlignOpts o = opts(slot_own_ops);
        o.summary = slot_sum;
        o.skipped = s.scalars + 2;
        o.listed = fuse;
        if (fuse && (flags & TREAT_B200_SKIP_DP)) return TREAT_B200_OK;
        rc = do_align(ctx, &pk, ncap, d_rc, flags, d_rec, d_ops, ops_stride, st, s.scratch, o);
        if (rc) return rc;
        CK(ctx, cudaMemcpyAsync(s.h_scalars, s.scalars, kScalarBytes, cudaMemcpyDeviceToHost, st));
        return TREAT_B200_OK;
    }
    PEEK(ctx, s.h_scalars, s.scalars, 8, st);
    CK(ctx, cudaStreamSynchronize(st));  // 8 bytes: max n sizes the align kernel's shared memory
    const uint32_t max_n = s.h_scalars[0];
    ctx->ncap_hint = std::max(ctx->ncap_hint, max_n);
    if (saturated) *saturated = s.h_scalars[1] != 0;
    if (fuse && !(flags & TREAT_B200_NO_SUMMARY)) {
        if (s.h_scalars[1]) {  // the chunk goes to the wide path: what the fused kernel counted is dropped
            CK(ctx, cudaMemsetAsync(s.summary.p, 0, treat_b200_summary_len(ctx) * 8, st));
        } else {
            int e = launch_merge_summary(ctx->d_summary, (uint64_t *)s.summary.p, (uint32_t)treat_b200_summary_len(ctx), st);
            if (e) return fail(ctx, TREAT_B200_ERR_CUDA, "k_merge_summary launch failed");
            ctx->launches++;
        }
    }
    if (!wide && s.h_scalars[1]) return TREAT_B200_OK;  // caller re-runs this chunk on the wide path
    if ((flags & TREAT_B200_WANT_OPS) && d_ops == nullptr) {
        if (ops_stride == 0) ops_stride = a16(((uint32_t)ctx->td.m + max_n + 3) / 4);
        CK(ctx, s.ops.reserve(N * (size_t)ops_stride));
        d_ops = (uint8_t *)s.ops.p;
    }
    if (ops_stride_used) *ops_stride_used = ops_stride;
    AlignOpts oa = opts(slot_own_ops);
    oa.listed = fuse;
    if (fuse && (flags & TREAT_B200_SKIP_DP)) return TREAT_B200_OK;
    rc = do_align(ctx, &pk, max_n, d_rc, flags, d_rec, d_ops, ops_stride, st, s.scratch, oa);
    if (rc) return rc;
    return add_heatmap(ctx, flags, d_rec, N, st);  // a speculative launch is added once it is validated
}

// pinned landing buffers for `rows` compacted ops rows of a slot
static int reserve_gap_rows(treat_b200_ctx *ctx, Slot &s, uint32_t rows)
{
    const size_t need_rows = (size_t)rows * s.dev_stride, need_idx = (size_t)rows * 4;
    if (need_rows > s.gap_rows_cap) {
        CK(ctx, cudaStreamSynchronize(s.stream));
        if (s.h_gap_rows) cudaFreeHost(s.h_gap_rows);
        s.h_gap_rows = nullptr;
        s.gap_rows_cap = 0;
        CK(ctx, cudaMallocHost((void **)&s.h_gap_rows, need_rows + need_rows / 4 + 4096));
        s.gap_rows_cap = need_rows + need_rows / 4 + 4096;
    }
    if (need_idx > s.gap_idx_cap) {
        CK(ctx, cudaStreamSynchronize(s.stream));
        if (s.h_gap_idx) cudaFreeHost(s.h_gap_idx);
        s.h_gap_idx = nullptr;
        s.gap_idx_cap = 0;
        CK(ctx, cudaMallocHost((void **)&s.h_gap_idx, need_idx + need_idx / 4 + 4096));
        s.gap_idx_cap = need_idx + need_idx / 4 + 4096;
    }
    return TREAT_B200_OK;
}

int treat_b200_align_batch_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off,
                                  const uint32_t *d_read_count, uint64_t N, uint32_t max_len, uint32_t flags,
                                  treat_b200_record *d_rec, uint8_t *d_ops, uint32_t ops_stride, void *stream)
{
    Nvtx nv_call("treat_b200_align_batch_device");
    if (!ctx || (N && (!d_seqs || !d_seq_off || !d_rec))) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    if (misaligned(d_rec) || misaligned(d_ops)) return fail(ctx, TREAT_B200_ERR_INVALID, "align: rec/ops must be 16-byte aligned");
    if (N == 0) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    bool sat = false;
    int rc = run_chunk(ctx, ctx->dev_slot, d_seqs, d_seq_off, 0, d_read_count, N, max_len, flags, d_rec, d_ops,
                       ops_stride, st, false, &sat);
    if (rc) return rc;
    if (sat && !ctx->tmpl_wide)
        rc = run_chunk(ctx, ctx->dev_slot, d_seqs, d_seq_off, 0, d_read_count, N, max_len, flags, d_rec, d_ops,
                       ops_stride, st, true, nullptr);
    return rc;
}

int treat_b200_align_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off,
                           const uint32_t *read_count, uint64_t N, uint32_t flags, treat_b200_record *rec, uint8_t *ops,
                           uint32_t ops_stride)
{
    if (!ctx || (N && (!seqs || !seq_off || !rec))) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    if ((flags & TREAT_B200_WANT_OPS) && !ops) return fail(ctx, TREAT_B200_ERR_INVALID, "align: WANT_OPS without ops buffer");
    if (N == 0) return TREAT_B200_OK;
    Nvtx nv_call("treat_b200_align_batch");
    CK(ctx, cudaSetDevice(ctx->device));
    const bool want_ops = (flags & TREAT_B200_WANT_OPS) != 0;
    // OPS_GAPPED_ONLY: only gapped reads' rows cross the bus, compacted on the device and placed by the host.  Placing
    // rows costs host time per row, so a chunk is run that way only while few reads (< 1/32 so far) have gaps;
    // otherwise every row is copied (rows of gapless reads then arrive as zeros, which the contract allows).
    const bool sparse_ok = want_ops && (flags & TREAT_B200_OPS_GAPPED_ONLY);
    for (int i = 0; i < kSlots; i++) {
        if (!ctx->slot[i].stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking));
        ctx->slot[i].pending = ctx->slot[i].scatter = ctx->slot[i].unstage = false;  // nothing carries over from a call that failed half-way
    }
    // Pageable caller buffers (a cgo caller's Go slices, numpy arrays): cudaMemcpyAsync would stage them inside the driver
    // and block this thread, i.e. no overlap at all (35.7 instead of 6.2 ms per 1M reads).  Such buffers go through pinned
    // staging buffers of the slot instead, filled / emptied by a few host copy threads.
    auto pageable = [](const void *p) {
        if (!p) return false;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return true;
        }
        return at.type == cudaMemoryTypeUnregistered;
    };
    static const bool no_staging = getenv("TREAT_B200_NO_STAGING") != nullptr;
    const bool stage_in = !no_staging && (pageable(seqs) || pageable(seq_off) || pageable(read_count));
    const bool stage_out = !no_staging && (pageable(rec) || (want_ops && pageable(ops)));
    if (stage_in || stage_out) {
        static const int copy_threads = getenv("TREAT_B200_COPY_THREADS") ? std::max(0, atoi(getenv("TREAT_B200_COPY_THREADS")))
                                                                         : (int)std::min(5u, std::max(1u, std::thread::hardware_concurrency() / 2));
        ctx->copy_pool.start(copy_threads);
    }
    auto reserve_pinned = [&](uint8_t *&buf, size_t &cap, size_t need) -> int {
        if (need <= cap) return TREAT_B200_OK;
        if (buf) cudaFreeHost(buf);
        buf = nullptr;
        cap = 0;
        CK(ctx, cudaMallocHost((void **)&buf, need + need / 4 + 4096));
        cap = need + need / 4 + 4096;
        return TREAT_B200_OK;
    };

    // seq_off and read_count of the whole call in two copies (instead of two small ones per chunk); the chunks' streams
    // wait for them through an event
    {
        const size_t ob = (N + 1) * 8, cb = read_count ? N * 4 : 0;
        CK(ctx, cudaStreamSynchronize(ctx->stream));  // a previous call's copies out of h_meta are done
        CK(ctx, ctx->d_off_all.reserve(ob));
        if (cb) CK(ctx, ctx->d_rc_all.reserve(cb));
        const void *src_off = seq_off, *src_rc = read_count;
        if (stage_in) {
            int r = reserve_pinned(ctx->h_meta, ctx->h_meta_cap, ob + cb);
            if (r) return r;
            ctx->copy_pool.copy(ctx->h_meta, seq_off, ob, true);
            if (cb) ctx->copy_pool.copy(ctx->h_meta + ob, read_count, cb, true);
            src_off = ctx->h_meta;
            src_rc = ctx->h_meta + ob;
        }
        CK(ctx, cudaMemcpyAsync(ctx->d_off_all.p, src_off, ob, cudaMemcpyHostToDevice, ctx->stream));
        if (cb) CK(ctx, cudaMemcpyAsync(ctx->d_rc_all.p, src_rc, cb, cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->meta_done) CK(ctx, cudaEventCreateWithFlags(&ctx->meta_done, cudaEventDisableTiming));
        CK(ctx, cudaEventRecord(ctx->meta_done, ctx->stream));
    }
    int rc = TREAT_B200_OK;
    uint64_t chunk_idx = 0;
    static const bool trace = getenv("TREAT_B200_TRACE") != nullptr;
    // tuning knobs (defaults: 65536 reads per chunk, 4 chunks in flight)
    static const uint64_t chunk_reads = getenv("TREAT_B200_CHUNK") ? std::max(1024ull, strtoull(getenv("TREAT_B200_CHUNK"), nullptr, 10)) : kChunkReads;
    static const int n_slots = getenv("TREAT_B200_SLOTS") ? std::min(kSlots, std::max(1, atoi(getenv("TREAT_B200_SLOTS")))) : kSlots;
    static const bool no_spec = getenv("TREAT_B200_NO_SPECULATION") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_wait_slot = 0, t_enqueue = 0;
    int redone = 0;

    // copy a finished chunk's results to the caller's buffers (async on the slot stream)
    auto copy_out = [&](Slot &s) -> int {
        const bool sparse = s.sparse;
        if (stage_out) {  // into the slot's pinned landing buffer; unstage_out() moves it on once the stream is idle
            const size_t rb = s.cn * sizeof(treat_b200_record), ob = want_ops && !sparse ? s.cn * (size_t)s.dev_stride : 0;
            int r = reserve_pinned(s.land, s.land_cap, rb + ob);
            if (r) return r;
            CK(ctx, cudaMemcpyAsync(s.land, s.rec.p, rb, cudaMemcpyDeviceToHost, s.stream));
            if (ob) CK(ctx, cudaMemcpyAsync(s.land + rb, s.ops.p, ob, cudaMemcpyDeviceToHost, s.stream));
            s.unstage = true;
        } else {
            CK(ctx, cudaMemcpyAsync(rec + s.c0, s.rec.p, s.cn * sizeof(treat_b200_record), cudaMemcpyDeviceToHost, s.stream));
            // only the bytes the batch can use cross the bus: rows of dev_stride <= ops_stride bytes
            if (want_ops && !sparse)
                CK(ctx, cudaMemcpy2DAsync(ops + s.c0 * (size_t)ops_stride, ops_stride, s.ops.p, s.dev_stride, s.dev_stride, s.cn,
                                          cudaMemcpyDeviceToHost, s.stream));
        }
        if (sparse) {
            // the number of gapped reads is only known on the device: copy a guess now, the rest (if any) in place_rows
            const uint32_t head = (uint32_t)std::min<uint64_t>(s.cn, std::max<uint32_t>(4096u, 2u * ctx->gap_hint));
            int r = reserve_gap_rows(ctx, s, head);
            if (r) return r;
            CK(ctx, cudaMemcpyAsync(s.h_gap_idx, s.ops_idx.p, head * 4ull, cudaMemcpyDeviceToHost, s.stream));
            CK(ctx, cudaMemcpyAsync(s.h_gap_rows, s.ops.p, head * (size_t)s.dev_stride, cudaMemcpyDeviceToHost, s.stream));
            s.gap_head = head;
        }
        if (want_ops) {  // the number of gapped reads of the chunk, for place_rows and for the next chunks' choice
            CK(ctx, cudaMemcpyAsync(s.h_scalars, s.scalars, kScalarBytes, cudaMemcpyDeviceToHost, s.stream));
            s.scatter = true;
        }
        return TREAT_B200_OK;
    };
    // pageable outputs: the chunk's records (and full ops rows) from the landing buffer to the caller (stream idle)
    auto unstage_out = [&](Slot &s) -> int {
        if (!s.unstage) return TREAT_B200_OK;
        CK(ctx, cudaStreamSynchronize(s.stream));
        s.unstage = false;
        const size_t rb = s.cn * sizeof(treat_b200_record);
        ctx->copy_pool.copy(rec + s.c0, s.land, rb);
        if (want_ops && !s.sparse) {
            if (s.dev_stride == ops_stride) {
                ctx->copy_pool.copy(ops + s.c0 * (size_t)ops_stride, s.land + rb, s.cn * (size_t)ops_stride);
            } else {
                for (uint64_t i = 0; i < s.cn; i++)
                    memcpy(ops + (s.c0 + i) * (size_t)ops_stride, s.land + rb + i * (size_t)s.dev_stride, s.dev_stride);
            }
        }
        return TREAT_B200_OK;
    };
    // OPS_GAPPED_ONLY: move the compacted rows of a finished chunk (stream idle) into the caller's ops buffer
    auto place_rows = [&](Slot &s) -> int {
        if (!s.scatter) return TREAT_B200_OK;
        s.scatter = false;
        const uint32_t cnt = s.h_scalars[4];
        ctx->gap_hint = std::max(ctx->gap_hint, (uint32_t)(cnt * (uint64_t)chunk_reads / std::max<uint64_t>(s.cn, 1)));
        if (!s.sparse) return TREAT_B200_OK;  // every row was copied
        if (cnt > s.gap_head) {  // more gapped reads than guessed: fetch the remaining rows now
            const uint32_t head = s.gap_head;
            std::vector<uint8_t> keep_rows(s.h_gap_rows, s.h_gap_rows + head * (size_t)s.dev_stride);
            std::vector<uint32_t> keep_idx(s.h_gap_idx, s.h_gap_idx + head);
            int r = reserve_gap_rows(ctx, s, cnt);
            if (r) return r;
            memcpy(s.h_gap_rows, keep_rows.data(), keep_rows.size());
            memcpy(s.h_gap_idx, keep_idx.data(), keep_idx.size() * 4);
            CK(ctx, cudaMemcpyAsync(s.h_gap_idx + head, (uint32_t *)s.ops_idx.p + head, (cnt - head) * 4ull, cudaMemcpyDeviceToHost, s.stream));
            CK(ctx, cudaMemcpyAsync(s.h_gap_rows + head * (size_t)s.dev_stride, (uint8_t *)s.ops.p + head * (size_t)s.dev_stride,
                                    (cnt - head) * (size_t)s.dev_stride, cudaMemcpyDeviceToHost, s.stream));
            CK(ctx, cudaStreamSynchronize(s.stream));
        }
        for (uint32_t k = 0; k < cnt; k++) {
            uint8_t *dst = ops + (s.c0 + s.h_gap_idx[k]) * (size_t)ops_stride;
            memcpy(dst, s.h_gap_rows + k * (size_t)s.dev_stride, s.dev_stride);
        }
        return TREAT_B200_OK;
    };
    // pack + align with the exact max n (one 8-byte sync), wide re-run if an edit-base run saturated
    auto run_exact = [&](Slot &s) -> int {
        const uint32_t *d_rc = s.drc;
        bool sat = false;
        const uint32_t cflags = s.sparse ? flags : flags & ~TREAT_B200_OPS_GAPPED_ONLY;
        int r = run_chunk(ctx, s, (const uint8_t *)s.seqs.p, s.doff, s.b0, d_rc, s.cn, s.max_len, cflags,
                          (treat_b200_record *)s.rec.p, nullptr, 0, s.stream, false, &sat, &s.dev_stride, 0, nullptr, true);
        if (r == TREAT_B200_OK && sat && !ctx->tmpl_wide)
            r = run_chunk(ctx, s, (const uint8_t *)s.seqs.p, s.doff, s.b0, d_rc, s.cn, s.max_len, cflags,
                          (treat_b200_record *)s.rec.p, nullptr, 0, s.stream, true, nullptr, &s.dev_stride, 0, nullptr, true);
        return r;
    };
    // a speculative chunk is final once its stream is idle and nothing was skipped or saturated
    auto settle = [&](Slot &s) -> int {
        if (!s.pending && !s.scatter) return TREAT_B200_OK;
        CK(ctx, cudaStreamSynchronize(s.stream));
        if (!s.pending) return place_rows(s);
        s.pending = false;
        const uint32_t max_n = s.h_scalars[0], sat = s.h_scalars[1], skipped = s.h_scalars[2];
        ctx->ncap_hint = std::max(ctx->ncap_hint, max_n);
        const bool bad = skipped != 0 || (sat && !ctx->tmpl_wide);
        if (!bad) {
            if (!(flags & TREAT_B200_NO_SUMMARY)) {
                int e = launch_merge_summary(ctx->d_summary, (uint64_t *)s.summary.p, (uint32_t)treat_b200_summary_len(ctx), s.stream);
                if (e) return fail(ctx, TREAT_B200_ERR_CUDA, "k_merge_summary launch failed");
                ctx->launches++;
            }
            int hr = add_heatmap(ctx, flags, (const treat_b200_record *)s.rec.p, s.cn, s.stream);
            if (hr) return hr;
            return place_rows(s);
        }
        redone++;
        s.scatter = false;  // the rows copied so far belong to the discarded launch
        if (!(flags & TREAT_B200_NO_SUMMARY))
            CK(ctx, cudaMemsetAsync(s.summary.p, 0, treat_b200_summary_len(ctx) * 8, s.stream));
        int r = run_exact(s);
        if (r) return r;
        if ((r = copy_out(s))) return r;
        if (!s.scatter) return TREAT_B200_OK;
        CK(ctx, cudaStreamSynchronize(s.stream));
        return place_rows(s);
    };

    // one chunk: wait for its slot, copy the reads in, launch, start the copies out.  Every failure returns from the
    // lambda only; the caller drains the streams before it gives the error back (copies into the caller's buffers may
    // still be in flight).
    auto enqueue_chunk = [&](uint64_t c0) -> int {
        int rc = TREAT_B200_OK;
        const uint64_t cn = std::min<uint64_t>(chunk_reads, N - c0);
        Slot &s = ctx->slot[chunk_idx % n_slots];
        cudaStream_t st = s.stream;
        double t0 = now();
        {
            Nvtx nv("chunk: wait for slot + validate + deliver");
            CK(ctx, cudaStreamSynchronize(st));  // the slot's previous chunk (incl. its D2H) is done
            if ((rc = settle(s))) return rc;
            if ((rc = unstage_out(s))) return rc;
        }
        Nvtx nv_enq("chunk: H2D + pack/align launches + D2H");
        double t1 = now();
        t_wait_slot += t1 - t0;
        const uint64_t b0 = seq_off[c0], b1 = seq_off[c0 + cn];
        uint32_t max_len = 0;
        for (uint64_t i = c0; i < c0 + cn; i++) {
            const uint64_t l = seq_off[i + 1] - seq_off[i];
            if (seq_off[i + 1] < seq_off[i] || l > kMaxReadLen)
                return fail(ctx, l > kMaxReadLen ? TREAT_B200_ERR_TOO_LONG : TREAT_B200_ERR_INVALID,
                            "align: read offsets not monotone or read longer than 16384");
            max_len = std::max<uint32_t>(max_len, (uint32_t)l);
        }
        const uint32_t need_stride = treat_b200_ops_stride(ctx, max_len);
        if (want_ops && (ops_stride < need_stride || ops_stride % 16))
            return fail(ctx, TREAT_B200_ERR_INVALID, "align: ops_stride too small (see treat_b200_ops_stride)");
        CK(ctx, s.seqs.reserve(b1 - b0 + 32));
        CK(ctx, s.rec.reserve(cn * sizeof(treat_b200_record)));
        const uint8_t *src_seqs = seqs + b0;
        if (stage_in) {  // the slot's previous chunk is done, so its staging buffer is free
            const size_t sb = (size_t)(b1 - b0);
            if ((rc = reserve_pinned(s.h_in, s.h_in_cap, sb))) return rc;
            ctx->copy_pool.copy(s.h_in, src_seqs, sb, true);
            src_seqs = s.h_in;
        }
        if (b1 > b0) CK(ctx, cudaMemcpyAsync(s.seqs.p, src_seqs, b1 - b0, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaStreamWaitEvent(st, ctx->meta_done, 0));  // the call's seq_off / read_count are on the device
        s.doff = (const uint64_t *)ctx->d_off_all.p + c0;
        s.drc = read_count ? (const uint32_t *)ctx->d_rc_all.p + c0 : nullptr;
        s.c0 = c0;
        s.cn = cn;
        s.b0 = b0;
        s.max_len = max_len;
        s.sparse = sparse_ok && (uint64_t)ctx->gap_hint * 32 <= chunk_reads;
        if (ctx->ncap_hint && !ctx->tmpl_wide && !no_spec) {
            // no host sync: size the align kernel from the largest n seen so far (+ slack) and validate later
            rc = run_chunk(ctx, s, (const uint8_t *)s.seqs.p, s.doff, b0, s.drc, cn, max_len,
                           s.sparse ? flags : flags & ~TREAT_B200_OPS_GAPPED_ONLY, (treat_b200_record *)s.rec.p,
                           nullptr, 0, st, false, nullptr, &s.dev_stride, ctx->ncap_hint + 8, nullptr, true);
            s.pending = rc == TREAT_B200_OK;
        } else {
            rc = run_exact(s);
        }
        if (rc) return rc;
        rc = copy_out(s);
        t_enqueue += now() - t1;
        return rc;
    };
    for (uint64_t c0 = 0; c0 < N && rc == TREAT_B200_OK; c0 += chunk_reads, chunk_idx++) rc = enqueue_chunk(c0);
    const double t_loop = now();
    for (int i = 0; i < kSlots && rc == TREAT_B200_OK; i++) {
        rc = settle(ctx->slot[i]);
        if (rc == TREAT_B200_OK) rc = unstage_out(ctx->slot[i]);
    }
    if (rc != TREAT_B200_OK) {
        // error exit: nothing of this call may still be writing into the caller's buffers when it returns, and what
        // unsettled speculative chunks counted privately must not reach the context's summary in a later call
        const std::string keep = ctx->err;
        for (int i = 0; i < kSlots; i++) {
            Slot &s = ctx->slot[i];
            if (s.stream) cudaStreamSynchronize(s.stream);
            if (s.summary.p && s.stream) cudaMemsetAsync(s.summary.p, 0, s.summary.cap, s.stream);
            s.pending = s.scatter = s.unstage = false;
        }
        cudaStreamSynchronize(ctx->stream);
        cudaGetLastError();
        ctx->err = keep;
    }
    for (int i = 0; i < kSlots; i++) {
        cudaError_t e = cudaStreamSynchronize(ctx->slot[i].stream);
        if (e != cudaSuccess && rc == TREAT_B200_OK) {
            ctx->err = std::string("stream sync: ") + cudaGetErrorString(e);
            rc = TREAT_B200_ERR_CUDA;
        }
    }
    if (trace)
        fprintf(stderr, "[treat_b200] align_batch N=%llu chunks=%llu total %.2f ms: wait_slot %.2f enqueue %.2f drain %.2f, chunks re-run %d\n",
                (unsigned long long)N, (unsigned long long)chunk_idx, now() - t_begin, t_wait_slot, t_enqueue, now() - t_loop, redone);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// exact-duplicate collapse + alignment of the distinct sequences (collapse_kernel.cu)
// ---------------------------------------------------------------------------------------------
int treat_b200_collapse_align_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off, const uint32_t *read_count,
                                    uint64_t N, uint32_t flags, treat_b200_record *rec, uint8_t *ops, uint32_t ops_stride,
                                    uint32_t *first, uint32_t *member, uint64_t *n_unique)
{
    if (!ctx || !n_unique || (N && (!seqs || !seq_off || !rec || !first))) return TREAT_B200_ERR_INVALID;
    *n_unique = 0;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "collapse: no template");
    const bool want_ops = (flags & TREAT_B200_WANT_OPS) != 0;
    if (want_ops && !ops) return fail(ctx, TREAT_B200_ERR_INVALID, "collapse: WANT_OPS without ops buffer");
    if (N == 0) return TREAT_B200_OK;
    if (N > 0x7fffffffull) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "collapse: more than 2^31 reads in one call");
    Nvtx nv_call("treat_b200_collapse_align_batch");
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto &c = ctx->col;
    const uint64_t b0 = seq_off[0], total = seq_off[N] - b0;
    for (uint64_t i = 0; i < N; i++) {
        const uint64_t l = seq_off[i + 1] - seq_off[i];
        if (seq_off[i + 1] < seq_off[i] || l > kMaxReadLen)
            return fail(ctx, l > kMaxReadLen ? TREAT_B200_ERR_TOO_LONG : TREAT_B200_ERR_INVALID,
                        "collapse: read offsets not monotone or read longer than 16384");
    }
    uint32_t cap = 1024;
    while (cap < 2 * N) cap <<= 1;
    const uint32_t n32 = (uint32_t)N;
    CK(ctx, cudaStreamSynchronize(st));
    CK(ctx, c.seqs.reserve(total + 64));
    CK(ctx, ctx->d_off_all.reserve((N + 1) * 8));
    if (read_count) CK(ctx, ctx->d_rc_all.reserve(N * 4));
    CK(ctx, c.key.reserve(N * 8));
    CK(ctx, c.table.reserve((size_t)cap * 8));
    CK(ctx, c.head.reserve((size_t)cap * 4));
    for (DevBuf *b : {&c.slot, &c.rep, &c.is_rep, &c.sum, &c.first, &c.len, &c.count, &c.member}) CK(ctx, b->reserve(N * 4));
    CK(ctx, c.uidx.reserve((N + 1) * 8));
    CK(ctx, c.begin.reserve(N * 8));
    CK(ctx, c.sums.reserve((N / 2048 + 4) * 8));
    CK(ctx, c.scal.reserve(64));
    if (!c.h) CK(ctx, cudaMallocHost((void **)&c.h, 64));
    if (total) CK(ctx, cudaMemcpyAsync(c.seqs.p, seqs + b0, total, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(ctx->d_off_all.p, seq_off, (N + 1) * 8, cudaMemcpyHostToDevice, st));
    if (read_count) CK(ctx, cudaMemcpyAsync(ctx->d_rc_all.p, read_count, N * 4, cudaMemcpyHostToDevice, st));
    const uint8_t *d_seqs = (const uint8_t *)c.seqs.p;
    const uint64_t *d_off = (const uint64_t *)ctx->d_off_all.p;
    const uint32_t *d_rc = read_count ? (const uint32_t *)ctx->d_rc_all.p : nullptr;
    uint32_t *d_scal = (uint32_t *)c.scal.p;  // [0] collisions, [1] longest unique read
    uint64_t nu = 0;
    uint32_t max_len = 0;
    for (uint32_t seed = 0x2545f491u, tries = 0;; seed = seed * 0x9e3779b1u + 1u, tries++) {
        if (tries == 8) return fail(ctx, TREAT_B200_ERR_CUDA, "collapse: eight hash seeds in a row gave a collision");
        CK(ctx, cudaMemsetAsync(c.table.p, 0, (size_t)cap * 8, st));
        CK(ctx, cudaMemsetAsync(c.head.p, 0xff, (size_t)cap * 4, st));
        CK(ctx, cudaMemsetAsync(c.sum.p, 0, N * 4, st));
        CK(ctx, cudaMemsetAsync(c.scal.p, 0, 64, st));
        int e = launch_collapse_hash(d_seqs, d_off, b0, n32, seed, (uint64_t *)c.key.p, ctx->num_sms, st);
        if (!e) e = launch_collapse_table((const uint64_t *)c.key.p, n32, (uint64_t *)c.table.p, cap, (uint32_t *)c.slot.p, (uint32_t *)c.head.p,
                                          ctx->num_sms, st);
        if (!e) e = launch_collapse_verify(d_seqs, d_off, b0, d_rc, n32, (const uint32_t *)c.slot.p, (const uint32_t *)c.head.p, (uint32_t *)c.rep.p,
                                           (uint32_t *)c.is_rep.p, (uint32_t *)c.sum.p, d_scal, d_scal + 1, ctx->num_sms, st);
        if (!e) e = launch_scan_u32((const uint32_t *)c.is_rep.p, N, (uint64_t *)c.uidx.p, (uint64_t *)c.sums.p, st);
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("collapse launch: ") + cudaGetErrorString((cudaError_t)e));
        ctx->launches += 7;
        PEEK(ctx, &c.h[0], d_scal, 8, st);
        PEEK(ctx, &c.h[1], (uint64_t *)c.uidx.p + N, 8, st);
        CK(ctx, cudaStreamSynchronize(st));
        if ((uint32_t)(c.h[0] & 0xffffffffu) == 0) {
            max_len = (uint32_t)(c.h[0] >> 32);
            nu = c.h[1];
            break;
        }
    }
    int e = launch_collapse_compact(d_off, b0, n32, (const uint32_t *)c.rep.p, (const uint32_t *)c.is_rep.p, (const uint64_t *)c.uidx.p,
                                    (const uint32_t *)c.sum.p, (uint32_t *)c.first.p, (uint64_t *)c.begin.p, (uint32_t *)c.len.p,
                                    (uint32_t *)c.count.p, member ? (uint32_t *)c.member.p : nullptr, ctx->num_sms, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_collapse_compact launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    const uint32_t need_stride = treat_b200_ops_stride(ctx, max_len);
    if (want_ops && (ops_stride < need_stride || ops_stride % 16))
        return fail(ctx, TREAT_B200_ERR_INVALID, "collapse: ops_stride too small (see treat_b200_ops_stride)");
    CK(ctx, c.rec.reserve(nu * sizeof(treat_b200_record)));
    if (want_ops) CK(ctx, c.ops.reserve(nu * (size_t)ops_stride));
    uint8_t *d_ops = want_ops ? (uint8_t *)c.ops.p : nullptr;
    // the distinct reads are aligned where they lie, as (begin, length) spans ...
    const ReadSpan span{(const uint64_t *)c.begin.p, (const uint32_t *)c.len.p, total};
    bool sat = false;
    int rc = TREAT_B200_OK;
    if (!ctx->tmpl_wide)
        rc = run_chunk(ctx, ctx->dev_slot, d_seqs, nullptr, 0, (const uint32_t *)c.count.p, nu, max_len, flags, (treat_b200_record *)c.rec.p, d_ops,
                       ops_stride, st, false, &sat, nullptr, 0, &span, true);
    if (rc) return rc;
    if (ctx->tmpl_wide || sat) {
        // ... unless the wide path is needed (a template outside the packed alphabet, an edit-base run of 255 or more):
        // its kernels read back-to-back reads, so the distinct ones are gathered first
        CK(ctx, c.goff.reserve((nu + 1) * 8));
        CK(ctx, c.gather.reserve(total + 64));
        e = launch_scan_u32((const uint32_t *)c.len.p, nu, (uint64_t *)c.goff.p, (uint64_t *)c.sums.p, st);
        if (!e)
            e = launch_collapse_gather(d_seqs, (const uint64_t *)c.begin.p, (const uint32_t *)c.len.p, (const uint64_t *)c.goff.p, (uint32_t)nu,
                                       (uint8_t *)c.gather.p, ctx->num_sms, st);
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("collapse gather launch: ") + cudaGetErrorString((cudaError_t)e));
        ctx->launches += 4;
        rc = run_chunk(ctx, ctx->dev_slot, (const uint8_t *)c.gather.p, (const uint64_t *)c.goff.p, 0, (const uint32_t *)c.count.p, nu, max_len, flags,
                       (treat_b200_record *)c.rec.p, d_ops, ops_stride, st, true, nullptr, nullptr, 0, nullptr, false);
        if (rc) return rc;
    }
    CK(ctx, cudaMemcpyAsync(rec, c.rec.p, nu * sizeof(treat_b200_record), cudaMemcpyDeviceToHost, st));
    if (want_ops) CK(ctx, cudaMemcpyAsync(ops, c.ops.p, nu * (size_t)ops_stride, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(first, c.first.p, nu * 4, cudaMemcpyDeviceToHost, st));
    if (member) CK(ctx, cudaMemcpyAsync(member, c.member.p, N * 4, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    *n_unique = nu;
    return TREAT_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// FASTA bytes in, records out (gofasta.SimpleParser + parseMergeCount + NewFragment + NewAlignment on the device)
// ---------------------------------------------------------------------------------------------
// stage A: copy the bytes, count the record starts of every tile, scan; the two numbers the host needs land in fa.h
static bool host_pageable(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// stage_pageable: the piece is small (stream / text calls) and the slot's previous piece is finished, so pageable source
// bytes (an mmap'ed file, a Go slice) are first copied into the slot's pinned buffer by the host copy threads
static int fasta_stage_copy_count(treat_b200_ctx *ctx, FastaDev &fa, const uint8_t *fasta, uint64_t len, cudaStream_t st,
                                  bool stage_pageable = false)
{
    Nvtx nv("fasta: H2D + record-start count");
    static const bool no_staging = getenv("TREAT_B200_NO_STAGING") != nullptr;
    const uint64_t piece = 32ull << 20;  // a multiple of the tile size
    const bool staged = stage_pageable && len && !no_staging && host_pageable(fasta);
    if (staged) {  // two halves of one pinned buffer, a 32 MB piece each
        const size_t need = (size_t)std::min<uint64_t>(len, 2 * piece);
        if (need > fa.h_stage_cap) {
            CK(ctx, cudaStreamSynchronize(st));
            if (fa.h_stage) cudaFreeHost(fa.h_stage);
            fa.h_stage = nullptr;
            fa.h_stage_cap = 0;
            const size_t want = need >= 2 * piece ? need : std::min<size_t>(2 * piece, need + need / 4 + 4096);
            CK(ctx, cudaMallocHost((void **)&fa.h_stage, want));
            fa.h_stage_cap = want;
        }
        for (cudaEvent_t &ev : fa.stage_done)
            if (!ev) CK(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        static const int copy_threads = getenv("TREAT_B200_COPY_THREADS") ? std::max(0, atoi(getenv("TREAT_B200_COPY_THREADS")))
                                                                         : (int)std::min(5u, std::max(1u, std::thread::hardware_concurrency() / 2));
        ctx->copy_pool.start(copy_threads);
    }
    fa.open = false;
    fa.compacted = false;
    fa.len = len;
    fa.N = 0;
    fa.max_len = 0;
    fa.seq_bytes = 0;
    fa.multi = false;
    if (!fa.h) CK(ctx, cudaMallocHost((void **)&fa.h, 64));
    const uint64_t ntiles = (len + kFastaTileBytes - 1) / kFastaTileBytes;
    CK(ctx, fa.file.reserve(len + 64));
    CK(ctx, fa.tile_count.reserve((ntiles + 1) * 4));
    CK(ctx, fa.tile_off.reserve((ntiles + 2) * 8));
    CK(ctx, fa.sums.reserve((ntiles / 2048 + 4) * 8));
    CK(ctx, fa.scal.reserve(64));
    uint64_t *d_first = (uint64_t *)fa.scal.p;
    CK(ctx, cudaMemsetAsync(fa.scal.p, 0xff, 8, st));  // first '>' = none
    CK(ctx, cudaMemsetAsync((uint8_t *)fa.scal.p + 8, 0, 56, st));
    fa.h[0] = 0;
    fa.h[1] = ~0ull;
    if (len == 0) return TREAT_B200_OK;
    // in pieces: the record starts of a piece are counted while the next piece is on the bus
    for (uint64_t o = 0, k = 0; o < len; o += piece, k++) {
        const uint64_t nb = std::min(piece, len - o);
        const uint8_t *src = fasta + o;
        if (staged) {  // pageable bytes: through the half of the pinned buffer whose previous copy is done
            uint8_t *half = fa.h_stage + (k & 1) * piece;
            if (k >= 2) CK(ctx, cudaEventSynchronize(fa.stage_done[k & 1]));
            ctx->copy_pool.copy(half, src, nb, true);
            src = half;
        }
        CK(ctx, cudaMemcpyAsync((uint8_t *)fa.file.p + o, src, nb, cudaMemcpyHostToDevice, st));
        if (staged) CK(ctx, cudaEventRecord(fa.stage_done[k & 1], st));
        int e = launch_fasta_mark((const uint8_t *)fa.file.p, len, o / kFastaTileBytes, (nb + kFastaTileBytes - 1) / kFastaTileBytes,
                                  (uint32_t *)fa.tile_count.p, d_first, ctx->num_sms, st);
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_fasta_mark launch: ") + cudaGetErrorString((cudaError_t)e));
        ctx->launches++;
    }
    int e = launch_scan_u32((const uint32_t *)fa.tile_count.p, ntiles, (uint64_t *)fa.tile_off.p, (uint64_t *)fa.sums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("scan launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches += 3;
    PEEK(ctx, &fa.h[0], (uint64_t *)fa.tile_off.p + ntiles, 8, st);  // line-start '>' count
    PEEK(ctx, &fa.h[1], d_first, 8, st);                             // first '>' position
    return TREAT_B200_OK;
}

// stage B (stage A complete on the host): record starts, ids, merge counts, sequence begin / length, scan; scalars land in fa.h
static int fasta_stage_index(treat_b200_ctx *ctx, FastaDev &fa, const uint8_t *fasta, cudaStream_t st)
{
    Nvtx nv("fasta: record index");
    const uint64_t len = fa.len, first = fa.h[1];
    const bool extra = first < len && !(first == 0 || fasta[first - 1] == '\n');  // the first '>' sits inside a line
    const uint64_t N = len ? fa.h[0] + (extra ? 1 : 0) : 0;
    if (N > 0x7fffffffull) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "fasta: more than 2^31 records");
    fa.N = N;
    fa.h[2] = fa.h[3] = 0;
    if (!N) return TREAT_B200_OK;
    const uint64_t ntiles = (len + kFastaTileBytes - 1) / kFastaTileBytes;
    uint64_t *d_first = (uint64_t *)fa.scal.p;
    uint32_t *d_scal = (uint32_t *)((uint8_t *)fa.scal.p + 16);
    CK(ctx, fa.start.reserve(N * 8));
    CK(ctx, fa.id_off.reserve(N * 8));
    CK(ctx, fa.id_len.reserve(N * 4));
    CK(ctx, fa.rc.reserve(N * 4));
    CK(ctx, fa.seq_begin.reserve(N * 8));
    CK(ctx, fa.seq_len.reserve(N * 4));
    CK(ctx, fa.seq_off.reserve((N + 1) * 8));
    CK(ctx, fa.sums.reserve((std::max(N, ntiles) / 2048 + 4) * 8));
    int e = launch_fasta_starts((const uint8_t *)fa.file.p, len, ntiles, (const uint64_t *)fa.tile_off.p, d_first, (uint64_t *)fa.start.p,
                                ctx->num_sms, st);
    if (!e)
        e = launch_fasta_records((const uint8_t *)fa.file.p, len, (const uint64_t *)fa.start.p, N, (uint64_t *)fa.id_off.p,
                                 (uint32_t *)fa.id_len.p, (uint32_t *)fa.rc.p, (uint64_t *)fa.seq_begin.p, (uint32_t *)fa.seq_len.p, d_scal,
                                 ctx->num_sms, st);
    if (!e) e = launch_scan_u32((const uint32_t *)fa.seq_len.p, N, (uint64_t *)fa.seq_off.p, (uint64_t *)fa.sums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("fasta index launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches += 5;
    PEEK(ctx, &fa.h[2], d_scal, 8, st);                          // longest sequence, multi-line flag
    PEEK(ctx, &fa.h[3], (uint64_t *)fa.seq_off.p + N, 8, st);   // sequence bytes
    return TREAT_B200_OK;
}

// stage B complete on the host
static void fasta_index_done(FastaDev &fa)
{
    fa.max_len = (uint32_t)(fa.h[2] & 0xffffffffu);
    fa.multi = (fa.h[2] >> 32) != 0;
    fa.seq_bytes = fa.h[3];
    fa.open = true;
}

// stage C: align records [c0, c0 + cn) of an indexed file on slot s: records (and ops) end up in the slot's device buffers.
// sparse: OPS_GAPPED_ONLY rows compacted in s.ops with s.ops_idx and the count in s.scalars[4]; else rows at ops_stride.
// spec_ncap != 0 (in-place reads only): no host sync -- the align kernel is sized for spec_ncap, accumulates into the slot's
// private summary and counts what it could not hold; the caller validates s.h_scalars later (see run_chunk).
static int fasta_align_chunk(treat_b200_ctx *ctx, FastaDev &fa, Slot &s, uint64_t c0, uint64_t cn, uint32_t flags, uint32_t ops_stride,
                             bool *in_place, uint32_t spec_ncap = 0)
{
    Nvtx nv("fasta: pack + align chunk");
    cudaStream_t st = s.stream;
    const bool want_ops = (flags & TREAT_B200_WANT_OPS) != 0;
    const bool sparse = want_ops && (flags & TREAT_B200_OPS_GAPPED_ONLY);
    CK(ctx, s.rec.reserve(cn * sizeof(treat_b200_record)));
    if (want_ops && !sparse) CK(ctx, s.ops.reserve(cn * (size_t)ops_stride));
    uint8_t *d_ops = want_ops && !sparse ? (uint8_t *)s.ops.p : nullptr;
    const uint32_t d_stride = sparse ? 0u : ops_stride;
    const uint32_t *d_rc = (const uint32_t *)fa.rc.p + c0;
    // reads with line breaks inside (or a template that needs the wide path) are first copied into one contiguous buffer
    auto compact = [&]() -> int {
        if (fa.compacted) return TREAT_B200_OK;
        CK(ctx, fa.seqs.reserve(fa.seq_bytes + 64));
        int e = launch_fasta_compact((const uint8_t *)fa.file.p, fa.len, (const uint64_t *)fa.start.p, fa.N, (const uint64_t *)fa.seq_begin.p,
                                     (const uint64_t *)fa.seq_off.p, (uint8_t *)fa.seqs.p, ctx->num_sms, st);
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_fasta_compact launch: ") + cudaGetErrorString((cudaError_t)e));
        ctx->launches++;
        fa.compacted = true;
        return TREAT_B200_OK;
    };
    int rc = TREAT_B200_OK;
    bool sat = false;
    if (*in_place) {
        const ReadSpan span{(const uint64_t *)fa.seq_begin.p + c0, (const uint32_t *)fa.seq_len.p + c0, fa.len};
        rc = run_chunk(ctx, s, (const uint8_t *)fa.file.p, nullptr, 0, d_rc, cn, fa.max_len, flags, (treat_b200_record *)s.rec.p, d_ops,
                       d_stride, st, false, &sat, &s.dev_stride, spec_ncap, &span, true);
        if (spec_ncap) return rc;
        if (rc == TREAT_B200_OK && sat) *in_place = false;  // an edit-base run >= 255: the wide kernels read contiguous reads
    }
    if (rc == TREAT_B200_OK && !*in_place) {
        if ((rc = compact())) return rc;
        rc = run_chunk(ctx, s, (const uint8_t *)fa.seqs.p, (const uint64_t *)fa.seq_off.p + c0, 0, d_rc, cn, fa.max_len, flags,
                       (treat_b200_record *)s.rec.p, d_ops, d_stride, st, false, &sat, &s.dev_stride, 0, nullptr, true);
        if (rc == TREAT_B200_OK && sat && !ctx->tmpl_wide)
            rc = run_chunk(ctx, s, (const uint8_t *)fa.seqs.p, (const uint64_t *)fa.seq_off.p + c0, 0, d_rc, cn, fa.max_len, flags,
                           (treat_b200_record *)s.rec.p, d_ops, d_stride, st, true, nullptr, &s.dev_stride, 0, nullptr, true);
    }
    return rc;
}

static int fasta_check_align(treat_b200_ctx *ctx, const FastaDev &fa, uint32_t flags, uint32_t ops_stride)
{
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    if (fa.max_len > kMaxReadLen) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "align: read longer than 16384");
    if ((flags & TREAT_B200_WANT_OPS) && (ops_stride < treat_b200_ops_stride(ctx, fa.max_len) || ops_stride % 16))
        return fail(ctx, TREAT_B200_ERR_INVALID, "align: ops_stride too small (see treat_b200_ops_stride)");
    return TREAT_B200_OK;
}

int treat_b200_fasta_upload(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint64_t *n_records, uint32_t *max_seq_len)
{
    Nvtx nv_call("treat_b200_fasta_upload");
    if (!ctx || (len && !fasta) || !n_records) return TREAT_B200_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    FastaDev &fa = ctx->fasta;
    cudaStream_t st = ctx->stream;
    *n_records = 0;
    if (max_seq_len) *max_seq_len = 0;
    int rc = fasta_stage_copy_count(ctx, fa, fasta, len, st, true);
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(st));
    if ((rc = fasta_stage_index(ctx, fa, fasta, st))) return rc;
    CK(ctx, cudaStreamSynchronize(st));
    fasta_index_done(fa);
    *n_records = fa.N;
    if (max_seq_len) *max_seq_len = fa.max_len;
    return TREAT_B200_OK;
}

int treat_b200_fasta_align(treat_b200_ctx *ctx, uint32_t flags, treat_b200_record *rec, uint8_t *ops, uint32_t ops_stride,
                           uint64_t *id_off, uint32_t *id_len, uint32_t *read_count, uint64_t *seq_off, uint32_t *seq_len)
{
    Nvtx nv_call("treat_b200_fasta_align");
    if (!ctx) return TREAT_B200_ERR_INVALID;
    FastaDev &fa = ctx->fasta;
    if (!fa.open) return fail(ctx, TREAT_B200_ERR_INVALID, "fasta_align: no file uploaded (treat_b200_fasta_upload)");
    const uint64_t N = fa.N;
    if (N == 0) return TREAT_B200_OK;
    if (!rec) return TREAT_B200_ERR_INVALID;
    const bool want_ops = (flags & TREAT_B200_WANT_OPS) != 0;
    const bool sparse = want_ops && (flags & TREAT_B200_OPS_GAPPED_ONLY);
    if (want_ops && !ops) return fail(ctx, TREAT_B200_ERR_INVALID, "align: WANT_OPS without ops buffer");
    int rc = fasta_check_align(ctx, fa, flags, ops_stride);
    if (rc) return rc;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (id_off) CK(ctx, cudaMemcpyAsync(id_off, fa.id_off.p, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (id_len) CK(ctx, cudaMemcpyAsync(id_len, fa.id_len.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (read_count) CK(ctx, cudaMemcpyAsync(read_count, fa.rc.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (seq_off) CK(ctx, cudaMemcpyAsync(seq_off, fa.seq_begin.p, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (seq_len) CK(ctx, cudaMemcpyAsync(seq_len, fa.seq_len.p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    for (int i = 0; i < 2; i++)
        if (!ctx->slot[i].stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking));
    bool in_place = !fa.multi && !ctx->tmpl_wide;
    static const uint64_t chunk = getenv("TREAT_B200_FASTA_CHUNK") ? std::max(1024ull, strtoull(getenv("TREAT_B200_FASTA_CHUNK"), nullptr, 10)) : (1ull << 20);
    uint64_t k = 0;
    for (uint64_t c0 = 0; c0 < N && rc == TREAT_B200_OK; c0 += chunk, k++) {
        const uint64_t cn = std::min(chunk, N - c0);
        Slot &s = ctx->slot[k & 1];
        cudaStream_t st = s.stream;
        CK(ctx, cudaStreamSynchronize(st));  // the slot's previous chunk, including its copies to the host
        if ((rc = fasta_align_chunk(ctx, fa, s, c0, cn, flags, ops_stride, &in_place))) break;
        CK(ctx, cudaMemcpyAsync(rec + c0, s.rec.p, cn * sizeof(treat_b200_record), cudaMemcpyDeviceToHost, st));
        if (want_ops && !sparse)
            CK(ctx, cudaMemcpyAsync(ops + c0 * (size_t)ops_stride, s.ops.p, cn * (size_t)ops_stride, cudaMemcpyDeviceToHost, st));
        if (sparse) {  // fetch the compacted rows of the gapped reads and put them where they belong
            PEEK(ctx, s.h_scalars, s.scalars, kScalarBytes, st);
            CK(ctx, cudaStreamSynchronize(st));
            const uint32_t cnt = s.h_scalars[4];
            if (cnt) {
                if ((rc = reserve_gap_rows(ctx, s, cnt))) break;
                CK(ctx, cudaMemcpyAsync(s.h_gap_idx, s.ops_idx.p, cnt * 4ull, cudaMemcpyDeviceToHost, st));
                CK(ctx, cudaMemcpyAsync(s.h_gap_rows, s.ops.p, cnt * (size_t)s.dev_stride, cudaMemcpyDeviceToHost, st));
                CK(ctx, cudaStreamSynchronize(st));
                for (uint32_t q = 0; q < cnt; q++)
                    memcpy(ops + (c0 + s.h_gap_idx[q]) * (size_t)ops_stride, s.h_gap_rows + q * (size_t)s.dev_stride, s.dev_stride);
            }
        }
    }
    for (int i = 0; i < 2; i++) {
        cudaError_t e = cudaStreamSynchronize(ctx->slot[i].stream);
        if (e != cudaSuccess && rc == TREAT_B200_OK) {
            ctx->err = std::string("stream sync: ") + cudaGetErrorString(e);
            rc = TREAT_B200_ERR_CUDA;
        }
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

static int fasta_blobs_sizes(treat_b200_ctx *ctx, Slot &s, uint64_t n);
static int fasta_blobs_write(treat_b200_ctx *ctx, Slot &s, const FastaDev &fa, bool in_place, uint64_t n);

// The same as one call that never holds the whole file on the device: the bytes are cut into pieces at record starts,
// four pieces are in flight (copy + count, index, align, deliver), and every finished piece is handed to `sink` in file order.
int treat_b200_fasta_stream(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint32_t flags, uint32_t ops_stride,
                            treat_b200_fasta_sink sink, void *user, uint64_t *n_records)
{
    Nvtx nv_call("treat_b200_fasta_stream");
    if (!ctx || (len && !fasta) || !sink) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    CK(ctx, cudaSetDevice(ctx->device));
    const bool want_ops = (flags & TREAT_B200_WANT_OPS) != 0;
    const bool sparse = want_ops && (flags & TREAT_B200_OPS_GAPPED_ONLY);
    if (n_records) *n_records = 0;
    for (int i = 0; i < kFastaSlots; i++)
        if (!ctx->fslot[i].stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->fslot[i].stream, cudaStreamNonBlocking));
    static const uint64_t piece_bytes = getenv("TREAT_B200_FASTA_PIECE") ? std::max(4096ull, strtoull(getenv("TREAT_B200_FASTA_PIECE"), nullptr, 10)) : (32ull << 20);
    // cut points: every piece but the first starts at a '>' that opens a line
    std::vector<uint64_t> cut{0};
    while (cut.back() < len) {
        uint64_t e = cut.back() + piece_bytes;
        if (e >= len) {
            cut.push_back(len);
            break;
        }
        const uint64_t at = treat::FastaNextRecordStart(fasta, len, e);  // the first record start at or behind the nominal end
        cut.push_back(at);
    }
    const int64_t npieces = (int64_t)cut.size() - 1;
    struct Piece {
        uint64_t first = 0;   // number of its first record in the file
        uint32_t stride = 0;  // bytes per ops row (ops_stride, or chosen from the piece's longest read when that is 0)
        bool in_place = true;
        bool spec = false;    // aligned without waiting for the pack kernel's longest read: validated before delivery
    };
    std::vector<Piece> pc((size_t)std::max<int64_t>(npieces, 0));
    uint64_t total = 0;
    int rc = TREAT_B200_OK;
    // pinned landing buffers of one slot (grow only)
    auto landing = [&](Slot &s, uint64_t n, uint32_t stride) -> int {
        const size_t need = n * (sizeof(treat_b200_record) + 8 + 4 + 4 + 8 + 4) + (want_ops && !sparse ? n * (size_t)stride : 0) + 256;
        if (need > s.land_cap) {
            if (s.land) cudaFreeHost(s.land);
            s.land = nullptr;
            s.land_cap = 0;
            CK(ctx, cudaMallocHost((void **)&s.land, need + need / 4));
            s.land_cap = need + need / 4;
        }
        return TREAT_B200_OK;
    };
    // copies of a piece's results into its slot's landing buffer (async on the slot stream)
    auto enqueue_results = [&](int64_t k) -> int {
        Slot &s = ctx->fslot[k % kFastaSlots];
        FastaDev &fa = ctx->fasta_slot[k % kFastaSlots];
        int r = landing(s, fa.N, pc[k].stride);
        if (r) return r;
        const uint64_t n = fa.N;
        uint8_t *l = s.land;
        CK(ctx, cudaMemcpyAsync(l, s.rec.p, n * sizeof(treat_b200_record), cudaMemcpyDeviceToHost, s.stream));
        l += n * sizeof(treat_b200_record);
        CK(ctx, cudaMemcpyAsync(l, fa.id_off.p, n * 8, cudaMemcpyDeviceToHost, s.stream));
        l += n * 8;
        CK(ctx, cudaMemcpyAsync(l, fa.seq_begin.p, n * 8, cudaMemcpyDeviceToHost, s.stream));
        l += n * 8;
        CK(ctx, cudaMemcpyAsync(l, fa.id_len.p, n * 4, cudaMemcpyDeviceToHost, s.stream));
        l += n * 4;
        CK(ctx, cudaMemcpyAsync(l, fa.rc.p, n * 4, cudaMemcpyDeviceToHost, s.stream));
        l += n * 4;
        CK(ctx, cudaMemcpyAsync(l, fa.seq_len.p, n * 4, cudaMemcpyDeviceToHost, s.stream));
        l += n * 4;
        if (want_ops && !sparse) CK(ctx, cudaMemcpyAsync(l, s.ops.p, n * (size_t)pc[k].stride, cudaMemcpyDeviceToHost, s.stream));
        CK(ctx, cudaMemcpyAsync(s.h_scalars, s.scalars, kScalarBytes, cudaMemcpyDeviceToHost, s.stream));
        return TREAT_B200_OK;
    };
    static const bool trace = getenv("TREAT_B200_TRACE") != nullptr;
    static const bool trace2 = trace && atoi(getenv("TREAT_B200_TRACE")) >= 2;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tA = 0, tB = 0, tC = 0, tD = 0, tSink = 0;
    const double t_begin = now();
    for (int64_t t = 0; t < npieces + 4 && rc == TREAT_B200_OK; t++) {
        double t0 = now();
        // stage A: piece t
        if (t < npieces) {
            Slot &s = ctx->fslot[t % kFastaSlots];
            if ((rc = h2d_in_order_begin(ctx, s, t ? &ctx->fslot[(t - 1) % kFastaSlots] : nullptr))) break;
            rc = fasta_stage_copy_count(ctx, ctx->fasta_slot[t % kFastaSlots], fasta + cut[t], cut[t + 1] - cut[t], s.stream, true);
            if (rc) break;
            if ((rc = h2d_in_order_end(ctx, s))) break;
        }
        tA += now() - t0;
        const double sA = now() - t_begin;
        t0 = now();
        // stage B: piece t - 2 (two pieces are on the bus or queued for it)
        if (t - 2 >= 0 && t - 2 < npieces) {
            const int64_t k = t - 2;
            Slot &s = ctx->fslot[k % kFastaSlots];
            CK(ctx, cudaStreamSynchronize(s.stream));
            if ((rc = fasta_stage_index(ctx, ctx->fasta_slot[k % kFastaSlots], fasta + cut[k], s.stream))) break;
        }
        tB += now() - t0;
        const double sB = now() - t_begin;
        t0 = now();
        // stage C: piece t - 3
        if (t - 3 >= 0 && t - 3 < npieces) {
            const int64_t k = t - 3;
            Slot &s = ctx->fslot[k % kFastaSlots];
            FastaDev &fa = ctx->fasta_slot[k % kFastaSlots];
            CK(ctx, cudaStreamSynchronize(s.stream));
            fasta_index_done(fa);
            pc[k].first = total;
            total += fa.N;
            if (fa.N) {
                pc[k].stride = ops_stride ? ops_stride : treat_b200_ops_stride(ctx, fa.max_len);
                if ((rc = fasta_check_align(ctx, fa, flags, pc[k].stride))) break;
                pc[k].in_place = !fa.multi && !ctx->tmpl_wide;
                pc[k].spec = pc[k].in_place && ctx->ncap_hint != 0;
                if ((rc = fasta_align_chunk(ctx, fa, s, 0, fa.N, flags, pc[k].stride, &pc[k].in_place, pc[k].spec ? ctx->ncap_hint + 8 : 0))) break;
                if ((rc = enqueue_results(k))) break;
                // MarshalBinary sizes behind the align kernels (valid unless a speculative piece has to be re-run)
                if ((flags & TREAT_B200_WANT_BLOBS) && (rc = fasta_blobs_sizes(ctx, s, fa.N))) break;
            }
        }
        tC += now() - t0;
        const double sC = now() - t_begin;
        t0 = now();
        // stage D: piece t - 4
        if (t - 4 >= 0 && t - 4 < npieces) {
            const int64_t k = t - 4;
            Slot &s = ctx->fslot[k % kFastaSlots];
            FastaDev &fa = ctx->fasta_slot[k % kFastaSlots];
            CK(ctx, cudaStreamSynchronize(s.stream));
            const uint64_t n = fa.N;
            if (n && pc[k].spec) {  // validate the speculative launch (as treat_b200_align_batch does at slot reuse)
                const uint32_t max_n = s.h_scalars[0], sat = s.h_scalars[1], skipped = s.h_scalars[2];
                ctx->ncap_hint = std::max(ctx->ncap_hint, max_n);
                if (skipped != 0 || (sat && !ctx->tmpl_wide)) {
                    if (!(flags & TREAT_B200_NO_SUMMARY)) CK(ctx, cudaMemsetAsync(s.summary.p, 0, treat_b200_summary_len(ctx) * 8, s.stream));
                    if ((rc = fasta_align_chunk(ctx, fa, s, 0, n, flags, pc[k].stride, &pc[k].in_place))) break;
                    if ((rc = enqueue_results(k))) break;
                    if ((flags & TREAT_B200_WANT_BLOBS) && (rc = fasta_blobs_sizes(ctx, s, n))) break;  // the records changed
                    CK(ctx, cudaStreamSynchronize(s.stream));
                } else {
                    if (!(flags & TREAT_B200_NO_SUMMARY)) {
                        int e = launch_merge_summary(ctx->d_summary, (uint64_t *)s.summary.p, (uint32_t)treat_b200_summary_len(ctx), s.stream);
                        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, "k_merge_summary launch failed");
                        ctx->launches++;
                    }
                    if ((rc = add_heatmap(ctx, flags, (const treat_b200_record *)s.rec.p, n, s.stream))) break;
                }
            }
            if (n && (flags & TREAT_B200_WANT_BLOBS)) {  // the records are final and the blobs' total size is on the host
                if ((rc = fasta_blobs_write(ctx, s, fa, pc[k].in_place, n))) break;
                CK(ctx, cudaStreamSynchronize(s.stream));
            }
            if (n) {
                treat_b200_fasta_piece out;
                memset(&out, 0, sizeof out);
                if (flags & TREAT_B200_WANT_BLOBS) {
                    out.blob_off = (const uint64_t *)s.tland;
                    out.blob = s.tland + (n + 1) * 8;
                }
                uint8_t *l = s.land;
                out.first_record = pc[k].first;
                out.n = n;
                out.rec = (const treat_b200_record *)l;
                l += n * sizeof(treat_b200_record);
                uint64_t *ido = (uint64_t *)l;
                l += n * 8;
                uint64_t *sqo = (uint64_t *)l;
                l += n * 8;
                for (uint64_t i = 0; i < n; i++) {  // piece offsets -> file offsets
                    ido[i] += cut[k];
                    sqo[i] += cut[k];
                }
                out.id_off = ido;
                out.seq_off = sqo;
                out.id_len = (const uint32_t *)l;
                l += n * 4;
                out.read_count = (const uint32_t *)l;
                l += n * 4;
                out.seq_len = (const uint32_t *)l;
                l += n * 4;
                out.seq_in_place = fa.multi ? 0 : 1;
                if (want_ops && !sparse) {
                    out.ops = l;
                    out.ops_stride = pc[k].stride;
                    out.n_ops_rows = n;
                } else if (sparse) {
                    const uint32_t cnt = s.h_scalars[4];
                    if (cnt) {
                        if ((rc = reserve_gap_rows(ctx, s, cnt))) break;
                        CK(ctx, cudaMemcpyAsync(s.h_gap_idx, s.ops_idx.p, cnt * 4ull, cudaMemcpyDeviceToHost, s.stream));
                        CK(ctx, cudaMemcpyAsync(s.h_gap_rows, s.ops.p, cnt * (size_t)s.dev_stride, cudaMemcpyDeviceToHost, s.stream));
                        CK(ctx, cudaStreamSynchronize(s.stream));
                    }
                    out.ops = s.h_gap_rows;
                    out.ops_stride = s.dev_stride;
                    out.n_ops_rows = cnt;
                    out.ops_index = s.h_gap_idx;
                }
                const double ts = now();
                const int stop = sink(user, &out);
                tSink += now() - ts;
                if (stop) {
                    rc = fail(ctx, TREAT_B200_ERR_INVALID, "fasta_stream: stopped by the sink");
                    break;
                }
            }
        }
        tD += now() - t0;
        if (trace2) fprintf(stderr, "[treat_b200]   step %lld: A done %.2f, B done %.2f, C done %.2f, D done %.2f ms\n", (long long)t, sA, sB, sC, now() - t_begin);
    }
    for (int i = 0; i < kFastaSlots; i++) cudaStreamSynchronize(ctx->fslot[i].stream);
    if (trace)
        fprintf(stderr, "[treat_b200] fasta_stream %lld pieces, %llu records, %.2f ms: copy+count %.2f, index %.2f, align %.2f, deliver %.2f (sink %.2f)\n",
                (long long)npieces, (unsigned long long)total, now() - t_begin, tA, tB, tC, tD, tSink);
    if (n_records) *n_records = total;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Alignment.WriteTo (align.go:388-520) for a batch on the device
// ---------------------------------------------------------------------------------------------
static void text_template(const treat_b200_ctx *ctx, TextParams &tp)
{
    tp.tb = (const uint8_t *)ctx->d_tb_wide.p;
    tp.tc32 = (const uint32_t *)ctx->d_tc32.p;
    tp.tmax = (const uint32_t *)ctx->d_tmax.p;
    tp.m = ctx->td.m;
    tp.len = ctx->td.len;
    tp.len_pad = ctx->td.len_pad;
    tp.K = ctx->td.K;
    tp.edit_base = ctx->td.edit_base;
}

// pass 1 + scan: text_off[N + 1] on the device (tp holds the reads, names, records and ops); the total lands in *h_total
static int text_sizes(treat_b200_ctx *ctx, Slot &s, TextParams &tp, uint64_t *d_text_off, cudaStream_t st)
{
    Nvtx nv("k_text: sizes + scan");
    const uint64_t N = tp.N;
    CK(ctx, s.tlen.reserve(N * 4));
    CK(ctx, s.gtot.reserve(N * 4));
    CK(ctx, s.tsums.reserve((N / 2048 + 4) * 8));
    if (!s.h_text_total) CK(ctx, cudaMallocHost((void **)&s.h_text_total, 16));
    int sr = ensure_scalars(ctx, s);
    if (sr) return sr;
    tp.tlen = (uint32_t *)s.tlen.p;
    tp.gtot = (uint32_t *)s.gtot.p;
    tp.gmax = s.scalars + 7;  // the slot's last scalar word: longest row of the batch
    CK(ctx, cudaMemsetAsync(tp.gmax, 0, 4, st));
    int e = launch_text(tp, false, ctx->num_sms, ctx->max_smem, st);
    if (e == -1000) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "text: read/template geometry exceeds shared memory");
    if (!e) e = launch_scan_u32((const uint32_t *)s.tlen.p, N, d_text_off, (uint64_t *)s.tsums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_text launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches += 4;
    PEEK(ctx, s.h_text_total, d_text_off + N, 8, st);
    PEEK(ctx, s.h_text_total + 1, tp.gmax, 4, st);
    return TREAT_B200_OK;
}

static int text_write(treat_b200_ctx *ctx, Slot &s, TextParams &tp, const uint64_t *d_text_off, uint8_t *d_text, cudaStream_t st)
{
    Nvtx nv("k_text: write");
    tp.text_off = d_text_off;
    tp.text = d_text;
    tp.g_longest = (uint32_t)s.h_text_total[1];  // pass 1's longest row (the stream was synchronised since)
    int e = launch_text(tp, true, ctx->num_sms, ctx->max_smem, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_text launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}

static int text_tw(int32_t tw) { return tw <= 0 ? 80 : tw; }  // align.go:389-391

int treat_b200_text_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off, const uint8_t *d_names,
                           const uint64_t *d_name_off, const treat_b200_record *d_rec, const uint8_t *d_ops, uint32_t ops_stride,
                           uint64_t N, uint32_t max_len, int32_t tw, uint64_t *d_text_off, uint8_t *d_text, uint64_t text_cap,
                           uint64_t *text_bytes, void *stream)
{
    Nvtx nv_call("treat_b200_text_device");
    if (!ctx || !text_bytes || (N && (!d_seqs || !d_seq_off || !d_names || !d_name_off || !d_rec || !d_ops || !d_text_off)))
        return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "text: no template");
    tw = text_tw(tw);
    if (tw < 5) return fail(ctx, TREAT_B200_ERR_INVALID, "text: tw must be at least 5 (rows wrap at tw - 4 columns)");
    if (max_len > kMaxReadLen) return fail(ctx, TREAT_B200_ERR_TOO_LONG, "text: read longer than 16384");
    if (ops_stride % 4 || ops_stride < a16(((uint32_t)ctx->td.m + max_len + 3) / 4) || misaligned(d_ops))
        return fail(ctx, TREAT_B200_ERR_INVALID, "text: ops rows as written by the align calls (see treat_b200_ops_stride)");
    *text_bytes = 0;
    if (N == 0) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    TextParams tp{};
    tp.seqs = d_seqs;
    tp.off = d_seq_off;
    tp.names = d_names;
    tp.name_off = d_name_off;
    tp.rec = d_rec;
    tp.ops = d_ops;
    tp.ops_stride = ops_stride;
    tp.N = N;
    tp.max_len = max_len;
    tp.tw = tw;
    text_template(ctx, tp);
    int rc = text_sizes(ctx, ctx->dev_slot, tp, d_text_off, st);
    if (rc) return rc;
    CK(ctx, cudaStreamSynchronize(st));
    *text_bytes = *ctx->dev_slot.h_text_total;
    if (!d_text) return TREAT_B200_OK;  // sizing call
    if (*text_bytes > text_cap) return fail(ctx, TREAT_B200_ERR_INVALID, "text: text buffer too small (call with d_text = NULL to size)");
    return text_write(ctx, ctx->dev_slot, tp, d_text_off, d_text, st);
}

// MarshalBinary blobs of the records in s.rec for the reads of an indexed FASTA piece: sizes + scan (total lands in
// s.h_text_total after a stream sync), then the write into s.text and the copy of offsets + blobs to the pinned s.tland
static int fasta_blobs_sizes(treat_b200_ctx *ctx, Slot &s, uint64_t n)
{
    cudaStream_t st = s.stream;
    CK(ctx, s.tlen.reserve(n * 4));
    CK(ctx, s.tsums.reserve((n / 2048 + 4) * 8));
    CK(ctx, s.text_off.reserve((n + 1) * 8));
    if (!s.h_text_total) CK(ctx, cudaMallocHost((void **)&s.h_text_total, 16));
    int e = launch_marshal_sizes((const treat_b200_record *)s.rec.p, n, (uint32_t *)s.tlen.p, ctx->num_sms, st);
    if (!e) e = launch_scan_u32((const uint32_t *)s.tlen.p, n, (uint64_t *)s.text_off.p, (uint64_t *)s.tsums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_marshal launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches += 4;
    PEEK(ctx, s.h_text_total, (uint64_t *)s.text_off.p + n, 8, st);
    return TREAT_B200_OK;
}
static int fasta_blobs_write(treat_b200_ctx *ctx, Slot &s, const FastaDev &fa, bool in_place, uint64_t n)
{
    cudaStream_t st = s.stream;
    const uint64_t total = *s.h_text_total, off_bytes = (n + 1) * 8;
    CK(ctx, s.text.reserve(total + 16));
    if (off_bytes + total > s.tland_cap) {
        if (s.tland) cudaFreeHost(s.tland);
        s.tland = nullptr;
        s.tland_cap = 0;
        const size_t want = off_bytes + total + (off_bytes + total) / 4 + 4096;
        CK(ctx, cudaMallocHost((void **)&s.tland, want));
        s.tland_cap = want;
    }
    MarshalParams mp{};
    mp.seqs = in_place ? (const uint8_t *)fa.file.p : (const uint8_t *)fa.seqs.p;
    if (in_place) mp.begin = (const uint64_t *)fa.seq_begin.p;
    else mp.off = (const uint64_t *)fa.seq_off.p;
    mp.rec = (const treat_b200_record *)s.rec.p;
    mp.N = n;
    mp.blob_off = (const uint64_t *)s.text_off.p;
    mp.blob = (uint8_t *)s.text.p;
    int e = launch_marshal_write(mp, ctx->num_sms, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_marshal launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(s.tland, s.text_off.p, off_bytes, cudaMemcpyDeviceToHost, st));
    if (total) CK(ctx, cudaMemcpyAsync(s.tland + off_bytes, s.text.p, total, cudaMemcpyDeviceToHost, st));
    return TREAT_B200_OK;
}

int treat_b200_marshal_device(treat_b200_ctx *ctx, const uint8_t *d_seqs, const uint64_t *d_seq_off, const treat_b200_record *d_rec,
                              uint64_t N, uint64_t *d_blob_off, uint8_t *d_blob, uint64_t blob_cap, uint64_t *blob_bytes, void *stream)
{
    Nvtx nv_call("treat_b200_marshal_device");
    if (!ctx || !blob_bytes || (N && (!d_seqs || !d_seq_off || !d_rec || !d_blob_off))) return TREAT_B200_ERR_INVALID;
    *blob_bytes = 0;
    if (N == 0) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    Slot &s = ctx->dev_slot;
    CK(ctx, s.tlen.reserve(N * 4));
    CK(ctx, s.tsums.reserve((N / 2048 + 4) * 8));
    if (!s.h_text_total) CK(ctx, cudaMallocHost((void **)&s.h_text_total, 16));
    int e = launch_marshal_sizes(d_rec, N, (uint32_t *)s.tlen.p, ctx->num_sms, st);
    if (!e) e = launch_scan_u32((const uint32_t *)s.tlen.p, N, d_blob_off, (uint64_t *)s.tsums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_marshal launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches += 4;
    PEEK(ctx, s.h_text_total, d_blob_off + N, 8, st);
    CK(ctx, cudaStreamSynchronize(st));
    *blob_bytes = *s.h_text_total;
    if (!d_blob) return TREAT_B200_OK;  // sizing call
    if (*blob_bytes > blob_cap) return fail(ctx, TREAT_B200_ERR_INVALID, "marshal: blob buffer too small (call with d_blob = NULL to size)");
    MarshalParams mp{};
    mp.seqs = d_seqs;
    mp.off = d_seq_off;
    mp.rec = d_rec;
    mp.N = N;
    mp.blob_off = d_blob_off;
    mp.blob = d_blob;
    e = launch_marshal_write(mp, ctx->num_sms, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("k_marshal launch: ") + cudaGetErrorString((cudaError_t)e));
    ctx->launches++;
    return TREAT_B200_OK;
}

// The record filter of Storage.Search (cmd/treat/storage.go:162-188, 257-286) for a batch in HBM: a mark per record, an
// exclusive scan (the match's rank, in record order), and the indices of the ranks [offset, offset + limit).
namespace {
__device__ __forceinline__ bool search_match_dev(const treat_b200_search &f, const treat_b200_record &a)
{
    if (!f.all) {
        if (f.has_mutation && a.has_mutation == 0) return false;
        if (!f.has_mutation && a.has_mutation == 1) return false;
    }
    if (f.edit_stop >= 0 && f.edit_stop != a.edit_stop) return false;
    if (f.junc_len >= 0 && f.junc_len != a.junc_len) return false;
    if (f.junc_end >= 0 && f.junc_end != a.junc_end) return false;
    if (f.has_alt && a.alt_editing == 0) return false;
    if (f.alt_region > 0 && (uint8_t)f.alt_region != a.alt_editing) return false;
    if (!f.has_alt && a.alt_editing > 0) return false;
    return true;
}
__global__ void __launch_bounds__(256) k_search_mark(const __grid_constant__ treat_b200_search f, const treat_b200_record *rec, uint64_t N,
                                                     uint32_t *mark)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x)
        mark[i] = search_match_dev(f, rec[i]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_search_write(const uint32_t *mark, const uint64_t *rank, uint64_t N, uint64_t offset, uint64_t limit,
                                                      uint64_t cap, uint64_t *idx)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!mark[i] || rank[i] < offset) continue;
        const uint64_t k = rank[i] - offset;
        if ((limit == 0 || k < limit) && k < cap) idx[k] = i;
    }
}
}  // namespace

int treat_b200_search_device(treat_b200_ctx *ctx, const treat_b200_search *f, const treat_b200_record *d_rec, uint64_t N,
                             uint64_t *d_idx, uint64_t cap, uint64_t *count, void *stream)
{
    if (!ctx || !f || !count || (N && !d_rec)) return TREAT_B200_ERR_INVALID;
    *count = 0;
    if (N == 0) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    Slot &s = ctx->dev_slot;
    CK(ctx, s.tlen.reserve(N * 4));
    CK(ctx, s.text_off.reserve((N + 1) * 8));
    CK(ctx, s.tsums.reserve((N / 2048 + 4) * 8));
    if (!s.h_text_total) CK(ctx, cudaMallocHost((void **)&s.h_text_total, 16));
    const uint64_t blocks = std::min<uint64_t>((N + 255) / 256, (uint64_t)ctx->num_sms * 8);
    k_search_mark<<<(unsigned)blocks, 256, 0, st>>>(*f, d_rec, N, (uint32_t *)s.tlen.p);
    CK(ctx, cudaGetLastError());
    int e = launch_scan_u32((const uint32_t *)s.tlen.p, N, (uint64_t *)s.text_off.p, (uint64_t *)s.tsums.p, st);
    if (e) return fail(ctx, TREAT_B200_ERR_CUDA, std::string("search scan launch: ") + cudaGetErrorString((cudaError_t)e));
    const uint64_t off = f->offset > 0 ? (uint64_t)f->offset : 0, lim = f->limit > 0 ? (uint64_t)f->limit : 0;
    if (d_idx && cap) {
        k_search_write<<<(unsigned)blocks, 256, 0, st>>>((const uint32_t *)s.tlen.p, (const uint64_t *)s.text_off.p, N, off, lim, cap, d_idx);
        CK(ctx, cudaGetLastError());
    }
    ctx->launches += 5;
    PEEK(ctx, s.h_text_total, (uint64_t *)s.text_off.p + N, 8, st);
    CK(ctx, cudaStreamSynchronize(st));
    const uint64_t total = *s.h_text_total;
    uint64_t n = total > off ? total - off : 0;
    if (lim && n > lim) n = lim;
    *count = n;
    return TREAT_B200_OK;
}

// `treat norm` for a batch in HBM (cmd/treat/storage.go:845-870): Norm = scale * float64(ReadCount) for the reads without a
// mutation, as a float64 per record and, when the batch's MarshalBinary blobs are there, into bytes 24..32 of each (big-endian).
// One multiplication in double precision: the device's IEEE result is the host's.
namespace {
__global__ void __launch_bounds__(256) k_normalize(const treat_b200_record *rec, uint64_t N, double scale, double *norm, uint8_t *blob,
                                                   const uint64_t *blob_off)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t rc = rec[i].read_count;
        const bool standard = rec[i].has_mutation == 0;
        const double v = standard ? __dmul_rn(scale, (double)rc) : 0.0;
        if (norm) norm[i] = v;
        if (blob && standard) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
            uint8_t *dst = blob + blob_off[i] + 24;
#pragma unroll
            for (int k = 0; k < 8; k++) dst[k] = (uint8_t)(bits >> (56 - 8 * k));
        }
    }
}
}  // namespace

int treat_b200_normalize_device(treat_b200_ctx *ctx, const treat_b200_record *d_rec, uint64_t N, double scale, double *d_norm,
                                uint8_t *d_blob, const uint64_t *d_blob_off, void *stream)
{
    if (!ctx || (N && !d_rec) || (d_blob && !d_blob_off)) return TREAT_B200_ERR_INVALID;
    if (N == 0 || (!d_norm && !d_blob)) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    const uint64_t blocks = std::min<uint64_t>((N + 255) / 256, (uint64_t)ctx->num_sms * 8);
    k_normalize<<<(unsigned)blocks, 256, 0, st>>>(d_rec, N, scale, d_norm, d_blob, d_blob_off);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    return TREAT_B200_OK;
}

// The loop of cmd/treat/align.go:114-120 for a caller that already holds the records (rec.Id, rec.Seq from gofasta):
// NewFragment + NewAlignment + WriteTo for reads [0, N), chunk by chunk on two slots; the text of each chunk goes to the
// sink in read order while the next chunk is being processed.
int treat_b200_text_batch(treat_b200_ctx *ctx, const uint8_t *seqs, const uint64_t *seq_off, const uint8_t *names,
                          const uint64_t *name_off, uint64_t N, uint32_t flags, int32_t tw, treat_b200_text_sink sink, void *user)
{
    Nvtx nv_call("treat_b200_text_batch");
    if (!ctx || !sink || (N && (!seqs || !seq_off || !names || !name_off))) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    tw = text_tw(tw);
    if (tw < 5) return fail(ctx, TREAT_B200_ERR_INVALID, "text: tw must be at least 5 (rows wrap at tw - 4 columns)");
    if (N == 0) return TREAT_B200_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 2; i++)
        if (!ctx->slot[i].stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking));
    const uint32_t aflags = (flags & ~(uint32_t)TREAT_B200_OPS_GAPPED_ONLY) | TREAT_B200_WANT_OPS;
    const uint64_t chunk_reads = 32768;  // ~90 MB of text per chunk for RPS12-sized reads
    struct Pending {
        bool live = false;
        uint64_t first = 0, n = 0, bytes = 0;
    } pend[2];
    auto deliver = [&](int i) -> int {
        if (!pend[i].live) return TREAT_B200_OK;
        pend[i].live = false;
        CK(ctx, cudaStreamSynchronize(ctx->slot[i].stream));
        if (sink(user, ctx->slot[i].tland, pend[i].bytes, pend[i].first, pend[i].n))
            return fail(ctx, TREAT_B200_ERR_INVALID, "text_batch: stopped by the sink");
        return TREAT_B200_OK;
    };
    int rc = TREAT_B200_OK;
    uint64_t k = 0;
    for (uint64_t c0 = 0; c0 < N && rc == TREAT_B200_OK; c0 += chunk_reads, k++) {
        const uint64_t cn = std::min(chunk_reads, N - c0);
        const int i = (int)(k & 1);
        Slot &s = ctx->slot[i];
        cudaStream_t st = s.stream;
        if ((rc = deliver(i))) break;  // chunk k - 2: its buffers are reused below
        const uint64_t b0 = seq_off[c0], b1 = seq_off[c0 + cn], nb0 = name_off[c0], nb1 = name_off[c0 + cn];
        uint32_t max_len = 0;
        for (uint64_t r = c0; r < c0 + cn; r++) {
            const uint64_t l = seq_off[r + 1] - seq_off[r];
            if (seq_off[r + 1] < seq_off[r] || name_off[r + 1] < name_off[r] || l > kMaxReadLen)
                return fail(ctx, l > kMaxReadLen ? TREAT_B200_ERR_TOO_LONG : TREAT_B200_ERR_INVALID,
                            "text_batch: offsets not monotone or read longer than 16384");
            max_len = std::max<uint32_t>(max_len, (uint32_t)l);
        }
        const uint32_t stride = treat_b200_ops_stride(ctx, max_len);
        CK(ctx, s.seqs.reserve(b1 - b0 + 32));
        CK(ctx, s.off.reserve((cn + 1) * 8));
        CK(ctx, s.names.reserve(nb1 - nb0 + 16));
        CK(ctx, s.name_off.reserve((cn + 1) * 8));
        CK(ctx, s.rec.reserve(cn * sizeof(treat_b200_record)));
        CK(ctx, s.ops.reserve(cn * (size_t)stride));
        CK(ctx, s.text_off.reserve((cn + 1) * 8));
        if (b1 > b0) CK(ctx, cudaMemcpyAsync(s.seqs.p, seqs + b0, b1 - b0, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(s.off.p, seq_off + c0, (cn + 1) * 8, cudaMemcpyHostToDevice, st));
        if (nb1 > nb0) CK(ctx, cudaMemcpyAsync(s.names.p, names + nb0, nb1 - nb0, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaMemcpyAsync(s.name_off.p, name_off + c0, (cn + 1) * 8, cudaMemcpyHostToDevice, st));
        bool sat = false;
        rc = run_chunk(ctx, s, (const uint8_t *)s.seqs.p, (const uint64_t *)s.off.p, b0, nullptr, cn, max_len, aflags,
                       (treat_b200_record *)s.rec.p, (uint8_t *)s.ops.p, stride, st, false, &sat);
        if (rc == TREAT_B200_OK && sat && !ctx->tmpl_wide)
            rc = run_chunk(ctx, s, (const uint8_t *)s.seqs.p, (const uint64_t *)s.off.p, b0, nullptr, cn, max_len, aflags,
                           (treat_b200_record *)s.rec.p, (uint8_t *)s.ops.p, stride, st, true, nullptr);
        if (rc) break;
        TextParams tp{};
        tp.seqs = (const uint8_t *)s.seqs.p;
        tp.off = (const uint64_t *)s.off.p;
        tp.off_bias = b0;
        tp.names = (const uint8_t *)s.names.p - nb0;  // only ever indexed at name_off[r] >= nb0
        tp.name_off = (const uint64_t *)s.name_off.p;
        tp.rec = (const treat_b200_record *)s.rec.p;
        tp.ops = (const uint8_t *)s.ops.p;
        tp.ops_stride = stride;
        tp.N = cn;
        tp.max_len = max_len;
        tp.tw = tw;
        text_template(ctx, tp);
        if ((rc = text_sizes(ctx, s, tp, (uint64_t *)s.text_off.p, st))) break;
        CK(ctx, cudaStreamSynchronize(st));
        const uint64_t bytes = *s.h_text_total;
        CK(ctx, s.text.reserve(bytes));
        if (bytes > s.tland_cap) {
            if (s.tland) cudaFreeHost(s.tland);
            s.tland = nullptr;
            s.tland_cap = 0;
            CK(ctx, cudaMallocHost((void **)&s.tland, bytes + bytes / 4 + 4096));
            s.tland_cap = bytes + bytes / 4 + 4096;
        }
        if ((rc = text_write(ctx, s, tp, (const uint64_t *)s.text_off.p, (uint8_t *)s.text.p, st))) break;
        CK(ctx, cudaMemcpyAsync(s.tland, s.text.p, bytes, cudaMemcpyDeviceToHost, st));
        pend[i].live = true;
        pend[i].first = c0;
        pend[i].n = cn;
        pend[i].bytes = bytes;
    }
    for (int j = 0; j < 2; j++) {
        const int i = (int)((k + j) & 1);
        const int r = rc == TREAT_B200_OK ? deliver(i) : TREAT_B200_OK;
        if (r) rc = r;
    }
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(ctx->slot[i].stream);
    return rc;
}

// `treat align -t templates.fa -f reads.fa` (cmd/treat/align.go:96-121) as one streaming call: the file is cut at record
// starts, every piece is indexed, aligned and formatted on the device, and its text is handed to `sink` in file order
// while the next piece is being processed (two pieces in flight).
int treat_b200_fasta_text(treat_b200_ctx *ctx, const uint8_t *fasta, uint64_t len, uint32_t flags, int32_t tw,
                          treat_b200_text_sink sink, void *user, uint64_t *n_records)
{
    Nvtx nv_call("treat_b200_fasta_text");
    if (!ctx || (len && !fasta) || !sink) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "align: no template");
    tw = text_tw(tw);
    if (tw < 5) return fail(ctx, TREAT_B200_ERR_INVALID, "text: tw must be at least 5 (rows wrap at tw - 4 columns)");
    CK(ctx, cudaSetDevice(ctx->device));
    if (n_records) *n_records = 0;
    for (int i = 0; i < 2; i++)
        if (!ctx->fslot[i].stream) CK(ctx, cudaStreamCreateWithFlags(&ctx->fslot[i].stream, cudaStreamNonBlocking));
    static const uint64_t piece_bytes = getenv("TREAT_B200_TEXT_PIECE") ? std::max(4096ull, strtoull(getenv("TREAT_B200_TEXT_PIECE"), nullptr, 10)) : (8ull << 20);
    const uint32_t aflags = (flags & ~(uint32_t)TREAT_B200_OPS_GAPPED_ONLY) | TREAT_B200_WANT_OPS;  // every read's ops row stays on the device
    struct Pending {
        bool live = false;
        uint64_t first = 0, n = 0, bytes = 0;
    } pend[2];
    int rc = TREAT_B200_OK;
    auto deliver = [&](int i) -> int {
        if (!pend[i].live) return TREAT_B200_OK;
        pend[i].live = false;
        CK(ctx, cudaStreamSynchronize(ctx->fslot[i].stream));
        if (sink(user, ctx->fslot[i].tland, pend[i].bytes, pend[i].first, pend[i].n))
            return fail(ctx, TREAT_B200_ERR_INVALID, "fasta_text: stopped by the sink");
        return TREAT_B200_OK;
    };
    uint64_t total = 0, k = 0;
    static const bool trace = getenv("TREAT_B200_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tDeliver = 0, tCopy = 0, tIndex = 0, tAlign = 0, tSizes = 0, tWrite = 0;
    const double t_begin = now();
    for (uint64_t at = 0; at < len && rc == TREAT_B200_OK; k++) {
        uint64_t end = at + piece_bytes;
        end = end >= len ? len : treat::FastaNextRecordStart(fasta, len, end);
        const int i = (int)(k & 1);
        Slot &s = ctx->fslot[i];
        FastaDev &fa = ctx->fasta_slot[i];
        cudaStream_t st = s.stream;
        double t0 = now();
        if ((rc = deliver(i))) break;  // piece k - 2: its landing buffer is reused below
        tDeliver += now() - t0, t0 = now();
        // (no host-side staging of pageable bytes here: measured slower than the driver's own staging, 32 against 25 ms per
        // 400,000 reads -- this call is bound by the text going the other way; pinned bytes take 21.6 ms)
        if ((rc = fasta_stage_copy_count(ctx, fa, fasta + at, end - at, st, false))) break;
        CK(ctx, cudaStreamSynchronize(st));
        tCopy += now() - t0, t0 = now();
        if ((rc = fasta_stage_index(ctx, fa, fasta + at, st))) break;
        CK(ctx, cudaStreamSynchronize(st));
        fasta_index_done(fa);
        tIndex += now() - t0, t0 = now();
        at = end;
        const uint64_t n = fa.N;
        if (!n) continue;
        const uint32_t stride = treat_b200_ops_stride(ctx, fa.max_len);
        if ((rc = fasta_check_align(ctx, fa, aflags, stride))) break;
        bool in_place = !fa.multi && !ctx->tmpl_wide;
        if ((rc = fasta_align_chunk(ctx, fa, s, 0, n, aflags, stride, &in_place))) break;
        tAlign += now() - t0, t0 = now();
        TextParams tp{};
        tp.seqs = in_place ? (const uint8_t *)fa.file.p : (const uint8_t *)fa.seqs.p;
        if (in_place) {
            tp.begin = (const uint64_t *)fa.seq_begin.p;
            tp.len_ = (const uint32_t *)fa.seq_len.p;
        } else {
            tp.off = (const uint64_t *)fa.seq_off.p;
        }
        tp.names = (const uint8_t *)fa.file.p;
        tp.name_off = (const uint64_t *)fa.id_off.p;
        tp.name_len = (const uint32_t *)fa.id_len.p;
        tp.rec = (const treat_b200_record *)s.rec.p;
        tp.ops = (const uint8_t *)s.ops.p;
        tp.ops_stride = stride;
        tp.N = n;
        tp.max_len = fa.max_len;
        tp.tw = tw;
        text_template(ctx, tp);
        CK(ctx, s.text_off.reserve((n + 1) * 8));
        if ((rc = text_sizes(ctx, s, tp, (uint64_t *)s.text_off.p, st))) break;
        CK(ctx, cudaStreamSynchronize(st));
        const uint64_t bytes = *s.h_text_total;
        tSizes += now() - t0, t0 = now();
        CK(ctx, s.text.reserve(bytes));
        if (bytes > s.tland_cap) {
            if (s.tland) cudaFreeHost(s.tland);
            s.tland = nullptr;
            s.tland_cap = 0;
            CK(ctx, cudaMallocHost((void **)&s.tland, bytes + bytes / 4 + 4096));
            s.tland_cap = bytes + bytes / 4 + 4096;
        }
        if ((rc = text_write(ctx, s, tp, (const uint64_t *)s.text_off.p, (uint8_t *)s.text.p, st))) break;
        CK(ctx, cudaMemcpyAsync(s.tland, s.text.p, bytes, cudaMemcpyDeviceToHost, st));
        pend[i].live = true;
        pend[i].first = total;
        pend[i].n = n;
        pend[i].bytes = bytes;
        total += n;
        tWrite += now() - t0;
    }
    const double t_loop = now();
    // the last two pieces, in file order
    for (int j = 0; j < 2; j++) {
        const int i = (int)((k + j) & 1);
        const int r = rc == TREAT_B200_OK ? deliver(i) : TREAT_B200_OK;
        if (r) rc = r;
    }
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(ctx->fslot[i].stream);
    if (trace)
        fprintf(stderr, "[treat_b200] fasta_text %llu pieces, %llu records, %.2f ms: wait+sink %.2f, copy+count %.2f, index %.2f, align %.2f, sizes %.2f, write(enqueue) %.2f, drain %.2f\n",
                (unsigned long long)k, (unsigned long long)total, now() - t_begin, tDeliver, tCopy, tIndex, tAlign, tSizes, tWrite, now() - t_loop);
    if (n_records) *n_records = total;
    return rc;
}

uint32_t treat_b200_summary_bins(const treat_b200_ctx *ctx) { return ctx ? ctx->bins : 0; }
uint64_t treat_b200_summary_len(const treat_b200_ctx *ctx)
{
    return ctx && ctx->have_template ? TREAT_B200_SUMMARY_HEADER + 3ull * ctx->bins : 0;
}

int treat_b200_get_summary(treat_b200_ctx *ctx, uint64_t *out, uint64_t len)
{
    if (!ctx || !out) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "summary: no template");
    if (len < treat_b200_summary_len(ctx)) return fail(ctx, TREAT_B200_ERR_INVALID, "summary: buffer too small");
    int rc = treat_b200_synchronize(ctx);
    if (rc) return rc;
    CK(ctx, cudaMemcpy(out, ctx->d_summary, treat_b200_summary_len(ctx) * 8, cudaMemcpyDeviceToHost));
    return TREAT_B200_OK;
}

int treat_b200_reset_summary(treat_b200_ctx *ctx)
{
    if (!ctx) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return TREAT_B200_OK;
    int rc = treat_b200_synchronize(ctx);
    if (rc) return rc;
    CK(ctx, cudaMemset(ctx->d_summary, 0, treat_b200_summary_len(ctx) * 8));
    CK(ctx, cudaMemset(ctx->d_heat, 0, (size_t)ctx->td.len * ctx->td.len * 8));
    return TREAT_B200_OK;
}

uint32_t treat_b200_heatmap_dim(const treat_b200_ctx *ctx) { return ctx && ctx->have_template ? (uint32_t)ctx->td.len : 0; }

int treat_b200_get_heatmap(treat_b200_ctx *ctx, uint64_t *out, uint64_t len)
{
    if (!ctx || !out) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "heatmap: no template");
    const uint64_t n = (uint64_t)ctx->td.len * ctx->td.len;
    if (len < n) return fail(ctx, TREAT_B200_ERR_INVALID, "heatmap: buffer too small (dim * dim)");
    int rc = treat_b200_synchronize(ctx);
    if (rc) return rc;
    CK(ctx, cudaMemcpy(out, ctx->d_heat, n * 8, cudaMemcpyDeviceToHost));
    return TREAT_B200_OK;
}

int treat_b200_heatmap_device_ptr(treat_b200_ctx *ctx, uint64_t **d_ptr)
{
    if (!ctx || !d_ptr) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "heatmap: no template");
    *d_ptr = ctx->d_heat;
    return TREAT_B200_OK;
}

int treat_b200_summary_device_ptr(treat_b200_ctx *ctx, uint64_t **d_ptr)
{
    if (!ctx || !d_ptr) return TREAT_B200_ERR_INVALID;
    if (!ctx->have_template) return fail(ctx, TREAT_B200_ERR_NO_TEMPLATE, "summary: no template");
    *d_ptr = ctx->d_summary;
    return TREAT_B200_OK;
}

int treat_b200_measure_int32_peak(treat_b200_ctx *ctx, double *lane_ops_per_s, double *sm_mhz_effective)
{
    if (!ctx || !lane_ops_per_s) return TREAT_B200_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    const int iters = 20000;
    const size_t threads = (size_t)ctx->num_sms * 8 * 256;
    int *d_out = nullptr;
    CK(ctx, cudaMalloc((void **)&d_out, threads * 4));
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        CK(ctx, cudaEventRecord(e0, ctx->stream));
        int e = launch_int32_peak(ctx->num_sms, iters, d_out, ctx->stream);
        if (e) return fail(ctx, TREAT_B200_ERR_CUDA, "k_int32_peak launch failed");
        ctx->launches++;
        CK(ctx, cudaEventRecord(e1, ctx->stream));
        CK(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        CK(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    const double ops = (double)threads * iters * 128.0;  // 8 x 8 x (add + max) per iteration
    *lane_ops_per_s = ops / (best * 1e-3);
    if (sm_mhz_effective) *sm_mhz_effective = *lane_ops_per_s / ((double)ctx->num_sms * 128.0) / 1e6;
    return TREAT_B200_OK;
}

int treat_b200_measure_h2d(treat_b200_ctx *ctx, const void *host, uint64_t bytes, int32_t reps, double *ms_per_copy)
{
    if (!ctx || !host || !bytes || reps < 1 || !ms_per_copy) return TREAT_B200_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    void *d = nullptr;
    CK(ctx, cudaMalloc(&d, bytes));
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    CK(ctx, cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));  // warm-up (page tables, clocks)
    CK(ctx, cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < reps; i++) CK(ctx, cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaEventRecord(e1, ctx->stream));
    CK(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *ms_per_copy = (double)ms / reps;
    return TREAT_B200_OK;
}

int treat_b200_measure_transfer(treat_b200_ctx *ctx, const void *host_in, uint64_t in_bytes, void *host_out, uint64_t out_bytes,
                                int32_t reps, double *ms_per_step)
{
    if (!ctx || !host_in || !in_bytes || !host_out || !out_bytes || reps < 1 || !ms_per_step) return TREAT_B200_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    void *d_in = nullptr, *d_out = nullptr;
    CK(ctx, cudaMalloc(&d_in, in_bytes));
    CK(ctx, cudaMalloc(&d_out, out_bytes));
    cudaStream_t s2;
    CK(ctx, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    CK(ctx, cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    CK(ctx, cudaMemcpyAsync(d_in, host_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));  // warm-up
    CK(ctx, cudaMemcpyAsync(host_out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaEventRecord(e0, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(s2, e0, 0));
    for (int i = 0; i < reps; i++) {
        CK(ctx, cudaMemcpyAsync(d_in, host_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaMemcpyAsync(host_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s2));
    }
    CK(ctx, cudaEventRecord(e2, s2));
    CK(ctx, cudaStreamWaitEvent(ctx->stream, e2, 0));
    CK(ctx, cudaEventRecord(e1, ctx->stream));
    CK(ctx, cudaEventSynchronize(e1));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(e2);
    cudaStreamDestroy(s2);
    cudaFree(d_in);
    cudaFree(d_out);
    *ms_per_step = (double)ms / reps;
    return TREAT_B200_OK;
}

}  // extern "C"

namespace treat_b200 {
int ctx_device(const treat_b200_ctx *ctx) { return ctx->device; }
cudaStream_t ctx_stream(treat_b200_ctx *ctx) { return ctx->stream; }
}  // namespace treat_b200
