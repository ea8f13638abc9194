// band_kernel.cu -- K2b: the reads that need a DP, before the full-matrix kernels see them (sm_100a).
//
// k_align2 / k_align fill all m x n cells of every read they are given.  The reads TREAT sees are PCR amplicons of one
// mRNA: a read that differs from the template's bases at all differs by a few substitutions, one or two indels, or a
// truncated end.  Two exact shortcuts take those, one warp per batch of 64 reads:
//
//  1. Pure deletions (the read's bases are a subsequence of the template's: truncated reads, single-base deletions).
//     Then F[i][k] <= 2k - i with equality iff read[0:k] embeds in template[0:i], so on the traceback path "Up" is
//     optimal iff the read prefix still embeds in the shorter template prefix, "Left" never is, and the diagonal step is a
//     match: nwalgo's traceback (any of the three tie orders) returns the LEFTMOST embedding, score 2n - m.  One thread
//     per read walks a next-occurrence table of the template (n dependent shared-memory loads); no DP at all.
//
//  2. Banded fill, one THREAD per read pair: rows are read bases j, the thread keeps the 32 cells G[i][j], i - j in
//     [-16, 15], of the current row in 32 registers (two reads in the 16-bit halves: VIADDMNMX.S16x2 / VIMNMX.S16x2), so a
//     cell costs four instructions and no lane ever waits for another:
//         PRMT      w  = (wA | wB << 16): byte tc of the read bases' "score quads" (w against template code 0 / 1 / 2, and 0
//                        for a pad), selected by a per-template-base selector that is the same for all threads (one
//                        broadcast LDS.128 per four cells),
//         VIADDMNMX e  = max(G[i-1][j-1] + w, G[i][j-1]),     VIMNMX  f = max(e, G[i-1][j]),
//         IMAD      the row's 32 two-bit fields (f_c - f_{c-1}; field 0: f_0(j) - f_0(j-1)) by the telescoped sum of
//                   align_kernel.cu, 8 fields per 16-bit half-word, 4 words per row and read pair.
//     G = F + i + j as in align_kernel.cu; template / read positions outside the matrix are pads with w = 0, which
//     reproduces the zero boundary and makes everything behind (m, n) a copy of G[m][n].  The direction words go to a
//     per-warp global scratch (16 B per thread and row, coalesced); afterwards the warp takes its reads one at a time:
//     certificate, traceback (32 cells of a diagonal run per step, as in align_kernel.cu) and the editing-site assembly.
//     Certificate: a path that leaves the region touches diagonal i - j = 16 or -17 and then holds at least
//     g = min(16 + |16 - (m-n)|, 17 + |-17 - (m-n)|) gap steps, so it scores at most (m + n - g)/2 - g; if the region's
//     score beats that, every optimal path lies inside, the region's values equal the full matrix's on them and on the
//     neighbours the traceback tests, and the ops are nwalgo's (tests/test_band_certificate.py).
//
// Everything else (and every read whose certificate fails) is appended to list2 for launch_align.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "assembly.cuh"
#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

namespace {

#ifndef TREAT_B200_BAND_WARPS
#define TREAT_B200_BAND_WARPS 8
#endif
#ifndef TREAT_B200_BAND_MIN_CTAS
#define TREAT_B200_BAND_MIN_CTAS 3
#endif
constexpr int kBandWarps = TREAT_B200_BAND_WARPS;
constexpr int kBandLo = 16;                 // cell c of a row <-> template index i = j + c - kBandLo, c = 0..31
constexpr uint32_t kSelPad = 0x7733u;       // PRMT selector of a pad template position: byte 3 of both quads (zero)
enum : int { kCatNone = 0, kCatPass = 1, kCatSubseq = 2, kCatBand = 3 };

// PRMT without __byte_perm's selector masking (every selector nibble here is <= 7: no sign replication)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t dsum64(unsigned long long x)
{
    return (uint32_t)__popcll(x & 0x5555555555555555ull) + 2u * (uint32_t)__popcll(x & 0xaaaaaaaaaaaaaaaaull);
}

// nwalgo's traceback on the band's direction rows.  row(j) returns the 32 two-bit fields of row j of the read being
// traced (row 0 = 0); field c >= 1 = G(i,j) - G(i-1,j) with i = j + c - 16, field 0 = G(j-16,j) - G(j-17,j-1).
template <typename RowFn, typename IsMatch>
__device__ __forceinline__ int traceback_band(const AlignParams &p, RowFn row, IsMatch is_match, uint32_t tie, uint32_t *ops2, int m, int n,
                                              int lane, int &ngaps)
{
    int pos = m + n, i = m, j = n;
    ngaps = 0;
    while (i > 0 && j > 0) {
        const int ii = i - lane, jj = j - lane;
        const int c = i - j + kBandLo;  // the same for the 32 cells of this diagonal run
        TB_CHECK(p, c >= 0 && c <= 31, 10);
        bool stop = true, is_up = false;
        if (ii >= 1 && jj >= 1) {
            const unsigned long long d1 = row(jj);
            const bool pU = c >= 1 && ((d1 >> (2 * c)) & 3ull) == 0ull;  // G[i][j] == G[i-1][j]: Up is optimal
            if (pU && tie != TREAT_B200_TIE_LUD) {
                is_up = true;
            } else {
                bool pL = false;
                unsigned long long d0 = 0ull;
                if (c <= 30) {
                    // G(i,j) - G(i,j-1) = sum_{k<=c} field_j(k) - sum_{1<=k<=c+1} field_{j-1}(k)
                    d0 = row(jj - 1);
                    const unsigned long long m1 = (4ull << (2 * c)) - 1ull;
                    const unsigned long long m0 = (c == 30 ? ~0ull : (16ull << (2 * c)) - 1ull) & ~3ull;
                    pL = ((dsum64(d1 & m1) - dsum64(d0 & m0)) & 3u) == 0u;
                }
                if (tie == TREAT_B200_TIE_ULD) {
                    stop = pL;
                } else if (tie == TREAT_B200_TIE_LUD) {
                    stop = pL || pU;
                    is_up = !pL && pU;
                } else {  // UDL: with Left optimal the diagonal is too iff G[i][j-1] - G[i-1][j-1] == w(i,j)
                    stop = pL && (uint32_t)((d0 >> (2 * (c + 1))) & 3ull) != (is_match(ii - 1, jj - 1) ? 3u : 1u);
                }
            }
        }
        const uint32_t sm = __ballot_sync(FULL_MASK, stop);
        const int run = sm ? __ffs(sm) - 1 : 32;
        pos -= run;
        i -= run;
        j -= run;
        if (run < 32 && i > 0 && j > 0) {
            const bool up = __shfl_sync(FULL_MASK, is_up, run);
            --pos;
            ngaps++;
            if (lane == 0) ops2[pos >> 4] |= (up ? TREAT_B200_OP_DEL : TREAT_B200_OP_INS) << (2 * (pos & 15));
            if (up) i--; else j--;
        }
    }
    __syncwarp();
    const int rest = i + j;  // first row / column: Left on row 0, Up on column 0
    if (rest) {
        const uint32_t op = i ? TREAT_B200_OP_DEL : TREAT_B200_OP_INS;
        for (int q = lane; q < rest; q += 32) {
            const int pp = pos - 1 - q;
            atomicOr(&ops2[pp >> 4], op << (2 * (pp & 15)));
        }
        pos -= rest;
        ngaps += rest;
    }
    __syncwarp();
    return pos;
}

}  // namespace

struct BandParams {
    uint32_t *list2;        // reads left for launch_align
    uint32_t *list2_count;
    uint32_t *sub_list;     // reads whose bases are a subsequence of the template's (pure deletions)
    uint32_t *sub_slot;     // ... and the slot of masks that holds each one's embedding
    uint32_t *sub_count;
    uint32_t *band_list;    // reads for the banded fill
    uint32_t *band_count;
    uint4 *scratch;         // [batch of the round][rows][32] direction words
    uint32_t *meta;         // [read of band_list] diagonal-only flag << 8 | score << 16
    uint32_t *masks;        // [read k_band_classify looked at][mwords] template positions the read's bases are matched to
    uint32_t first;         // k_band_fill / k_band_finish: the round covers band_list[first, first + round_reads)
    uint32_t round_reads;   // multiple of 64
    uint32_t mode;          // k_band_finish: 0 = the reads of sub_list, 1 = the band reads of the round
    uint32_t rows;          // rows of one batch's scratch region (largest row index + 1)
    uint32_t ncapb;         // largest n the fill is sized for (m + 15, or the launch's ncap if smaller)
    uint32_t subseq;        // 1: the pure-deletion shortcut is on
    uint32_t min_reads;     // fewer reads to align than this: no banded fill (k_band_classify)
    unsigned long long *stats;  // [3] reads finished as pure deletions / by the banded fill / left for launch_align (may be null)
    uint32_t sel_words;     // words of one copy of the selector table
    uint32_t pk_stride;     // words per packed read in shared memory (odd)
    uint32_t mwords;        // words of one read's embedding mask
    // k_band_fill: CTA-wide sections, then 64 packed reads per warp
    uint32_t a_off_sel, a_off_quad, a_cta_bytes, a_warp_bytes;
    // k_band_finish: 17 byte masks behind the template section (sm_tmpl_bytes), then the sections of a warp
    uint32_t b_cta_bytes, b_off_pk, b_off_mask, b_off_dir, b_off_cnt, b_off_ops, b_off_t2f, b_warp_bytes;
};

// ------------------------------------------------------------------------------------------
// K2b-0: one THREAD per read that launch_align would get: pure deletion (leftmost embedding found -> sub_list + mask),
// band candidate (| m - n | < 16 -> band_list), or neither (list2).
// ------------------------------------------------------------------------------------------
constexpr int kClassifyThreads = 128;

__global__ void __launch_bounds__(kClassifyThreads) k_band_classify(const __grid_constant__ AlignParams p, const __grid_constant__ BandParams b)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t total = p.list ? *p.list_count : (uint32_t)p.N;
    if (total == 0u) return;
    const int m = p.t.m;
    uint16_t *s_next = reinterpret_cast<uint16_t *>(smem);  // [3][m + 2]: next position >= i of code c (m: none)
    uint32_t *s_tpk = reinterpret_cast<uint32_t *>(smem + align16(6u * ((uint32_t)m + 2)));  // template, 16 codes per word (+ 2 zero words)
    if (b.subseq) {
        uint8_t *s_tb = reinterpret_cast<uint8_t *>(s_tpk) + align16(4u * (((uint32_t)m + 15) / 16 + 2));  // [m] template codes
        for (int i = threadIdx.x; i < m; i += blockDim.x) s_tb[i] = p.t.tb[i];
        __syncthreads();
        // every entry on its own: scan forward to the next occurrence (three steps on average)
        for (int x = threadIdx.x; x < 3 * (m + 2); x += blockDim.x) {
            const int c = x / (m + 2);
            int i = x - c * (m + 2);
            while (i < m && (int)s_tb[i] != c) i++;
            s_next[x] = (uint16_t)min(i, m);
        }
        for (int w = threadIdx.x; w < (m + 15) / 16 + 2; w += blockDim.x) {
            uint32_t v = 0;
            for (int k = 0; k < 16; k++) {
                const int i = 16 * w + k;
                if (i < m) v |= (uint32_t)s_tb[i] << (2 * k);
            }
            s_tpk[w] = v;
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const bool small = total < b.min_reads;
    const int rwords = (int)(p.bases_stride / 4);  // words of a packed read's row
    uint32_t st_pass = 0;
    // whole warps stay in the loop: the three list appends are one atomic per warp and list
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t x0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); x0 < total; x0 += stride) {
        const uint32_t x = x0 + lane;
        uint32_t r = 0;
        int n = 0, cat = kCatNone;
        if (x < total) {
            r = p.list ? p.list[x] : x;
            n = (int)p.n[r];
            cat = kCatPass;
            if (n >= 1 && n <= (int)b.ncapb && n <= (int)p.ncap) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
                // (with n = m an embedding means equal bases: the kernels in front of this one finish those reads; if one does
                //  arrive here the banded fill takes it)
                // a short list: the banded fill's fixed latency (one warp walks ~m rows for its 64 reads) is not worth it, the
                // full-matrix kernel spreads each read over a warp; truncated reads still take the pure-deletion route
                bool sub = b.subseq && n < m && (!small || m - n >= kBandLo);
                if (sub) {
                    // Leftmost embedding of the read in the template, if there is one, 16 bases per step: the next read base goes
                    // to the template's next occurrence of its code (it may sit at most m - n positions late: an embedding skips
                    // exactly that many template bases), then both sequences are compared word-wise as far as they agree.  The
                    // matched template positions are written as a bit mask into slot x as the walk goes (a failed walk leaves a
                    // partial mask nobody reads).
                    const int slack = m - n;
                    uint32_t *mk = b.masks + (size_t)x * b.mwords;
                    int pos = 0, k = 0, wi = 0;
                    uint32_t cur = 0;
                    while (k < n) {
                        const int rw = k >> 4, rs = 2 * (k & 15);
                        const uint32_t r0 = src[rw], r1 = rw + 1 < rwords ? src[rw + 1] : 0u;
                        const uint32_t rbits = __funnelshift_r(r0, r1, rs);  // read bases k .. k + 15
                        const uint32_t c = rbits & 3u;
                        if (c == 3u) {
                            sub = false;
                            break;
                        }
                        const int q = (int)s_next[c * (m + 2) + pos];
                        if (q > k + slack) {  // includes q = m: no further occurrence
                            sub = false;
                            break;
                        }
                        const int tw = q >> 4;
                        const uint32_t tbits = __funnelshift_r(s_tpk[tw], s_tpk[tw + 1], 2 * (q & 15));  // template bases q .. q + 15
                        const uint32_t df = tbits ^ rbits;  // base 0 agrees
                        int e = df ? (__ffs(df) - 1) >> 1 : 16;
                        e = min(e, min(n - k, m - q));
                        while (wi < (q >> 5)) {
                            mk[wi++] = cur;
                            cur = 0u;
                        }
                        const uint32_t bits = (1u << e) - 1u, sh = (uint32_t)q & 31u;
                        cur |= bits << sh;
                        if (sh + (uint32_t)e > 32u) {
                            mk[wi++] = cur;
                            cur = bits >> (32u - sh);
                        }
                        pos = q + e;
                        k += e;
                    }
                    if (sub) {
                        while (wi < (int)b.mwords) {
                            mk[wi++] = cur;
                            cur = 0u;
                        }
                    }
                }
                if (sub) cat = kCatSubseq;
                else if (abs(m - n) < kBandLo && !small) cat = kCatBand;
            }
        }
        __syncwarp();
        const uint32_t m_sub = __ballot_sync(FULL_MASK, cat == kCatSubseq), m_band = __ballot_sync(FULL_MASK, cat == kCatBand);
        const uint32_t m_pass = __ballot_sync(FULL_MASK, cat == kCatPass);
        uint32_t b_sub = 0, b_band = 0, b_pass = 0;
        if (lane == 0) {
            if (m_sub) b_sub = atomicAdd(b.sub_count, (uint32_t)__popc(m_sub));
            if (m_band) b_band = atomicAdd(b.band_count, (uint32_t)__popc(m_band));
            if (m_pass) b_pass = atomicAdd(b.list2_count, (uint32_t)__popc(m_pass));
            st_pass += (uint32_t)__popc(m_pass);
        }
        b_sub = __shfl_sync(FULL_MASK, b_sub, 0);
        b_band = __shfl_sync(FULL_MASK, b_band, 0);
        b_pass = __shfl_sync(FULL_MASK, b_pass, 0);
        if (cat == kCatBand) b.band_list[b_band + __popc(m_band & lt)] = r;
        if (cat == kCatPass) b.list2[b_pass + __popc(m_pass & lt)] = r;
        if (cat == kCatSubseq) {
            const uint32_t z = b_sub + __popc(m_sub & lt);
            b.sub_list[z] = r;
            b.sub_slot[z] = x;  // where its mask is
        }
    }
    if (b.stats && st_pass) atomicAdd(&b.stats[2], (unsigned long long)st_pass);
}

// ------------------------------------------------------------------------------------------
// K2b-1: banded fill.  One warp per batch of 64 reads of band_list (thread t: reads 2t and 2t + 1 of the batch).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBandWarps * 32, TREAT_B200_BAND_MIN_CTAS) k_band_fill(const __grid_constant__ AlignParams p, const __grid_constant__ BandParams b)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t total_all = *b.band_count;
    if (total_all <= b.first) return;
    const uint32_t total = min(total_all - b.first, b.round_reads);  // reads of this round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m;
    uint8_t *s_tb = smem;                                                   // [m] template codes
    uint32_t *s_sel = reinterpret_cast<uint32_t *>(smem + b.a_off_sel);     // [4][sel_words]: copy s holds T[k + s]
    uint32_t *s_quad = reinterpret_cast<uint32_t *>(smem + b.a_off_quad);   // [8] score quads by read code (4: pad)
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_tb[i] = p.t.tb[i];
    __syncthreads();
    // selector of template index i = e - 15 (table index e = j + c - 1 for row j, cell c)
    for (uint32_t x = threadIdx.x; x < 4u * b.sel_words; x += blockDim.x) {
        const uint32_t s = x / b.sel_words, k = x - s * b.sel_words;
        const int i = (int)(k + s) - (kBandLo - 1);
        s_sel[x] = (i >= 1 && i <= m) ? 0x7430u + 0x0101u * (uint32_t)s_tb[i - 1] : kSelPad;
    }
    if (threadIdx.x < 8) {
        const uint32_t c = threadIdx.x;
        s_quad[c] = c < 3 ? 0x00010101u + (2u << (8 * c)) : c == 3 ? 0x00010101u : 0u;
    }
    uint32_t *s_pk = reinterpret_cast<uint32_t *>(smem + b.a_cta_bytes + (size_t)warp * b.a_warp_bytes);
    __syncthreads();

    // batches are dealt to the CTAs round robin, so that a short list still spreads over every SM
    const uint32_t gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * nwarps;
    const uint32_t sel32 = smem_u32(s_sel), quad32 = smem_u32(s_quad);
    const uint32_t pkA32 = smem_u32(s_pk + (2 * lane) * b.pk_stride), pkB32 = pkA32 + 4u * b.pk_stride;

    for (uint32_t base = gw * 64u; base < total; base += nw * 64u) {
        const uint32_t cnt = min(64u, total - base);
        __syncwarp();
        int nn[2] = {0, 0};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t x = 2u * lane + h;
            if (x < cnt) {
                const uint32_t r = b.band_list[b.first + base + x];
                const int n = (int)p.n[r];
                nn[h] = n;
                const uint32_t *src = reinterpret_cast<const uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
                uint32_t *dst = s_pk + x * b.pk_stride;
                const int words = (n + 15) >> 4;
                for (int w = 0; w < words; w++) dst[w] = src[w];
            }
        }
        const int nA = nn[0], nB = nn[1];
        const int nmax = __reduce_max_sync(FULL_MASK, max(nA, nB));
        const int J = max(nmax, m);
        TB_CHECK(p, (uint32_t)J < b.rows, 11);
        uint32_t B[32];
#pragma unroll
        for (int c = 0; c < 32; c++) B[c] = 0u;
        uint32_t wa = 0, wbv = 0;
        uint32_t dmin = 0xffffffffu;  // per half: the smallest of G[i][i] - G[i-1][i] and G[i][i] - G[i][i-1] over the rows
        uint4 *srow = b.scratch + ((size_t)(base >> 6) * b.rows + 1u) * 32u + lane;
#pragma unroll 1
        for (int j = 1; j <= J; j++, srow += 32) {
            const int k0 = j - 1;
            if ((k0 & 15) == 0) {
                wa = lds_u32(pkA32 + 4u * (uint32_t)(k0 >> 4));
                wbv = lds_u32(pkB32 + 4u * (uint32_t)(k0 >> 4));
            }
            const uint32_t sh = 2u * (uint32_t)(k0 & 15);
            const uint32_t ca = j <= nA ? (wa >> sh) & 3u : 4u, cb = j <= nB ? (wbv >> sh) & 3u : 4u;
            const uint32_t QA = lds_u32(quad32 + 4u * ca), QB = lds_u32(quad32 + 4u * cb);
            const uint32_t padm = (j <= nA ? 0u : 0xffffu) | (j <= nB ? 0u : 0xffff0000u);
            const uint32_t s = (uint32_t)k0 & 3u;
            const uint32_t sa = sel32 + 4u * (s * b.sel_words + ((uint32_t)k0 - s));
            uint32_t acc[4];
            uint32_t fprev = B[0];  // field 0: f_0(j) - f_0(j-1)
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 sv = lds_v4(sa + 16u * q);
                const uint32_t se[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = 4 * q + k;
                    const uint32_t w = prmt(QA, QB, se[k]);
                    // the add on the FMA pipe (both halves stay below 2^15: no carry), one three-input maximum on the ALU pipe
                    const uint32_t dw = B[c] + w;
                    const uint32_t f = c == 0 ? __vmaxs2(dw, B[c + 1]) : c < 31 ? __vimax3_s16x2(dw, B[c + 1], fprev) : __vmaxs2(dw, fprev);
                    if (c == kBandLo)  // B[c + 1] still is row j - 1; rows behind a read do not count
                        dmin = __vminu2(dmin, __vminu2(f - fprev, f - B[c + 1]) | padm);
                    const int fi = c & 7;
                    if (fi == 0) acc[c >> 3] = 0u - fprev;
                    acc[c >> 3] += f * (fi == 7 ? (1u << 14) : 0u - (3u << (2 * fi)));
                    B[c] = f;
                    fprev = f;
                }
            }
            __stcs(srow, make_uint4(acc[0], acc[1], acc[2], acc[3]));
        }
        // G[m][n] of both reads: every cell (i >= m, j >= n) holds it; cell c* of row J is (m, J)
        uint32_t fin = 0;
        const int cstar = m - J + kBandLo;
#pragma unroll
        for (int c = 1; c <= kBandLo; c++)
            if (c == cstar) fin = B[c];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t x = 2u * lane + h;
            if (x < cnt) {
                const int score = (int)((fin >> (16 * h)) & 0xffffu) - m - nn[h];
                // with n = m: neither Up nor Left is optimal anywhere on the main diagonal, so the traceback is the diagonal
                const uint32_t diag_only = (nn[h] == m && ((dmin >> (16 * h)) & 0xffffu) != 0u) ? 1u : 0u;
                b.meta[b.first + base + x] = (diag_only << 8) | ((uint32_t)(score & 0xffff) << 16);
            }
        }
    }
}

// everything but the gap-free fast path of k_band_finish (kept out of line: it would set the register count of the loop)
static __device__ __noinline__ unsigned long long band_finish_slow(const AlignParams &p, const BandParams &b, uint8_t *smem, uint8_t *wb, int lane,
                                                                   uint32_t y, bool band, bool diag_only, bool mine, uint32_t r, int n,
                                                                   uint32_t rflags, int score, uint4 c, const uint32_t *mk)
{
    const CtaSmem<uint8_t> cs = carve_cta<uint8_t>(smem, p, (int)p.cpl);
    unsigned long long ctr = 0;
    const int m = p.t.m;
    uint32_t *s_pk = reinterpret_cast<uint32_t *>(wb + b.b_off_pk);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(wb + b.b_off_mask);
    uint2 *s_dir = reinterpret_cast<uint2 *>(wb + b.b_off_dir);
    uint8_t *s_cnt = wb + b.b_off_cnt;
    WarpSmem ws{};
    ws.ops2 = reinterpret_cast<uint32_t *>(wb + b.b_off_ops);
    ws.t2f = reinterpret_cast<uint16_t *>(wb + b.b_off_t2f);
    const int ops_words = (int)(p.sm_ops / 4);
    int pos = n, ngaps = 0;
    __syncwarp();
    for (int i = lane; i < ops_words; i += 32) ws.ops2[i] = 0u;
    {
        const uint4 *gc = reinterpret_cast<const uint4 *>(p.counts + r * (uint64_t)p.counts_stride);
        if (n + 1 <= 512) {
            if (mine) reinterpret_cast<uint4 *>(s_cnt)[lane] = c;
        } else {
            for (int i = lane; i * 16 < n + 1; i += 32) reinterpret_cast<uint4 *>(s_cnt)[i] = gc[i];
        }
        const uint32_t *gb = reinterpret_cast<const uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
        for (int i = lane; i * 16 < n; i += 32) s_pk[i] = gb[i];
    }
    auto code = [&](int fi) -> int { return (int)((s_pk[fi >> 4] >> (2 * (fi & 15))) & 3u); };
    int mism_known = -1;
    if (diag_only) {
        mism_known = (m - score) >> 1;  // the diagonal, templates too long for finish_same
        __syncwarp();
    } else if (band && !(score == m && n == m)) {
        const uint4 *rows = b.scratch + (size_t)(y >> 6) * b.rows * 32u + ((y & 63u) >> 1);
        const uint32_t sel = (y & 1u) ? 0x7632u : 0x5410u;
        if (lane == 0) s_dir[0] = make_uint2(0u, 0u);
        for (int jj = 1 + lane; jj <= n; jj += 32) {
            const uint4 v = __ldcs(rows + jj * 32);
            s_dir[jj] = make_uint2(__byte_perm(v.x, v.y, sel), __byte_perm(v.z, v.w, sel));
        }
        __syncwarp();
        pos = traceback_band(
            p,
            [&](int jj) {
                const uint2 d = s_dir[jj];
                return (unsigned long long)d.x | ((unsigned long long)d.y << 32);
            },
            [&](int ti, int fi) { return code(fi) == (int)cs.tb[ti]; }, p.tie, ws.ops2, m, n, lane, ngaps);
    } else if (!band) {
        for (uint32_t i = lane; i < b.mwords; i += 32) s_mask[i] = mk[i];
        __syncwarp();
        for (int ti = lane; ti < m; ti += 32) {
            if (!((s_mask[ti >> 5] >> (ti & 31)) & 1u)) {
                const int pp = n + ti;
                atomicOr(&ws.ops2[pp >> 4], TREAT_B200_OP_DEL << (2 * (pp & 15)));
            }
        }
        ngaps = m - n;
        __syncwarp();
    } else {
        __syncwarp();
    }
    finish_read<uint8_t>(p, cs, ws, lane, r, n, rflags, score, pos, ngaps, s_cnt, code, ctr, mism_known);
    return ctr;
}

// ------------------------------------------------------------------------------------------
// K2b-2: one warp per read: certificate, traceback on the stored direction rows (or the ops of the leftmost embedding),
// editing-site assembly.  Reads whose certificate fails go to list2.
// ------------------------------------------------------------------------------------------
constexpr int kBandFinishWarps = 8;

__global__ void __launch_bounds__(kBandFinishWarps * 32, 3) k_band_finish(const __grid_constant__ AlignParams p, const __grid_constant__ BandParams b)
{
    extern __shared__ __align__(128) uint8_t smem[];
    // mode 0: the reads of sub_list; 1: the band reads of the round; 2: both (items [0, S) are sub reads, the rest band reads)
    const uint32_t S = b.mode != 1u ? *b.sub_count : 0u;
    uint32_t Bn = 0;
    if (b.mode != 0u) {
        const uint32_t all = *b.band_count;
        Bn = all > b.first ? min(all - b.first, b.round_reads) : 0u;
    }
    const uint32_t total = S + Bn;
    if (total == 0u) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m;
    const CtaSmem<uint8_t> cs = carve_cta<uint8_t>(smem, p, (int)p.cpl);
    load_template(cs, p, (int)p.cpl);
    uint4 *s_bmask = reinterpret_cast<uint4 *>(smem + p.sm_tmpl_bytes);  // [17]: t lowest bytes 0xff
    if (threadIdx.x < 17) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = min(max((int)threadIdx.x - 4 * i, 0), 4);
            w[i] = k >= 4 ? 0xffffffffu : (1u << (8 * k)) - 1u;
        }
        s_bmask[threadIdx.x] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint8_t *wb = smem + b.b_cta_bytes + (size_t)warp * b.b_warp_bytes;
    uint8_t *s_cnt = wb + b.b_off_cnt;
    WarpSmem ws{};
    ws.ops2 = reinterpret_cast<uint32_t *>(wb + b.b_off_ops);
    ws.t2f = reinterpret_cast<uint16_t *>(wb + b.b_off_t2f);
    __syncthreads();
    const SameLane sl = load_same_lane(cs, p, lane);
    const bool same_ok = p.t.len <= 512;  // finish_same<32>: 32 lanes x 16 sites
    const bool mine = (uint32_t)(16 * lane) < p.counts_stride;  // this lane's 16 count bytes lie inside a read's row

    // items are dealt to the CTAs round robin (a short list still reaches every SM)
    const uint32_t gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * nwarps;
    const int ops_words = (int)(p.sm_ops / 4);
    unsigned long long ctr = 0;
    uint32_t st_sub = 0, st_band = 0, st_pass = 0;
    // software pipeline: (read index, meta) two items ahead; (n, flags, this lane's count bytes) one item ahead
    auto fetch1 = [&](uint32_t x, uint32_t &r, uint32_t &me) {
        r = 0u;
        me = 0u;
        if (x < S) {
            r = b.sub_list[x];
            me = b.sub_slot[x];  // pure deletions carry their mask slot where band reads carry score and marks
        } else if (x < total) {
            r = b.band_list[b.first + (x - S)];
            me = b.meta[b.first + (x - S)];
        }
    };
    auto fetch2 = [&](uint32_t x, uint32_t r, uint32_t &nf, uint32_t &rcn, uint4 &c) {
        nf = 0u;
        rcn = 1u;
        c = make_uint4(0, 0, 0, 0);
        if (x < total) {
            nf = (uint32_t)p.n[r] | ((uint32_t)p.rflags[r] << 16);
            if (p.read_count) rcn = p.read_count[r];
            if (mine) c = reinterpret_cast<const uint4 *>(p.counts + r * (uint64_t)p.counts_stride)[lane];
        }
    };
    uint32_t r_cur, me_cur, nf_cur, rc_cur, r_nxt, me_nxt;
    uint4 c_cur;
    fetch1(gw, r_cur, me_cur);
    fetch2(gw, r_cur, nf_cur, rc_cur, c_cur);
    fetch1(gw + nw, r_nxt, me_nxt);
    for (uint32_t x = gw; x < total; x += nw) {
        uint32_t r_nn, me_nn, nf_nxt, rc_nxt;
        uint4 c_nxt;
        fetch2(x + nw, r_nxt, nf_nxt, rc_nxt, c_nxt);
        fetch1(x + 2 * nw, r_nn, me_nn);
        const uint32_t r = r_cur, me = me_cur, rflags = nf_cur >> 16;
        const int n = (int)(nf_cur & 0xffffu);
        const uint4 c = c_cur;
        const uint32_t read_count = rc_cur;
        r_cur = r_nxt, me_cur = me_nxt, nf_cur = nf_nxt, rc_cur = rc_nxt, c_cur = c_nxt;
        r_nxt = r_nn, me_nxt = me_nn;

        const bool band = x >= S;
        int score = (int)(short)(me >> 16);
        if (band) {
            const int dmn = m - n;
            const int g = min(kBandLo + abs(kBandLo - dmn), kBandLo + 1 + abs(-(kBandLo + 1) - dmn));
            if (!(2 * score + 3 * g > m + n)) {  // not certified: the full matrix decides
                if (lane == 0) {
                    b.list2[atomicAdd(b.list2_count, 1u)] = r;
                    st_pass++;
                }
                continue;
            }
            st_band++;
        } else {
            score = 2 * n - m;
            st_sub++;
        }
        const bool diag_only = band && n == m && ((me >> 8) & 1u);
        if (diag_only && same_ok) {
            // no gap: the assembly on count bytes in registers (lane l: sites 16 l .. 16 l + 15), (m - score) / 2 mismatches
            const int mism = (m - score) >> 1;
            if (!finish_same<32>(p, cs, s_bmask, lane, FULL_MASK, 0, r, rflags, c, sl, ctr, 0xffffffffu, mism, read_count)) {
                __syncwarp();
                reinterpret_cast<uint4 *>(s_cnt)[lane] = c;  // 512 bytes: b_off_cnt holds at least that (band_plan)
                for (int i = lane; i < ops_words; i += 32) ws.ops2[i] = 0u;
                __syncwarp();
                unsigned long long ctr2 = 0;  // its address goes to the out-of-line call: keep `ctr` itself in a register
                finish_same_general(p, cs, ws, lane, r, rflags, s_cnt, ctr2, 0xffffffffu, mism);
                ctr += ctr2;
            }
            continue;
        }
        ctr += band_finish_slow(p, b, smem, wb, lane, x - S, band, diag_only, mine, r, n, rflags, score, c,
                                band ? nullptr : b.masks + (size_t)me * b.mwords);
    }
    if (b.stats && lane == 0) {
        if (st_sub) atomicAdd(&b.stats[0], (unsigned long long)st_sub);
        if (st_band) atomicAdd(&b.stats[1], (unsigned long long)st_band);
        if (st_pass) atomicAdd(&b.stats[2], (unsigned long long)st_pass);
    }
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
constexpr uint64_t kBandScratchCap = 1536ull << 20;  // direction words of one round

static bool band_plan(const AlignParams &planned, AlignParams &q, BandParams &b, int max_smem, int &warps_a, int &warps_b)
{
    q = planned;
    const int m = q.t.m;
    if (m < 1 || q.t.len > 1024) return false;
    const uint32_t sum_len = q.summary ? TREAT_B200_SUMMARY_HEADER + 3u * q.bins : 0u;
    q.cpl = (uint32_t)((m + 31) / 32);  // sizes the template-base section (32 x cpl bytes)
    q.sm_prof_bytes = 0;
    q.sm_tmpl_bytes = align16(32u * q.cpl) + align16((uint32_t)(q.t.K * q.t.len_pad)) + align16(8u * (q.t.n_alt ? q.t.n_alt : 1)) +
                      align16(8u * sum_len);
    q.sm_tmpl_bytes = (q.sm_tmpl_bytes + 127u) & ~127u;
    b.ncapb = std::min<uint32_t>((uint32_t)m + kBandLo - 1, std::max<uint32_t>(planned.ncap, 1u));
    const uint32_t nfin = std::max<uint32_t>(b.ncapb, (uint32_t)m);  // rows of a fill; largest n k_band_finish sees
    b.rows = nfin + 1u;
    q.sm_ops = align16((((uint32_t)m + nfin + 15) / 16 + 2) * 4);
    q.sm_t2f = align16(2u * (uint32_t)q.t.len_pad);
    if (q.ncap < nfin) q.ncap = nfin;  // only bounds checks read it in k_band_finish; eligibility uses the planned value
    b.sel_words = ((uint32_t)m + 2 * kBandLo + 32 + 8 + 3) & ~3u;
    b.pk_stride = ((nfin + 15) / 16) | 1u;
    b.mwords = ((uint32_t)m + 31) / 32;
    b.a_off_sel = (align16((uint32_t)m) + 127u) & ~127u;
    b.a_off_quad = b.a_off_sel + 16u * b.sel_words;
    b.a_cta_bytes = (b.a_off_quad + 32u + 127u) & ~127u;
    b.a_warp_bytes = (64u * b.pk_stride * 4u + 127u) & ~127u;
    b.b_cta_bytes = q.sm_tmpl_bytes + 17u * 16u;
    b.b_cta_bytes = (b.b_cta_bytes + 127u) & ~127u;
    b.b_off_pk = 0;
    b.b_off_mask = align16(b.pk_stride * 4u);
    b.b_off_dir = b.b_off_mask + align16(b.mwords * 4u);
    b.b_off_cnt = b.b_off_dir + align16(8u * (nfin + 2));
    b.b_off_ops = b.b_off_cnt + std::max<uint32_t>(align16(nfin + 1) + 16u, 512u);
    b.b_off_t2f = b.b_off_ops + q.sm_ops;
    b.b_warp_bytes = (b.b_off_t2f + q.sm_t2f + 127u) & ~127u;
    warps_a = (int)std::min<long>(kBandWarps, ((long)max_smem - (long)b.a_cta_bytes) / (long)b.a_warp_bytes);
    warps_b = (int)std::min<long>(kBandFinishWarps, ((long)max_smem - (long)b.b_cta_bytes) / (long)b.b_warp_bytes);
    if (warps_a < 1 || warps_b < 1) return false;
    if (8u * ((uint32_t)m + 2) + 64u > 48u * 1024u) return false;  // k_band_classify's next-occurrence table + template codes
    // reads per round: what the scratch cap holds, in whole batches
    const uint64_t per_batch = (uint64_t)b.rows * 512ull;
    uint64_t batches = std::max<uint64_t>(1, kBandScratchCap / per_batch);
    batches = std::min<uint64_t>(batches, (q.N + 63) / 64);
    b.round_reads = (uint32_t)(batches * 64);
    return true;
}

// global memory of a launch: direction words of a round, meta, the embedding masks, the two lists and their counters
struct BandLayout {
    uint64_t dir, meta, masks, sub_list, sub_slot, band_list, counters, total;
};
static BandLayout band_layout(const AlignParams &q, const BandParams &b)
{
    BandLayout L;
    uint64_t at = (uint64_t)(b.round_reads / 64) * b.rows * 512ull;
    L.dir = 0;
    L.meta = at;
    at += q.N * 4ull;
    L.masks = at;
    at += q.N * (uint64_t)b.mwords * 4ull;
    L.sub_list = at;
    at += q.N * 4ull;
    L.sub_slot = at;
    at += q.N * 4ull;
    L.band_list = at;
    at += q.N * 4ull;
    L.counters = 0;
    L.total = (at + 15ull) & ~15ull;
    return L;
}

// the kernels apply to the packed alphabet; returns the global bytes they need (0: they do not apply)
uint64_t band_scratch_bytes(const AlignParams &planned, int num_sms, int max_smem)
{
    (void)num_sms;
    AlignParams q;
    BandParams b{};
    int wa = 0, wb = 0;
    if (planned.N == 0 || !band_plan(planned, q, b, max_smem, wa, wb)) return 0;
    return band_layout(q, b).total;
}

int occupancy_cached(const void *kern, int threads, size_t smem)
{
    struct Entry {
        const void *kern;
        int threads;
        size_t smem;
        int per_sm;
    };
    static std::mutex mu;
    static Entry cache[64];
    static int used = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (int i = 0; i < used; i++)
            if (cache[i].kern == kern && cache[i].threads == threads && cache[i].smem == smem) return cache[i].per_sm;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (used < 64) cache[used++] = Entry{kern, threads, smem, per_sm};
    return per_sm;
}

template <typename Kern>
static uint64_t resident_ctas(Kern kern, int threads, size_t smem, int num_sms, int cap_per_sm)
{
    return (uint64_t)num_sms * std::max(1, std::min(occupancy_cached((const void *)kern, threads, smem), cap_per_sm));
}

// A second stream next to a caller's stream (per device), with the two events that fork from and join into it: the pure
// deletions are finished there while k_band_fill, which keeps only a few warps per SM busy on a short list, runs on the main one.
namespace {
struct AuxLane {
    cudaStream_t aux = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
AuxLane *aux_lane(cudaStream_t st)
{
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, AuxLane> lanes;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    AuxLane &l = lanes[std::make_pair(dev, st)];
    if (!l.aux) {
        if (cudaStreamCreateWithFlags(&l.aux, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            if (l.aux) cudaStreamDestroy(l.aux);
            l = AuxLane{};
            return nullptr;
        }
    }
    return &l;
}
}  // namespace

int launch_band(const AlignParams &planned, uint32_t *list2, uint32_t *list2_count, void *scratch, bool subseq, uint32_t min_reads,
                unsigned long long *stats, int num_sms, int max_smem, void *stream, int *kernels_launched)
{
    AlignParams q;
    BandParams b{};
    int launched = 0;
    int wa = 0, wb = 0;
    if (!band_plan(planned, q, b, max_smem, wa, wb)) return -1000;
    cudaStream_t st = (cudaStream_t)stream;
    const BandLayout L = band_layout(q, b);
    uint8_t *base = reinterpret_cast<uint8_t *>(scratch);
    b.list2 = list2;
    b.list2_count = list2_count;
    b.subseq = subseq ? 1u : 0u;
    b.min_reads = min_reads;
    b.stats = stats;
    b.scratch = reinterpret_cast<uint4 *>(base + L.dir);
    b.meta = reinterpret_cast<uint32_t *>(base + L.meta);
    b.masks = reinterpret_cast<uint32_t *>(base + L.masks);
    b.sub_list = reinterpret_cast<uint32_t *>(base + L.sub_list);
    b.sub_slot = reinterpret_cast<uint32_t *>(base + L.sub_slot);
    b.band_list = reinterpret_cast<uint32_t *>(base + L.band_list);
    b.sub_count = list2_count + 1;  // the caller's three counters: reads left for launch_align, pure deletions, band reads
    b.band_count = list2_count + 2;
    const size_t smem_c = subseq ? align16(6u * ((uint32_t)q.t.m + 2)) + align16(4u * (((uint32_t)q.t.m + 15) / 16 + 2)) + align16((uint32_t)q.t.m) : 0;
    const size_t smem_a = b.a_cta_bytes + (size_t)wa * b.a_warp_bytes, smem_b = b.b_cta_bytes + (size_t)wb * b.b_warp_bytes;
    static const int cap_a = getenv("TREAT_B200_BAND_CTAS") ? atoi(getenv("TREAT_B200_BAND_CTAS")) : TREAT_B200_BAND_MIN_CTAS;
    const uint64_t res_a = resident_ctas(k_band_fill, wa * 32, smem_a, num_sms, cap_a);
    const uint64_t res_b = resident_ctas(k_band_finish, wb * 32, smem_b, num_sms, 8);
    {
        const uint64_t blocks = std::min<uint64_t>((uint64_t)num_sms * 16, (q.N + kClassifyThreads - 1) / kClassifyThreads);
        k_band_classify<<<(unsigned)blocks, kClassifyThreads, smem_c, st>>>(q, b);
        launched++;
    }
    // large launches: the pure deletions (which need no fill) are finished on a second stream while the fill runs -- on a list
    // of a few ten thousand reads k_band_fill is the latency of one thread's rows, not throughput, and leaves the SMs to others
    static const bool no_overlap = getenv("TREAT_B200_NO_BAND_OVERLAP") != nullptr;
    AuxLane *lane = (subseq && !no_overlap && q.N >= 100000) ? aux_lane(st) : nullptr;
    if (lane) {
        if (cudaEventRecord(lane->fork, st) != cudaSuccess || cudaStreamWaitEvent(lane->aux, lane->fork, 0) != cudaSuccess) return (int)cudaGetLastError();
        b.first = 0u;
        b.mode = 0u;
        const uint64_t blocks_b = std::min<uint64_t>(res_b, (q.N + wb - 1) / wb);
        k_band_finish<<<(unsigned)blocks_b, wb * 32, smem_b, lane->aux>>>(q, b);
        launched++;
        if (cudaEventRecord(lane->join, lane->aux) != cudaSuccess) return (int)cudaGetLastError();
    }
    // one round holds every band read in the usual case: then a single k_band_finish takes the pure deletions and the
    // band reads together; further rounds (more band reads than the scratch cap holds) only have band reads
    for (uint64_t first = 0; first < q.N; first += b.round_reads) {
        b.first = (uint32_t)first;
        b.mode = first == 0 ? (subseq && !lane ? 2u : 1u) : 1u;
        const uint64_t reads = std::min<uint64_t>(b.round_reads, q.N - first);
        const uint64_t items = first == 0 ? q.N : reads;  // round 0 also finishes the pure deletions (at most N in all)
        const uint64_t blocks_a = std::min<uint64_t>(res_a, ((reads + 63) / 64 + wa - 1) / wa);
        const uint64_t blocks_b = std::min<uint64_t>(res_b, (items + wb - 1) / wb);
        k_band_fill<<<(unsigned)blocks_a, wa * 32, smem_a, st>>>(q, b);
        k_band_finish<<<(unsigned)blocks_b, wb * 32, smem_b, st>>>(q, b);
        launched += 2;
    }
    if (kernels_launched) *kernels_launched = launched;
    if (lane && cudaStreamWaitEvent(st, lane->join, 0) != cudaSuccess) return (int)cudaGetLastError();
    return (int)cudaGetLastError();
}

int configure_band_kernels(int max_smem_optin)
{
    cudaError_t e = cudaFuncSetAttribute(k_band_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_band_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)e;
}

}  // namespace treat_b200
