// align_kernel.cu -- K2 of the TREAT alignment hot path (sm_100a): nwalgo.Align + traceback +
// computeT / findJSS / findJES / computeAltEditing / NewAlignment (align.go:95-279) on the device.
//
// One warp per read-template pair (k_align) or per TWO pairs (k_align2).  The DP is an
// anti-diagonal wavefront: lane l owns template bases [l*CPL, (l+1)*CPL), the read streams through
// as rows, lane l works on row t-l at step t and hands its right-edge cell to lane l+1 with
// __shfl_up_sync.  Template, edit-site counts and the scoring profile live in shared memory;
// packed reads are staged by cp.async.bulk (TMA) behind an mbarrier, one batch ahead.
//
// Score transform.  nwalgo fills F[i][j] = max(F[i-1][j-1] + (a==b ? 1 : -1), F[i-1][j] - 1,
// F[i][j-1] - 1), F[i][0] = -i, F[0][j] = -j.  With G = F + i + j the gap terms vanish:
//     G[i][j] = max(G[i-1][j-1] + w, G[i-1][j], G[i][j-1]),  w = 3 on a match, 1 otherwise,
//     G[i][0] = G[0][j] = 0,  F[m][n] = G[m][n] - m - n,
// every horizontal / vertical difference of G lies in {0,1,2,3}, and nwalgo's pointer tests are
//     max == hgap ("Up",   consumes the template base)  <=>  G[i][j] == G[i-1][j]
//     max == vgap ("Left", consumes the read base)      <=>  G[i][j] == G[i][j-1].
// A row with w = 0 leaves the previous row unchanged, so idle wavefront steps need no branch.
//
// Direction storage.  Per lane and step ONE word: the CPL horizontal differences as base-4 digits
// (built with IMADs: sum_k (f_k - f_{k-1}) 4^k = 4^CPL f_last - 3 sum_k f_k 4^k - left) plus the
// vertical difference of the lane's left boundary cell on top.  From two such words (rows j, j-1) the
// traceback gets dH of the cell and dV = dV(boundary) + sum of (dH_j - dH_{j-1}) over the lane's
// columns up to the cell, hence both pointer tests; it walks 32 cells of a diagonal run per iteration.
//
// Band.  Only words whose cells lie within B diagonals of the main diagonal are kept: the fill runs
// in chunks of P = CPL+1 steps and lane l keeps chunk u iff |u - l| <= C (a window of (2C+1) P rows
// per lane in shared memory, B = C P - 1).  Any optimal path with D deletions, I insertions
// and X mismatches has 2D + I + 2X = m - S and D + 2I + 2X = n - S (S = F[m][n]), so it stays inside
// the band whenever (m - S)/2 <= B and (n - S)/2 <= B; the warp checks this after the fill and
// otherwise repeats the fill with every word written to an L2-resident global scratch.  Exact either
// way; the band only shrinks shared memory so that more warps fit on an SM.
//
// k_align2 packs two reads into the 16-bit halves of every register (VIADDMNMX.S16x2 /
// VIMNMX.S16x2, native DPX on sm_100a); the digit words of both reads come out of the same 32-bit
// IMAD chain because the packing is a ring homomorphism mod 2^32.
//
// Integer dynamic programming: no tensor cores on purpose.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "assembly.cuh"
#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

// resident warps per SM the register budget allows in k_align2 (64 registers up to 7 columns per lane)
__host__ __device__ constexpr int max_warps_pair(int cpl) { return cpl <= 7 ? 32 : cpl <= 10 ? 20 : 16; }
// k_align2's direction words: both reads' 16-bit words in one uint32 up to 7 columns per lane; beyond that two words per
// read (columns 0..6 + boundary dV, columns 7..CPL-1), i.e. a uint2 per step
template <int CPL>
using pair_word_t = typename std::conditional<(CPL <= 7), uint32_t, uint2>::type;
// one-read kernel: more columns per lane need more registers, so fewer warps
__host__ __device__ constexpr int max_warps_one(int cpl) { return cpl <= 7 ? 24 : cpl <= 12 ? 20 : 16; }

// ------------------------------------------------------------------------------------------
// direction words
// ------------------------------------------------------------------------------------------
template <int CPL>
struct Dir {
    using word_t = typename std::conditional<(CPL <= 7), uint16_t, uint32_t>::type;
    static constexpr int kTop = (CPL <= 7) ? 14 : 30;                     // bit position of dV(left boundary)
    static constexpr uint32_t kOdd = (CPL <= 7) ? 0x1555u : 0x15555555u;  // low bits of the digits
    static_assert(CPL <= 15, "direction word holds at most 15 base-4 digits");
};

// sum of the base-4 digits of x (the caller only needs it mod 4)
__device__ __forceinline__ uint32_t digit_sum(uint32_t x, uint32_t odd)
{
    return __popc(x & odd) + 2u * __popc(x & (odd << 1));
}

// ------------------------------------------------------------------------------------------
// scoring profile in shared memory: w for (read code row, lane, column of the lane) as 32-bit words,
// laid out so that every lane fetches its CPL words with conflict-free vector loads
// ------------------------------------------------------------------------------------------
template <int CPL>
struct Prof {
    static constexpr int N4 = CPL / 4, REM = CPL % 4;
    static constexpr int REMW = REM == 3 ? 4 : REM;  // words of the remainder chunk (3 -> padded to 4)
    __host__ __device__ static constexpr uint32_t bytes(int rows) { return (uint32_t)rows * 32u * 4u * (N4 * 4 + REMW); }
    // word index of (row, lane, column k)
    __device__ static int index(int rows, int row, int lane, int k)
    {
        if (k < N4 * 4) return (((k >> 2) * rows + row) * 32 + lane) * 4 + (k & 3);
        return N4 * rows * 128 + (row * 32 + lane) * REMW + (k - N4 * 4);
    }
    // per-lane view in the 32-bit shared window: both bases hold the lane term, a row costs one multiply-add each
    struct Lane {
        uint32_t a, b;
        __device__ __forceinline__ Lane(const uint32_t *base, int rows, int lane)
            : a(smem_u32(base + lane * 4)), b(smem_u32(base + N4 * rows * 128 + lane * REMW)) {}
    };
    template <int ROWS>
    __device__ __forceinline__ static void load(const Lane &pl, int row, uint32_t (&w)[CPL])
    {
#pragma unroll
        for (int q = 0; q < N4; q++) {
            const uint4 v = lds_v4(pl.a + (uint32_t)(q * ROWS + row) * 512u);
            w[4 * q] = v.x;
            w[4 * q + 1] = v.y;
            w[4 * q + 2] = v.z;
            w[4 * q + 3] = v.w;
        }
        const uint32_t rp = pl.b + (uint32_t)row * (128u * REMW);
        if (REM == 1) {
            w[4 * N4] = lds_u32(rp);
        } else if (REM == 2) {
            const uint2 v = lds_v2(rp);
            w[4 * N4] = v.x;
            w[4 * N4 + 1] = v.y;
        } else if (REM == 3) {
            const uint4 v = lds_v4(rp);
            w[4 * N4] = v.x;
            w[4 * N4 + 1] = v.y;
            w[4 * N4 + 2] = v.z;
        }
    }
};

// w of template column `col` against read code c (0..2 bases, 3 = other, 4 = pad row)
__device__ __forceinline__ uint32_t score_w(const AlignParams &p, int c, int col)
{
    if (c == kPadCode) return 0u;
    return (c < 3 && col < p.t.m && p.t.tb[col] == c) ? 3u : 1u;
}

// ------------------------------------------------------------------------------------------
// traceback (nwalgo: Up, then Left, then Diagonal).  getw(step, lane) returns the direction word
// of the read being traced.  Ops are OR-ed into the zeroed 2-bit array
// ops2 at index p (matches are 0, so only gaps are written).  Returns pos (alignment = [pos, m+n)).
// ------------------------------------------------------------------------------------------
template <int CPL, typename GetW, typename IsMatch>
__device__ __forceinline__ int traceback(GetW getw, IsMatch is_match, uint32_t tie, uint32_t *ops2, int m, int n, int lane, int &ngaps)
{
    using DT = Dir<CPL>;
    int pos = m + n, i = m, j = n;
    ngaps = 0;
    while (i > 0 && j > 0) {
        const int ii = i - lane, jj = j - lane;
        bool stop = true, is_up = false;
        if (ii >= 1 && jj >= 1) {
            const int ln = (ii - 1) / CPL, k = (ii - 1) - ln * CPL;
            const int tt = jj - 1 + ln;
            const uint32_t w1 = getw(tt, ln);
            const bool pU = ((w1 >> (2 * k)) & 3u) == 0u;  // G[i][j] == G[i-1][j]: Up is optimal
            if (pU && tie != TREAT_B200_TIE_LUD) {
                is_up = true;  // Up first (ULD, UDL)
            } else {
                // dV(i,j) = dV(left boundary of the lane, row j) + sum_{columns <= k} (dH(row j) - dH(row j-1))
                const uint32_t msk = (4u << (2 * k)) - 1u;
                const uint32_t w0 = jj >= 2 ? getw(tt - 1, ln) : 0u;  // row 0 is all zero
                const uint32_t v = (w1 >> DT::kTop) + digit_sum(w1 & msk, DT::kOdd) - digit_sum(w0 & msk, DT::kOdd);
                const bool pL = (v & 3u) == 0u;  // G[i][j] == G[i][j-1]: Left is optimal
                if (tie == TREAT_B200_TIE_ULD) {
                    stop = pL;
                } else if (tie == TREAT_B200_TIE_LUD) {
                    stop = pL || pU;
                    is_up = !pL && pU;
                } else {  // UDL: Diagonal beats Left; with Left optimal the diagonal is too iff G[i-1][j-1] + w == G[i][j-1],
                          // i.e. dH(i, j-1) == w(i, j)
                    stop = pL && ((w0 >> (2 * k)) & 3u) != (is_match(ii - 1, jj - 1) ? 3u : 1u);
                }
            }
        }
        const uint32_t sm = __ballot_sync(FULL_MASK, stop);
        const int run = sm ? __ffs(sm) - 1 : 32;  // diagonal moves before the first gap / matrix edge
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
    // first row / column: pointers are Left on row 0 and Up on column 0
    const int rest = i + j;  // at most one of them is non-zero
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

// ------------------------------------------------------------------------------------------
// fill loops.  FULL = false: words inside the band go to the shared-memory window (row t - tlo);
// FULL = true: every word goes to the warp's global scratch (row t).
// ------------------------------------------------------------------------------------------
template <int CPL, bool WIDE, bool FULL, typename dir_t>
__device__ __forceinline__ void fill_one(const uint32_t *prof, const uint8_t *rbl, const int (&tcr)[CPL], int n, int steps,
                                         int lane, int (&H)[CPL], dir_t *out, uint32_t *dump, int C)
{
    using DT = Dir<CPL>;
    constexpr int ROWS = 5, P = CPL + 1;
#pragma unroll
    for (int k = 0; k < CPL; k++) H[k] = 0;  // G[i][0] = 0
    int diag = 0;                            // G[i0-1][j-1]
    uint32_t pwc[CPL];
    int cnext = 0;
    const typename Prof<CPL>::Lane pl(prof, ROWS, lane);
    if (!WIDE) {
        Prof<CPL>::template load<ROWS>(pl, rbl[0], pwc);
        cnext = rbl[1];
    }
    const int nchunks = (steps + P - 1) / P;
    const uint32_t rbl32 = smem_u32(rbl);
    const uint32_t dump32 = smem_u32(dump + lane), out32 = FULL ? 0u : smem_u32(out + lane);
    for (int u = 0; u < nchunks; u++) {
        // lanes outside their band window write into a CTA-wide dump row: no predicate inside the chunk
        const bool keep = FULL || (unsigned)(u - lane + C) <= (unsigned)(2 * C);
        dir_t *grow = out + u * (P * 32) + lane;  // FULL: global scratch row
        const uint32_t srow = keep ? out32 + (uint32_t)(u - lane + C) * (P * 32u * (uint32_t)sizeof(dir_t)) : dump32;
        const uint32_t rbu32 = rbl32 + (uint32_t)(u * P);
#pragma unroll
        for (int v = 0; v < P; v++) {
            uint32_t pwn[CPL];
            int wv[CPL];
            if (!WIDE) {
                Prof<CPL>::template load<ROWS>(pl, cnext, pwn);
                cnext = (int)lds_u8(rbu32 + v + 2);
#pragma unroll
                for (int k = 0; k < CPL; k++) wv[k] = (int)pwc[k];
            } else {
                const int idx = u * P + v - lane;
                const bool live = idx >= 0 && idx < n;
                const int rc = live ? (int)lds_u8(rbu32 + v) : 0x200;
#pragma unroll
                for (int k = 0; k < CPL; k++) wv[k] = live ? (rc == tcr[k] ? 3 : 1) : 0;
            }
            const int recv = __shfl_up_sync(FULL_MASK, H[CPL - 1], 1);
            const int left = lane == 0 ? 0 : recv;  // G[i0-1][j]
            // sum_k (f_k - f_{k-1}) 4^k + (left - diag) 2^kTop, as multiply-adds on the f values
            uint32_t acc = ((uint32_t)(left - diag) << DT::kTop) - (uint32_t)left;
            int dg = diag, lf = left;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int e = max(dg + wv[k], H[k]);
                const int f = max(e, lf);
                dg = H[k];
                lf = f;
                H[k] = f;
                acc += (uint32_t)f * (k == CPL - 1 ? (1u << (2 * (CPL - 1))) : 0u - (3u << (2 * k)));
            }
            diag = left;
            if (FULL) grow[v * 32] = (dir_t)acc;
            else if (sizeof(dir_t) == 2) sts_u16(srow + v * 64u, acc);
            else sts_u32(srow + v * 128u, acc);
            if (!WIDE) {
#pragma unroll
                for (int k = 0; k < CPL; k++) pwc[k] = pwn[k];
            }
        }
    }
    __syncwarp();
}

template <int CPL, bool FULL>
__device__ __forceinline__ void fill_two(const uint32_t *prof, const uint8_t *pcl, int steps, int lane, uint32_t (&H)[CPL],
                                         pair_word_t<CPL> *out, uint32_t *dump, int C)
{
    constexpr int ROWS = 25, P = CPL + 1;
    constexpr bool TWO = CPL > 7;                      // two words per step
    constexpr int G0 = TWO ? 7 : CPL;                  // columns in the first word
    constexpr uint32_t kTop = 14;                      // bit position of dV(left boundary) in the first word
    constexpr uint32_t ROWB = 32u * (uint32_t)sizeof(pair_word_t<CPL>);  // bytes of one step's words
#pragma unroll
    for (int k = 0; k < CPL; k++) H[k] = 0u;
    uint32_t nd14 = 0u;  // -(G[i0-1][j-1] << kTop), carried from the previous step
    uint32_t diag = 0u;
    uint32_t pwc[CPL];
    const typename Prof<CPL>::Lane pl(prof, ROWS, lane);
    Prof<CPL>::template load<ROWS>(pl, pcl[0], pwc);
    int cnext = pcl[1];
    const int nchunks = (steps + P - 1) / P;
    const uint32_t pcl32 = smem_u32(pcl);
    const uint32_t dump32 = smem_u32(dump) + (uint32_t)lane * (uint32_t)sizeof(pair_word_t<CPL>);
    const uint32_t out32 = FULL ? 0u : smem_u32(out + lane);
    for (int u = 0; u < nchunks; u++) {
        // lanes outside their band window write into a CTA-wide dump row: no predicate inside the chunk
        const bool keep = FULL || (unsigned)(u - lane + C) <= (unsigned)(2 * C);
        pair_word_t<CPL> *grow = out + u * (P * 32) + lane;                                  // FULL: global scratch row
        const uint32_t srow = keep ? out32 + (uint32_t)(u - lane + C) * (P * ROWB) : dump32;  // banded: shared row
        const uint32_t pcu32 = pcl32 + (uint32_t)(u * P);
#pragma unroll
        for (int v = 0; v < P; v++) {
            uint32_t pwn[CPL];
            Prof<CPL>::template load<ROWS>(pl, cnext, pwn);
            cnext = (int)lds_u8(pcu32 + v + 2);
            const uint32_t recv = __shfl_up_sync(FULL_MASK, H[CPL - 1], 1);
            const uint32_t left = lane == 0 ? 0u : recv;
            // digits + dV(boundary) for both halves at once (mod 2^32 keeps each 16-bit word exact; left >= diag per half):
            //   sum_{k<G0} (f_k - f_{k-1}) 4^k + (left - diag) 2^kTop
            // = f_{G0-1} 4^(G0-1) - 3 sum_{k<G0-1} f_k 4^k - left + left 2^kTop - diag 2^kTop
            // and, for the columns of the second word (its "left" is f_{G0-1}):
            //   sum_{k>=G0} (f_k - f_{k-1}) 4^(k-G0) = f_last 4^(CPL-1-G0) - 3 sum_{G0<=k<CPL-1} f_k 4^(k-G0) - f_{G0-1}
            const uint32_t nl14 = left * (0u - (1u << kTop));
            uint32_t acc = nd14 - nl14 - left, acc2 = 0u;
            nd14 = nl14;
            uint32_t dg = diag, lf = left;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const uint32_t e = __viaddmax_s16x2(dg, pwc[k], H[k]);
                const uint32_t f = __vmaxs2(e, lf);
                dg = H[k];
                lf = f;
                H[k] = f;
                if (k < G0) {
                    acc += f * (k == G0 - 1 ? (1u << (2 * (G0 - 1))) : 0u - (3u << (2 * k)));
                    if (TWO && k == G0 - 1) acc2 -= f;
                } else {
                    constexpr int last2 = CPL - 1 - G0 > 0 ? CPL - 1 - G0 : 0;
                    const int k2 = k >= G0 ? k - G0 : 0;
                    acc2 += f * (k == CPL - 1 ? (1u << (2 * last2)) : 0u - (3u << (2 * k2)));
                }
            }
            diag = left;
            if constexpr (TWO) {
                if (FULL) grow[v * 32] = make_uint2(acc, acc2);
                else sts_v2(srow + v * ROWB, acc, acc2);
            } else {
                if (FULL) grow[v * 32] = acc;
                else sts_u32(srow + v * ROWB, acc);
            }
#pragma unroll
            for (int k = 0; k < CPL; k++) pwc[k] = pwn[k];
        }
    }
    __syncwarp();
}

// the 16-bit (CPL <= 7) or 32-bit direction word of read h (0 / 1) out of a pair word, in the layout traceback<CPL> reads
template <int CPL>
__device__ __forceinline__ uint32_t pair_dir_word(const pair_word_t<CPL> &w, int h)
{
    if constexpr (CPL <= 7) {
        return (w >> (16 * h)) & 0xffffu;
    } else {
        const uint32_t w0 = (w.x >> (16 * h)) & 0xffffu, w1 = (w.y >> (16 * h)) & 0xffffu;
        return (w0 & 0x3fffu) | (w1 << 14) | ((w0 >> 14) << 30);  // digits 0..6, digits 7.., dV(boundary) at Dir<CPL>::kTop = 30
    }
}

// the optimal path stays within B diagonals of the main one (see the header comment)
__device__ __forceinline__ bool band_ok(int m, int n, int score, int B) { return B >= 0 && m - score <= 2 * B && n - score <= 2 * B; }

// ------------------------------------------------------------------------------------------
// K2a: one read per warp (any CPL <= 15; the only kernel of the wide path)
// ------------------------------------------------------------------------------------------
template <int CPL, bool WIDE>
__global__ void __launch_bounds__(max_warps_one(CPL) * 32, 1) k_align(const __grid_constant__ AlignParams p)
{
    using DT = Dir<CPL>;
    using dir_t = typename DT::word_t;
    using cnt_t = typename std::conditional<WIDE, uint32_t, uint8_t>::type;
    constexpr int ROWS = 5;

    extern __shared__ __align__(128) uint8_t smem[];
    if (p.list && *p.list_count == 0u) return;  // nothing was left for the DP: skip the template / profile set-up
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m;

    const CtaSmem<cnt_t> cs = carve_cta<cnt_t>(smem, p, CPL);
    load_template(cs, p, CPL);
    if (!WIDE) {
        for (int i = threadIdx.x; i < ROWS * 32 * CPL; i += blockDim.x) {
            const int k = i % CPL, ln = (i / CPL) & 31, c = i / (CPL * 32);
            cs.prof[Prof<CPL>::index(ROWS, c, ln, k)] = score_w(p, c, ln * CPL + k);
        }
    }
    const WarpSmem ws = carve_warp(smem, p, warp);
    const uint32_t stage_bytes = p.sm_stage_b + p.sm_stage_c;
    dir_t *s_dir = reinterpret_cast<dir_t *>(ws.dir);
    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        fence_mbar_init();
    }
    ws.rb[lane] = kPadCode;  // rows "before" the read: no-op rows for lanes that have not started
    __syncthreads();

    int tcr[CPL];  // template bases of this lane's columns (the wide path compares instead of a profile)
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const int idx = lane * CPL + k;
        tcr[k] = idx < m ? (int)cs.tb[idx] : 0x100;  // 0x100 never equals a read byte
    }
    const int lanes_used = (m + CPL - 1) / CPL;
    const int ops_words = (int)(p.sm_ops / 4);
    constexpr int P = CPL + 1;
    const int C = (int)p.band_chunks, B = C * P - 1;  // B < 0: nothing is certified, every read takes the full path
    const uint32_t gw = blockIdx.x * nwarps + warp, nw = gridDim.x * nwarps;  // a launch holds < 2^32 reads (host check)
    dir_t *scratch = reinterpret_cast<dir_t *>(p.scratch) + gw * (uint64_t)p.scratch_stride;
    unsigned long long ctr = 0;

    auto issue = [&](uint32_t r, int slot) {
        if (lane == 0) {
            const uint32_t n = p.n[r];
            const uint32_t cb = WIDE ? align16(n) : align16((n + 3) / 4);
            const uint32_t cc = WIDE ? align16(4 * (n + 1)) : align16(n + 1);
            uint8_t *dst = ws.stage + slot * stage_bytes;
            fence_proxy_async();
            mbar_expect_tx(&ws.bar[slot], cb + cc);
            if (cb) bulk_g2s(dst, p.bases + r * (uint64_t)p.bases_stride, cb, &ws.bar[slot]);
            bulk_g2s(dst + p.sm_stage_b, p.counts + r * (uint64_t)p.counts_stride, cc, &ws.bar[slot]);
        }
    };

    // the reads of this launch: all of them, or the ones k_same_bases listed
    const uint32_t total = p.list ? *p.list_count : (uint32_t)p.N;
    auto read_at = [&](uint32_t u) -> uint32_t { return p.list ? p.list[u] : u; };
    uint32_t it = 0;
    if (gw < total) issue(read_at(gw), 0);
    for (uint32_t u = gw; u < total; u += nw, it++) {
        const int slot = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const uint32_t r = read_at(u);
        const int n = (int)p.n[r];
        const uint32_t rflags = p.rflags[r];
        __syncwarp();
        if (nw < total - u) issue(read_at(u + nw), slot ^ 1);  // prefetch the next read of this warp
        mbar_wait(&ws.bar[slot], parity);
        if (n > (int)p.ncap) {  // only on a speculative launch (ncap guessed): the host re-runs the chunk
            if (lane == 0 && p.skipped) atomicAdd(p.skipped, 1u);
            continue;
        }

        const uint8_t *st_b = ws.stage + slot * stage_bytes;
        const cnt_t *rcount = reinterpret_cast<const cnt_t *>(st_b + p.sm_stage_b);

        // ---- unpack read bases to one byte each; rows behind the read are no-op rows -----------
        uint8_t *rb = ws.rb + 32;
        if (!WIDE) {
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(st_b);
            for (int wi = lane; wi * 16 < n; wi += 32) {
                const uint32_t w = pw[wi];
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t x = (w >> (8 * q)) & 0xffu;
                    o[q] = (x & 3u) | ((x & 0xcu) << 6) | ((x & 0x30u) << 12) | ((x & 0xc0u) << 18);
                }
                *reinterpret_cast<uint4 *>(rb + wi * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
            rb[n + lane] = kPadCode;
            rb[n + 32 + lane] = kPadCode;
            if (lane < 16) rb[n + 64 + lane] = kPadCode;
        } else {
            for (int i = lane; i < n; i += 32) rb[i] = st_b[i];
        }
        for (int i = lane; i < ops_words; i += 32) ws.ops2[i] = 0u;
        __syncwarp();

        // ---- fill (band in shared memory), score, certificate ------------------------------------
        int H[CPL];
        const int steps = n > 0 ? n + lanes_used - 1 : 0;
        const uint8_t *rbl = rb - lane;  // rbl[t] = read base of this lane's row at step t
        auto final_score = [&]() {
            // F[m][n] = G[m][n] - m - n; G[m][n] lives in lane (m-1)/CPL, column (m-1)%CPL
            int fin = 0;
            const int kl = (m - 1) % CPL;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kl) fin = H[k];
            return __shfl_sync(FULL_MASK, fin, (m - 1) / CPL) - m - n;
        };
        // S <= min(m, n) - |m - n|, so the certificate (m - S <= 2B and n - S <= 2B) cannot hold when |m - n| > B:
        // such a read (a truncated one, typically) goes straight to the full fill
        bool full = p.force_full || abs(m - n) > B;
        int score;
        for (;;) {
            if (full) fill_one<CPL, WIDE, true>(cs.prof, rbl, tcr, n, steps, lane, H, scratch, cs.dump, 0);  // every word kept (global scratch, L2)
            else fill_one<CPL, WIDE, false>(cs.prof, rbl, tcr, n, steps, lane, H, s_dir, cs.dump, C);
            score = final_score();
            if (full || (score == m && n == m) || band_ok(m, n, score, B)) break;
            full = true;  // the path may leave the band: repeat the fill
        }
        int ngaps = 0, pos = n;  // alignment = ops2[pos, m + n)
        auto same_base = [&](int ti, int fi) { return (int)rb[fi] == (int)cs.tb[ti]; };  // template base ti == read base fi
        if (score == m && n == m) {
            // every base matches: the only optimal path is the main diagonal (m zero ops, already zeroed)
        } else if (!full) {
            pos = traceback<CPL>(
                [&](int tt, int ln) {
                    TB_CHECK(p, (unsigned)(tt - (ln - C) * P) < p.band_rows && (unsigned)ln < 32u, 0);
                    return (uint32_t)s_dir[(tt - (ln - C) * P) * 32 + ln];
                },
                same_base, p.tie, ws.ops2, m, n, lane, ngaps);
        } else {
            pos = traceback<CPL>(
                [&](int tt, int ln) {
                    TB_CHECK(p, (unsigned)(tt * 32 + ln) < p.scratch_stride, 1);
                    return (uint32_t)scratch[tt * 32 + ln];
                },
                same_base, p.tie, ws.ops2, m, n, lane, ngaps);
        }
        finish_read<cnt_t>(p, cs, ws, lane, r, n, rflags, score, pos, ngaps, rcount, [&](int fi) { return (int)rb[fi]; }, ctr);
    }
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// ------------------------------------------------------------------------------------------
// K2c: templates longer than one pass of the wavefront (32 lanes x 15 columns = 480 bases).  One read per warp; the
// template is cut into column strips of 480 bases and the wavefront runs once per strip over all rows of the read.
// The strip's right-edge column G[480 (s+1)][j] stays in a per-warp shared-memory array and is the left boundary of
// strip s+1 (lane 0 reads edge[j] where the one-pass kernel has the constant 0; lane 31 writes its last cell back in
// place, 31 steps behind the read).  Every direction word goes to the warp's global scratch, strip after strip, so
// the traceback and everything behind it are the one-pass kernel's with a strip term in the word address.
// ------------------------------------------------------------------------------------------
constexpr int kStripCpl = 15, kStripCols = 32 * kStripCpl;

template <bool WIDE>
__device__ __forceinline__ void fill_strip(const uint32_t *prof, const uint8_t *rbl, const int (&tcr)[kStripCpl], int n, int steps,
                                           int lane, int (&H)[kStripCpl], uint32_t *out, int *edge, bool first, bool has_next)
{
    constexpr int CPL = kStripCpl;
    using DT = Dir<CPL>;
    constexpr int ROWS = 5, P = CPL + 1;
#pragma unroll
    for (int k = 0; k < CPL; k++) H[k] = 0;  // G[i][0] = 0
    int diag = 0;                            // G[i0-1][j-1]; column i0-1 of row 0 is 0 in every strip
    uint32_t pwc[CPL];
    int cnext = 0;
    const typename Prof<CPL>::Lane pl(prof, ROWS, lane);
    if (!WIDE) {
        Prof<CPL>::template load<ROWS>(pl, rbl[0], pwc);
        cnext = rbl[1];
    }
    const int nchunks = (steps + P - 1) / P;
    const uint32_t rbl32 = smem_u32(rbl);
    // lane 0's left boundary: 0 in the first strip, else the previous strip's right edge; rows behind the read keep
    // the value of row n (a no-op row needs left <= G of its own first column, and the edge is being overwritten)
    int e_cur = (first || n == 0) ? 0 : edge[1];
    for (int u = 0; u < nchunks; u++) {
        uint32_t *grow = out + u * (P * 32) + lane;
        const uint32_t rbu32 = rbl32 + (uint32_t)(u * P);
#pragma unroll
        for (int v = 0; v < P; v++) {
            const int t = u * P + v;
            uint32_t pwn[CPL];
            int wv[CPL];
            if (!WIDE) {
                Prof<CPL>::template load<ROWS>(pl, cnext, pwn);
                cnext = (int)lds_u8(rbu32 + v + 2);
#pragma unroll
                for (int k = 0; k < CPL; k++) wv[k] = (int)pwc[k];
            } else {
                const int idx = t - lane;
                const bool live = idx >= 0 && idx < n;
                const int rc = live ? (int)lds_u8(rbu32 + v) : 0x200;
#pragma unroll
                for (int k = 0; k < CPL; k++) wv[k] = live ? (rc == tcr[k] ? 3 : 1) : 0;
            }
            int e_nxt = e_cur;
            if (!first && lane == 0 && t + 2 <= n) e_nxt = edge[t + 2];  // row t + 2 of the next step, one step ahead
            const int recv = __shfl_up_sync(FULL_MASK, H[CPL - 1], 1);
            const int left = lane == 0 ? e_cur : recv;  // G[i0-1][j]
            uint32_t acc = ((uint32_t)(left - diag) << DT::kTop) - (uint32_t)left;
            int dg = diag, lf = left;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int e = max(dg + wv[k], H[k]);
                const int f = max(e, lf);
                dg = H[k];
                lf = f;
                H[k] = f;
                acc += (uint32_t)f * (k == CPL - 1 ? (1u << (2 * (CPL - 1))) : 0u - (3u << (2 * k)));
            }
            diag = left;
            grow[v * 32] = acc;
            // right edge of a full strip: lane 31 finished row t - 30 in this step
            if (has_next && lane == 31 && t >= 31 && t - 30 <= n) edge[t - 30] = H[CPL - 1];
            e_cur = e_nxt;
            if (!WIDE) {
#pragma unroll
                for (int k = 0; k < CPL; k++) pwc[k] = pwn[k];
            }
        }
    }
    __syncwarp();
}

template <bool WIDE>
__global__ void __launch_bounds__(max_warps_one(kStripCpl) * 32, 1) k_align_strips(const __grid_constant__ AlignParams p)
{
    constexpr int CPL = kStripCpl, ROWS = 5;
    using cnt_t = typename std::conditional<WIDE, uint32_t, uint8_t>::type;

    extern __shared__ __align__(128) uint8_t smem[];
    if (p.list && *p.list_count == 0u) return;  // nothing was left for the DP
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m, strips = (int)p.strips;

    const CtaSmem<cnt_t> cs = carve_cta<cnt_t>(smem, p, CPL * strips);
    load_template(cs, p, CPL * strips);
    constexpr uint32_t kProfWords = Prof<CPL>::bytes(ROWS) / 4;
    if (!WIDE) {
        for (int i = threadIdx.x; i < ROWS * 32 * CPL * strips; i += blockDim.x) {
            const int k = i % CPL, ln = (i / CPL) & 31, c = (i / (CPL * 32)) % ROWS, s = i / (CPL * 32 * ROWS);
            cs.prof[s * kProfWords + Prof<CPL>::index(ROWS, c, ln, k)] = score_w(p, c, s * kStripCols + ln * CPL + k);
        }
    }
    const WarpSmem ws = carve_warp(smem, p, warp);
    int *edge = reinterpret_cast<int *>(ws.bar + 2);  // [ncap + 2] right edge of the strip filled last
    const uint32_t stage_bytes = p.sm_stage_b + p.sm_stage_c;
    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        fence_mbar_init();
    }
    ws.rb[lane] = kPadCode;
    __syncthreads();

    const int ops_words = (int)(p.sm_ops / 4);
    const uint32_t gw = blockIdx.x * nwarps + warp, nw = gridDim.x * nwarps;
    uint32_t *scratch = reinterpret_cast<uint32_t *>(p.scratch) + gw * (uint64_t)p.scratch_stride;
    const int strip_rows = (int)(p.scratch_stride / (32u * p.strips));  // scratch rows of one strip
    unsigned long long ctr = 0;

    auto issue = [&](uint32_t r, int slot) {
        if (lane == 0) {
            const uint32_t n = p.n[r];
            const uint32_t cb = WIDE ? align16(n) : align16((n + 3) / 4);
            const uint32_t cc = WIDE ? align16(4 * (n + 1)) : align16(n + 1);
            uint8_t *dst = ws.stage + slot * stage_bytes;
            fence_proxy_async();
            mbar_expect_tx(&ws.bar[slot], cb + cc);
            if (cb) bulk_g2s(dst, p.bases + r * (uint64_t)p.bases_stride, cb, &ws.bar[slot]);
            bulk_g2s(dst + p.sm_stage_b, p.counts + r * (uint64_t)p.counts_stride, cc, &ws.bar[slot]);
        }
    };

    // the reads of this launch: all of them, or the ones k_same_long listed
    const uint32_t total = p.list ? *p.list_count : (uint32_t)p.N;
    auto read_at = [&](uint32_t u) -> uint32_t { return p.list ? p.list[u] : u; };
    uint32_t it = 0;
    if (gw < total) issue(read_at(gw), 0);
    for (uint32_t u = gw; u < total; u += nw, it++) {
        const int slot = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const uint32_t r = read_at(u);
        const int n = (int)p.n[r];
        const uint32_t rflags = p.rflags[r];
        __syncwarp();
        if (nw < total - u) issue(read_at(u + nw), slot ^ 1);
        mbar_wait(&ws.bar[slot], parity);
        if (n > (int)p.ncap) {  // speculative launch (ncap guessed): the host re-runs the chunk
            if (lane == 0 && p.skipped) atomicAdd(p.skipped, 1u);
            continue;
        }
        const uint8_t *st_b = ws.stage + slot * stage_bytes;
        const cnt_t *rcount = reinterpret_cast<const cnt_t *>(st_b + p.sm_stage_b);

        uint8_t *rb = ws.rb + 32;
        if (!WIDE) {
            const uint32_t *pw = reinterpret_cast<const uint32_t *>(st_b);
            for (int wi = lane; wi * 16 < n; wi += 32) {
                const uint32_t w = pw[wi];
                uint32_t o[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t x = (w >> (8 * q)) & 0xffu;
                    o[q] = (x & 3u) | ((x & 0xcu) << 6) | ((x & 0x30u) << 12) | ((x & 0xc0u) << 18);
                }
                *reinterpret_cast<uint4 *>(rb + wi * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
            rb[n + lane] = kPadCode;
            rb[n + 32 + lane] = kPadCode;
            if (lane < 16) rb[n + 64 + lane] = kPadCode;
        } else {
            for (int i = lane; i < n; i += 32) rb[i] = st_b[i];
        }
        for (int i = lane; i < ops_words; i += 32) ws.ops2[i] = 0u;
        __syncwarp();

        // ---- fill, strip after strip ------------------------------------------------------------
        int H[CPL];
        const uint8_t *rbl = rb - lane;
        for (int s = 0; s < strips; s++) {
            const int cols = min(kStripCols, m - s * kStripCols);
            const int lanes_used = (cols + CPL - 1) / CPL;
            const int steps = n > 0 ? n + lanes_used - 1 : 0;
            int tcr[CPL];
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int idx = s * kStripCols + lane * CPL + k;
                tcr[k] = idx < m ? (int)cs.tb[idx] : 0x100;
            }
            TB_CHECK(p, ((steps + CPL) / (CPL + 1)) * (CPL + 1) <= strip_rows, 7);
            fill_strip<WIDE>(cs.prof + s * kProfWords, rbl, tcr, n, steps, lane, H, scratch + (size_t)s * strip_rows * 32, edge,
                             s == 0, s + 1 < strips);
        }
        // F[m][n] = G[m][n] - m - n; G[m][n] lives in the last strip
        int score;
        {
            const int ml = m - 1 - (strips - 1) * kStripCols;
            int fin = 0;
            const int kl = ml % CPL;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kl) fin = H[k];
            score = __shfl_sync(FULL_MASK, fin, ml / CPL) - m - n;
        }
        int ngaps = 0, pos = n;
        if (!(score == m && n == m)) {
            pos = traceback<CPL>(
                [&](int tt, int ln) {  // ln counts lanes across strips; the word of (strip, lane, step tt - 32 strip)
                    const int s = ln >> 5, l = ln & 31;
                    const int row = s * strip_rows + (tt - ln + l);
                    TB_CHECK(p, (unsigned)(row * 32 + l) < p.scratch_stride, 1);
                    return scratch[row * 32 + l];
                },
                [&](int ti, int fi) { return (int)rb[fi] == (int)cs.tb[ti]; }, p.tie, ws.ops2, m, n, lane, ngaps);
        }
        finish_read<cnt_t>(p, cs, ws, lane, r, n, rflags, score, pos, ngaps, rcount, [&](int fi) { return (int)rb[fi]; }, ctr);
    }
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// ------------------------------------------------------------------------------------------
// K2b: TWO reads per warp in the 16-bit halves of every register (packed alphabet, CPL <= 7)
// ------------------------------------------------------------------------------------------
template <int CPL>
__global__ void __launch_bounds__(max_warps_pair(CPL) * 32, 1) k_align2(const __grid_constant__ AlignParams p)
{
    static_assert(CPL <= 15, "the second direction word holds at most 8 columns");
    using cnt_t = uint8_t;
    constexpr int ROWS = 25;  // (code of read A) * 5 + (code of read B)

    extern __shared__ __align__(128) uint8_t smem[];
    if (p.list && *p.list_count == 0u) return;  // nothing was left for the DP: skip the template / profile set-up
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m;

    const CtaSmem<cnt_t> cs = carve_cta<cnt_t>(smem, p, CPL);
    load_template(cs, p, CPL);
    for (int i = threadIdx.x; i < ROWS * 32 * CPL; i += blockDim.x) {
        const int k = i % CPL, ln = (i / CPL) & 31, row = i / (CPL * 32);
        const int col = ln * CPL + k;
        cs.prof[Prof<CPL>::index(ROWS, row, ln, k)] = score_w(p, row / 5, col) | (score_w(p, row % 5, col) << 16);
    }
    const WarpSmem ws = carve_warp(smem, p, warp);
    const uint32_t read_bytes = p.sm_stage_b + p.sm_stage_c;  // one read inside a stage slot
    const uint32_t slot_bytes = 2 * read_bytes;
    pair_word_t<CPL> *s_dir = reinterpret_cast<pair_word_t<CPL> *>(ws.dir);
    if (lane == 0) {
        mbar_init(&ws.bar[0], 1);
        mbar_init(&ws.bar[1], 1);
        fence_mbar_init();
    }
    ws.rb[lane] = kPadCode * 5 + kPadCode;
    __syncthreads();

    const int lanes_used = (m + CPL - 1) / CPL;
    const int ops_words = (int)(p.sm_ops / 4);
    constexpr int P = CPL + 1;
    const int C = (int)p.band_chunks, B = C * P - 1;
    // the reads of this launch: all of them, or the ones k_same_bases listed; pair q = reads 2q, 2q+1 of that sequence
    const uint32_t total = p.list ? *p.list_count : (uint32_t)p.N;  // a launch holds < 2^32 reads (host check)
    auto read_at = [&](uint32_t u) -> uint32_t { return p.list ? p.list[u] : u; };
    const uint32_t npairs = total / 2 + (total & 1u);
    const uint32_t gw = blockIdx.x * nwarps + warp, nw = gridDim.x * nwarps;
    pair_word_t<CPL> *scratch = reinterpret_cast<pair_word_t<CPL> *>(p.scratch) + gw * (uint64_t)p.scratch_stride;
    unsigned long long ctr = 0;

    auto issue = [&](uint32_t q, int slot) {
        if (lane == 0) {
            uint8_t *dst = ws.stage + slot * slot_bytes;
            const uint32_t r0 = read_at(2 * q);
            const uint32_t n0 = p.n[r0];
            const bool has1 = 2 * q + 1 < total;
            const uint32_t r1 = has1 ? read_at(2 * q + 1) : r0;
            const uint32_t n1 = has1 ? p.n[r1] : 0u;
            const uint32_t cb0 = align16((n0 + 3) / 4), cc0 = align16(n0 + 1);
            const uint32_t cb1 = has1 ? align16((n1 + 3) / 4) : 0u, cc1 = has1 ? align16(n1 + 1) : 0u;
            fence_proxy_async();
            mbar_expect_tx(&ws.bar[slot], cb0 + cc0 + cb1 + cc1);
            if (cb0) bulk_g2s(dst, p.bases + r0 * (uint64_t)p.bases_stride, cb0, &ws.bar[slot]);
            bulk_g2s(dst + p.sm_stage_b, p.counts + r0 * (uint64_t)p.counts_stride, cc0, &ws.bar[slot]);
            if (cb1) bulk_g2s(dst + read_bytes, p.bases + r1 * (uint64_t)p.bases_stride, cb1, &ws.bar[slot]);
            if (cc1) bulk_g2s(dst + read_bytes + p.sm_stage_b, p.counts + r1 * (uint64_t)p.counts_stride, cc1, &ws.bar[slot]);
        }
    };

    uint32_t it = 0;
    if (gw < npairs) issue(gw, 0);
    for (uint32_t q = gw; q < npairs; q += nw, it++) {
        const int slot = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const uint32_t r0 = read_at(2 * q);
        const bool has1 = 2 * q + 1 < total;
        const uint32_t r1 = has1 ? read_at(2 * q + 1) : r0;
        const int n0 = (int)p.n[r0], n1 = has1 ? (int)p.n[r1] : 0;
        __syncwarp();
        if (nw < npairs - q) issue(q + nw, slot ^ 1);  // prefetch the next pair of this warp
        mbar_wait(&ws.bar[slot], parity);
        if (n0 > (int)p.ncap || n1 > (int)p.ncap) {  // only on a speculative launch: the host re-runs the chunk
            if (lane == 0 && p.skipped) atomicAdd(p.skipped, 1u);
            continue;
        }

        const uint8_t *st0 = ws.stage + slot * slot_bytes, *st1 = st0 + read_bytes;
        const int nmax = max(n0, n1);

        // ---- pair codes: pc[j] = 5 * codeA(j) + codeB(j), pad code behind either read; 4 per lane ---
        uint8_t *pc = ws.rb + 32;
        {
            const uint32_t *pw0 = reinterpret_cast<const uint32_t *>(st0), *pw1 = reinterpret_cast<const uint32_t *>(st1);
            auto four = [](const uint32_t *pw, int j0, int n) {
                // codes j0..j0+3 as bytes; positions >= n hold the pad code
                const uint32_t x = j0 < n ? (pw[j0 >> 4] >> (2 * (j0 & 15))) & 0xffu : 0u;
                const uint32_t c = (x & 3u) | ((x & 0xcu) << 6) | ((x & 0x30u) << 12) | ((x & 0xc0u) << 18);
                const int live = min(max(n - j0, 0), 4);
                const uint32_t keep = live >= 4 ? 0xffffffffu : (1u << (8 * live)) - 1u;
                return (c & keep) | (0x04040404u & ~keep);
            };
            for (int j0 = 4 * lane; j0 < nmax + 80; j0 += 128)
                *reinterpret_cast<uint32_t *>(pc + j0) = four(pw0, j0, n0) * 5u + four(pw1, j0, n1);
        }
        __syncwarp();

        // ---- fill: both reads at once, G values in the 16-bit halves ----------------------------
        uint32_t H[CPL];
        const int steps = nmax > 0 ? nmax + lanes_used - 1 : 0;
        const uint8_t *pcl = pc - lane;
        // S <= min(m, n) - |m - n|: a read with |m - n| > B can never pass the band certificate, so such a pair goes
        // straight to the full fill (every word kept in the global scratch)
        bool full = p.force_full || abs(m - n0) > B || (has1 && abs(m - n1) > B);
        int score0, score1;
        for (;;) {
            if (full) fill_two<CPL, true>(cs.prof, pcl, steps, lane, H, scratch, cs.dump, 0);
            else fill_two<CPL, false>(cs.prof, pcl, steps, lane, H, s_dir, cs.dump, C);
            uint32_t fin = 0;
            const int kl = (m - 1) % CPL;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kl) fin = H[k];
            fin = __shfl_sync(FULL_MASK, fin, (m - 1) / CPL);
            score0 = (int)(fin & 0xffffu) - m - n0;
            score1 = (int)(fin >> 16) - m - n1;
            if (full || (band_ok(m, n0, score0, B) && (!has1 || band_ok(m, n1, score1, B)))) break;
            full = true;  // a path may leave the band: repeat the fill
        }
        const bool banded = !full;

        // ---- per read: traceback + editing-site assembly ---------------------------------------
        for (int h = 0; h < (has1 ? 2 : 1); h++) {
            const int n = h ? n1 : n0;
            const uint8_t *st = h ? st1 : st0;
            const cnt_t *rcount = reinterpret_cast<const cnt_t *>(st + p.sm_stage_b);
            for (int i = lane; i < ops_words; i += 32) ws.ops2[i] = 0u;
            __syncwarp();
            int ngaps = 0, pos = n;  // alignment = ops2[pos, m + n)
            const int score = h ? score1 : score0;
            auto same_base = [&](int ti, int fi) {  // template base ti == base fi of read h
                const int c = pc[fi];
                return (h ? c % 5 : c / 5) == (int)cs.tb[ti];
            };
            if (score == m && n == m) {
                // every base matches: the only optimal path is the main diagonal (m zero ops, already zeroed)
            } else if (banded)
                pos = traceback<CPL>(
                    [&](int tt, int ln) {
                        TB_CHECK(p, (unsigned)(tt - (ln - C) * P) < p.band_rows && (unsigned)ln < 32u, 0);
                        return pair_dir_word<CPL>(s_dir[(tt - (ln - C) * P) * 32 + ln], h);
                    },
                    same_base, p.tie, ws.ops2, m, n, lane, ngaps);
            else
                pos = traceback<CPL>(
                    [&](int tt, int ln) {
                        TB_CHECK(p, (unsigned)(tt * 32 + ln) < p.scratch_stride, 1);
                        return pair_dir_word<CPL>(scratch[tt * 32 + ln], h);
                    },
                    same_base, p.tie, ws.ops2, m, n, lane, ngaps);
            const uint32_t r = h ? r1 : r0;
            finish_read<cnt_t>(p, cs, ws, lane, r, n, p.rflags[r], score, pos, ngaps, rcount,
                               [&](int fi) {
                                   const int c = pc[fi];
                                   return h ? c % 5 : c / 5;
                               },
                               ctr);
        }
    }
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// ------------------------------------------------------------------------------------------
// K2s: reads whose stripped bases equal the template's (kFlagSameBases, set by k_pack) need no DP.
// The identity alignment scores m, every other path scores at most m - 2 (one gap costs a base of
// each sequence, a mismatch costs 2), so it is the unique optimum and nwalgo's traceback returns it
// whatever its tie order: score = m, no gap, no mismatch, read base i sits on template base i.
// Such a read goes straight to the editing-site assembly, reading its counts from HBM; every other
// read is appended to p.list for the DP kernels.  One warp per read, next read's loads in flight.
// ------------------------------------------------------------------------------------------
constexpr int kSameWarps = 16;

template <int G>
__global__ void __launch_bounds__(kSameWarps * 32, 2) k_same_bases(const __grid_constant__ AlignParams p)
{
    using cnt_t = uint8_t;
    constexpr int RPW = 32 / G;  // reads per warp
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int gl = lane & (G - 1), grp = lane / G, lead = grp * G;
    const unsigned gmask = G == 32 ? FULL_MASK : 0xffffu << lead;
    const int m = p.t.m;

    const CtaSmem<cnt_t> cs = carve_cta<cnt_t>(smem, p, (int)(p.cpl * p.strips));  // no profile, no dump rows in this plan
    load_template(cs, p, (int)(p.cpl * p.strips));
    uint32_t *zero = reinterpret_cast<uint32_t *>(smem + p.sm_tmpl_bytes);  // the all-match ops of such a read
    for (uint32_t i = threadIdx.x; i < p.sm_ops / 4; i += blockDim.x) zero[i] = 0u;
    uint4 *s_mask = reinterpret_cast<uint4 *>(smem + p.sm_tmpl_bytes + p.sm_ops);  // [17]: t lowest bytes 0xff
    if (threadIdx.x < 17) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = min(max((int)threadIdx.x - 4 * i, 0), 4);
            w[i] = k >= 4 ? 0xffffffffu : (1u << (8 * k)) - 1u;
        }
        s_mask[threadIdx.x] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint4 *stage = reinterpret_cast<uint4 *>(smem + p.sm_tmpl_bytes + p.sm_ops + 17 * 16 + (size_t)warp * p.sm_cnt_stage);
    WarpSmem ws{};
    ws.ops2 = zero;
    __syncthreads();

    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, nw = (uint64_t)gridDim.x * nwarps;
    const bool mine = gl * 16 < m + 1;  // this lane's 16 counts exist
    const SameLane sl = load_same_lane(cs, p, gl);
    unsigned long long ctr = 0;
    auto fetch = [&](uint64_t r, uint32_t &f, uint4 &c) {
        f = 0u;
        c = make_uint4(0, 0, 0, 0);
        if (r < p.N) {
            f = p.rflags[r];
            if (mine) c = reinterpret_cast<const uint4 *>(p.counts + r * (uint64_t)p.counts_stride)[gl];
        }
    };
    uint32_t f = 0, fn = 0;
    uint4 c = make_uint4(0, 0, 0, 0), cn = c;
    fetch(gw * RPW + grp, f, c);
    for (uint64_t r0 = gw * RPW; r0 < p.N; r0 += nw * RPW) {
        const uint64_t r = r0 + grp;
        fetch(r + nw * RPW, fn, cn);
        bool general = false;  // the group's read needs the general assembly
        if (r < p.N) {
            if (f & kFlagsInternal) general = !finish_same<G>(p, cs, s_mask, gl, gmask, lead, r, f, c, sl, ctr);
            else if (gl == 0) p.list[atomicAdd(p.list_count, 1u)] = (uint32_t)r;
        }
        // an alternative template's region: the whole warp runs the general assembly, one read at a time
        unsigned todo = __ballot_sync(FULL_MASK, general && gl == 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            __syncwarp();
            if (lead == src) stage[gl] = c;
            __syncwarp();
            finish_same_general(p, cs, ws, lane, r0 + src / G, __shfl_sync(FULL_MASK, f, src), reinterpret_cast<const cnt_t *>(stage), ctr);
        }
        f = fn;
        c = cn;
    }
    if (G == 16) ctr += __shfl_down_sync(FULL_MASK, ctr, 16);  // lane L and lane L + 16 both hold counter L
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// The same for templates of more than 511 bases (up to 1024 edit sites): one warp per read, the read's counts staged in
// shared memory, the general assembly on the identity alignment.  A few hundred instructions against the ~10^6 cell
// updates of the DP such a read would otherwise take.
__global__ void __launch_bounds__(kSameWarps * 32, 2) k_same_long(const __grid_constant__ AlignParams p)
{
    using cnt_t = uint8_t;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m, tcpl = (int)(p.cpl * p.strips);
    const CtaSmem<cnt_t> cs = carve_cta<cnt_t>(smem, p, tcpl);
    load_template(cs, p, tcpl);
    uint32_t *zero = reinterpret_cast<uint32_t *>(smem + p.sm_tmpl_bytes);  // the all-match ops of such a read
    for (uint32_t i = threadIdx.x; i < p.sm_ops / 4; i += blockDim.x) zero[i] = 0u;
    uint4 *stage = reinterpret_cast<uint4 *>(smem + p.sm_tmpl_bytes + p.sm_ops + (size_t)warp * p.sm_cnt_stage);
    WarpSmem ws{};
    ws.ops2 = zero;
    __syncthreads();
    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, nw = (uint64_t)gridDim.x * nwarps;
    const int groups = (m + 1 + 15) / 16;
    unsigned long long ctr = 0;
    for (uint64_t r = gw; r < p.N; r += nw) {
        const uint32_t f = p.rflags[r];
        if (!(f & kFlagsInternal)) {
            if (lane == 0) p.list[atomicAdd(p.list_count, 1u)] = (uint32_t)r;
            continue;
        }
        const uint4 *src = reinterpret_cast<const uint4 *>(p.counts + r * (uint64_t)p.counts_stride);
        __syncwarp();
        for (int g = lane; g < groups; g += 32) stage[g] = src[g];
        __syncwarp();
        finish_same_general(p, cs, ws, lane, r, f, reinterpret_cast<const cnt_t *>(stage), ctr);
    }
    flush_summary(p, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// ------------------------------------------------------------------------------------------
// host side: geometry + launch
// ------------------------------------------------------------------------------------------
static const int kCplSet[] = {1, 2, 4, 6, 7, 10, 12, 15};

static int pick_cpl(int m)
{
    const int need = (m + 31) / 32;
    for (int c : kCplSet)
        if (c >= need) return c;
    return -1;
}

static uint32_t prof_bytes(int cpl, int rows)
{
    const int n4 = cpl / 4, rem = cpl % 4, remw = rem == 3 ? 4 : rem;
    return (uint32_t)rows * 32u * 4u * (uint32_t)(n4 * 4 + remw);
}

int plan_align(AlignParams &p, bool wide)
{
    const int m = p.t.m, K = p.t.K;
    if (m < 1 || m > kMaxTemplateBases) return -1;
    // longer than one pass of the wavefront: column strips of 480 bases, one read per warp (k_align_strips)
    const int strips = m > kStripCols ? (m + kStripCols - 1) / kStripCols : 1;
    const int cpl = strips > 1 ? kStripCpl : pick_cpl(m);
    if (cpl < 0) return -1;
    p.strips = (uint32_t)strips;
    const bool pair = !wide && !p.force_single && strips == 1;
    p.pair = pair ? 1 : 0;
    p.cpl = (uint32_t)cpl;
    const uint32_t ncap = p.ncap;
    const uint32_t csz = wide ? 4u : 1u;
    const uint32_t dsz = pair ? (cpl <= 7 ? 4u : 8u) : (cpl <= 7 ? 2u : 4u);
    const uint32_t sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    p.sm_prof_bytes = wide ? 0u : prof_bytes(cpl, pair ? 25 : 5) * (uint32_t)strips;
    p.sm_tmpl_bytes = align16(32 * cpl * strips) + align16((uint32_t)(K * p.t.len_pad) * csz) +
                      align16(8u * (p.t.n_alt ? p.t.n_alt : 1)) + align16(8u * sum_len) + p.sm_prof_bytes;
    p.sm_tmpl_bytes += 16u * 32u * 8u;  // + dump rows (up to 16 steps of 8-byte words)
    p.sm_tmpl_bytes = (p.sm_tmpl_bytes + 127u) & ~127u;
    p.sm_stage_b = wide ? align16(ncap) : align16((ncap + 3) / 4);
    if (p.sm_stage_b == 0) p.sm_stage_b = 16;
    p.sm_stage_c = wide ? align16(4 * (ncap + 1)) : align16(ncap + 1);
    p.sm_stage_total = 2 * (pair ? 2 : 1) * (p.sm_stage_b + p.sm_stage_c);
    p.sm_rb = align16(32 + ncap + 80 + 16);
    // band window: (2C+1) chunks of P = cpl+1 rows per lane; the fill pads its steps to whole chunks
    const uint32_t P = (uint32_t)cpl + 1;
    const uint32_t full_rows = ncap + 48;
    const uint32_t want = p.band_rows ? p.band_rows : (pair ? 35u : 64u);
    uint32_t C = want / P >= 1 ? (want / P - 1) / 2 : 0;
    while (C > 0 && (2 * C + 1) * P > full_rows) C--;
    p.band_chunks = C;
    p.band_rows = (2 * C + 1) * P;
    p.sm_dir = strips > 1 ? 0u : align16(p.band_rows * 32 * dsz);  // strips: every word goes to the scratch
    p.scratch_stride = full_rows * 32 * (uint32_t)strips;  // direction words of one warp's full fill
    p.sm_ops = align16((((uint32_t)m + ncap + 15) / 16 + 2) * 4);
    p.sm_t2f = align16(2u * (uint32_t)p.t.len_pad);
    p.sm_warp_bytes = p.sm_stage_total + p.sm_rb + p.sm_dir + p.sm_ops + p.sm_t2f + 16;
    if (strips > 1) p.sm_warp_bytes += align16(4u * (ncap + 2));  // right-edge column handed from strip to strip
    p.sm_warp_bytes = (p.sm_warp_bytes + 127u) & ~127u;
    return cpl;
}

// warps per CTA for this geometry (0: does not fit)
int align_warps(const AlignParams &p, int max_smem)
{
    long warps = ((long)max_smem - (long)p.sm_tmpl_bytes) / (long)p.sm_warp_bytes;
    const long cap = p.pair ? max_warps_pair((int)p.cpl) : max_warps_one((int)p.cpl);
    if (warps > cap) warps = cap;
    return warps < 1 ? 0 : (int)warps;
}

// bytes of global scratch the launch needs (num_sms CTAs x warps x scratch_stride words)
uint64_t align_scratch_bytes(const AlignParams &p, int num_sms, int max_smem)
{
    const uint32_t dsz = p.pair ? (p.cpl <= 7 ? 4u : 8u) : (p.cpl <= 7 ? 2u : 4u);
    return (uint64_t)num_sms * (uint64_t)align_warps(p, max_smem) * (uint64_t)p.scratch_stride * dsz;
}

template <typename Kern>
static int launch_k(Kern kern, const AlignParams &p, uint64_t units, int num_sms, int max_smem, void *stream)
{
    const int warps = align_warps(p, max_smem);
    if (warps < 1) return -1000;  // geometry does not fit shared memory
    const uint64_t need_blocks = (units + warps - 1) / warps;
    uint64_t blocks = (uint64_t)num_sms;  // persistent: one CTA per SM
    if (need_blocks < blocks) blocks = need_blocks;
    if (blocks == 0) return 0;
    const size_t smem = p.sm_tmpl_bytes + (size_t)warps * p.sm_warp_bytes;
    kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

template <int CPL>
static int launch_cpl(const AlignParams &p, bool wide, int num_sms, int max_smem, void *stream)
{
    if (p.pair) return launch_k(k_align2<CPL>, p, (p.N + 1) / 2, num_sms, max_smem, stream);
    return wide ? launch_k(k_align<CPL, true>, p, p.N, num_sms, max_smem, stream)
                : launch_k(k_align<CPL, false>, p, p.N, num_sms, max_smem, stream);
}

int launch_align(const AlignParams &p, bool wide, int num_sms, int max_smem_optin, void *stream)
{
    if (p.strips > 1)
        return wide ? launch_k(k_align_strips<true>, p, p.N, num_sms, max_smem_optin, stream)
                    : launch_k(k_align_strips<false>, p, p.N, num_sms, max_smem_optin, stream);
    switch (pick_cpl(p.t.m)) {
    case 1: return launch_cpl<1>(p, wide, num_sms, max_smem_optin, stream);
    case 2: return launch_cpl<2>(p, wide, num_sms, max_smem_optin, stream);
    case 4: return launch_cpl<4>(p, wide, num_sms, max_smem_optin, stream);
    case 6: return launch_cpl<6>(p, wide, num_sms, max_smem_optin, stream);
    case 7: return launch_cpl<7>(p, wide, num_sms, max_smem_optin, stream);
    case 10: return launch_cpl<10>(p, wide, num_sms, max_smem_optin, stream);
    case 12: return launch_cpl<12>(p, wide, num_sms, max_smem_optin, stream);
    case 15: return launch_cpl<15>(p, wide, num_sms, max_smem_optin, stream);
    default: return -1001;
    }
}

template <typename Kern>
static int kernel_regs(Kern kern)
{
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, kern) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return a.numRegs;
}

template <int CPL>
static int regs_cpl(const AlignParams &p, bool wide)
{
    if (p.pair) return kernel_regs(k_align2<CPL>);
    return wide ? kernel_regs(k_align<CPL, true>) : kernel_regs(k_align<CPL, false>);
}

int align_kernel_regs(const AlignParams &p, bool wide)
{
    if (p.strips > 1) return wide ? kernel_regs(k_align_strips<true>) : kernel_regs(k_align_strips<false>);
    switch (pick_cpl(p.t.m)) {
    case 1: return regs_cpl<1>(p, wide);
    case 2: return regs_cpl<2>(p, wide);
    case 4: return regs_cpl<4>(p, wide);
    case 6: return regs_cpl<6>(p, wide);
    case 7: return regs_cpl<7>(p, wide);
    case 10: return regs_cpl<10>(p, wide);
    case 12: return regs_cpl<12>(p, wide);
    case 15: return regs_cpl<15>(p, wide);
    default: return 0;
    }
}

int launch_same_bases(const AlignParams &planned, int num_sms, void *stream)
{
    AlignParams p = planned;
    const int m = p.t.m;
    const uint32_t sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    p.sm_prof_bytes = 0;
    p.sm_tmpl_bytes = align16(32 * p.cpl * p.strips) + align16((uint32_t)(p.t.K * p.t.len_pad)) + align16(8u * (p.t.n_alt ? p.t.n_alt : 1)) +
                      align16(8u * sum_len);
    p.sm_tmpl_bytes = (p.sm_tmpl_bytes + 127u) & ~127u;
    p.sm_ops = align16((((uint32_t)(2 * m) + 15) / 16 + 2) * 4);
    if (p.t.len > 512) {  // k_same_long: the counts of one read per warp, 16 bytes past the last site
        p.sm_cnt_stage = align16((uint32_t)p.t.len) + 16u;
        if (p.ncap < (uint32_t)m) p.ncap = (uint32_t)m;
        const size_t smem_l = p.sm_tmpl_bytes + p.sm_ops + (size_t)kSameWarps * p.sm_cnt_stage;
        int per = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_same_long, kSameWarps * 32, smem_l) != cudaSuccess || per < 1) per = 1;
        uint64_t nb = (p.N + kSameWarps - 1) / kSameWarps;
        if (nb > (uint64_t)num_sms * per) nb = (uint64_t)num_sms * per;
        if (nb == 0) return 0;
        k_same_long<<<(unsigned)nb, kSameWarps * 32, smem_l, (cudaStream_t)stream>>>(p);
        return (int)cudaGetLastError();
    }
    p.sm_cnt_stage = 512;  // 32 lanes x 16 counts >= m + 1
    if (p.ncap < (uint32_t)m) p.ncap = (uint32_t)m;
    const size_t smem = p.sm_tmpl_bytes + p.sm_ops + 17 * 16 + (size_t)kSameWarps * p.sm_cnt_stage;
    const bool two = p.t.len <= 256;  // 16 lanes x 16 sites hold the read: two reads per warp
    int per_sm = 0;
    cudaError_t oe = two ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_same_bases<16>, kSameWarps * 32, smem)
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_same_bases<32>, kSameWarps * 32, smem);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint64_t per_block = (uint64_t)kSameWarps * (two ? 2 : 1);
    uint64_t blocks = (p.N + per_block - 1) / per_block;
    if (blocks > (uint64_t)num_sms * per_sm) blocks = (uint64_t)num_sms * per_sm;
    if (blocks == 0) return 0;
    if (two)
        k_same_bases<16><<<(unsigned)blocks, kSameWarps * 32, smem, (cudaStream_t)stream>>>(p);
    else
        k_same_bases<32><<<(unsigned)blocks, kSameWarps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

template <int CPL>
static int set_attr(int max_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_align<CPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_align<CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_align2<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    return (int)e;
}

__global__ void k_merge_summary(unsigned long long *dst, unsigned long long *src, uint32_t len)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
        const unsigned long long v = src[i];
        if (v) {
            atomicAdd(&dst[i], v);
            src[i] = 0ull;
        }
    }
}

// ESS x junction-length heat map (cmd/treat/handlers.go:578-589): heat[EditStop - EditOffset][JuncLen] += ReadCount for
// records with EditStop >= EditOffset (the reference adds Norm, the ReadCount-derived weight of `treat norm`).
// One thread per record; lanes of a warp that hit the same cell combine first, so hot cells see one atomic per warp.
__global__ void __launch_bounds__(256) k_heatmap(const treat_b200_record *rec, uint64_t N, int edit_offset, uint32_t n,
                                                 unsigned long long *heat)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (N + stride - 1) / stride;  // whole warps stay in the loop for the warp votes
    for (uint64_t k = 0; k < rounds; k++) {
        const uint64_t i = k * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t cell = 0xffffffffu, rc = 0;
        if (i < N) {
            const int es = rec[i].edit_stop - edit_offset;
            const uint32_t jl = rec[i].junc_len;
            if (es >= 0 && (uint32_t)es < n && jl < n) {
                cell = (uint32_t)es * n + jl;
                rc = rec[i].read_count;
            }
        }
        const unsigned peers = __match_any_sync(FULL_MASK, cell);
        // REDUX adds 32-bit values; 32 read counts may need 37 bits: add the halves separately
        unsigned long long tot;
        const uint32_t lo = __reduce_add_sync(peers, rc & 0xffffu), hi = __reduce_add_sync(peers, rc >> 16);
        tot = (unsigned long long)lo + ((unsigned long long)hi << 16);
        if (cell != 0xffffffffu && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&heat[cell], tot);
    }
}

int launch_heatmap(const treat_b200_record *d_rec, uint64_t N, int edit_offset, uint32_t n, uint64_t *d_heat, int num_sms, void *stream)
{
    if (!N) return 0;
    uint64_t blocks = (N + 255) / 256;
    if (blocks > (uint64_t)num_sms * 8) blocks = (uint64_t)num_sms * 8;
    k_heatmap<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_rec, N, edit_offset, n, reinterpret_cast<unsigned long long *>(d_heat));
    return (int)cudaGetLastError();
}

int launch_merge_summary(uint64_t *dst, uint64_t *src, uint32_t len, void *stream)
{
    k_merge_summary<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *>(dst),
                                                         reinterpret_cast<unsigned long long *>(src), len);
    return (int)cudaGetLastError();
}

int configure_align_kernels(int max_smem_optin)
{
    int e;
    if ((e = set_attr<1>(max_smem_optin))) return e;
    if ((e = set_attr<2>(max_smem_optin))) return e;
    if ((e = set_attr<4>(max_smem_optin))) return e;
    if ((e = set_attr<6>(max_smem_optin))) return e;
    if ((e = set_attr<7>(max_smem_optin))) return e;
    if ((e = set_attr<10>(max_smem_optin))) return e;
    if ((e = set_attr<12>(max_smem_optin))) return e;
    if ((e = set_attr<15>(max_smem_optin))) return e;
    if ((e = (int)cudaFuncSetAttribute(k_same_long, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin))) return e;
    if ((e = (int)cudaFuncSetAttribute(k_align_strips<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin))) return e;
    if ((e = (int)cudaFuncSetAttribute(k_align_strips<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin))) return e;
    return 0;
}

}  // namespace treat_b200
