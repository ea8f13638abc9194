// frontier_kernel.cu -- K0: the reads that are "the template, edited up to a frontier" are finished straight from their
// raw characters; no base is stripped, no edit-site count is formed.
//
// With R the upper-cased raw read and F / P the fully edited / pre-edited template re-inflated to raw strings (the rows
// Template.EditSite[0] / [1] of template.go:41-49 around Template.Bases), NewFragment + NewAlignment (fragment.go:91-121,
// align.go:95-279) of a read whose stripped bases equal the template's reduce to two string compares:
//   * findJSS (align.go:110-120): the highest site whose count differs from FE is m - bs, bs = bases inside the longest
//     common SUFFIX of R and F; T[0] is all ones iff R == F;
//   * findJES (align.go:95-108): the lowest site whose count differs from PE is bp, bp = bases inside the longest common
//     PREFIX of R and P; T[1] is all ones iff R == P;
//   * the bases ARE the template's iff the characters between that prefix and that suffix, edit bases dropped, are
//     Bases[bp .. m - bs) (and, when prefix and suffix overlap, iff bp + the bases behind the prefix add up to m);
//   * JuncSeq (align.go:247-258) is the window of R from just behind base bp - 1 to just behind base m - bs: both offsets
//     are table look-ups on P and F.
// tests/test_frontier_model.py states the same in Python and checks it against the CPU restatement of the Go path.
//
// One group of G lanes per read, 32 / G reads per warp; lane gl holds the 16-byte blocks gl, G + gl, 2 G + gl, ... (eight of
// them) of the read's 128 G-byte window, so every load, of the read and of the tables, is a run of consecutive 16-byte blocks
// across the group: coalesced in HBM, and -- with the windows of a warp an odd multiple of 64 bytes apart when two reads share
// a quarter-warp -- conflict-free in shared memory.  Per block the four words are XORed against P (aligned to the read's
// start) or F (aligned to its end) -- 16 shifted copies of each in shared memory make every operand an aligned LDS.128 --
// with case folding ((w ^ P) & 0xDF.. per byte is exact for letters); P is walked upwards and F downwards, and a lane stops
// loading on a side at its first / last differing block (predicated loads: its registers keep that block, from which the byte
// position follows); one min / max butterfly per side gives the prefix length t and the suffix length s.  The characters in
// between (the junction: 33 on average on RPS12, with a long tail) are checked by the whole warp, 32 characters per round from
// the staged copy of the blocks, against the template's bases (rank = ballot prefix count); the first round of four reads is
// in flight together.
// Everything else -- a substitution, an indel, a foreign character, a junction start on an alternative template's region
// (align.go:186-205), a read whose last run differs from FE's (findJSS may answer -1), a read longer than a pass -- is
// appended to f.rest and taken by k_pack_finish2 / k_pack_finish exactly as before.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "assembly.cuh"
#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

namespace {
constexpr uint32_t kFrNB = 64;      // a read's window starts at a 64-byte boundary of memory
constexpr uint32_t kFrGuardP = 64;  // zero bytes in front of P inside a shifted copy (a block of P starts up to 48 bytes early)
constexpr uint32_t kFrGuardF = 16;  // zero block in front of F (block offsets are clamped into the copy)
// at most this many warps per CTA (each owns two staging halves of 32 / G windows; the launcher takes as many as fit next to
// the image) (one CTA per SM: one copy of the template image, the rest of the shared memory stages reads -- 28 warps with RPS12)
#ifndef TREAT_B200_FR_WARPS
#define TREAT_B200_FR_WARPS 28
#endif
#ifndef TREAT_B200_FR_MIN_CTAS
#define TREAT_B200_FR_MIN_CTAS 1
#endif
constexpr int kFrWarps = TREAT_B200_FR_WARPS;
__host__ __device__ constexpr uint32_t fr_copy_bytes(uint32_t guard, uint32_t len) { return guard + ((15u + len + 15u) & ~15u) + 16u; }
}  // namespace

// ---- host: the shared-memory image of a template (built once per treat_b200_set_template) ----------------------
// bases: m upper-case letters; fe / pe: m + 1 counts each; alt_start: n_alt site numbers.  Returns false when the kernel
// does not apply (characters that are not upper-case letters, strings too long for a pass, counts near saturation).
bool build_frontier_image(const uint8_t *bases, int m, const uint32_t *fe, const uint32_t *pe, const int32_t *alt_start, int n_alt,
                          uint8_t edit_base, std::vector<uint8_t> &image, FrontierGeom &g)
{
    g = FrontierGeom{};
    if (m < 1 || m > 511 || edit_base < 'A' || edit_base > 'Z') return false;
    uint32_t max_fe = 0, max_pe = 0;
    std::vector<uint8_t> F, P;
    std::vector<uint16_t> ppos(m + 1, 0), ftail(m + 1, 0);
    for (int i = 0; i <= m; i++) {
        max_fe = std::max(max_fe, fe[i]);
        max_pe = std::max(max_pe, pe[i]);
        if (fe[i] > 254u || pe[i] > 254u) return false;
        F.insert(F.end(), fe[i], edit_base);
        P.insert(P.end(), pe[i], edit_base);
        if (i < m) {
            if (bases[i] < 'A' || bases[i] > 'Z' || bases[i] == edit_base) return false;
            F.push_back(bases[i]);
            P.push_back(bases[i]);
            ppos[i + 1] = (uint16_t)P.size();  // raw offset just behind base i of P
        }
    }
    if (max_fe + max_pe > 200u) return false;  // a run in the read = a piece of a P run + the junction + a piece of an F run: stays below 255
    const uint32_t Lf = (uint32_t)F.size(), Lp = (uint32_t)P.size();
    if (std::max(Lf, Lp) + 16u > 2048u - (kFrNB - 1u)) return false;
    {
        int k = 0;
        for (int i = (int)Lf - 1; i >= 0; i--)
            if (F[i] != edit_base) ftail[++k] = (uint16_t)(Lf - 1 - i);  // characters of F behind base m - k
    }
    g.Lf = Lf, g.Lp = Lp, g.tm = (uint32_t)m;
    // G lanes per read, 8 blocks of 16 bytes per lane: the smallest window (64-byte aligned, so 63 bytes may be lost in
    // front) that holds reads a little longer than the longer template string.  Windows of one warp lie ws bytes apart in the
    // staging area: enough for such a read, and an odd multiple of 64 when two reads share a quarter-warp (G = 4), so that
    // their 64-byte rows fall into different halves of the 32 banks
    {
        const uint32_t need = std::max(Lf, Lp) + 16u + (kFrNB - 1u);
        g.G = 0;
        for (uint32_t G : {4u, 8u, 16u})
            if (need <= G * 128u) {
                g.G = G;
                break;
            }
        if (g.G == 0) return false;
        g.ws = (need + 63u) & ~63u;
        if (g.G == 4u && ((g.ws >> 6) & 1u) == 0u) g.ws += 64u;
    }
    g.mid_cap = 254u - max_fe - max_pe;
    g.sp = fr_copy_bytes(kFrGuardP, Lp), g.sf = fr_copy_bytes(kFrGuardF, Lf);
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) {
        const uint32_t at = o;
        o += (bytes + 15u) & ~15u;
        return at;
    };
    g.off_keep = take(17 * 16);  // row c: the c lowest bytes 0xDF
    g.off_p = take(16 * g.sp);
    g.off_f = take(16 * g.sf);
    g.off_pcum = take(2 * (Lp + 1));
    g.off_fsuf = take(2 * (Lf + 1));
    g.off_ppos = take(2 * (m + 1));
    g.off_ftail = take(2 * (m + 1));
    g.off_bases = take((uint32_t)m + 48);  // zeros behind the bases: a rank past the template never matches
    g.off_alt = take((uint32_t)m + 1);
    g.image_bytes = o;
    image.assign(o, 0);
    for (uint32_t c = 0; c <= 16; c++) memset(&image[g.off_keep + c * 16], 0xdf, c);
    for (uint32_t c = 0; c < 16; c++) {
        memcpy(&image[g.off_p + c * g.sp + kFrGuardP + c], P.data(), Lp);
        memcpy(&image[g.off_f + c * g.sf + kFrGuardF + c], F.data(), Lf);
    }
    uint16_t *pcum = reinterpret_cast<uint16_t *>(&image[g.off_pcum]), *fsuf = reinterpret_cast<uint16_t *>(&image[g.off_fsuf]);
    for (uint32_t i = 0; i < Lp; i++) pcum[i + 1] = (uint16_t)(pcum[i] + (P[i] != edit_base));
    for (uint32_t i = 0; i < Lf; i++) fsuf[i + 1] = (uint16_t)(fsuf[i] + (F[Lf - 1 - i] != edit_base));
    memcpy(&image[g.off_ppos], ppos.data(), 2 * (m + 1));
    memcpy(&image[g.off_ftail], ftail.data(), 2 * (m + 1));
    memcpy(&image[g.off_bases], bases, m);
    for (int k = 0; k < n_alt; k++)
        if (alt_start[k] >= 0 && alt_start[k] <= m) image[g.off_alt + alt_start[k]] = 1;
    return true;
}

namespace {

}  // namespace

template <int G, bool SPAN>
__global__ void __launch_bounds__(kFrWarps * 32, TREAT_B200_FR_MIN_CTAS) k_frontier(const __grid_constant__ PackParams p, const __grid_constant__ AlignParams q,
                                                                                    const __grid_constant__ FrontierParams f)
{
    constexpr int RPW = 32 / G;
    constexpr int NBL = 8;                // 16-byte blocks per lane
    const uint32_t ws = f.g.ws;           // bytes between two reads' windows in the staging area (a window is 64-byte aligned in memory)
    const uint32_t half = (uint32_t)RPW * ws;  // one half of a warp's staging area: the windows of its 32 / G reads
    const uint32_t wcap = min(ws, 16u * G * NBL);  // a read is taken when it ends inside its window (G lanes x NBL blocks)
    constexpr uint32_t JS = 16u * G;      // a lane's blocks are JS bytes apart
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1), grp = lane / G;
    const uint32_t lt = (1u << lane) - 1u;
    const FrontierGeom &g = f.g;
    // ---- the template's image, this CTA's histograms, one 2 KB staging area per warp ---------------------------------
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(f.image);
        uint4 *dst = reinterpret_cast<uint4 *>(smem);
        for (uint32_t i = threadIdx.x; i < g.image_bytes / 16; i += blockDim.x) dst[i] = src[i];
    }
    const uint32_t sum_len = q.summary ? TREAT_B200_SUMMARY_HEADER + 3u * q.bins : 0u;
    const uint32_t sum_bytes = align16(8u * sum_len);
    uint32_t *sum_lo = reinterpret_cast<uint32_t *>(smem + g.image_bytes), *sum_hi = sum_lo + sum_len;
    for (uint32_t i = threadIdx.x; i < 2 * sum_len; i += blockDim.x) sum_lo[i] = 0u;
    __syncthreads();
    const uint32_t s32 = smem_u32(smem);
    const uint32_t warp32 = s32 + g.image_bytes + sum_bytes + (uint32_t)warp * (2u * half);  // two halves: the reads being worked on, the next ones
    const uint32_t lane32 = warp32 + (uint32_t)grp * ws + 16u * (uint32_t)gl;     // this lane's first block inside a half
    const uint32_t keep32 = s32 + g.off_keep, pc32 = s32 + g.off_p, fc32 = s32 + g.off_f;
    const uint16_t *s_pcum = reinterpret_cast<const uint16_t *>(smem + g.off_pcum), *s_fsuf = reinterpret_cast<const uint16_t *>(smem + g.off_fsuf);
    const uint16_t *s_ppos = reinterpret_cast<const uint16_t *>(smem + g.off_ppos), *s_ftail = reinterpret_cast<const uint16_t *>(smem + g.off_ftail);
    const uint32_t bases32 = s32 + g.off_bases;
    const uint8_t *s_alt = smem + g.off_alt;
    const uint32_t Lp = g.Lp, Lf = g.Lf, tm = g.tm;
    const uint32_t eb = p.edit_base;
    const int K = q.t.K;
    const int p_top = (int)g.sp - 16, f_top = (int)g.sf - 16;  // the zero block that ends a copy

    const uint32_t base_lo = (uint32_t)reinterpret_cast<uintptr_t>(p.seqs);
    const uint32_t N32 = (uint32_t)p.N;
    const int64_t lo = SPAN ? 0 : p.N ? (int64_t)(p.off[0] - p.off_bias) : 0;
    const int64_t hi = SPAN ? (int64_t)p.span : p.N ? (int64_t)(p.off[p.N] - p.off_bias) : 0;
    auto where = [&](uint32_t r, uint64_t &x0, uint64_t &x1) {
        x0 = x1 = SPAN ? 0ull : p.off_bias;
        if (r >= N32) return;
        if (SPAN) {
            x0 = p.begin[r];
            x1 = p.len[r];
        } else {
            x0 = p.off[r];
            x1 = p.off[r + 1];
        }
    };
    auto resolve = [&](uint64_t x0, uint64_t x1, uint64_t &b, uint32_t &L) {
        b = SPAN ? x0 : x0 - p.off_bias;
        L = SPAN ? (uint32_t)x1 : (uint32_t)(x1 - x0);
    };
    // this lane's four 16-byte blocks of a read's window go straight to the staging area (cp.async, L2 only); blocks behind
    // the read are not copied -- whatever the area holds there is masked out or never looked at.  A window whose used
    // blocks stick out of [lo, hi) -- at the ends of a caller's buffer -- is read byte by byte.
    auto stage_blocks = [&](uint64_t b, uint32_t L, uint32_t dst32) {
        const uint32_t a = (base_lo + (uint32_t)b) & (kFrNB - 1u);
        const uint32_t nb16 = (a + L + 15u) >> 4;  // 16-byte blocks the read touches
        if (L == 0u || a + L > wcap) return;
        const int64_t w0 = (int64_t)b - a;  // the window starts here
        if (w0 >= lo && w0 + (int64_t)(16u * nb16) <= hi) {
            const uint8_t *src = p.seqs + w0 + 16 * gl;
#pragma unroll
            for (int j = 0; j < NBL; j++)
                if ((uint32_t)(G * j + gl) < nb16)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst32 + JS * j), "l"(src + JS * j) : "memory");
            return;
        }
#pragma unroll 1
        for (int j = 0; j < NBL; j++) {
            if ((uint32_t)(G * j + gl) >= nb16) break;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; c4++) {
                uint32_t w = 0u;
#pragma unroll 1
                for (int c = 0; c < 4; c++) {
                    const int64_t y = w0 + (int64_t)(JS * j + 16u * gl) + 4 * c4 + c;
                    if (y >= lo && y < hi) w |= (uint32_t)p.seqs[y] << (8 * c);
                }
                sts_u32(dst32 + JS * j + 4u * c4, w);
            }
        }
    };

    uint32_t n_done = 0;  // reads this warp finished (warp-uniform)
    unsigned long long rc_done = 0;  // ... and the sum of their ReadCount
    const uint32_t units = (uint32_t)((p.N + RPW - 1) / RPW);
    const uint32_t wpc = blockDim.x >> 5;  // warps per CTA: as many as the shared memory holds next to the image
    const uint32_t gw = blockIdx.x * wpc + warp, nw = gridDim.x * wpc;
    uint64_t b_nxt = 0, x0_nxt = 0, x1_nxt = 0;
    uint32_t L_cur = 0, L_nxt = 0, a_cur = 0;
    if (gw < units) {
        uint64_t x0, x1;
        where(gw * RPW + grp, x0, x1);
        resolve(x0, x1, b_nxt, L_cur);
        stage_blocks(b_nxt, L_cur, lane32);
        a_cur = (base_lo + (uint32_t)b_nxt) & (kFrNB - 1u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    where((gw + nw) * RPW + grp, x0_nxt, x1_nxt);
    uint32_t buf = 0;  // 0 / half: which half of the warp's staging area holds the current reads
    for (uint32_t u = gw; u < units; u += nw) {
        // the next reads' blocks start their way into the other half (its readers of two iterations ago are past their
        // __syncwarp); the current ones have arrived when all but that newest group are complete
        resolve(x0_nxt, x1_nxt, b_nxt, L_nxt);
        stage_blocks(b_nxt, L_nxt, lane32 + (buf ^ half));
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const uint32_t stage32 = warp32 + buf;
        const uint32_t mystage32 = lane32 + buf;  // this lane's block j sits JS j bytes further
        const uint32_t a_nxt = (base_lo + (uint32_t)b_nxt) & (kFrNB - 1u);
        where((u + 2 * nw) * RPW + grp, x0_nxt, x1_nxt);

        const uint32_t r = u * RPW + grp;
        const uint32_t a = a_cur;  // the read starts at byte a of its window
        const bool in_batch = r < N32;
        const uint32_t L = (in_batch && a + L_cur <= wcap) ? L_cur : 0u;
        const uint32_t read_count = (in_batch && q.read_count) ? q.read_count[r] : 1u;

        // ---- R against P from the read's start and against F from its end ---------------------------------------
        // window byte X is P[X - a]; P[i] sits at byte kFrGuardP + c + i of copy c = a & 15
        const int po = (int)(kFrGuardP + 16u * (uint32_t)gl) - (int)(a & ~15u);  // >= 16
        const uint32_t pa32 = pc32 + (a & 15u) * g.sp;
        // window byte X is F[X - d], d = a + L - Lf; F[i] sits at byte kFrGuardF + c + i of copy c = d & 15
        const int d = (int)(a + L) - (int)Lf;
        const int fo = (int)(kFrGuardF + 16u * (uint32_t)gl) - (d & ~15);
        const uint32_t fa32 = fc32 + ((uint32_t)d & 15u) * g.sf;
        const int e_hi = (int)(a + L) - 16 * gl, e_lo = (int)a - 16 * gl;  // the read covers bytes [e_lo, e_hi) counted from this lane's first block
        // blocks past the longest read of the warp are skipped by all lanes; rows (one block per lane) that lie inside every
        // read of the warp need no mask but the case fold
        const uint32_t nb_max = __reduce_max_sync(FULL_MASK, (a + L + 15u) >> 4);
        const uint32_t nb_min = __reduce_min_sync(FULL_MASK, L ? (a + L) >> 4 : 0xffffu);
        // window offsets of the group's first difference from P and last difference from F
        uint32_t gp = 0x7fffffffu;
        int gf = -1;
        {
            // P is walked from the lane's first block upwards, F from its last block downwards, and a lane stops loading on a
            // side once one of its blocks differs there (only its first difference from P and its last one from F count): the
            // registers then keep the words of exactly that block, and the two walks together touch each row about once.
            // bit j of mp / mf: block j differs (bits set after the stop are leftovers and lie on the far side)
            constexpr uint32_t DF = 0xdfdfdfdfu;
            uint32_t mp = 0u, mf = 0u;
            uint4 wp = make_uint4(0u, 0u, 0u, 0u), pw = wp, wf = wp, fw = wp;
#pragma unroll
            for (int i = 0; i < NBL; i++) {
                {
                    const int j = i;
                    if (!(j >= 2 && (uint32_t)(G * j) >= nb_max)) {
                        lds_v4x2_if_zero(mp, mystage32 + JS * j, wp, pa32 + (uint32_t)min(po + (int)(JS * j), p_top), pw);
                        uint32_t o;
                        if (j == 0) {  // only the first block of a lane can start in front of the read (a < 64)
                            const uint4 kl = lds_v4(keep32 + 16u * (uint32_t)min(max(e_lo, 0), 16));
                            o = ((wp.x ^ pw.x) & ~kl.x) | ((wp.y ^ pw.y) & ~kl.y) | ((wp.z ^ pw.z) & ~kl.z) | ((wp.w ^ pw.w) & ~kl.w);
                        } else {
                            o = (wp.x ^ pw.x) | (wp.y ^ pw.y) | (wp.z ^ pw.z) | (wp.w ^ pw.w);
                        }
                        mp |= (o & DF) ? (1u << j) : 0u;
                    }
                }
                {
                    const int j = NBL - 1 - i;
                    if (!(j >= 2 && (uint32_t)(G * j) >= nb_max)) {
                        lds_v4x2_if_zero(mf, mystage32 + JS * j, wf, fa32 + (uint32_t)min(max(fo + (int)(JS * j), 0), f_top), fw);
                        uint32_t o;
                        if ((uint32_t)(G * (j + 1)) <= nb_min) {  // (warp-uniform) the row ends inside every read
                            o = ((wf.x ^ fw.x) | (wf.y ^ fw.y) | (wf.z ^ fw.z) | (wf.w ^ fw.w)) & DF;
                        } else {
                            const uint4 kp = lds_v4(keep32 + 16u * (uint32_t)min(max(e_hi - (int)(JS * j), 0), 16));
                            o = ((wf.x ^ fw.x) & kp.x) | ((wf.y ^ fw.y) & kp.y) | ((wf.z ^ fw.z) & kp.z) | ((wf.w ^ fw.w) & kp.w);
                        }
                        mf |= o ? (1u << j) : 0u;
                    }
                }
            }
            if (mp) {  // (bytes behind the read's end are not masked on this side: a difference there gives t = L below)
                const int j = __ffs((int)mp) - 1;
                const uint4 kl = lds_v4(keep32 + 16u * (uint32_t)min(max(e_lo - (int)JS * j, 0), 16));
                const uint32_t x0 = (wp.x ^ pw.x) & DF & ~kl.x, x1 = (wp.y ^ pw.y) & DF & ~kl.y, x2 = (wp.z ^ pw.z) & DF & ~kl.z, x3 = (wp.w ^ pw.w) & DF & ~kl.w;
                const bool b0 = x0 != 0u, b1 = x1 != 0u, b2 = x2 != 0u;
                const uint32_t k = b0 ? 0u : b1 ? 4u : b2 ? 8u : 12u;
                const uint32_t w = b0 ? x0 : b1 ? x1 : b2 ? x2 : x3;
                gp = JS * (uint32_t)j + 16u * (uint32_t)gl + k + ((uint32_t)(__ffs((int)w) - 1) >> 3);
            }
            if (mf) {  // (bytes in front of the read's start are not masked on this side: a difference there gives s = L below)
                const int j = 31 - __clz((int)mf);
                const uint4 kp = lds_v4(keep32 + 16u * (uint32_t)min(max(e_hi - (int)JS * j, 0), 16));
                const uint32_t x0 = (wf.x ^ fw.x) & kp.x, x1 = (wf.y ^ fw.y) & kp.y, x2 = (wf.z ^ fw.z) & kp.z, x3 = (wf.w ^ fw.w) & kp.w;
                const bool b3 = x3 != 0u, b2 = x2 != 0u, b1 = x1 != 0u;
                const uint32_t k = b3 ? 12u : b2 ? 8u : b1 ? 4u : 0u;
                const uint32_t w = b3 ? x3 : b2 ? x2 : b1 ? x1 : x0;
                gf = (int)(JS * (uint32_t)j + 16u * (uint32_t)gl + k + ((uint32_t)(31 - __clz((int)w)) >> 3));
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {  // (a REDUX over part of a warp runs once per group)
                gp = min(gp, __shfl_xor_sync(FULL_MASK, gp, o));
                gf = max(gf, __shfl_xor_sync(FULL_MASK, gf, o));
            }
        }
        const uint32_t t = min(gp != 0x7fffffffu ? gp - a : L, min(L, Lp));  // common prefix of R and P
        // (a lane without a read -- L = 0 -- sees unmasked rows of the others' extent: its difference may lie behind a + L)
        const uint32_t s = (uint32_t)max(min(gf >= 0 ? (int)(a + L) - 1 - gf : (int)L, (int)min(L, Lf)), 0);  // common suffix of R and F
        const uint32_t rstage32 = stage32 + (uint32_t)grp * ws + a;  // the read's first character in the staging area
        const uint32_t bp = s_pcum[t], bs = s_fsuf[s];
        const bool t0_all = s == L && L == Lf, pe_all = t == L && L == Lp;
        const bool overlap = t + s >= L;
        const uint32_t n_mid = overlap ? 0u : L - s - t;
        // findJSS answers len - 1 when every count equals FE's, else bs (-1 is possible only with bs == 0: left to the general
        // path); a junction start on the first site of an alternative template's region is decided by computeAltEditing
        // (align.go:183-235); a junction longer than mid_cap may hold a saturating run
        const int junc_start = t0_all ? (int)tm : (int)bs;
        const bool cand = L != 0u && (t0_all || bs >= 1u) && !(K > 2 && s_alt[junc_start] != 0) && n_mid <= g.mid_cap;

        // ---- the characters between prefix and suffix against Bases[bp ..): the whole warp, one read after the other ----
        bool mid_ok = true;
        uint32_t mid_bases = 0;
        {
            // (junction length | bases in front of it << 16) and the staged address of its first character
            const uint32_t my_nb = (cand ? n_mid : 0u) | (bp << 16);
            const uint32_t my_at = rstage32 + t;
            // the first 32 characters of KB reads at a time: their loads are in flight together (the two dependent
            // shared-memory loads of a round are the latency of this kernel); a read without a junction runs through with
            // an empty ballot.  bit k of more: read k's junction goes on and has matched so far
            constexpr int KB = RPW < 4 ? RPW : 4;
            uint32_t more = 0u;
#pragma unroll
            for (int k0 = 0; k0 < RPW; k0 += KB) {
                uint32_t nb[KB], cu[KB], bal[KB], want[KB];
#pragma unroll
                for (int i = 0; i < KB; i++) {
                    nb[i] = __shfl_sync(FULL_MASK, my_nb, (k0 + i) * G);
                    const uint32_t at = __shfl_sync(FULL_MASK, my_at, (k0 + i) * G) + (uint32_t)lane;  // (inside the staging area + its pad)
                    cu[i] = lds_u8(at) & 0xdfu;
                }
#pragma unroll
                for (int i = 0; i < KB; i++) {
                    bal[i] = __ballot_sync(FULL_MASK, (uint32_t)lane < (nb[i] & 0xffffu) && cu[i] != eb);
                    want[i] = lds_u8(bases32 + (nb[i] >> 16) + (uint32_t)__popc(bal[i] & lt));
                }
#pragma unroll
                for (int i = 0; i < KB; i++) {
                    const bool bad = __any_sync(FULL_MASK, ((bal[i] >> lane) & 1u) != 0u && want[i] != cu[i]);
                    if (grp == k0 + i) {
                        mid_ok = !bad;
                        mid_bases = (uint32_t)__popc(bal[i]);
                    }
                    if (!bad && (nb[i] & 0xffffu) > 32u) more |= 1u << (k0 + i);
                }
            }
            while (more) {  // the rest of the long junctions, one read after the other
                const int k = __ffs((int)more) - 1;
                more &= more - 1u;
                const uint32_t nb = __shfl_sync(FULL_MASK, my_nb, k * G);
                int left = (int)(nb & 0xffffu) - 32;
                uint32_t at = __shfl_sync(FULL_MASK, my_at, k * G) + 32u + (uint32_t)lane;
                const uint32_t rank0 = bases32 + (nb >> 16);
                uint32_t rank = rank0 + __shfl_sync(FULL_MASK, mid_bases, k * G);
                bool okk = true;
                do {
                    const uint32_t cu = lds_u8(at) & 0xdfu;
                    const bool isb = lane < left && cu != eb;
                    const uint32_t bal = __ballot_sync(FULL_MASK, isb);
                    const uint32_t want = lds_u8(rank + (uint32_t)__popc(bal & lt));
                    rank += (uint32_t)__popc(bal);
                    if (__any_sync(FULL_MASK, isb && want != cu)) {
                        okk = false;
                        break;
                    }
                    at += 32u;
                    left -= 32;
                } while (left > 0);
                if (grp == k) {
                    mid_ok = okk;
                    mid_bases = rank - rank0;
                }
            }
        }
        __syncwarp();
        const bool ident = overlap ? bp + (uint32_t)s_fsuf[L - t] == tm : (mid_ok && bp + mid_bases + bs == tm);
        const bool ok = cand && ident;

        if (ok) {
            int junc_end = pe_all ? -1 : (int)(tm - bp);
            int edit_stop = junc_start - 1, junc_len = 0;
            uint32_t js_off = 0, js_len = 0;
            if (junc_end > edit_stop) {
                junc_len = junc_end - edit_stop;
                if (!t0_all) {
                    const uint32_t o0 = s_ppos[bp], o1 = L - (uint32_t)s_ftail[bs];
                    js_len = o1 - o0;
                    js_off = js_len ? o0 : 0u;
                }
            } else if (junc_end < edit_stop) {
                junc_end = edit_stop;
            }
            if (t0_all) {
                junc_end = junc_start;
                edit_stop = junc_start;
                junc_len = 0;
            }
            const int off = q.t.edit_offset;
            if (gl == 0) {
                uint4 *dst = reinterpret_cast<uint4 *>(q.rec + r);
                dst[0] = make_uint4((uint32_t)(edit_stop + off), (uint32_t)(junc_start + off), (uint32_t)(junc_end + off), (uint32_t)junc_len);
                dst[1] = make_uint4(read_count, 0u, tm, js_off);
                dst[2] = make_uint4(js_len, tm | (tm << 16), 0u, L);
            }
            // cmd/treat/handlers.go:203-215: three histograms, one lane each (lanes 0..2); the sums the header counters are made
            // of (cmd/treat/stats.go:140-187) -- the reads and their ReadCount -- are kept per warp in registers (below)
            if (sum_len) {
                const bool hist = gl < 3 && !(edit_stop + off == q.t.tmpl_edit_stop && junc_len == 0);
                const uint32_t bin = gl == 0 ? (uint32_t)(edit_stop + 2) : gl == 1 ? (uint32_t)junc_len : (uint32_t)(junc_end + 2);
                if (hist && bin < q.bins) {
                    const uint32_t idx = TREAT_B200_SUMMARY_HEADER + (uint32_t)gl * q.bins + bin;
                    const uint32_t old = atomicAdd(&sum_lo[idx], read_count);
                    if (old + read_count < old) atomicAdd(&sum_hi[idx], 1u);
                }
            }
            if (q.ops && !q.ops_gapped_only) {  // m match columns: an all-zero row
                uint32_t *go = reinterpret_cast<uint32_t *>(q.ops + r * (uint64_t)q.ops_stride);
                for (int wi = gl; wi < (int)(q.ops_stride / 4); wi += G) go[wi] = 0u;
            }
        }
        n_done += (uint32_t)__popc(__ballot_sync(FULL_MASK, ok && gl == 0));
        {   // (two 16-bit halves: a warp-wide add of 32-bit values must not wrap)
            const uint32_t v = (ok && gl == 0) ? read_count : 0u;
            rc_done += (unsigned long long)__reduce_add_sync(FULL_MASK, v & 0xffffu) + ((unsigned long long)__reduce_add_sync(FULL_MASK, v >> 16) << 16);
        }
        // everything else goes to the general path, one list append per warp
        {
            const uint32_t todo = __ballot_sync(FULL_MASK, in_batch && !ok && gl == 0);
            if (todo) {
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(f.rest_count, (uint32_t)__popc(todo));
                at = __shfl_sync(FULL_MASK, at, 0);
                if (in_batch && !ok && gl == 0) f.rest[at + (uint32_t)__popc(todo & lt)] = r;
            }
        }
        L_cur = L_nxt, a_cur = a_nxt;
        buf ^= half;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0 && n_done) {
        atomicMax(&p.scalars[0], tm);
        if (sum_len) {
            atomicAdd(&sum_lo[1], n_done);  // (a launch holds < 2^31 reads)
            const uint32_t lo32 = (uint32_t)rc_done, old = atomicAdd(&sum_lo[0], lo32);
            atomicAdd(&sum_hi[0], (uint32_t)(rc_done >> 32) + (old + lo32 < old ? 1u : 0u));
        }
    }
    // the header counters of the reads finished above (cmd/treat/stats.go:140-187), from the CTA's two sums: Total and Std
    // weighted by ReadCount, then per read, reads aligned, DP cells; none of these reads has a gap or a mismatch
    unsigned long long ctr = 0;
    __syncthreads();
    if (sum_len && threadIdx.x < 32) {
        const unsigned long long A = (unsigned long long)sum_lo[0] | ((unsigned long long)sum_hi[0] << 32);
        const unsigned long long Nn = (unsigned long long)sum_lo[1] | ((unsigned long long)sum_hi[1] << 32);
        ctr = (lane == 0 || lane == 1) ? A : (lane == 7 || lane == 8 || lane == 14) ? Nn
              : lane == 15 ? Nn * (unsigned long long)tm * (unsigned long long)tm : 0ull;
    }
    flush_summary(q, sum_len, sum_lo, sum_hi, lane, ctr);
}

bool frontier_applies(const PackParams &p, const AlignParams &q, const FrontierParams &f)
{
    return f.image != nullptr && f.rest != nullptr && f.rest_count != nullptr && f.g.image_bytes != 0 && f.g.tm == (uint32_t)q.t.m && p.tpk != nullptr &&
           p.edit_base >= 'A' && p.edit_base <= 'Z';
}

template <typename Kern>
static int launch_frontier_kernel(Kern kern, const PackParams &p, const AlignParams &q, const FrontierParams &f, int rpw, size_t fixed, int max_smem,
                                  int num_sms, void *stream)
{
    // as many warps per CTA as the planned number of CTAs per SM can stage next to their images (one CTA when the image is large)
    const size_t share = ((size_t)max_smem + 1024) / TREAT_B200_FR_MIN_CTAS;  // (the driver keeps 1 KB per resident CTA)
    const size_t per_warp = 2 * (size_t)rpw * f.g.ws;  // two staging halves
    int warps = (int)std::min<size_t>(kFrWarps, (share > fixed + 1024 ? (share - 1024 - fixed) / per_warp : 0));
    if (warps < 4) warps = (int)std::min<size_t>(kFrWarps, ((size_t)max_smem - fixed) / per_warp);
    if (warps < 1) return -1000;
    if (warps > 4) warps &= ~3;  // the same number of warps on each of the SM's four schedulers
    const size_t smem = fixed + (size_t)warps * per_warp;
    const uint64_t units = (p.N + rpw - 1) / rpw;
    uint64_t blocks = (units + warps - 1) / warps;
    const int per_sm = occupancy_cached((const void *)kern, warps * 32, smem);
    const uint64_t cap = (uint64_t)num_sms * per_sm;  // persistent grid
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p, q, f);
    return (int)cudaGetLastError();
}

int launch_frontier(const PackParams &p, const AlignParams &q, const FrontierParams &f, int num_sms, int max_smem_optin, void *stream)
{
    const uint32_t sum_len = q.summary ? TREAT_B200_SUMMARY_HEADER + 3u * q.bins : 0u;
    const size_t fixed = (size_t)f.g.image_bytes + align16(8u * sum_len) + 512u;  // + pad: the junction check reads whole rounds of 32 characters
    const int G = (int)f.g.G;
    if (G == 0 || f.g.ws == 0 || fixed + 2 * (size_t)(32 / G) * f.g.ws > (size_t)max_smem_optin) return -1000;
    const bool span = p.begin != nullptr;
    switch (G) {
    case 4:
        return span ? launch_frontier_kernel(k_frontier<4, true>, p, q, f, 8, fixed, max_smem_optin, num_sms, stream)
                    : launch_frontier_kernel(k_frontier<4, false>, p, q, f, 8, fixed, max_smem_optin, num_sms, stream);
    case 8:
        return span ? launch_frontier_kernel(k_frontier<8, true>, p, q, f, 4, fixed, max_smem_optin, num_sms, stream)
                    : launch_frontier_kernel(k_frontier<8, false>, p, q, f, 4, fixed, max_smem_optin, num_sms, stream);
    default:
        return span ? launch_frontier_kernel(k_frontier<16, true>, p, q, f, 2, fixed, max_smem_optin, num_sms, stream)
                    : launch_frontier_kernel(k_frontier<16, false>, p, q, f, 2, fixed, max_smem_optin, num_sms, stream);
    }
}

int configure_frontier_kernels(int max_smem_optin)
{
    cudaError_t ce = cudaFuncSetAttribute(k_frontier<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_frontier<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_frontier<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_frontier<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_frontier<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_frontier<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
