// kernels.cu -- sm_100a kernels of the TREAT alignment hot path.
//
//   K0 k_int32_peak : dependent-free INT32 add/max loop, the roofline denominator.
//   K1 k_pack       : NewFragment for a batch (fragment.go:91-121): upper-case, strip the edit
//                     base, per-site edit-base run lengths; writes 2-bit bases + uint8 counts
//                     (or uint8 bases + uint32 counts on the wide path) in fixed-stride slots.
//   K2 k_align      : one warp per read-template pair.  nwalgo.Align (align.go:129) as an
//                     anti-diagonal wavefront: lane l owns template bases [l*CPL, (l+1)*CPL), the
//                     read streams through as rows, lane l works on row t-l at step t and hands
//                     its right-edge cell to lane l+1 with __shfl_up_sync.  Direction predicates
//                     (max==up+gap, max==left+gap) are kept as packed bits in shared memory, the
//                     traceback applies nwalgo's Up > Left > Diagonal order, and computeT /
//                     findJSS / findJES / computeAltEditing / NewAlignment (align.go:95-279) run
//                     on the warp so that only a 48-byte record (+ optional 2-bit ops) leaves.
//                     Packed reads are staged into shared memory with cp.async.bulk (TMA) behind
//                     an mbarrier, one read ahead of the one being aligned.
//
// Integer dynamic programming: no tensor cores on purpose.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "device_format.h"

namespace treat_b200 {

#define FULL_MASK 0xffffffffu

__host__ __device__ constexpr uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA, 1-D)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// K0: INT32 peak
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_int32_peak(int *out, int iters)
{
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int c = blockIdx.x | 1, d = -(int)blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {  // 8 x 8 x (add, max) = 128 lane-ops per iteration
            a0 = max(a0 + c, d);
            a1 = max(a1 + c, d);
            a2 = max(a2 + c, d);
            a3 = max(a3 + c, d);
            a4 = max(a4 + c, d);
            a5 = max(a5 + c, d);
            a6 = max(a6 + c, d);
            a7 = max(a7 + c, d);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

int launch_int32_peak(int num_sms, int iters, int *d_out, void *stream)
{
    k_int32_peak<<<num_sms * 8, 256, 0, (cudaStream_t)stream>>>(d_out, iters);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K1: pack (NewFragment, fragment.go:91-121), one warp per read
// ------------------------------------------------------------------------------------------
constexpr int kPackWarpsMax = 8;

template <bool WIDE>
__global__ void __launch_bounds__(kPackWarpsMax * 32) k_pack(const __grid_constant__ PackParams p)
{
    using cnt_t = typename std::conditional<WIDE, uint32_t, uint8_t>::type;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_lut = smem;  // 256
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cap = p.stage_cap;  // multiple of 16
    uint8_t *s_code = smem + 256 + (size_t)warp * (cap + cap * sizeof(cnt_t));
    cnt_t *s_cnt = reinterpret_cast<cnt_t *>(s_code + cap);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = p.lut[i];
    __syncthreads();

    const uint32_t lt = (1u << lane) - 1u;
    const uint8_t eb = p.edit_base;
    uint32_t warp_max_n = 0, warp_sat = 0;
    const uint64_t pw = blockDim.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * pw + warp, nw = (uint64_t)gridDim.x * pw;
    for (uint64_t r = gw; r < p.N; r += nw) {
        const uint64_t b = p.off[r] - p.off_bias;
        const uint32_t L = (uint32_t)(p.off[r + 1] - p.off[r]);
        const uint8_t *seq = p.seqs + b;
        uint32_t nb = 0;
        int lastpos = -1;
        bool sat = false, widec = false;
        for (uint32_t c0 = 0; c0 < L; c0 += 32) {
            const uint32_t j = c0 + lane;
            uint8_t ch = eb;
            if (j < L) {
                ch = seq[j];
                if (ch >= 'a' && ch <= 'z') ch -= 32;  // strings.ToUpper, ASCII
            }
            const bool isb = ch != eb;
            const uint32_t mask = __ballot_sync(FULL_MASK, isb);
            if (isb) {
                const uint32_t below = mask & lt;
                const uint32_t idx = nb + __popc(below);
                const int prev = below ? (int)(c0 + 31 - __clz(below)) : lastpos;
                uint32_t cnt = (uint32_t)((int)j - prev - 1);
                uint8_t code = ch;
                if (cnt >= 255u) sat = true;  // informational on the wide path, which keeps exact counts
                if (!WIDE) {
                    code = s_lut[ch];
                    widec |= (code == 3);
                    if (cnt >= 255u) cnt = 255u;
                }
                s_code[idx] = code;
                s_cnt[idx] = (cnt_t)cnt;
            }
            if (mask) lastpos = (int)(c0 + 31 - __clz(mask));
            nb += __popc(mask);
        }
        if (lane == 0) {
            uint32_t cnt = (uint32_t)((int)L - 1 - lastpos);
            if (cnt >= 255u) sat = true;
            if (!WIDE && cnt >= 255u) cnt = 255u;
            s_cnt[nb] = (cnt_t)cnt;
        }
        __syncwarp();
        const uint32_t n = nb;
        const bool any_sat = __any_sync(FULL_MASK, sat), any_wide = __any_sync(FULL_MASK, widec);
        // bases
        uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
        if (!WIDE) {
            const uint32_t nwords = align16((n + 3) / 4) / 4;
            for (uint32_t wi = lane; wi < nwords; wi += 32) {
                uint32_t w = 0;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const uint32_t bi = wi * 16 + q;
                    const uint32_t c = bi < n ? s_code[bi] : 0u;
                    w |= c << (2 * q);
                }
                gb[wi] = w;
            }
        } else {
            const uint32_t nwords = align16(n) / 4;
            for (uint32_t wi = lane; wi < nwords; wi += 32) {
                uint32_t w = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t bi = wi * 4 + q;
                    const uint32_t c = bi < n ? s_code[bi] : 0u;
                    w |= c << (8 * q);
                }
                gb[wi] = w;
            }
        }
        // counts (n + 1 entries)
        uint32_t *gc = reinterpret_cast<uint32_t *>(p.counts + r * (uint64_t)p.counts_stride);
        if (!WIDE) {
            const uint32_t nwords = align16(n + 1) / 4;
            for (uint32_t wi = lane; wi < nwords; wi += 32) {
                uint32_t w = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t ci = wi * 4 + q;
                    const uint32_t c = ci <= n ? (uint32_t)s_cnt[ci] : 0u;
                    w |= c << (8 * q);
                }
                gc[wi] = w;
            }
        } else {
            const uint32_t nwords = align16(4 * (n + 1)) / 4;
            for (uint32_t wi = lane; wi < nwords; wi += 32) gc[wi] = wi <= n ? (uint32_t)s_cnt[wi] : 0u;
        }
        if (lane == 0) {
            p.n[r] = (uint16_t)n;
            p.flags[r] = (uint8_t)((any_sat ? TREAT_B200_ST_SATURATED : 0) | (any_wide ? TREAT_B200_ST_WIDE_BASE : 0));
            p.raw_len[r] = L;
        }
        warp_max_n = max(warp_max_n, n);
        warp_sat |= any_sat ? 1u : 0u;
        __syncwarp();
    }
    if (lane == 0) {
        if (warp_max_n) atomicMax(&p.scalars[0], warp_max_n);
        if (warp_sat) atomicOr(&p.scalars[1], 1u);
    }
}

int launch_pack(const PackParams &p, bool wide, int num_sms, int max_smem_optin, void *stream)
{
    const size_t per_warp = (size_t)p.stage_cap * (wide ? 5 : 2);
    int warps = (int)(((size_t)max_smem_optin - 256) / per_warp);
    if (warps < 1) return -1000;
    if (warps > kPackWarpsMax) warps = kPackWarpsMax;
    const size_t smem = 256 + per_warp * warps;
    uint64_t blocks = (p.N + warps - 1) / warps;
    const uint64_t cap = (uint64_t)num_sms * (smem > 100 * 1024 ? 1 : 2048 / (warps * 32));
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    if (wide)
        k_pack<true><<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    else
        k_pack<false><<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K2: align
// ------------------------------------------------------------------------------------------
// Score transform.  nwalgo fills F[i][j] = max(F[i-1][j-1] + (a==b ? 1 : -1), F[i-1][j] - 1,
// F[i][j-1] - 1) with F[i][0] = -i, F[0][j] = -j.  With G = F + i + j the gap terms vanish:
//     G[i][j] = max(G[i-1][j-1] + w, G[i-1][j], G[i][j-1]),  w = 3 on a match, 1 otherwise,
//     G[i][0] = G[0][j] = 0,  F[m][n] = G[m][n] - m - n,
// every horizontal / vertical difference of G lies in {0,1,2,3}, and nwalgo's pointer tests become
//     max == hgap ("Up",   consumes the template base)  <=>  G[i][j] == G[i-1][j]
//     max == vgap ("Left", consumes the read base)      <=>  G[i][j] == G[i][j-1].
// A row with w = 0 leaves the previous row unchanged, so idle wavefront steps need no branch.
//
// Per lane and step the kernel stores ONE word: the CPL horizontal differences as base-4 digits
// plus (left boundary value mod 4) on top.  From two such words (rows j and j-1) the traceback
// recovers G mod 4 of any cell and therefore both pointer tests.
template <int CPL>
struct Dir {
    using word_t = typename std::conditional<(CPL <= 7), uint16_t, uint32_t>::type;
    static constexpr int kTop = (CPL <= 7) ? 14 : 30;            // bit position of (left & 3)
    static constexpr uint32_t kOdd = (CPL <= 7) ? 0x1555u : 0x15555555u;   // low bits of the digits
    static_assert(CPL <= 15, "direction word holds at most 15 base-4 digits");
};

// sum of the base-4 digits of x (mod 4 is all the caller needs)
__device__ __forceinline__ uint32_t digit_sum(uint32_t x, uint32_t odd)
{
    return __popc(x & odd) + 2u * __popc(x & (odd << 1));
}

// Bit rows over template sites in array order ti = 0..len-1 (site number = len-1-ti), one
// 32-bit word per lane, kept in shared memory: the willf/bitset sets of align.go:122-181.
struct BitRows {
    const uint32_t *T;  // [K][tw]
    int tw, len, lane;
    // highest ti <= limit whose bit is clear in row k, or -1
    __device__ __forceinline__ int highest_clear_le(int k, int limit) const
    {
        uint32_t w = 0;
        const int lo = lane * 32;
        if (lane < tw && lo <= limit) {
            w = ~T[k * tw + lane];
            if (limit - lo < 31) w &= (2u << (limit - lo)) - 1u;
        }
        const uint32_t any = __ballot_sync(FULL_MASK, w != 0);
        if (!any) return -1;
        const int hw = 31 - __clz(any);
        const uint32_t wv = __shfl_sync(FULL_MASK, w, hw);
        return hw * 32 + 31 - __clz(wv);
    }
    // lowest ti whose bit is clear in row k, or -1
    __device__ __forceinline__ int lowest_clear(int k) const
    {
        uint32_t w = 0;
        const int lo = lane * 32;
        if (lane < tw) {
            w = ~T[k * tw + lane];
            if (len - 1 - lo < 31) w &= (2u << (len - 1 - lo)) - 1u;
        }
        const uint32_t any = __ballot_sync(FULL_MASK, w != 0);
        if (!any) return -1;
        const int lw = __ffs(any) - 1;
        const uint32_t wv = __shfl_sync(FULL_MASK, w, lw);
        return lw * 32 + __ffs(wv) - 1;
    }
    __device__ __forceinline__ bool any_set(int k) const
    {
        uint32_t w = 0;
        const int lo = lane * 32;
        if (lane < tw) {
            w = T[k * tw + lane];
            if (len - 1 - lo < 31) w &= (2u << (len - 1 - lo)) - 1u;
        }
        return __any_sync(FULL_MASK, w != 0);
    }
    // bitset.Test on a SITE number (out of range -> false, as willf/bitset does)
    __device__ __forceinline__ bool test_site(int k, int site) const
    {
        if (site < 0 || site >= len) return false;
        const int ti = len - 1 - site;
        return (T[k * tw + (ti >> 5)] >> (ti & 31)) & 1u;
    }
};

constexpr int kPadCode = 4;  // profile row with w = 0: a no-op row of the wavefront

template <int CPL, bool WIDE>
__global__ void __launch_bounds__(512, 1) k_align(const __grid_constant__ AlignParams p)
{
    using DT = Dir<CPL>;
    using dir_t = typename DT::word_t;
    using cnt_t = typename std::conditional<WIDE, uint32_t, uint8_t>::type;
    constexpr int NCH = (CPL + 3) / 4;  // uint4 chunks of the scoring profile per lane

    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int m = p.t.m, len = p.t.len, len_pad = p.t.len_pad, K = p.t.K, n_alt = p.t.n_alt;
    const int tw = len_pad >> 5;

    // ---- CTA-wide template section ---------------------------------------------------------
    uint8_t *s_tb = smem;                                                       // [32*CPL]
    cnt_t *s_tcount = reinterpret_cast<cnt_t *>(smem + align16(32 * CPL));      // [K][len_pad]
    int32_t *s_alt = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(s_tcount) +
                                                 align16((uint32_t)(K * len_pad * sizeof(cnt_t))));
    unsigned long long *s_sum =
        reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(s_alt) + align16(8u * (n_alt ? n_alt : 1)));
    const uint32_t sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    // scoring profile: s_prof[(q*5 + code)*32 + lane] = w of columns lane*CPL + 4q .. +3
    uint4 *s_prof = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(s_sum) + align16(8u * sum_len));

    for (int i = threadIdx.x; i < 32 * CPL; i += blockDim.x) s_tb[i] = i < m ? p.t.tb[i] : 0;
    for (int i = threadIdx.x; i < K * len_pad; i += blockDim.x)
        s_tcount[i] = reinterpret_cast<const cnt_t *>(p.t.tcount)[i];
    for (int i = threadIdx.x; i < 2 * n_alt; i += blockDim.x) s_alt[i] = p.t.alt[i];
    for (uint32_t i = threadIdx.x; i < sum_len; i += blockDim.x) s_sum[i] = 0ull;
    if (!WIDE) {
        uint32_t *pw = reinterpret_cast<uint32_t *>(s_prof);
        for (int i = threadIdx.x; i < NCH * 5 * 32 * 4; i += blockDim.x) {
            const int e = i & 3, ln = (i >> 2) & 31, c = (i >> 7) % 5, q = (i >> 7) / 5;
            const int col = ln * CPL + q * 4 + e;
            uint32_t w = c == kPadCode ? 0u : 1u;
            if (c < 3 && q * 4 + e < CPL && col < m && p.t.tb[col] == c) w = 3u;
            pw[i] = w;
        }
    }

    // ---- per-warp section ------------------------------------------------------------------
    uint8_t *wbase = smem + p.sm_tmpl_bytes + (size_t)warp * p.sm_warp_bytes;
    const uint32_t stage_bytes = p.sm_stage_b + p.sm_stage_c;
    uint8_t *s_stage = wbase;                                  // two stage buffers back to back
    uint8_t *s_rb = wbase + 2 * stage_bytes;                   // read bases as bytes, 32 pad bytes in front
    dir_t *s_dir = reinterpret_cast<dir_t *>(s_rb + p.sm_rb);
    uint8_t *s_ops = reinterpret_cast<uint8_t *>(s_dir) + p.sm_dir;
    uint16_t *s_t2f = reinterpret_cast<uint16_t *>(s_ops + p.sm_ops);
    uint32_t *s_tbits = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_t2f) + p.sm_t2f);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(s_tbits) + p.sm_tbits);  // [2]

    if (lane == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    s_rb[lane] = kPadCode;  // rows "before" the read: no-op rows for lanes that have not started
    __syncthreads();

    int tcr[CPL];  // template bases of this lane's columns (wide path compares instead of a profile)
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const int idx = lane * CPL + k;
        tcr[k] = idx < m ? (int)s_tb[idx] : 0x100;  // 0x100 never equals a read byte
    }
    const int lanes_used = (m + CPL - 1) / CPL;
    const uint32_t lt = (1u << lane) - 1u;
    const uint4 *prof_lane = s_prof + lane;

    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, nw = (uint64_t)gridDim.x * nwarps;

    // TMA: stage the packed read r into buffer `slot`
    auto issue = [&](uint64_t r, int slot) {
        if (lane == 0) {
            const uint32_t n = p.n[r];
            const uint32_t cb = WIDE ? align16(n) : align16((n + 3) / 4);
            const uint32_t cc = WIDE ? align16(4 * (n + 1)) : align16(n + 1);
            uint8_t *dst = s_stage + slot * stage_bytes;
            fence_proxy_async();
            mbar_expect_tx(&s_bar[slot], cb + cc);
            if (cb) bulk_g2s(dst, p.bases + r * (uint64_t)p.bases_stride, cb, &s_bar[slot]);
            bulk_g2s(dst + p.sm_stage_b, p.counts + r * (uint64_t)p.counts_stride, cc, &s_bar[slot]);
        }
    };

    uint32_t it = 0;
    if (gw < p.N) issue(gw, 0);

    for (uint64_t r = gw; r < p.N; r += nw, it++) {
        const int slot = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const int n = (int)p.n[r];
        const uint8_t rflags = p.rflags[r];
        // prefetch the next read of this warp while this one is aligned
        __syncwarp();
        if (r + nw < p.N) issue(r + nw, slot ^ 1);
        mbar_wait(&s_bar[slot], parity);
        if (n > (int)p.ncap) continue;  // cannot happen when ncap comes from K1's max; keeps the launch safe

        const uint8_t *st_b = s_stage + slot * stage_bytes;
        const cnt_t *rcount = reinterpret_cast<const cnt_t *>(st_b + p.sm_stage_b);

        // ---- unpack read bases to one byte each; pad rows behind the read are no-op rows -----
        uint8_t *rb = s_rb + 32;
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
        } else {
            for (int i = lane; i < n; i += 32) rb[i] = st_b[i];
        }
        __syncwarp();

        // ---- fill: anti-diagonal wavefront ------------------------------------------------------
        // lane owns template bases i = lane*CPL+1 .. +CPL; at step t it works on read row j = t-lane+1.
        int H[CPL];  // G[i][j-1] for this lane's columns
#pragma unroll
        for (int k = 0; k < CPL; k++) H[k] = 0;
        int diag = 0;  // G[i0-1][j-1]
        const int steps = n > 0 ? n + lanes_used - 1 : 0;
        const uint8_t *rbl = rb - lane;  // rbl[t] = read base of this lane's row at step t
        uint4 pwc[NCH];
        int cnext = 0;
        if (!WIDE) {
#pragma unroll
            for (int q = 0; q < NCH; q++) pwc[q] = prof_lane[(q * 5 + rbl[0]) * 32];
            cnext = rbl[1];
        }
        dir_t *dirp = s_dir + lane;
#pragma unroll 2
        for (int t = 0; t < steps; t++) {
            uint4 pwn[NCH];
            int wv[CPL];
            if (!WIDE) {
#pragma unroll
                for (int q = 0; q < NCH; q++) pwn[q] = prof_lane[(q * 5 + cnext) * 32];
                cnext = rbl[t + 2];
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const uint4 v = pwc[k >> 2];
                    wv[k] = (int)((k & 3) == 0 ? v.x : (k & 3) == 1 ? v.y : (k & 3) == 2 ? v.z : v.w);
                }
            } else {
                const int idx = t - lane;
                const bool live = idx >= 0 && idx < n;
                const int rc = live ? (int)rbl[t] : 0x200;
#pragma unroll
                for (int k = 0; k < CPL; k++) wv[k] = live ? (rc == tcr[k] ? 3 : 1) : 0;
            }
            const int recv = __shfl_up_sync(FULL_MASK, H[CPL - 1], 1);
            const int left = lane == 0 ? 0 : recv;  // G[i0-1][j]
            int dg = diag, lf = left;
            uint32_t S = 0;
#pragma unroll
            for (int k = 0; k < CPL; k++) {
                const int e = max(dg + wv[k], H[k]);
                const int f = max(e, lf);
                dg = H[k];
                lf = f;
                H[k] = f;
                S += (uint32_t)f << (2 * k);
            }
            diag = left;
            // digits: sum_k (f_k - f_{k-1}) 4^k = 4^CPL f_last - 3 S - left
            const uint32_t word = ((uint32_t)lf << (2 * CPL)) - 3u * S - (uint32_t)left + (((uint32_t)left & 3u) << DT::kTop);
            dirp[t * 32] = (dir_t)word;
            if (!WIDE) {
#pragma unroll
                for (int q = 0; q < NCH; q++) pwc[q] = pwn[q];
            }
        }
        __syncwarp();

        // final score F[m][n] = G[m][n] - m - n; G[m][n] lives in lane (m-1)/CPL, column (m-1)%CPL
        int fin = 0;
        {
            const int kl = (m - 1) % CPL;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kl) fin = H[k];
            fin = __shfl_sync(FULL_MASK, fin, (m - 1) / CPL) - m - n;
        }

        // ---- traceback (nwalgo: Up, then Left, then Diagonal), 32 cells of a diagonal run at a time
        int pos = m + n;
        {
            int i = m, j = n;
            while (i > 0 && j > 0) {
                const int ii = i - lane, jj = j - lane;
                bool stop = true, is_up = false;
                if (ii >= 1 && jj >= 1) {
                    const int ln = (ii - 1) / CPL, k = (ii - 1) - ln * CPL;
                    const int tt = jj - 1 + ln;
                    const uint32_t w1 = s_dir[tt * 32 + ln];
                    const uint32_t a = (w1 >> (2 * k)) & 3u;
                    if (a == 0) {
                        is_up = true;  // G[i][j] == G[i-1][j]
                    } else {
                        const uint32_t msk = (4u << (2 * k)) - 1u;
                        const uint32_t g1 = (w1 >> DT::kTop) + digit_sum(w1 & msk, DT::kOdd);
                        uint32_t g0 = 0;  // G[i][0] = 0
                        if (jj >= 2) {
                            const uint32_t w0 = s_dir[(tt - 1) * 32 + ln];
                            g0 = (w0 >> DT::kTop) + digit_sum(w0 & msk, DT::kOdd);
                        }
                        stop = ((g1 - g0) & 3u) == 0;  // G[i][j] == G[i][j-1]: Left
                    }
                }
                const uint32_t sm = __ballot_sync(FULL_MASK, stop);
                const int run = sm ? __ffs(sm) - 1 : 32;  // diagonal moves before the first gap / edge
                if (lane < run) s_ops[pos - 1 - lane] = TREAT_B200_OP_MATCH;
                pos -= run;
                i -= run;
                j -= run;
                if (run < 32 && i > 0 && j > 0) {
                    const bool up = __shfl_sync(FULL_MASK, is_up, run);
                    --pos;
                    if (lane == 0) s_ops[pos] = up ? TREAT_B200_OP_DEL : TREAT_B200_OP_INS;
                    if (up) i--; else j--;
                }
            }
            // first row / column: pointers are Left on row 0 and Up on column 0
            for (int q = lane; q < i; q += 32) s_ops[pos - 1 - q] = TREAT_B200_OP_DEL;
            pos -= i;
            for (int q = lane; q < j; q += 32) s_ops[pos - 1 - q] = TREAT_B200_OP_INS;
            pos -= j;
        }
        const int alen = m + n - pos;
        const uint8_t *ops = s_ops + pos;
        __syncwarp();

        // ---- computeT part 1 (align.go:131-171): walk the alignment columns 32 at a time --------
        int ti_base = 0, fi_base = 0, mism = 0;
        bool indel = false;
        for (int c0 = 0; c0 < alen; c0 += 32) {
            const int c = c0 + lane;
            const int op = c < alen ? (int)ops[c] : 3;
            const bool cT = op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_DEL;
            const bool cR = op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_INS;
            const uint32_t mT = __ballot_sync(FULL_MASK, cT), mR = __ballot_sync(FULL_MASK, cR);
            const int ti = ti_base + __popc(mT & lt), fi = fi_base + __popc(mR & lt);
            if (cT) s_t2f[ti] = op == TREAT_B200_OP_MATCH ? (uint16_t)fi : (uint16_t)0xffff;
            const bool mm = op == TREAT_B200_OP_MATCH && rb[fi] != s_tb[ti];
            mism += __popc(__ballot_sync(FULL_MASK, mm));
            indel |= (mT != mR);
            ti_base += __popc(mT);
            fi_base += __popc(mR);
        }
        if (lane == 0) s_t2f[m] = (uint16_t)n;  // last edit site (align.go:173-178)
        __syncwarp();

        // ---- computeT part 2: per template site, compare edit-base counts, build the bit rows ---
        for (int wv = 0; wv < tw; wv++) {
            const int ti = wv * 32 + lane;
            const bool valid = ti < len;
            uint32_t count = 0;
            if (valid) {
                const uint32_t f = s_t2f[ti];
                count = f == 0xffffu ? 0u : (uint32_t)rcount[f];
            }
            for (int k = 0; k < K; k++) {
                const bool match = valid && (uint32_t)s_tcount[k * len_pad + ti] == count;
                const uint32_t word = __ballot_sync(FULL_MASK, match);
                if (lane == 0) s_tbits[k * tw + wv] = word;
            }
        }
        __syncwarp();

        // ---- findJSS / computeAltEditing / findJES / NewAlignment (align.go:95-120,183-279) -----
        BitRows B{s_tbits, tw, len, lane};
        int has_mutation = 0;
        if (indel) has_mutation = 1;
        if (((p.flags & TREAT_B200_EXCLUDE_SNPS) && mism >= 1) || mism >= 3) has_mutation = 1;

        const int hc0 = B.highest_clear_le(0, len - 1);
        const bool t0_all = hc0 < 0;
        int junc_start = t0_all ? len - 1 : (!B.any_set(0) ? -1 : len - 1 - hc0);
        int alt_editing = 0;
        {
            int shift = junc_start, alt = -1;
            for (int k = 2; k < K; k++) {
                if (B.test_site(k, shift)) {
                    alt = k - 2;
                    break;
                }
            }
            if (alt >= 0 && shift == s_alt[alt]) {
                const int hc = B.highest_clear_le(alt + 2, len - 1 - shift);
                const bool full_match = hc < 0;
                if (!full_match) shift = len - 1 - hc;
                if (full_match || shift >= s_alt[n_alt + alt]) {
                    alt_editing = alt + 1;
                    if (full_match) {
                        junc_start = len - 1;
                    } else {
                        const int h0 = B.highest_clear_le(0, len - 1 - shift);
                        if (h0 >= 0) junc_start = len - 1 - h0;
                    }
                }
            }
        }
        const int lc1 = B.lowest_clear(1);
        int junc_end = lc1 < 0 ? -1 : len - 1 - lc1;
        int edit_stop = junc_start - 1;
        int junc_len = 0;
        uint32_t js_off = 0, js_len = 0;
        uint32_t status = rflags;
        if (junc_end > edit_stop) {
            junc_len = junc_end - edit_stop;
            if (has_mutation == 0 && !t0_all) {
                const int from = (len - 1) - junc_end, to = (len - 1) - edit_stop;
                if (to > n + 1) {
                    status |= TREAT_B200_ST_REF_PANIC;  // Go: index out of range on frag.EditSite
                } else {
                    // JuncSeq = EditBase x EditSite[i] + Bases[i] for i in [from, to): a window of
                    // the upper-cased raw read starting at the edit-base run in front of base `from`
                    uint32_t sf = 0, st = 0;
                    for (int q = lane; q < to; q += 32) {
                        const uint32_t c = (uint32_t)rcount[q];
                        st += c;
                        if (q < from) sf += c;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        st += __shfl_xor_sync(FULL_MASK, st, o);
                        sf += __shfl_xor_sync(FULL_MASK, sf, o);
                    }
                    js_off = sf + (uint32_t)min(from, n);
                    js_len = st + (uint32_t)min(to, n) - js_off;
                    if (js_len == 0) js_off = 0;
                }
            }
        } else if (junc_end < edit_stop) {
            junc_end = edit_stop;
            junc_len = 0;
        }
        if (t0_all) {
            junc_end = junc_start;
            edit_stop = junc_start;
            junc_len = 0;
        }
        const uint32_t read_count = p.read_count ? p.read_count[r] : 1u;

        if (lane == 0) {
            treat_b200_record o;
            o.edit_stop = edit_stop + p.t.edit_offset;
            o.junc_start = junc_start + p.t.edit_offset;
            o.junc_end = junc_end + p.t.edit_offset;
            o.junc_len = junc_len;
            o.read_count = read_count;
            o.has_mutation = (uint8_t)has_mutation;
            o.alt_editing = (uint8_t)alt_editing;
            o.mismatches = (uint8_t)(mism & 0xff);  // uint8 wraps (align.go:51,148)
            o.indel = indel ? 1 : 0;
            o.score = fin;
            o.junc_seq_off = js_off;
            o.junc_seq_len = js_len;
            o.aln_len = (uint16_t)alen;
            o.n_bases = (uint16_t)n;
            o.status = (uint8_t)status;
            o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
            o.raw_len = p.raw_len[r];
            const uint4 *src = reinterpret_cast<const uint4 *>(&o);
            uint4 *dst = reinterpret_cast<uint4 *>(p.rec + r);
            dst[0] = src[0];
            dst[1] = src[1];
            dst[2] = src[2];

            if (sum_len) {
                // cmd/treat/stats.go:140-187
                const unsigned long long rcw = read_count;
                atomicAdd(&s_sum[0], rcw);
                atomicAdd(&s_sum[7], 1ull);
                const int cls = has_mutation ? 2 : 1;
                atomicAdd(&s_sum[cls], rcw);
                atomicAdd(&s_sum[7 + cls], 1ull);
                const int mm8 = mism & 0xff;
                const int mcls = indel ? 3 : mm8 == 1 ? 5 : mm8 == 2 ? 6 : mm8 > 2 ? 4 : 0;
                if (mcls) {
                    atomicAdd(&s_sum[mcls], rcw);
                    atomicAdd(&s_sum[7 + mcls], 1ull);
                }
                atomicAdd(&s_sum[14], 1ull);
                atomicAdd(&s_sum[15], (unsigned long long)m * (unsigned long long)n);
                // cmd/treat/handlers.go:203-215
                if (!(edit_stop + p.t.edit_offset == p.t.tmpl_edit_stop && junc_len == 0)) {
                    const uint32_t B3 = p.bins;
                    const uint32_t b_es = (uint32_t)(edit_stop + 2), b_jl = (uint32_t)junc_len, b_je = (uint32_t)(junc_end + 2);
                    if (b_es < B3) atomicAdd(&s_sum[TREAT_B200_SUMMARY_HEADER + b_es], rcw);
                    if (b_jl < B3) atomicAdd(&s_sum[TREAT_B200_SUMMARY_HEADER + B3 + b_jl], rcw);
                    if (b_je < B3) atomicAdd(&s_sum[TREAT_B200_SUMMARY_HEADER + 2 * B3 + b_je], rcw);
                }
            }
        }

        // ---- traceback ops, 2 bits per column, in alignment order --------------------------------
        if (p.ops) {
            uint32_t *go = reinterpret_cast<uint32_t *>(p.ops + r * (uint64_t)p.ops_stride);
            const int nwords = (int)(p.ops_stride / 4);
            for (int wi = lane; wi < nwords; wi += 32) {
                uint32_t w = 0;
                const int c0 = wi * 16;
                if (c0 < alen) {
#pragma unroll
                    for (int q = 0; q < 16; q++)
                        if (c0 + q < alen) w |= (uint32_t)ops[c0 + q] << (2 * q);
                }
                go[wi] = w;
            }
        }
        __syncwarp();
    }

    // ---- flush the CTA's summary counters -----------------------------------------------------
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < sum_len; i += blockDim.x) {
        const unsigned long long v = s_sum[i];
        if (v) atomicAdd(reinterpret_cast<unsigned long long *>(p.summary) + i, v);
    }
}

static const int kCplSet[] = {1, 2, 4, 6, 7, 10, 12, 15};

static int pick_cpl(int m)
{
    const int need = (m + 31) / 32;
    for (int c : kCplSet)
        if (c >= need) return c;
    return -1;
}

int plan_align(AlignParams &p, bool wide)
{
    const int m = p.t.m, K = p.t.K;
    const int cpl = pick_cpl(m);
    if (cpl < 0 || m < 1) return -1;
    const uint32_t ncap = p.ncap;
    const uint32_t csz = wide ? 4u : 1u;
    const uint32_t dsz = cpl <= 7 ? 2u : 4u;
    const uint32_t prof_bytes = wide ? 0u : (uint32_t)((cpl + 3) / 4) * 5u * 32u * 16u;
    const uint32_t sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    p.sm_tmpl_bytes = align16(32 * cpl) + align16((uint32_t)(K * p.t.len_pad) * csz) +
                      align16(8u * (p.t.n_alt ? p.t.n_alt : 1)) + align16(8u * sum_len) + prof_bytes;
    p.sm_tmpl_bytes = (p.sm_tmpl_bytes + 127u) & ~127u;
    p.sm_stage_b = wide ? align16(ncap) : align16((ncap + 3) / 4);
    if (p.sm_stage_b == 0) p.sm_stage_b = 16;
    p.sm_stage_c = wide ? align16(4 * (ncap + 1)) : align16(ncap + 1);
    p.sm_rb = align16(32 + ncap + 64 + 16);
    p.sm_dir = align16((ncap + 32) * 32 * dsz);
    p.sm_ops = align16((uint32_t)m + ncap + 4);
    p.sm_t2f = align16(2u * (uint32_t)p.t.len_pad);
    p.sm_tbits = align16(4u * (uint32_t)K * (uint32_t)(p.t.len_pad / 32));
    p.sm_warp_bytes = 2 * (p.sm_stage_b + p.sm_stage_c) + p.sm_rb + p.sm_dir + p.sm_ops + p.sm_t2f + p.sm_tbits + 16;
    p.sm_warp_bytes = (p.sm_warp_bytes + 127u) & ~127u;
    return cpl;
}

template <int CPL, bool WIDE>
static int launch_align_t(const AlignParams &p, int num_sms, int max_smem, void *stream)
{
    int warps = (int)((max_smem - (long)p.sm_tmpl_bytes) / (long)p.sm_warp_bytes);
    if (warps < 1) return -1000;  // geometry does not fit shared memory
    if (warps > 16) warps = 16;
    const uint64_t need_blocks = (p.N + warps - 1) / warps;
    uint64_t blocks = (uint64_t)num_sms;  // persistent: one CTA per SM
    if (need_blocks < blocks) blocks = need_blocks;
    if (blocks == 0) return 0;
    const size_t smem = p.sm_tmpl_bytes + (size_t)warps * p.sm_warp_bytes;
    k_align<CPL, WIDE><<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

template <int CPL>
static int launch_align_w(const AlignParams &p, bool wide, int num_sms, int max_smem, void *stream)
{
    return wide ? launch_align_t<CPL, true>(p, num_sms, max_smem, stream)
                : launch_align_t<CPL, false>(p, num_sms, max_smem, stream);
}

int launch_align(const AlignParams &p, bool wide, int num_sms, int max_smem_optin, void *stream)
{
    switch (pick_cpl(p.t.m)) {
    case 1: return launch_align_w<1>(p, wide, num_sms, max_smem_optin, stream);
    case 2: return launch_align_w<2>(p, wide, num_sms, max_smem_optin, stream);
    case 4: return launch_align_w<4>(p, wide, num_sms, max_smem_optin, stream);
    case 6: return launch_align_w<6>(p, wide, num_sms, max_smem_optin, stream);
    case 7: return launch_align_w<7>(p, wide, num_sms, max_smem_optin, stream);
    case 10: return launch_align_w<10>(p, wide, num_sms, max_smem_optin, stream);
    case 12: return launch_align_w<12>(p, wide, num_sms, max_smem_optin, stream);
    case 15: return launch_align_w<15>(p, wide, num_sms, max_smem_optin, stream);
    default: return -1001;
    }
}

template <int CPL>
static int set_attr(int max_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_align<CPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_align<CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    return (int)e;
}

int configure_kernels(int max_smem_optin)
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
    cudaError_t ce = cudaFuncSetAttribute(k_pack<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
