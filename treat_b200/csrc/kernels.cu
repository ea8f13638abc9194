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
template <int CPL>
struct Dir {
    using word_t = typename std::conditional<(CPL <= 8), uint16_t, uint32_t>::type;
    static constexpr int kLeftShift = (CPL <= 8) ? 8 : 16;  // pL bits sit above the pU bits
};

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

template <int CPL, bool WIDE>
__global__ void __launch_bounds__(512, 1) k_align(const __grid_constant__ AlignParams p)
{
    using dir_t = typename Dir<CPL>::word_t;
    using cnt_t = typename std::conditional<WIDE, uint32_t, uint8_t>::type;
    constexpr int LSH = Dir<CPL>::kLeftShift;

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

    for (int i = threadIdx.x; i < 32 * CPL; i += blockDim.x) s_tb[i] = i < m ? p.t.tb[i] : 0;
    for (int i = threadIdx.x; i < K * len_pad; i += blockDim.x)
        s_tcount[i] = reinterpret_cast<const cnt_t *>(p.t.tcount)[i];
    for (int i = threadIdx.x; i < 2 * n_alt; i += blockDim.x) s_alt[i] = p.t.alt[i];
    for (uint32_t i = threadIdx.x; i < sum_len; i += blockDim.x) s_sum[i] = 0ull;

    // ---- per-warp section ------------------------------------------------------------------
    uint8_t *wbase = smem + p.sm_tmpl_bytes + (size_t)warp * p.sm_warp_bytes;
    uint8_t *s_stage[2];
    s_stage[0] = wbase;
    s_stage[1] = wbase + p.sm_stage_b + p.sm_stage_c;
    uint8_t *s_rb = s_stage[1] + p.sm_stage_b + p.sm_stage_c;  // read bases as bytes, 32 bytes of pad in front
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
    __syncthreads();

    // ---- per-lane template registers ---------------------------------------------------------
    int tcr[CPL];
#pragma unroll
    for (int k = 0; k < CPL; k++) {
        const int idx = lane * CPL + k;
        tcr[k] = idx < m ? (int)s_tb[idx] : 0x100;  // 0x100 never equals a read byte
    }
    const int lanes_used = (m + CPL - 1) / CPL;
    const uint32_t lt = (1u << lane) - 1u;

    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, nw = (uint64_t)gridDim.x * nwarps;

    // TMA: stage the packed read r into buffer `slot`
    auto issue = [&](uint64_t r, int slot) {
        if (lane == 0) {
            const uint32_t n = p.n[r];
            const uint32_t cb = WIDE ? align16(n) : align16((n + 3) / 4);
            const uint32_t cc = WIDE ? align16(4 * (n + 1)) : align16(n + 1);
            fence_proxy_async();
            mbar_expect_tx(&s_bar[slot], cb + cc);
            if (cb) bulk_g2s(s_stage[slot], p.bases + r * (uint64_t)p.bases_stride, cb, &s_bar[slot]);
            bulk_g2s(s_stage[slot] + p.sm_stage_b, p.counts + r * (uint64_t)p.counts_stride, cc, &s_bar[slot]);
        }
    };

    uint32_t it = 0;
    if (gw < p.N) issue(gw, 0);

    for (uint64_t r = gw; r < p.N; r += nw, it++) {
        const int slot = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const int n = (int)p.n[r];
        const uint8_t rflags = p.rflags[r];
        if (n > (int)p.ncap) {  // cannot happen when ncap comes from K1's max; keep the launch safe
            if (r + nw < p.N) issue(r + nw, slot ^ 1);
            mbar_wait(&s_bar[slot], parity);
            continue;
        }
        // prefetch the next read of this warp while this one is aligned
        __syncwarp();
        if (r + nw < p.N) issue(r + nw, slot ^ 1);
        mbar_wait(&s_bar[slot], parity);

        const uint8_t *st_b = s_stage[slot];
        const cnt_t *rcount = reinterpret_cast<const cnt_t *>(s_stage[slot] + p.sm_stage_b);

        // ---- unpack read bases to bytes (index -32.. for idle lanes of the wavefront) -------
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
        } else {
            for (int i = lane; i < n; i += 32) rb[i] = st_b[i];
        }
        __syncwarp();

        // ---- fill: anti-diagonal wavefront ------------------------------------------------------
        // F[i][j]: i template bases, j read bases consumed; lane owns i = lane*CPL+1 .. +CPL.
        int H[CPL];  // F[i][j-1] for this lane's columns
#pragma unroll
        for (int k = 0; k < CPL; k++) H[k] = -(lane * CPL + k + 1);
        int diag = -(lane * CPL);  // F[i0-1][j-1]
        const int steps = n > 0 ? n + lanes_used - 1 : 0;
        for (int t = 0; t < steps; t++) {
            const int j = t - lane + 1;
            const int recv = __shfl_up_sync(FULL_MASK, H[CPL - 1], 1);
            const int left = lane == 0 ? -j : recv;  // F[i0-1][j]
            uint32_t w = 0;
            if (j >= 1 && j <= n) {
                const int rc = rb[j - 1];
                int dg = diag, lf = left;
#pragma unroll
                for (int k = 0; k < CPL; k++) {
                    const int x = H[k] - 1;                       // vgap: F[i][j-1] + gap   ("Left")
                    const int s = (rc == tcr[k]) ? 1 : -1;
                    const int e = max(dg + s, x);
                    const int f = max(e, lf - 1);                 // hgap: F[i-1][j] + gap   ("Up")
                    w |= (e < lf ? 1u : 0u) << k;                 // max == hgap
                    w |= (f == x ? 1u : 0u) << (LSH + k);         // max == vgap
                    dg = H[k];
                    lf = f;
                    H[k] = f;
                }
                diag = left;
            }
            s_dir[t * 32 + lane] = (dir_t)w;
        }
        __syncwarp();

        // final score F[m][n] lives in lane (m-1)/CPL, column (m-1)%CPL
        int fin = 0;
        {
            const int kl = (m - 1) % CPL;
#pragma unroll
            for (int k = 0; k < CPL; k++)
                if (k == kl) fin = H[k];
            fin = __shfl_sync(FULL_MASK, fin, (m - 1) / CPL);
            if (n == 0) fin = -m;
        }

        // ---- traceback (nwalgo: Up, then Left, then Diagonal) -----------------------------------
        int pos = m + n;
        {
            int i = m, j = n;
            int ln = (m - 1) / CPL, k = (m - 1) % CPL;
            while (i > 0 || j > 0) {
                int op;
                if (i == 0)
                    op = TREAT_B200_OP_INS;
                else if (j == 0)
                    op = TREAT_B200_OP_DEL;
                else {
                    const uint32_t w = s_dir[(j - 1 + ln) * 32 + ln];
                    op = ((w >> k) & 1u) ? TREAT_B200_OP_DEL : ((w >> (LSH + k)) & 1u) ? TREAT_B200_OP_INS : TREAT_B200_OP_MATCH;
                }
                --pos;
                if (lane == 0) s_ops[pos] = (uint8_t)op;
                if (op != TREAT_B200_OP_INS) {
                    i--;
                    if (--k < 0) {
                        k = CPL - 1;
                        ln--;
                    }
                }
                if (op != TREAT_B200_OP_DEL) j--;
            }
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

static const int kCplSet[] = {1, 2, 4, 6, 8, 12, 16};

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
    const uint32_t dsz = cpl <= 8 ? 2u : 4u;
    const uint32_t sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    p.sm_tmpl_bytes = align16(32 * cpl) + align16((uint32_t)(K * p.t.len_pad) * csz) +
                      align16(8u * (p.t.n_alt ? p.t.n_alt : 1)) + align16(8u * sum_len);
    p.sm_tmpl_bytes = (p.sm_tmpl_bytes + 127u) & ~127u;
    p.sm_stage_b = wide ? align16(ncap) : align16((ncap + 3) / 4);
    if (p.sm_stage_b == 0) p.sm_stage_b = 16;
    p.sm_stage_c = wide ? align16(4 * (ncap + 1)) : align16(ncap + 1);
    p.sm_rb = align16(ncap + 64);
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
    case 8: return launch_align_w<8>(p, wide, num_sms, max_smem_optin, stream);
    case 12: return launch_align_w<12>(p, wide, num_sms, max_smem_optin, stream);
    case 16: return launch_align_w<16>(p, wide, num_sms, max_smem_optin, stream);
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
    if ((e = set_attr<8>(max_smem_optin))) return e;
    if ((e = set_attr<12>(max_smem_optin))) return e;
    if ((e = set_attr<16>(max_smem_optin))) return e;
    cudaError_t ce = cudaFuncSetAttribute(k_pack<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
