// pack_kernel.cu -- sm_100a kernels of the TREAT alignment hot path, part 1 (K0, K1).
//
//   K0 k_int32_peak : dependent-free INT32 add/max loop, the roofline denominator.
//   K1 k_pack       : NewFragment for a batch (fragment.go:91-121): upper-case, strip the edit
//                     base, per-site edit-base run lengths; writes 2-bit bases + uint8 counts
//                     (or uint8 bases + uint32 counts on the wide path) in fixed-stride slots.
//   K2 lives in align_kernel.cu.
//
// Integer dynamic programming: no tensor cores on purpose.
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

// ------------------------------------------------------------------------------------------
// K0: INT32 peak
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_int32_peak(int *out, int iters)
{
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int c = blockIdx.x | 1, d = -(int)blockIdx.x;
    // asm volatile keeps the compiler from collapsing the recurrence; ptxas still fuses add+max
#define TREAT_ADDMAX(x) asm volatile("{\n\tadd.s32 %0, %0, %1;\n\tmax.s32 %0, %0, %2;\n\t}" : "+r"(x) : "r"(c), "r"(d))
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {  // 8 x 8 x (add, max) = 128 lane-ops per iteration
            TREAT_ADDMAX(a0);
            TREAT_ADDMAX(a1);
            TREAT_ADDMAX(a2);
            TREAT_ADDMAX(a3);
            TREAT_ADDMAX(a4);
            TREAT_ADDMAX(a5);
            TREAT_ADDMAX(a6);
            TREAT_ADDMAX(a7);
        }
    }
#undef TREAT_ADDMAX
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
    uint8_t *s_lut = smem;  // 256: byte -> 2-bit code (0x80 = edit base) or, on the wide path, the upper-cased byte
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cap = p.stage_cap;  // multiple of 16
    uint8_t *s_code = smem + 256 + (size_t)warp * (cap + cap * sizeof(cnt_t));
    cnt_t *s_cnt = reinterpret_cast<cnt_t *>(s_code + cap);
    const uint8_t eb = p.edit_base;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint8_t up = (i >= 'a' && i <= 'z') ? (uint8_t)(i - 32) : (uint8_t)i;  // strings.ToUpper, ASCII
        s_lut[i] = WIDE ? up : (up == eb ? (uint8_t)0x80 : p.lut[up]);
    }
    __syncthreads();

    const uint32_t lt = (1u << lane) - 1u;
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
        // one 32-character chunk: ballot the non-edit bases, compact them, run lengths from the gaps
        auto chunk = [&](uint32_t c0, uint32_t v, bool inside) {
            const bool isb = inside && (WIDE ? v != eb : !(v & 0x80u));
            const uint32_t mask = __ballot_sync(FULL_MASK, isb);
            if (isb) {
                const uint32_t below = mask & lt;
                const uint32_t idx = nb + __popc(below);
                const int prev = below ? (int)(c0 + 31 - __clz(below)) : lastpos;
                uint32_t cnt = (uint32_t)((int)(c0 + lane) - prev - 1);
                if (cnt >= 255u) sat = true;  // informational on the wide path, which keeps exact counts
                if (!WIDE) {
                    widec |= (v == 3u);
                    if (cnt >= 255u) cnt = 255u;
                }
                s_code[idx] = (uint8_t)v;
                s_cnt[idx] = (cnt_t)cnt;
            }
            if (mask) lastpos = (int)(c0 + 31 - __clz(mask));
            nb += __popc(mask);
        };
        for (uint32_t c0 = 0; c0 < L; c0 += 64) {
            // both loads first: two global round trips overlap
            const uint32_t j0 = c0 + lane, j1 = j0 + 32;
            const uint8_t r0 = j0 < L ? seq[j0] : (uint8_t)0, r1 = j1 < L ? seq[j1] : (uint8_t)0;
            const uint32_t v0 = s_lut[r0], v1 = s_lut[r1];
            chunk(c0, v0, j0 < L);
            if (c0 + 32 < L) chunk(c0 + 32, v1, j1 < L);
        }
        if (lane == 0) {
            uint32_t cnt = (uint32_t)((int)L - 1 - lastpos);
            if (cnt >= 255u) sat = true;
            if (!WIDE && cnt >= 255u) cnt = 255u;
            s_cnt[nb] = (cnt_t)cnt;
        }
        __syncwarp();
        const uint32_t n = nb;
        // zero the tails so that the output below can move whole words
        s_code[n + lane] = 0;
        s_cnt[n + 1 + lane] = 0;
        __syncwarp();
        const bool any_sat = __any_sync(FULL_MASK, sat), any_wide = __any_sync(FULL_MASK, widec);
        // bases
        uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
        bool same = false;  // stripped bases equal the template's: such a read needs no DP (see k_same_bases)
        if (!WIDE) {
            const uint32_t nwords = align16((n + 3) / 4) / 4;
            const uint32_t twords = (p.tm + 15) / 16;
            bool eq = p.tpk != nullptr && n == p.tm;
            for (uint32_t wi = lane; wi < nwords; wi += 32) {
                // 16 codes (one byte each) -> 16 x 2 bits: (x * M) >> 24 gathers the four 2-bit fields of a word
                const uint4 x = *reinterpret_cast<const uint4 *>(s_code + 16 * wi);
                constexpr uint32_t M = 0x01041040u;
                const uint32_t lo = __byte_perm(x.x * M, x.y * M, 0x0073), hi = __byte_perm(x.z * M, x.w * M, 0x0073);
                const uint32_t word = __byte_perm(lo, hi, 0x5410);
                gb[wi] = word;
                if (eq && wi < twords) eq = word == p.tpk[wi];  // codes past n are zero in both
            }
            same = __all_sync(FULL_MASK, eq);
        } else {
            const uint32_t nwords = align16(n) / 4;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_code);
            for (uint32_t wi = lane; wi < nwords; wi += 32) gb[wi] = src[wi];
        }
        // counts (n + 1 entries)
        uint32_t *gc = reinterpret_cast<uint32_t *>(p.counts + r * (uint64_t)p.counts_stride);
        {
            const uint32_t nwords = WIDE ? align16(4 * (n + 1)) / 4 : align16(n + 1) / 4;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_cnt);
            for (uint32_t wi = lane; wi < nwords; wi += 32) gc[wi] = src[wi];
        }
        if (lane == 0) {
            p.n[r] = (uint16_t)n;
            p.flags[r] = (uint8_t)((any_sat ? TREAT_B200_ST_SATURATED : 0) | (any_wide ? TREAT_B200_ST_WIDE_BASE : 0) |
                                   (same ? kFlagSameBases : 0));
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


int configure_pack_kernels(int max_smem_optin)
{
    cudaError_t ce = cudaFuncSetAttribute(k_pack<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
