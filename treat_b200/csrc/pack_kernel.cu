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
#include <cstdlib>
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

// read-back of a few scalars into pinned (device-mapped) host memory; see peek() in capi.cu
__global__ void k_peek(uint32_t *dst, const uint32_t *src, uint32_t nwords)
{
    if (threadIdx.x < nwords) dst[threadIdx.x] = src[threadIdx.x];
}

int launch_peek(void *h_dst, const void *d_src, uint32_t bytes, void *stream)
{
    if (bytes > 128u || (bytes & 3u)) return (int)cudaErrorInvalidValue;
    k_peek<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint32_t *>(h_dst), reinterpret_cast<const uint32_t *>(d_src), bytes / 4u);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K1, wide path (NewFragment, fragment.go:91-121): bases stay bytes, counts are uint32 -- for templates with a base
// outside the packed alphabet or a count >= 255, and for chunks with an edit-base run >= 255.  One warp per read,
// one character per lane: ballot the non-edit bases, compact them, run lengths from the gaps.
// ------------------------------------------------------------------------------------------
constexpr int kPackWarpsMax = 8;

__global__ void __launch_bounds__(kPackWarpsMax * 32) k_pack_wide(const __grid_constant__ PackParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *s_up = smem;           // 256: byte -> upper-cased byte (strings.ToUpper, ASCII)
    uint8_t *s_other = smem + 256;  // 256: 1 = a base outside the packed alphabet (status bit ST_WIDE_BASE, as k_pack16 sets it)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cap = p.stage_cap;  // multiple of 16
    uint8_t *s_code = smem + 512 + (size_t)warp * (5 * cap);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_code + cap);
    const uint8_t eb = p.edit_base;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint8_t up = (i >= 'a' && i <= 'z') ? (uint8_t)(i - 32) : (uint8_t)i;
        s_up[i] = up;
        s_other[i] = (up != eb && p.lut[up] == 3) ? 1 : 0;
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
        auto chunk = [&](uint32_t c0, uint32_t v, bool inside) {
            const bool isb = inside && v != eb;
            const uint32_t mask = __ballot_sync(FULL_MASK, isb);
            if (isb) {
                const uint32_t below = mask & lt;
                const uint32_t idx = nb + __popc(below);
                const int prev = below ? (int)(c0 + 31 - __clz(below)) : lastpos;
                const uint32_t cnt = (uint32_t)((int)(c0 + lane) - prev - 1);
                if (cnt >= 255u) sat = true;  // informational here: the counts stay exact
                widec |= s_other[v] != 0;
                s_code[idx] = (uint8_t)v;
                s_cnt[idx] = cnt;
            }
            if (mask) lastpos = (int)(c0 + 31 - __clz(mask));
            nb += __popc(mask);
        };
        for (uint32_t c0 = 0; c0 < L; c0 += 64) {
            // both loads first: two global round trips overlap
            const uint32_t j0 = c0 + lane, j1 = j0 + 32;
            const uint8_t r0 = j0 < L ? seq[j0] : (uint8_t)0, r1 = j1 < L ? seq[j1] : (uint8_t)0;
            const uint32_t v0 = s_up[r0], v1 = s_up[r1];
            chunk(c0, v0, j0 < L);
            if (c0 + 32 < L) chunk(c0 + 32, v1, j1 < L);
        }
        if (lane == 0) {
            const uint32_t cnt = (uint32_t)((int)L - 1 - lastpos);
            if (cnt >= 255u) sat = true;
            s_cnt[nb] = cnt;
        }
        __syncwarp();
        const uint32_t n = nb;
        // zero the tails so that the output below can move whole words
        s_code[n + lane] = 0;
        s_cnt[n + 1 + lane] = 0;
        __syncwarp();
        const bool any_sat = __any_sync(FULL_MASK, sat), any_wide = __any_sync(FULL_MASK, widec);
        uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
        {
            const uint32_t nwords = align16(n) / 4;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(s_code);
            for (uint32_t wi = lane; wi < nwords; wi += 32) gb[wi] = src[wi];
        }
        uint32_t *gc = reinterpret_cast<uint32_t *>(p.counts + r * (uint64_t)p.counts_stride);  // n + 1 counts
        {
            const uint32_t nwords = align16(4 * (n + 1)) / 4;
            for (uint32_t wi = lane; wi < nwords; wi += 32) gc[wi] = s_cnt[wi];
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

// ------------------------------------------------------------------------------------------
// K1, packed path: 16 raw characters per lane.  Lane l takes the l-th 16-byte aligned block that overlaps the
// read (one LDG.128, prefetched one read ahead), so a read of up to ~500 characters is one pass of the warp:
// 16 LUT look-ups (PRMT builds the shared address of each), the lane's base count from a byte sum, an exclusive
// scan over the lanes for the output index, and an unrolled emit of one (code, preceding edit-base run) pair per
// base.  Runs that cross lanes are patched into the lane's first base afterwards.  Characters of a block that lie
// outside the read are turned into edit bases first (mask table, no branch).
// ------------------------------------------------------------------------------------------
constexpr uint32_t kLutEdit = 0x80u;   // LUT value of the edit base
constexpr uint32_t kLutOther = 7u;     // outside the alphabet: code 3 in the low two bits, bit 2 flags it
constexpr uint32_t kPack16Hdr = 1280;  // LUT (256-aligned inside the first 512 bytes), byte-mask table at 768, stages from here

__device__ __forceinline__ void sts_u16r(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8n(uint32_t addr)  // not volatile: the LUT is read-only after the barrier
{
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// SPAN: reads given as (begin, length) inside one buffer (FASTA records in place) instead of back-to-back offsets
template <bool SPAN>
__global__ void __launch_bounds__(kPackWarpsMax * 32, 4) k_pack16(const __grid_constant__ PackParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cap = p.stage_cap;  // multiple of 16; one 2-byte pair per base
    const uint32_t smem32 = smem_u32(smem);
    const uint32_t lut32 = (smem32 + 255u) & ~255u;  // 256-aligned: PRMT(word, lut32) splices a character into the address
    uint8_t *s_lut = smem + (lut32 - smem32);
    uint4 *s_mask = reinterpret_cast<uint4 *>(smem + 768);  // [17]: t lowest bytes 0xff
    uint8_t *s_pair = smem + kPack16Hdr + (size_t)warp * (2 * cap);
    const uint32_t pair32 = smem_u32(s_pair);
    const uint8_t eb = p.edit_base;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint8_t up = (i >= 'a' && i <= 'z') ? (uint8_t)(i - 32) : (uint8_t)i;  // strings.ToUpper, ASCII
        const uint8_t c = p.lut[up];
        s_lut[i] = up == eb ? (uint8_t)kLutEdit : (c == 3 ? (uint8_t)kLutOther : c);
    }
    if (threadIdx.x < 17) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = min(max((int)threadIdx.x - 4 * i, 0), 4);
            w[i] = k >= 4 ? 0xffffffffu : (1u << (8 * k)) - 1u;
        }
        s_mask[threadIdx.x] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();

    const uint32_t eb4 = (uint32_t)eb * 0x01010101u;
    const uint32_t lt = (1u << lane) - 1u;
    const uintptr_t base_addr = reinterpret_cast<uintptr_t>(p.seqs);
    // the bytes of p.seqs this launch may touch: [lo, hi)
    const int64_t lo = SPAN ? 0 : p.N ? (int64_t)(p.off[0] - p.off_bias) : 0;
    const int64_t hi = SPAN ? (int64_t)p.span : p.N ? (int64_t)(p.off[p.N] - p.off_bias) : 0;
    // start and length of read r
    auto where = [&](uint64_t r, uint64_t &b, uint32_t &L) {
        if (SPAN) {
            b = p.begin[r];
            L = p.len[r];
        } else {
            const uint64_t o0 = p.off[r], o1 = p.off[r + 1];
            b = o0 - p.off_bias;
            L = (uint32_t)(o1 - o0);
        }
    };
    // the 16-byte aligned block at offset x of p.seqs (byte loads where it sticks out of [lo, hi))
    // (p.padded: the buffer is 16-byte aligned and readable 16 bytes past its end, so no block ever sticks out)
    auto load_block = [&](int64_t x) -> uint4 {
        if (p.padded || (x >= lo && x + 16 <= hi)) return *reinterpret_cast<const uint4 *>(p.seqs + x);
        uint32_t w[4] = {0u, 0u, 0u, 0u};  // first / last read of a caller's buffer: byte loads
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int64_t y = x + q;
            if (y >= lo && y < hi) w[q >> 2] |= (uint32_t)p.seqs[y] << (8 * (q & 3));
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    // this lane's block of the first 32 of a read (anything past the read's last block: it is masked out later)
    auto first_block = [&](uint64_t b, uint32_t L) -> uint4 {
        const uint32_t a = (uint32_t)((base_addr + b) & 15u);
        return (uint32_t)lane < ((a + L + 15u) >> 4) ? load_block((int64_t)b - a + 16 * lane) : make_uint4(0, 0, 0, 0);
    };

    uint32_t warp_max_n = 0, warp_sat = 0;
    const uint64_t pw = blockDim.x >> 5;
    const uint64_t gw = (uint64_t)blockIdx.x * pw + warp, nw = (uint64_t)gridDim.x * pw;
    // software pipeline: offsets two reads ahead, the first block one read ahead
    uint64_t b_cur = 0, b_nxt = 0;
    uint32_t L_cur = 0, L_nxt = 0;
    uint4 blk = make_uint4(0, 0, 0, 0);
    if (gw < p.N) {
        where(gw, b_cur, L_cur);
        blk = first_block(b_cur, L_cur);
    }
    if (gw + nw < p.N) where(gw + nw, b_nxt, L_nxt);
    for (uint64_t r = gw; r < p.N; r += nw) {
        uint4 blk_n = make_uint4(0, 0, 0, 0);
        uint64_t b_nn = 0;
        uint32_t L_nn = 0;
        if (r + nw < p.N) blk_n = first_block(b_nxt, L_nxt);
        if (r + 2 * nw < p.N) where(r + 2 * nw, b_nn, L_nn);

        const uint32_t L = L_cur;
        const uint32_t a = (uint32_t)((base_addr + b_cur) & 15u);
        const int64_t x0 = (int64_t)b_cur - a;
        const uint32_t blocks = (a + L + 15u) >> 4;
        const uint32_t nsb = blocks > 32 ? (blocks + 31) >> 5 : 1;  // passes of 32 blocks = 512 characters
        uint32_t nb_total = 0, vor = 0;
        int carry = -(int)a;  // characters since the last base in front of lane 0's block (the a masked-in ones do not count)
        bool sat = false;
        for (uint32_t sb = 0; sb < nsb; sb++) {
            const uint32_t j = sb * 32 + lane;
            if (sb > 0) blk = j < blocks ? load_block(x0 + 16 * (int64_t)j) : make_uint4(0, 0, 0, 0);
            // characters outside the read become edit bases: keep bytes [vlo, vhi) of the block
            const int pos0 = (int)(16 * j) - (int)a;  // read position of this lane's first character
            const int vlo = min(max(-pos0, 0), 16), vhi = min(max((int)L - pos0, 0), 16);
            const uint4 mh = s_mask[vhi], ml = s_mask[vlo];
            uint32_t w[4] = {blk.x, blk.y, blk.z, blk.w};
            const uint32_t keep[4] = {mh.x & ~ml.x, mh.y & ~ml.y, mh.z & ~ml.z, mh.w & ~ml.w};
#pragma unroll
            for (int i = 0; i < 4; i++) w[i] = (w[i] & keep[i]) | (eb4 & ~keep[i]);
            uint32_t v[16], sum = 0, orv = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                v[k] = lds_u8n(__byte_perm(w[k >> 2], lut32, 0x7650 + (k & 3)));  // lut32 | character
                sum += v[k];
                orv |= v[k];
            }
            vor |= orv;
            const uint32_t nbl = 16u - (sum >> 7);  // codes sum to < 128, every edit base adds 128
            uint32_t incl = nbl;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL_MASK, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t tot = __shfl_sync(FULL_MASK, incl, 31);
            const uint32_t first32 = pair32 + 2u * (nb_total + incl - nbl);
            // emit: pair = code | (edit-base characters since the lane's previous base) << 8
            uint32_t addr = first32;
            int last8 = -256;  // (position of the lane's last base) << 8
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (v[k] < kLutEdit) {
                    sts_u16r(addr, v[k] + (uint32_t)(((k - 1) << 8) - last8));
                    addr += 2u;
                    last8 = k << 8;
                }
            }
            const uint32_t run = (uint32_t)((15 << 8) - last8) >> 8;  // characters behind the lane's last base (16: it has none)
            // add the run in front of the block to the lane's first base
            const uint32_t E = __ballot_sync(FULL_MASK, nbl != 0u);
            const uint32_t below = E & lt;
            const int pl = below ? 31 - __clz(below) : 0;
            const uint32_t tail_pl = __shfl_sync(FULL_MASK, run, pl);
            if (nbl) {
                const int cin = below ? 16 * (lane - pl - 1) + (int)tail_pl : 16 * lane + carry;
                const uint32_t pr = lds_u16(first32);
                uint32_t c = (uint32_t)((int)(pr >> 8) + cin);
                if (c >= 255u) {
                    sat = true;
                    c = 255u;
                }
                sts_u16r(first32, (pr & 0xffu) | (c << 8));
            }
            const int hl = E ? 31 - __clz(E) : 0;
            const uint32_t tail_hl = __shfl_sync(FULL_MASK, run, hl);
            carry = E ? 16 * (31 - hl) + (int)tail_hl : carry + 512;
            nb_total += tot;
        }
        const uint32_t n = nb_total;
        if (lane == 0) {
            uint32_t c = (uint32_t)(carry - (int)(512u * nsb - a - L));  // minus the masked-in characters behind the read
            if (c >= 255u) {
                sat = true;
                c = 255u;
            }
            sts_u16r(pair32 + 2u * n, c << 8);  // EditSite[n]: the run behind the last base
        }
        if (lane < 16) sts_u16r(pair32 + 2u * (n + 1 + lane), 0u);  // whole 16-pair groups below
        __syncwarp();
        const bool any_sat = __any_sync(FULL_MASK, sat), any_wide = __any_sync(FULL_MASK, (vor & 4u) != 0u);
        // output: lane g turns pairs [16 g, 16 g + 16) into one word of 2-bit codes and 16 count bytes
        uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
        uint4 *gc = reinterpret_cast<uint4 *>(p.counts + r * (uint64_t)p.counts_stride);
        const uint32_t bwords = align16((n + 3) / 4) / 4, cgroups = (n + 16) / 16, twords = (p.tm + 15) / 16;
        uint32_t nd = (p.tpk != nullptr && n == p.tm) ? 0u : 2u;  // bases that differ from the template's (2: not comparable)
        for (uint32_t g = lane; g < max(bwords, cgroups); g += 32) {
            uint32_t word = 0;
            if (g < cgroups) {
                const uint4 x = *reinterpret_cast<const uint4 *>(s_pair + 32 * g), y = *reinterpret_cast<const uint4 *>(s_pair + 32 * g + 16);
                gc[g] = make_uint4(__byte_perm(x.x, x.y, 0x7531), __byte_perm(x.z, x.w, 0x7531), __byte_perm(y.x, y.y, 0x7531),
                                   __byte_perm(y.z, y.w, 0x7531));
                // 16 codes (one byte each) -> 16 x 2 bits: (c * M) >> 24 gathers the four 2-bit fields of a word
                constexpr uint32_t M = 0x01041040u, C3 = 0x03030303u;
                const uint32_t c0 = __byte_perm(x.x, x.y, 0x6420) & C3, c1 = __byte_perm(x.z, x.w, 0x6420) & C3;
                const uint32_t c2 = __byte_perm(y.x, y.y, 0x6420) & C3, c3 = __byte_perm(y.z, y.w, 0x6420) & C3;
                const uint32_t lo16 = __byte_perm(c0 * M, c1 * M, 0x0073), hi16 = __byte_perm(c2 * M, c3 * M, 0x0073);
                word = __byte_perm(lo16, hi16, 0x5410);
            }
            if (g < bwords) gb[g] = word;
            if (nd < 2u && g < twords) {  // codes past n are zero in both
                const uint32_t x = word ^ p.tpk[g];
                nd += __popc((x | (x >> 1)) & 0x55555555u);
            }
        }
        const uint32_t ndiff = __reduce_add_sync(FULL_MASK, nd);
        const bool same = ndiff == 0u, one = ndiff == 1u;
        if (lane == 0) {
            p.n[r] = (uint16_t)n;
            p.flags[r] = (uint8_t)((any_sat ? TREAT_B200_ST_SATURATED : 0) | (any_wide ? TREAT_B200_ST_WIDE_BASE : 0) |
                                   (same ? kFlagSameBases : 0) | (one ? kFlagOneMismatch : 0));
            p.raw_len[r] = L;
        }
        warp_max_n = max(warp_max_n, n);
        warp_sat |= any_sat ? 1u : 0u;
        __syncwarp();
        b_cur = b_nxt, L_cur = L_nxt, blk = blk_n;
        b_nxt = b_nn, L_nxt = L_nn;
    }
    if (lane == 0) {
        if (warp_max_n) atomicMax(&p.scalars[0], warp_max_n);
        if (warp_sat) atomicOr(&p.scalars[1], 1u);
    }
}

int launch_pack(const PackParams &p, bool wide, int num_sms, int max_smem_optin, void *stream)
{
    const size_t per_warp = (size_t)p.stage_cap * (wide ? 5 : 2);
    int warps = (int)(((size_t)max_smem_optin - kPack16Hdr) / per_warp);
    if (warps < 1) return -1000;
    if (warps > kPackWarpsMax) warps = kPackWarpsMax;
    const size_t smem = kPack16Hdr + per_warp * warps;  // k_pack_wide only needs a 512-byte header
    uint64_t blocks = (p.N + warps - 1) / warps;
    // persistent grid: as many CTAs as are resident at once
    int per_sm = 0;
    cudaError_t oe = wide        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pack_wide, warps * 32, smem)
                     : p.begin   ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pack16<true>, warps * 32, smem)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pack16<false>, warps * 32, smem);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    if (wide)
        k_pack_wide<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    else if (p.begin)
        k_pack16<true><<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    else
        k_pack16<false><<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}


int configure_pack_kernels(int max_smem_optin)
{
    cudaError_t ce = cudaFuncSetAttribute(k_pack_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack16<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack16<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
