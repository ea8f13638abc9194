// fused_kernel.cu -- K1f: NewFragment (fragment.go:91-121) and, for the reads that need no DP, the whole editing-site
// assembly (align.go:95-279) in ONE pass over the raw reads.
//
// k_pack16 (pack_kernel.cu) writes every read's packed bases + counts to HBM and k_same_bases reads them back only to
// finish ~92-97 % of an RPS12 batch with a few compares.  Here the warp that strips a read keeps going: its epilogue has
// the read's 2-bit codes and edit-site counts in registers (lane g: sites 16 g .. 16 g + 15), compares the codes with
// the template's (shared memory) and
//   * bases equal the template's, or same length and exactly one base differs (the diagonal is the unique optimum,
//     DESIGN.md 4): finish_same on the registers -> the 48-byte record; nothing else is written;
//   * otherwise: the packed read goes to HBM (as k_pack16 writes it) and its index to the DP list of k_align2 / k_align.
// Templates of up to 511 bases (32 lanes x 16 sites), packed alphabet; everything else takes k_pack16 + k_same_*.
#include <cuda_runtime.h>

#include <cstdint>

#include "assembly.cuh"
#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

namespace {

constexpr int kFusedWarps = 8;
constexpr uint32_t kLutEdit = 0x80u;   // LUT value of the edit base
constexpr uint32_t kLutOther = 7u;     // outside the alphabet: code 3 in the low two bits, bit 2 flags it
// shared-memory header: [0, 512) the LUT (256-aligned inside), then
constexpr uint32_t kHdrMask = 768;     // [17] uint4: t lowest bytes 0xff
constexpr uint32_t kHdrTpk = 1280;     // [32] template bases, 16 codes per word
constexpr uint32_t kHdrBytes = 1408;   // template section (carve_cta) starts here

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

}  // namespace

// p: the raw reads and where a packed read goes; q: template, outputs and summary of the align launch that follows
// (q.list / q.list_count receive the reads left for the DP; q.sm_tmpl_bytes / q.sm_ops as set by launch_pack_finish)
template <bool SPAN>
__global__ void __launch_bounds__(kFusedWarps * 32, 4) k_pack_finish(const __grid_constant__ PackParams p, const __grid_constant__ AlignParams q)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cap = p.stage_cap;  // multiple of 16; one 2-byte pair per base
    const uint32_t smem32 = smem_u32(smem);
    const uint32_t lut32 = (smem32 + 255u) & ~255u;  // 256-aligned: PRMT(word, lut32) splices a character into the address
    uint8_t *s_lut = smem + (lut32 - smem32);
    uint4 *s_mask = reinterpret_cast<uint4 *>(smem + kHdrMask);
    uint32_t *s_tpk = reinterpret_cast<uint32_t *>(smem + kHdrTpk);
    const CtaSmem<uint8_t> cs = carve_cta<uint8_t>(smem + kHdrBytes, q, 0);
    uint32_t *zero = reinterpret_cast<uint32_t *>(smem + kHdrBytes + q.sm_tmpl_bytes);  // the all-match ops of a read without a gap
    uint8_t *s_pair = smem + kHdrBytes + q.sm_tmpl_bytes + q.sm_ops + (size_t)warp * (2 * cap);
    const uint32_t pair32 = smem_u32(s_pair);
    const uint8_t eb = p.edit_base;
    const uint32_t tm = p.tm, twords = (tm + 15) / 16;
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
    if (threadIdx.x < 32) s_tpk[threadIdx.x] = threadIdx.x < twords ? p.tpk[threadIdx.x] : 0u;
    load_template(cs, q, 0);
    for (uint32_t i = threadIdx.x; i < q.sm_ops / 4; i += blockDim.x) zero[i] = 0u;
    WarpSmem ws{};
    ws.ops2 = zero;
    __syncthreads();

    const SameLane sl = load_same_lane(cs, q, lane);
    const uint32_t eb4 = (uint32_t)eb * 0x01010101u;
    const uint32_t lt = (1u << lane) - 1u;
    const uintptr_t base_addr = reinterpret_cast<uintptr_t>(p.seqs);
    // the bytes of p.seqs this launch may touch: [lo, hi)
    const int64_t lo = SPAN ? 0 : p.N ? (int64_t)(p.off[0] - p.off_bias) : 0;
    const int64_t hi = SPAN ? (int64_t)p.span : p.N ? (int64_t)(p.off[p.N] - p.off_bias) : 0;
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
    // the 16-byte aligned block at offset x of p.seqs (byte loads where it sticks out of [lo, hi); never with p.padded)
    auto load_block = [&](int64_t x) -> uint4 {
        if (p.padded || (x >= lo && x + 16 <= hi)) return *reinterpret_cast<const uint4 *>(p.seqs + x);
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int c = 0; c < 16; c++) {
            const int64_t y = x + c;
            if (y >= lo && y < hi) w[c >> 2] |= (uint32_t)p.seqs[y] << (8 * (c & 3));
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    };
    auto first_block = [&](uint64_t b, uint32_t L) -> uint4 {
        const uint32_t a = (uint32_t)((base_addr + b) & 15u);
        return (uint32_t)lane < ((a + L + 15u) >> 4) ? load_block((int64_t)b - a + 16 * lane) : make_uint4(0, 0, 0, 0);
    };

    uint32_t warp_max_n = 0, warp_sat = 0;
    unsigned long long ctr = 0;
    const uint64_t gw = (uint64_t)blockIdx.x * kFusedWarps + warp, nw = (uint64_t)gridDim.x * kFusedWarps;
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
        uint32_t rflags = (any_sat ? TREAT_B200_ST_SATURATED : 0u) | (any_wide ? TREAT_B200_ST_WIDE_BASE : 0u);
        // lane g turns pairs [16 g, 16 g + 16) into one word of 2-bit codes and 16 count bytes
        const uint32_t bwords = align16((n + 3) / 4) / 4, cgroups = (n + 16) / 16;
        auto group = [&](uint32_t g, uint4 &cnt, uint32_t &word) {
            const uint4 x = *reinterpret_cast<const uint4 *>(s_pair + 32 * g), y = *reinterpret_cast<const uint4 *>(s_pair + 32 * g + 16);
            cnt = make_uint4(__byte_perm(x.x, x.y, 0x7531), __byte_perm(x.z, x.w, 0x7531), __byte_perm(y.x, y.y, 0x7531),
                             __byte_perm(y.z, y.w, 0x7531));
            // 16 codes (one byte each) -> 16 x 2 bits: (c * M) >> 24 gathers the four 2-bit fields of a word
            constexpr uint32_t M = 0x01041040u, C3 = 0x03030303u;
            const uint32_t c0 = __byte_perm(x.x, x.y, 0x6420) & C3, c1 = __byte_perm(x.z, x.w, 0x6420) & C3;
            const uint32_t c2 = __byte_perm(y.x, y.y, 0x6420) & C3, c3 = __byte_perm(y.z, y.w, 0x6420) & C3;
            const uint32_t lo16 = __byte_perm(c0 * M, c1 * M, 0x0073), hi16 = __byte_perm(c2 * M, c3 * M, 0x0073);
            word = __byte_perm(lo16, hi16, 0x5410);
        };
        uint4 cnt = make_uint4(0, 0, 0, 0);
        uint32_t word = 0;
        if ((uint32_t)lane < cgroups) group((uint32_t)lane, cnt, word);
        // bases that differ from the template's (codes past n are zero in both words)
        uint32_t ndiff = 2u;
        if (n == tm) {
            const uint32_t x = word ^ s_tpk[lane];
            ndiff = __reduce_add_sync(FULL_MASK, __popc((x | (x >> 1)) & 0x55555555u));
        }
        if (ndiff <= 1u) {
            // no DP: the identity alignment is the unique optimum (score m, or m - 2 with the one mismatch)
            rflags |= ndiff ? kFlagOneMismatch : kFlagSameBases;
            if (!finish_same<32>(q, cs, s_mask, lane, FULL_MASK, 0, r, rflags, cnt, sl, ctr, L)) {
                // the junction start sits on an alternative template's region: the general assembly, counts as bytes
                __syncwarp();
                reinterpret_cast<uint4 *>(s_pair)[lane] = cnt;
                __syncwarp();
                finish_same_general(q, cs, ws, lane, r, rflags, s_pair, ctr, L);
            }
        } else {
            // left for the DP: the packed read goes to HBM, its index to the list
            uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
            uint4 *gc = reinterpret_cast<uint4 *>(p.counts + r * (uint64_t)p.counts_stride);
            for (uint32_t g = lane; g < max(bwords, cgroups); g += 32) {
                if (g >= 32u) {
                    cnt = make_uint4(0, 0, 0, 0);
                    word = 0;
                    if (g < cgroups) group(g, cnt, word);
                }
                if (g < cgroups) gc[g] = cnt;
                if (g < bwords) gb[g] = word;
            }
            if (lane == 0) {
                p.n[r] = (uint16_t)n;
                p.flags[r] = (uint8_t)rflags;
                p.raw_len[r] = L;
                q.list[atomicAdd(q.list_count, 1u)] = (uint32_t)r;
            }
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
    flush_summary(q, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// the fused launch applies to: packed alphabet, a template of at most 511 bases (32 lanes x 16 edit sites), marks on
bool pack_finish_applies(const PackParams &p, const AlignParams &q)
{
    return p.tpk != nullptr && q.t.len <= 512 && q.t.len >= 1 && q.list != nullptr && q.list_count != nullptr;
}

int launch_pack_finish(const PackParams &pack, const AlignParams &planned, int num_sms, int max_smem_optin, void *stream)
{
    PackParams p = pack;
    if (p.stage_cap < 256) p.stage_cap = 256;  // the general assembly stages 32 x 16 count bytes in the pair area
    AlignParams q = planned;
    const uint32_t sum_len = q.summary ? TREAT_B200_SUMMARY_HEADER + 3u * q.bins : 0u;
    q.sm_prof_bytes = 0;
    q.sm_tmpl_bytes = align16((uint32_t)(q.t.K * q.t.len_pad)) + align16(8u * (q.t.n_alt ? q.t.n_alt : 1)) + align16(8u * sum_len);
    q.sm_tmpl_bytes = (q.sm_tmpl_bytes + 127u) & ~127u;
    q.sm_ops = align16((((uint32_t)(2 * q.t.m) + 15) / 16 + 2) * 4);
    if (q.ncap < (uint32_t)q.t.m) q.ncap = (uint32_t)q.t.m;
    const size_t per_warp = (size_t)p.stage_cap * 2;
    const size_t fixed = (size_t)kHdrBytes + q.sm_tmpl_bytes + q.sm_ops;
    if (fixed + per_warp * kFusedWarps > (size_t)max_smem_optin) return -1000;
    const size_t smem = fixed + per_warp * kFusedWarps;
    uint64_t blocks = (p.N + kFusedWarps - 1) / kFusedWarps;
    int per_sm = 0;
    cudaError_t oe = p.begin ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pack_finish<true>, kFusedWarps * 32, smem)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pack_finish<false>, kFusedWarps * 32, smem);
    if (oe != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint64_t cap = (uint64_t)num_sms * per_sm;  // persistent grid: as many CTAs as are resident at once
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    if (p.begin)
        k_pack_finish<true><<<(unsigned)blocks, kFusedWarps * 32, smem, (cudaStream_t)stream>>>(p, q);
    else
        k_pack_finish<false><<<(unsigned)blocks, kFusedWarps * 32, smem, (cudaStream_t)stream>>>(p, q);
    return (int)cudaGetLastError();
}

int configure_fused_kernels(int max_smem_optin)
{
    cudaError_t ce = cudaFuncSetAttribute(k_pack_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce != cudaSuccess) return (int)ce;
    ce = cudaFuncSetAttribute(k_pack_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
