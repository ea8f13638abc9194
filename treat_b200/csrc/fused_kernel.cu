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

#include <algorithm>
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
    // a list shorter than the grid (the usual case: few reads are longer than a pass of k_pack_finish2): CTAs without a read
    // leave before the set-up
    if (p.sel && (uint64_t)blockIdx.x * (blockDim.x >> 5) >= (uint64_t)*p.sel_count) return;
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
    const uint64_t pw = blockDim.x >> 5;  // warps per CTA: fewer than kFusedWarps when the reads are very long
    const uint64_t gw = (uint64_t)blockIdx.x * pw + warp, nw = (uint64_t)gridDim.x * pw;
    // the reads of this launch: all of them, or the ones k_pack_finish2 left (longer than one of its passes)
    const uint64_t total = p.sel ? (uint64_t)*p.sel_count : p.N;
    auto read_at = [&](uint64_t i) -> uint64_t { return p.sel ? (uint64_t)p.sel[i] : i; };
    // software pipeline: offsets two reads ahead, the first block one read ahead
    uint64_t b_cur = 0, b_nxt = 0;
    uint32_t L_cur = 0, L_nxt = 0;
    uint4 blk = make_uint4(0, 0, 0, 0);
    if (gw < total) {
        where(read_at(gw), b_cur, L_cur);
        blk = first_block(b_cur, L_cur);
    }
    if (gw + nw < total) where(read_at(gw + nw), b_nxt, L_nxt);
    for (uint64_t ri = gw; ri < total; ri += nw) {
        const uint64_t r = read_at(ri);
        uint4 blk_n = make_uint4(0, 0, 0, 0);
        uint64_t b_nn = 0;
        uint32_t L_nn = 0;
        if (ri + nw < total) blk_n = first_block(b_nxt, L_nxt);
        if (ri + 2 * nw < total) where(read_at(ri + 2 * nw), b_nn, L_nn);

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

// ------------------------------------------------------------------------------------------
// K1f, second form: G lanes per read (G = 16: two reads per warp, templates of up to 255 bases; G = 32: up to 511),
// 32 raw characters per lane, so a read of up to 32 G - 31 characters is one pass and everything the warp does once per
// pass (offsets, block loads, the scan, the epilogue, the editing-site assembly) is shared by two reads.
//
//   * base mask: byte-SIMD compares on the raw words ((c & 0xDF) == edit base is exact for a letter) give one flag bit
//     per character; the eight words' flags are folded into ONE 32-bit mask of the lane's characters (bit 8 b + i =
//     character 4 i + b), ANDed with the lane's part of the read (a table of folded prefix masks) -- no masking of the
//     raw bytes, the lane's base count is one POPC;
//   * emit: one 16-bit pair per base, code | (position in the lane blocks + 1) << 2: LUT look-up (PRMT + LDS), one add,
//     bit test, predicated STS + pointer bump (six instructions per character).  The pairs of a read end up compacted in
//     shared memory; the read's offset in front and its end behind close the sequence;
//   * epilogue: lane g takes pairs [16 g, 16 g + 16): the codes go into one 2-bit word (PRMT + multiply-gather), the
//     edit-site counts are the differences of neighbouring positions minus one (byte-SIMD, two sites per step) -- no run
//     lengths are tracked while stripping and no lane ever patches another lane's count;
//   * bases equal to the template's (or one substitution): the editing-site assembly on the registers (finish_identity:
//     findJSS / findJES as ballots + one shuffle each, the JuncSeq window straight from the stored positions, the record,
//     the summary counters as four running sums); otherwise the packed read goes to HBM and its index to the DP list.
// Reads longer than a pass go to p.long_list for k_pack_finish (the multi-pass form).
// ------------------------------------------------------------------------------------------
#ifndef TREAT_B200_F2_WARPS
#define TREAT_B200_F2_WARPS 10
#endif
constexpr int kF2Warps = TREAT_B200_F2_WARPS;
constexpr uint32_t kF2Pad = 16;  // bytes in front of a group's pairs: pair -1 (the read's offset in its first block) sits at byte 14
__host__ __device__ constexpr uint32_t f2_group_bytes(int G) { return 2u * (32u * (uint32_t)G + 32u); }  // 8 pad + 32 G + 16 (+ 8) pairs
constexpr uint32_t kHdrRange = 768;  // [33] folded prefix masks (k_pack_finish keeps its byte-mask table here)

// bit of the folded base mask that stands for character k of the lane's 32
__host__ __device__ constexpr uint32_t f2_bit(int k) { return 8u * (uint32_t)(k & 3) + (uint32_t)(k >> 2); }

// store `val` at addr and advance addr iff the mask bit is set
template <uint32_t BIT, uint32_t ADD>
__device__ __forceinline__ void f2_emit(uint32_t &addr, uint32_t v, uint32_t vb, uint32_t mask)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t, val;\n\t"
        "and.b32 t, %3, %4;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "add.u32 val, %1, %2;\n\t"
        "add.u32 val, val, %5;\n\t"
        "@p st.shared.u16 [%0], val;\n\t"
        "@p add.u32 %0, %0, 2;\n\t}"
        : "+r"(addr)
        : "r"(v), "r"(vb), "r"(mask), "n"(1u << BIT), "n"(ADD)
        : "memory");
}

// the lane's 32 characters as four interleaved chains (characters 0..7, 8..15, 16..23, 24..31, each through its own pointer):
// a chain's pointer bump depends on the one before it; four of them in flight cut the time a warp waits on that dependency
template <int K>
struct F2Emit {
    __device__ __forceinline__ static void run(const uint32_t (&w)[8], uint32_t lut32, uint32_t vb, uint32_t mask, uint32_t (&addr)[4])
    {
        F2Emit<K - 1>::run(w, lut32, vb, mask, addr);
        constexpr int k0 = K - 1, k1 = k0 + 8, k2 = k0 + 16, k3 = k0 + 24;
        const uint32_t v0 = lds_u8n(__byte_perm(w[k0 >> 2], lut32, 0x7650 + (k0 & 3)));  // lut32 | character
        const uint32_t v1 = lds_u8n(__byte_perm(w[k1 >> 2], lut32, 0x7650 + (k1 & 3)));
        const uint32_t v2 = lds_u8n(__byte_perm(w[k2 >> 2], lut32, 0x7650 + (k2 & 3)));
        const uint32_t v3 = lds_u8n(__byte_perm(w[k3 >> 2], lut32, 0x7650 + (k3 & 3)));
        f2_emit<f2_bit(k0), (uint32_t)(k0 << 2)>(addr[0], v0, vb, mask);
        f2_emit<f2_bit(k1), (uint32_t)(k1 << 2)>(addr[1], v1, vb, mask);
        f2_emit<f2_bit(k2), (uint32_t)(k2 << 2)>(addr[2], v2, vb, mask);
        f2_emit<f2_bit(k3), (uint32_t)(k3 << 2)>(addr[3], v3, vb, mask);
    }
};
template <>
struct F2Emit<0> {
    __device__ __forceinline__ static void run(const uint32_t (&)[8], uint32_t, uint32_t, uint32_t, uint32_t (&)[4]) {}
};

#ifndef TREAT_B200_F2_MIN_CTAS
#define TREAT_B200_F2_MIN_CTAS 2  // 10 warps x 2 CTAs: 96 registers (no spilled prefetch registers), 20 resident warps per SM
#endif
template <int G, bool SPAN>
__global__ void __launch_bounds__(kF2Warps * 32, TREAT_B200_F2_MIN_CTAS) k_pack_finish2(const __grid_constant__ PackParams p, const __grid_constant__ AlignParams q)
{
    constexpr int RPW = 32 / G;
    constexpr uint32_t GB = f2_group_bytes(G);
    constexpr uint32_t GM = G == 32 ? 0xffffffffu : 0xffffu;
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (G - 1), grp = lane / G, lead = grp * G;
    const unsigned gmask = GM << lead;
    const uint32_t smem32 = smem_u32(smem);
    const uint32_t lut32 = (smem32 + 255u) & ~255u;  // 256-aligned: PRMT(word, lut32) splices a character into the address
    uint8_t *s_lut = smem + (lut32 - smem32);
    uint32_t *s_range = reinterpret_cast<uint32_t *>(smem + kHdrRange);
    uint32_t *s_tpk = reinterpret_cast<uint32_t *>(smem + kHdrTpk);
    const CtaSmem<uint8_t> cs = carve_cta<uint8_t>(smem + kHdrBytes, q, 0);
    uint32_t *zero = reinterpret_cast<uint32_t *>(smem + kHdrBytes + q.sm_tmpl_bytes);  // the all-match ops of a read without a gap
    uint8_t *s_alt = smem + kHdrBytes + q.sm_tmpl_bytes + q.sm_ops;  // [len_pad] 1: an alternative template's region starts at this site
    uint8_t *s_warp = s_alt + q.t.len_pad + (size_t)warp * (RPW * GB);
    uint8_t *s_grp = s_warp + grp * GB;  // this group's pairs
    const uint32_t grp32 = smem_u32(s_grp), pair32 = grp32 + kF2Pad;
    const uint8_t eb = p.edit_base;
    const int m = q.t.m, len = q.t.len, K = q.t.K;
    const uint32_t tm = p.tm, twords = (tm + 15) / 16;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint8_t up = (i >= 'a' && i <= 'z') ? (uint8_t)(i - 32) : (uint8_t)i;  // strings.ToUpper, ASCII
        s_lut[i] = up == eb ? (uint8_t)kLutEdit : p.lut[up];  // code 0..2, 3 = outside the alphabet
    }
    if (threadIdx.x < 33) {
        uint32_t v = 0;
        for (int k = 0; k < (int)threadIdx.x; k++) v |= 1u << f2_bit(k);
        s_range[threadIdx.x] = v;
    }
    if (threadIdx.x < 32) s_tpk[threadIdx.x] = threadIdx.x < twords ? p.tpk[threadIdx.x] : 0u;
    load_template(cs, q, 0);
    for (uint32_t i = threadIdx.x; i < q.sm_ops / 4; i += blockDim.x) zero[i] = 0u;
    for (int i = threadIdx.x; i < q.t.len_pad; i += blockDim.x) {
        uint8_t v = 0;
        for (int k = 0; k < q.t.n_alt; k++) v |= (q.t.alt[k] == len - 1 - i) ? 1 : 0;
        s_alt[i] = v;
    }
    WarpSmem ws{};
    ws.ops2 = zero;
    __syncthreads();

    const SameLane sl = load_same_lane(cs, q, gl);
    const uint32_t eb4 = (uint32_t)eb * 0x01010101u;
    uint32_t vb = (uint32_t)(32 * gl + 1) << 2;  // pair value of the lane's first character (code 0)
    asm volatile("" : "+r"(vb));  // one register, not an expression to rebuild at every use: the emit adds it once per character
    const uint32_t base_lo = (uint32_t)reinterpret_cast<uintptr_t>(p.seqs);
    const uint32_t N32 = (uint32_t)p.N;  // a launch holds < 2^31 reads (host check)
    // the bytes of p.seqs this launch may touch: [lo, hi)
    const int64_t lo = SPAN ? 0 : p.N ? (int64_t)(p.off[0] - p.off_bias) : 0;
    const int64_t hi = SPAN ? (int64_t)p.span : p.N ? (int64_t)(p.off[p.N] - p.off_bias) : 0;
    // where read r lies, as loaded (no arithmetic on the loaded values here: they are used one iteration later)
    // the reads of this launch: all of them, or the ones k_frontier left (p.sel: their indices, *p.sel_count of them)
    const uint32_t total = p.sel ? min(*p.sel_count, N32) : N32;
    // which read is item i of the launch (0xffffffff: past the end); with a list this is a load of its own, issued one pass
    // before the read's offsets are
    auto item = [&](uint32_t i) -> uint32_t { return i < total ? (p.sel ? p.sel[i] : i) : 0xffffffffu; };
    auto where = [&](uint32_t r, uint64_t &x0, uint64_t &x1) {
        x0 = x1 = SPAN ? 0ull : p.off_bias;
        if (r == 0xffffffffu) return;
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
    // this lane's two 16-byte blocks of a read (32-byte aligned in memory); zero when the lane is behind the read or the
    // read is too long for one pass.  Blocks that stick out of [lo, hi) -- only at the ends of a caller's buffer -- are
    // read byte by byte.
    auto load_blocks = [&](uint64_t b, uint32_t L, uint4 &A, uint4 &B) {
        A = make_uint4(0, 0, 0, 0);
        B = A;
        const uint32_t a = (base_lo + (uint32_t)b) & 31u;
        const uint32_t nblk = (a + L + 31u) >> 5;
        if (L == 0u || (uint32_t)gl >= nblk || nblk > (uint32_t)G) return;
        const int64_t x = (int64_t)b - a + 32 * gl;
        if (p.padded || (x >= lo && x + 32 <= hi)) {
            A = *reinterpret_cast<const uint4 *>(p.seqs + x);
            B = *reinterpret_cast<const uint4 *>(p.seqs + x + 16);
            return;
        }
        uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int c = 0; c < 32; c++) {
            const int64_t y = x + c;
            if (y >= lo && y < hi) w[c >> 2] |= (uint32_t)p.seqs[y] << (8 * (c & 3));
        }
        A = make_uint4(w[0], w[1], w[2], w[3]);
        B = make_uint4(w[4], w[5], w[6], w[7]);
    };

    uint32_t warp_max_n = 0, warp_sat = 0;
    unsigned long long ctr = 0;                  // counters of reads that took the general assembly (lane L: counter L)
    unsigned long long accA = 0, accC = 0;       // sum of ReadCount over the reads finished here / over those with one mismatch
    uint32_t accN = 0, accE = 0;                 // their numbers
    const uint32_t units = (total + RPW - 1) / RPW;
    const uint32_t gw = blockIdx.x * kF2Warps + warp, nw = gridDim.x * kF2Warps;
    // software pipeline: offsets two units ahead, the blocks one unit ahead; every load is issued a whole pass before its
    // result is touched
    uint64_t b_cur = 0, b_nxt = 0, x0_nxt = 0, x1_nxt = 0;
    uint32_t L_cur = 0, L_nxt = 0;
    uint4 blkA = make_uint4(0, 0, 0, 0), blkB = blkA;
    uint32_t r_cur = 0xffffffffu, r_nxt = 0xffffffffu, r_nn = 0xffffffffu;  // the reads of this pass, the next one, the one after
    if (gw < units) {
        uint64_t x0, x1;
        r_cur = item(gw * RPW + grp);
        where(r_cur, x0, x1);
        resolve(x0, x1, b_cur, L_cur);
        load_blocks(b_cur, L_cur, blkA, blkB);
    }
    r_nxt = item((gw + nw) * RPW + grp);
    where(r_nxt, x0_nxt, x1_nxt);  // past the end: an empty read
    r_nn = item((gw + 2 * nw) * RPW + grp);
    for (uint32_t u = gw; u < units; u += nw) {
        uint4 nA, nB;
        resolve(x0_nxt, x1_nxt, b_nxt, L_nxt);
        load_blocks(b_nxt, L_nxt, nA, nB);
        where(r_nn, x0_nxt, x1_nxt);
        const uint32_t r_ahead = item((u + 3 * nw) * RPW + grp);

        const bool in_batch = r_cur != 0xffffffffu;
        const uint32_t r = in_batch ? r_cur : 0u;
        const uint32_t a = (base_lo + (uint32_t)b_cur) & 31u;  // the read starts at character a of lane 0's blocks
        const bool too_long = ((a + L_cur + 31u) >> 5) > (uint32_t)G;
        const bool valid = in_batch && !too_long;
        const uint32_t L = valid ? L_cur : 0u;  // a read that is not handled here runs through as an empty one, without outputs
        if (too_long && in_batch && gl == 0) p.long_list[atomicAdd(p.long_count, 1u)] = r;
        const uint32_t read_count = (valid && q.read_count) ? q.read_count[r] : 1u;  // needed at the very end: in flight meanwhile

        // ---- which of the lane's 32 characters are bases of the read -------------------------------------------
        const uint32_t w[8] = {blkA.x, blkA.y, blkA.z, blkA.w, blkB.x, blkB.y, blkB.z, blkB.w};
        uint32_t f[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t x = (w[i] & 0xdfdfdfdfu) ^ eb4;  // a byte of x is zero iff the character is the edit base (either case)
            f[i] = (((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
        }
        const uint32_t folded = (f[0] >> 7) | (f[1] >> 6) | (f[2] >> 5) | (f[3] >> 4) | (f[4] >> 3) | (f[5] >> 2) | (f[6] >> 1) | f[7];
        const int c0 = (int)a - 32 * gl;  // lane characters [c0, c0 + L) belong to the read
        const uint32_t mask = folded & s_range[min(max(c0 + (int)L, 0), 32)] & ~s_range[min(max(c0, 0), 32)];
        const uint32_t nbl = __popc(mask);
        uint32_t incl = nbl;
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL_MASK, incl, d, G);
            if (gl >= d) incl += t;
        }
        const uint32_t n = __shfl_sync(FULL_MASK, incl, G - 1, G);
        // ---- emit: pair = code | (character index in the lane blocks + 1) << 2 ------------------------------------
        // characters 8 c .. 8 c + 7 of the lane are the bits 8 b + i with i in {2 c, 2 c + 1} of the folded mask
        uint32_t addr[4];
        addr[0] = pair32 + 2u * (incl - nbl);
        addr[1] = addr[0] + 2u * (uint32_t)__popc(mask & 0x03030303u);
        addr[2] = addr[0] + 2u * (uint32_t)__popc(mask & 0x0f0f0f0fu);
        addr[3] = addr[0] + 2u * (uint32_t)__popc(mask & 0x3f3f3f3fu);
        const uint32_t a1 = addr[1], a2 = addr[2], a3 = addr[3];
        F2Emit<8>::run(w, lut32, vb, mask, addr);
        // every chain filled exactly its slice of the lane's slice of the pair area
        TB_CHECK(q, addr[0] == a1 && addr[1] == a2 && addr[2] == a3 && addr[3] == pair32 + 2u * incl && n <= 32u * (uint32_t)G, 8);
        (void)a1, (void)a2, (void)a3;
        // in front of the first base: the read's offset a ("the base before the first one sat at character a - 1");
        // behind the last base: the read's end closes the last run, then consecutive positions (zero counts, zero codes)
        if (gl == 0) sts_u16r(pair32 - 2u, a << 2);
        if (gl < 16) sts_u16r(pair32 + 2u * (n + (uint32_t)gl), (a + L + 1u + (uint32_t)gl) << 2);
        __syncwarp();

        // ---- lane g turns pairs [16 g, 16 g + 16) into one word of 2-bit codes and 16 count bytes ---------------
        const uint32_t bwords = align16((n + 3) / 4) / 4, cgroups = (n + 16) / 16;
        auto group = [&](uint32_t g, uint4 &cnt, uint32_t &word, uint32_t &over) {
            const uint32_t at = pair32 + 32u * g;
            const uint4 x = lds_v4(at), y = lds_v4(at + 16u);
            const uint32_t prev = lds_u16(at - 2u);
            constexpr uint32_t M = 0x01041040u, C3 = 0x03030303u, PM = 0xfffcfffcu;
            // 16 codes (low two bits of every pair) -> 16 x 2 bits: (c * M) >> 24 gathers the four 2-bit fields of a word
            const uint32_t c0w = __byte_perm(x.x, x.y, 0x6420) & C3, c1w = __byte_perm(x.z, x.w, 0x6420) & C3;
            const uint32_t c2w = __byte_perm(y.x, y.y, 0x6420) & C3, c3w = __byte_perm(y.z, y.w, 0x6420) & C3;
            const uint32_t lo16 = __byte_perm(c0w * M, c1w * M, 0x0073), hi16 = __byte_perm(c2w * M, c3w * M, 0x0073);
            word = __byte_perm(lo16, hi16, 0x5410);
            // counts: (position - previous position - 1) for both halves of a word at once (positions grow: no borrow)
            const uint32_t mm[8] = {x.x & PM, x.y & PM, x.z & PM, x.w & PM, y.x & PM, y.y & PM, y.z & PM, y.w & PM};
            uint32_t d[8];
            d[0] = mm[0] - __byte_perm(prev & 0xfffcu, mm[0], 0x5410);
#pragma unroll
            for (int i = 1; i < 8; i++) d[i] = mm[i] - __byte_perm(mm[i - 1], mm[i], 0x5432);
            over |= (d[0] | d[1] | d[2] | d[3] | d[4] | d[5] | d[6] | d[7]) & 0xfc00fc00u;  // a run of 255 or more
            cnt = make_uint4(__byte_perm(d[0] >> 2, d[1] >> 2, 0x6420) - 0x01010101u, __byte_perm(d[2] >> 2, d[3] >> 2, 0x6420) - 0x01010101u,
                             __byte_perm(d[4] >> 2, d[5] >> 2, 0x6420) - 0x01010101u, __byte_perm(d[6] >> 2, d[7] >> 2, 0x6420) - 0x01010101u);
        };
        uint4 cnt = make_uint4(0, 0, 0, 0);
        uint32_t word = 0, over = 0;
        if ((uint32_t)gl < cgroups) group((uint32_t)gl, cnt, word, over);
        uint32_t wide = word & (word >> 1) & 0x55555555u;  // a code 3: a base outside the alphabet
        // reads of more than 16 G bases: the further groups only matter for the flags and the packed output below
        for (uint32_t g = gl + G; g < cgroups; g += G) {
            uint4 c2;
            uint32_t w2;
            group(g, c2, w2, over);
            wide |= w2 & (w2 >> 1) & 0x55555555u;
        }
        // bases that differ from the template's (codes past n are zero in both words): 0, 1 or "2 and more"
        uint32_t nd = 2u;
        if (n == tm) {
            const uint32_t x = word ^ s_tpk[gl];
            nd = __popc((x | (x >> 1)) & 0x55555555u);
        }
        const uint32_t b_over = (__ballot_sync(FULL_MASK, over != 0u) >> lead) & GM, b_wide = (__ballot_sync(FULL_MASK, wide != 0u) >> lead) & GM;
        const uint32_t b_d1 = (__ballot_sync(FULL_MASK, nd != 0u) >> lead) & GM, b_d2 = (__ballot_sync(FULL_MASK, nd >= 2u) >> lead) & GM;
        uint32_t rflags = (b_over ? (uint32_t)TREAT_B200_ST_SATURATED : 0u) | (b_wide ? (uint32_t)TREAT_B200_ST_WIDE_BASE : 0u);
        const bool identity = valid && b_d2 == 0u && (b_d1 & (b_d1 - 1u)) == 0u;  // at most one lane with exactly one differing base
        const int mism = b_d1 ? 1 : 0;

        // ---- no DP: the identity alignment is the unique optimum (score m, or m - 2 with the one mismatch) --------
        // computeT / findJSS / findJES / NewAlignment (align.go:95-279) on the count bytes against the FE / PE rows
        const uint32_t cw[4] = {cnt.x, cnt.y, cnt.z, cnt.w};
        uint32_t df[4], dp[4];  // non-zero byte: the read's count differs from FE / PE at that site
#pragma unroll
        for (int i = 0; i < 4; i++) {
            df[i] = (cw[i] ^ sl.fe[i]) & sl.valid[i];
            dp[i] = (cw[i] ^ sl.pe[i]) & sl.valid[i];
        }
        // this lane's highest site that differs from FE / lowest that differs from PE
        const uint32_t hw = df[3] ? df[3] : df[2] ? df[2] : df[1] ? df[1] : df[0];
        const int hb = df[3] ? 12 : df[2] ? 8 : df[1] ? 4 : 0;
        const int hl = 16 * gl + hb + ((31 - __clz(hw)) >> 3);
        const uint32_t lw = dp[0] ? dp[0] : dp[1] ? dp[1] : dp[2] ? dp[2] : dp[3];
        const int lb = dp[0] ? 0 : dp[1] ? 4 : dp[2] ? 8 : 12;
        const int ll = 16 * gl + lb + ((__ffs(lw) - 1) >> 3);
        const uint32_t b_fe = (__ballot_sync(FULL_MASK, hw != 0u) >> lead) & GM, b_pe = (__ballot_sync(FULL_MASK, lw != 0u) >> lead) & GM;
        const bool t0_all = b_fe == 0u;  // every site matches FE
        const int hc0 = __shfl_sync(FULL_MASK, hl, lead + (b_fe ? 31 - __clz(b_fe) : 0));
        const int lc1v = __shfl_sync(FULL_MASK, ll, lead + (b_pe ? __ffs(b_pe) - 1 : 0));
        // findJSS (align.go:110-120): no site matching FE at all gives -1.  A differing site below the last one means the last
        // one matches; only when the last site itself differs do all sites have to be looked at.
        bool fe_none = false;
        if (__any_sync(FULL_MASK, identity && !t0_all && hc0 == len - 1)) {
            uint32_t hz = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t x = df[i] | ~sl.valid[i];
                hz |= (x - 0x01010101u) & ~x & 0x80808080u;  // non-zero iff some byte of x is zero: a real site equal to FE
            }
            fe_none = ((__ballot_sync(FULL_MASK, hz != 0u) >> lead) & GM) == 0u;
        }
        int junc_start = t0_all ? len - 1 : (fe_none ? -1 : len - 1 - hc0);
        // does the junction start sit on the first site of an alternative template's region?  Then the general assembly
        // decides (align.go:186-205); the table says for which sites that can be
        bool general = identity && K > 2 && junc_start >= 0 && junc_start < len && s_alt[len - 1 - junc_start] != 0;
        if (__any_sync(FULL_MASK, general)) {
            bool hit = false;
            if (general) {
                const int ti_s = len - 1 - junc_start, qd = (ti_s >> 2) & 3;
                const uint32_t wsel = qd == 0 ? cw[0] : qd == 1 ? cw[1] : qd == 2 ? cw[2] : cw[3];
                const uint32_t cstar = (__shfl_sync(gmask, wsel, lead + (ti_s >> 4)) >> (8 * (ti_s & 3))) & 0xffu;
                for (int k = 2; k < K; k++) {
                    if ((uint32_t)cs.tcount[k * q.t.len_pad + ti_s] == cstar) {
                        hit = junc_start == cs.alt[k - 2];
                        break;
                    }
                }
            }
            general = hit;
        }
        if (identity && !general) {
            const int has_mutation = ((q.flags & TREAT_B200_EXCLUDE_SNPS) && mism) ? 1 : 0;  // align.go:240-241
            int junc_end = b_pe ? len - 1 - lc1v : -1;
            int edit_stop = junc_start - 1, junc_len = 0;
            uint32_t js_off = 0, js_len = 0, status = rflags;
            if (junc_end > edit_stop) {
                junc_len = junc_end - edit_stop;
                if (has_mutation == 0 && !t0_all) {
                    const int from = (len - 1) - junc_end, to = (len - 1) - edit_stop;
                    if (to > m + 1) {
                        status |= TREAT_B200_ST_REF_PANIC;  // Go: index out of range on frag.EditSite (align.go:250-257)
                    } else {
                        // JuncSeq = raw[just behind base from-1, just behind base to-1) (the trailing run for to = n + 1):
                        // the stored positions are those offsets
                        TB_CHECK(q, from >= 0 && from < to && to <= (int)n + 1, 9);  // both positions lie inside the pair area
                        const uint32_t o0 = (lds_u16(pair32 + 2u * (uint32_t)from - 2u) >> 2) - a;
                        const uint32_t o1 = min((lds_u16(pair32 + 2u * (uint32_t)to - 2u) >> 2) - a, L);
                        js_len = o1 - o0;
                        js_off = js_len ? o0 : 0u;
                    }
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
                dst[1] = make_uint4(read_count, (uint32_t)has_mutation | ((uint32_t)mism << 16), (uint32_t)(m - 2 * mism), js_off);
                dst[2] = make_uint4(js_len, (uint32_t)m | ((uint32_t)m << 16), status, L);
            }
            accA += read_count;
            accN += 1u;
            if (mism) {
                accC += read_count;
                accE += 1u;
            }
            // cmd/treat/handlers.go:203-215: three histograms, one lane each
            if (cs.sum_len && gl < 3 && !(edit_stop + off == q.t.tmpl_edit_stop && junc_len == 0)) {
                const uint32_t bin = gl == 0 ? (uint32_t)(edit_stop + 2) : gl == 1 ? (uint32_t)junc_len : (uint32_t)(junc_end + 2);
                if (bin < q.bins) {
                    const uint32_t idx = TREAT_B200_SUMMARY_HEADER + (uint32_t)gl * q.bins + bin;
                    const uint32_t old = atomicAdd(&cs.sum_lo[idx], read_count);
                    if (old + read_count < old) atomicAdd(&cs.sum_hi[idx], 1u);
                }
            }
            if (q.ops && !q.ops_gapped_only) {  // m match columns: an all-zero row
                uint32_t *go = reinterpret_cast<uint32_t *>(q.ops + r * (uint64_t)q.ops_stride);
                for (int wi = gl; wi < (int)(q.ops_stride / 4); wi += G) go[wi] = 0u;
            }
        } else if (valid && !identity) {
            // left for the DP: the packed read goes to HBM, its index to the list
            uint32_t *gb = reinterpret_cast<uint32_t *>(p.bases + r * (uint64_t)p.bases_stride);
            uint4 *gc = reinterpret_cast<uint4 *>(p.counts + r * (uint64_t)p.counts_stride);
            for (uint32_t g = gl; g < max(bwords, cgroups); g += G) {
                if (g >= (uint32_t)G) {
                    cnt = make_uint4(0, 0, 0, 0);
                    word = 0;
                    if (g < cgroups) group(g, cnt, word, over);
                }
                if (g < cgroups) gc[g] = cnt;
                if (g < bwords) gb[g] = word;
            }
            if (gl == 0) {
                p.n[r] = (uint16_t)n;
                p.flags[r] = (uint8_t)rflags;
                p.raw_len[r] = L;
                q.list[atomicAdd(q.list_count, 1u)] = r;
            }
        }
        // the junction start sits on an alternative template's region: the whole warp runs the general assembly on the
        // read's counts (as bytes in the pair area), one read at a time
        unsigned todo = __ballot_sync(FULL_MASK, general && gl == 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            __syncwarp();
            uint8_t *stage = s_warp + (src / G) * GB + kF2Pad;
            if (lead == src) reinterpret_cast<uint4 *>(stage)[gl] = cnt;
            __syncwarp();
            const uint32_t rf = rflags | (mism ? kFlagOneMismatch : kFlagSameBases);
            finish_same_general(q, cs, ws, lane, (uint64_t)__shfl_sync(FULL_MASK, r, src), __shfl_sync(FULL_MASK, rf, src), stage, ctr,
                                __shfl_sync(FULL_MASK, L, src));
        }
        if (valid) {
            warp_max_n = max(warp_max_n, n);
            warp_sat |= rflags & TREAT_B200_ST_SATURATED;
        }
        __syncwarp();
        b_cur = b_nxt, L_cur = L_nxt, blkA = nA, blkB = nB;
        r_cur = r_nxt, r_nxt = r_nn, r_nn = r_ahead;
    }
    warp_max_n = __reduce_max_sync(FULL_MASK, warp_max_n);
    warp_sat = __reduce_or_sync(FULL_MASK, warp_sat);
    if (lane == 0) {
        if (warp_max_n) atomicMax(&p.scalars[0], warp_max_n);
        if (warp_sat) atomicOr(&p.scalars[1], 1u);
    }
    // the header counters of the reads finished above (cmd/treat/stats.go:140-187), from the four running sums: lane L of
    // a group holds counter L.  None of these reads has a gap; a mismatch makes a read a mutant only with EXCLUDE_SNPS.
    {
        const bool ex = (q.flags & TREAT_B200_EXCLUDE_SNPS) != 0;
        const unsigned long long B = ex ? accC : 0ull, D = ex ? accE : 0ull, Nn = accN;
        const unsigned long long v = gl == 0 ? accA : gl == 1 ? accA - B : gl == 2 ? B : gl == 5 ? accC : gl == 7 ? Nn : gl == 8 ? Nn - D
                                   : gl == 9 ? D : gl == 12 ? (unsigned long long)accE : gl == 14 ? Nn
                                   : gl == 15 ? Nn * (unsigned long long)m * (unsigned long long)m : 0ull;
        if (cs.sum_len) ctr += v;
    }
    if (G == 16) ctr += __shfl_down_sync(FULL_MASK, ctr, 16);  // lane L and lane L + 16 both hold counter L
    flush_summary(q, cs.sum_len, cs.sum_lo, cs.sum_hi, lane, ctr);
}

// the fused launch applies to: packed alphabet, a template of at most 511 bases (32 lanes x 16 edit sites), marks on
bool pack_finish_applies(const PackParams &p, const AlignParams &q)
{
    return p.tpk != nullptr && q.t.len <= 512 && q.t.len >= 1 && q.list != nullptr && q.list_count != nullptr;
}

static void fused_smem_plan(AlignParams &q)
{
    const uint32_t sum_len = q.summary ? TREAT_B200_SUMMARY_HEADER + 3u * q.bins : 0u;
    q.sm_prof_bytes = 0;
    q.sm_tmpl_bytes = align16((uint32_t)(q.t.K * q.t.len_pad)) + align16(8u * (q.t.n_alt ? q.t.n_alt : 1)) + align16(8u * sum_len);
    q.sm_tmpl_bytes = (q.sm_tmpl_bytes + 127u) & ~127u;
    q.sm_ops = align16((((uint32_t)(2 * q.t.m) + 15) / 16 + 2) * 4);
    if (q.ncap < (uint32_t)q.t.m) q.ncap = (uint32_t)q.t.m;
}

template <typename Kern>
static int launch_fused(Kern kern, const PackParams &p, const AlignParams &q, uint64_t units_per_cta, uint64_t units, int threads,
                        size_t smem, int num_sms, void *stream)
{
    uint64_t blocks = (units + units_per_cta - 1) / units_per_cta;
    const int per_sm = occupancy_cached((const void *)kern, threads, smem);
    const uint64_t cap = (uint64_t)num_sms * per_sm;  // persistent grid: as many CTAs as are resident at once
    if (blocks > cap) blocks = cap;
    if (blocks == 0) return 0;
    kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(p, q);
    return (int)cudaGetLastError();
}

// multi-pass form (one read per warp, reads of any length).  With pack.sel set it handles only the listed reads.
int launch_pack_finish(const PackParams &pack, const AlignParams &planned, int num_sms, int max_smem_optin, void *stream)
{
    PackParams p = pack;
    if (p.stage_cap < 256) p.stage_cap = 256;  // the general assembly stages 32 x 16 count bytes in the pair area
    AlignParams q = planned;
    fused_smem_plan(q);
    const size_t per_warp = (size_t)p.stage_cap * 2;
    const size_t fixed = (size_t)kHdrBytes + q.sm_tmpl_bytes + q.sm_ops;
    if (fixed + per_warp > (size_t)max_smem_optin) return -1000;
    const int warps = (int)std::min<size_t>(kFusedWarps, ((size_t)max_smem_optin - fixed) / per_warp);  // very long reads: fewer warps
    const size_t smem = fixed + per_warp * warps;
    // a list: its length is on the device and it is short (reads longer than a pass of k_pack_finish2): one CTA per SM, so that
    // the CTAs' set-up does not cost more than the reads
    const uint64_t units = p.sel ? std::min<uint64_t>(p.N, (uint64_t)num_sms * warps) : p.N;
    if (p.begin) return launch_fused(k_pack_finish<true>, p, q, warps, units, warps * 32, smem, num_sms, stream);
    return launch_fused(k_pack_finish<false>, p, q, warps, units, warps * 32, smem, num_sms, stream);
}

// one-pass form: two reads per warp for templates of up to 255 bases, one read per warp up to 511; the edit base must be a
// letter (byte-SIMD case folding).  Reads too long for a pass are appended to pack.long_list.
bool pack_finish2_applies(const PackParams &p, const AlignParams &q)
{
    const uint8_t eb = p.edit_base;
    return pack_finish_applies(p, q) && eb >= 'A' && eb <= 'Z' && p.long_list != nullptr && p.long_count != nullptr;
}

int launch_pack_finish2(const PackParams &p, const AlignParams &planned, int num_sms, int max_smem_optin, void *stream)
{
    AlignParams q = planned;
    fused_smem_plan(q);
    const bool two = q.t.len <= 256;
    const size_t per_warp = two ? 2 * (size_t)f2_group_bytes(16) : (size_t)f2_group_bytes(32);
    const size_t smem = (size_t)kHdrBytes + q.sm_tmpl_bytes + q.sm_ops + (size_t)q.t.len_pad + per_warp * kF2Warps;
    if (smem > (size_t)max_smem_optin) return -1000;
    const uint64_t units = two ? (p.N + 1) / 2 : p.N;
    if (two) {
        if (p.begin) return launch_fused(k_pack_finish2<16, true>, p, q, kF2Warps, units, kF2Warps * 32, smem, num_sms, stream);
        return launch_fused(k_pack_finish2<16, false>, p, q, kF2Warps, units, kF2Warps * 32, smem, num_sms, stream);
    }
    if (p.begin) return launch_fused(k_pack_finish2<32, true>, p, q, kF2Warps, units, kF2Warps * 32, smem, num_sms, stream);
    return launch_fused(k_pack_finish2<32, false>, p, q, kF2Warps, units, kF2Warps * 32, smem, num_sms, stream);
}

int configure_fused_kernels(int max_smem_optin)
{
    cudaError_t ce = cudaFuncSetAttribute(k_pack_finish<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_pack_finish<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_pack_finish2<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_pack_finish2<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_pack_finish2<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_pack_finish2<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return (int)ce;
}

}  // namespace treat_b200
