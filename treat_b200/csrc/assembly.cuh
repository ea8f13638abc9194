// assembly.cuh -- device code shared by the align kernels (align_kernel.cu) and the fused NewFragment + editing-site
// assembly kernel (fused_kernel.cu): shared-memory views of the template, the bit rows over edit sites, the 48-byte
// record + summary counters, and computeT / findJSS / findJES / computeAltEditing / NewAlignment (align.go:95-279) for a
// read whose alignment is known -- from the traceback ops (finish_read) or as the identity (finish_same).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

constexpr int kPadCode = 4;  // profile row with w = 0: a no-op row of the wavefront

// Optional bounds checks (build with -DTREAT_B200_CHECKS): compute-sanitizer is not available on the GPU pool,
// so index computations that depend on the band certificate or on n can report into a flag word instead.
#ifdef TREAT_B200_CHECKS
#define TB_CHECK(p, cond, bit)                                   \
    do {                                                         \
        if (!(cond) && (p).dbg) atomicOr((p).dbg, 1u << (bit)); \
    } while (0)
#else
#define TB_CHECK(p, cond, bit) \
    do {                       \
    } while (0)
#endif

// ------------------------------------------------------------------------------------------
// Bit rows over template sites in array order ti = 0..len-1 (site number = len-1-ti): the willf/bitset
// sets of align.go:122-181.  A row lives in registers, lane w holds bits ti = 32w .. 32w+31; queries are
// one warp reduction (REDUX) each.
// ------------------------------------------------------------------------------------------
struct RowGeom {
    int lane, len;
    uint32_t vmask;  // bits of this lane's word that are real sites
    __device__ __forceinline__ RowGeom(int lane_, int len_) : lane(lane_), len(len_)
    {
        const int rem = len - lane * 32;
        vmask = rem <= 0 ? 0u : rem >= 32 ? 0xffffffffu : (1u << rem) - 1u;
    }
    // highest ti <= limit whose bit is clear, or -1
    __device__ __forceinline__ int highest_clear_le(uint32_t row, int limit) const
    {
        uint32_t w = ~row & vmask;
        const int lo = lane * 32;
        if (lo > limit) w = 0;
        else if (limit - lo < 31) w &= (2u << (limit - lo)) - 1u;
        return __reduce_max_sync(FULL_MASK, w ? lo + 31 - __clz(w) : -1);
    }
    // lowest ti whose bit is clear, or -1
    __device__ __forceinline__ int lowest_clear(uint32_t row) const
    {
        const uint32_t w = ~row & vmask;
        const int r = __reduce_min_sync(FULL_MASK, w ? lane * 32 + __ffs(w) - 1 : 0x7fffffff);
        return r == 0x7fffffff ? -1 : r;
    }
    __device__ __forceinline__ bool any_set(uint32_t row) const { return __any_sync(FULL_MASK, (row & vmask) != 0); }
};

// ------------------------------------------------------------------------------------------
// shared-memory views
// ------------------------------------------------------------------------------------------
template <typename cnt_t>
struct CtaSmem {
    uint8_t *tb;               // [32*CPL] template bases (codes, or bytes on the wide path)
    cnt_t *tcount;             // [K][len_pad] EditSite
    int32_t *alt;              // [2*n_alt]
    uint32_t *sum_lo, *sum_hi; // summary histograms of this CTA (64-bit counters as two words)
    uint32_t sum_len;
    uint32_t *prof;            // scoring profile
    uint32_t *dump;            // [16][32] words: where lanes outside their band window park their direction words
};

struct WarpSmem {
    uint8_t *stage;    // 2 slots x reads-per-warp x (bases, counts)
    uint8_t *rb;       // read bases (or pair codes) as bytes; 32 pad bytes in front
    void *dir;         // direction words [steps][32]
    uint32_t *ops2;    // traceback ops, 2 bits each, index p = 0 .. m+n-1 (alignment = [pos, m+n))
    uint16_t *t2f;     // template site -> read base index (0xffff = deletion)
    uint64_t *bar;     // [2] mbarriers
};

template <typename cnt_t>
__device__ __forceinline__ CtaSmem<cnt_t> carve_cta(uint8_t *smem, const AlignParams &p, int cpl)
{
    CtaSmem<cnt_t> c;
    c.tb = smem;
    c.tcount = reinterpret_cast<cnt_t *>(smem + align16(32 * cpl));
    c.alt = reinterpret_cast<int32_t *>(reinterpret_cast<uint8_t *>(c.tcount) +
                                        align16((uint32_t)(p.t.K * p.t.len_pad * sizeof(cnt_t))));
    c.sum_lo = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(c.alt) + align16(8u * (p.t.n_alt ? p.t.n_alt : 1)));
    c.sum_len = p.summary ? TREAT_B200_SUMMARY_HEADER + 3u * p.bins : 0u;
    c.sum_hi = c.sum_lo + c.sum_len;
    c.prof = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(c.sum_lo) + align16(8u * c.sum_len));
    c.dump = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(c.prof) + p.sm_prof_bytes);
    return c;
}

__device__ __forceinline__ WarpSmem carve_warp(uint8_t *smem, const AlignParams &p, int warp)
{
    WarpSmem w;
    uint8_t *b = smem + p.sm_tmpl_bytes + (size_t)warp * p.sm_warp_bytes;
    w.stage = b;
    w.rb = b + p.sm_stage_total;
    w.dir = w.rb + p.sm_rb;
    w.ops2 = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(w.dir) + p.sm_dir);
    w.t2f = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(w.ops2) + p.sm_ops);
    w.bar = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(w.t2f) + p.sm_t2f);
    return w;
}

template <typename cnt_t>
__device__ __forceinline__ void load_template(const CtaSmem<cnt_t> &c, const AlignParams &p, int cpl)
{
    const int m = p.t.m;
    for (int i = threadIdx.x; i < 32 * cpl; i += blockDim.x) c.tb[i] = i < m ? p.t.tb[i] : 0;
    for (int i = threadIdx.x; i < p.t.K * p.t.len_pad; i += blockDim.x)
        c.tcount[i] = reinterpret_cast<const cnt_t *>(p.t.tcount)[i];
    for (int i = threadIdx.x; i < 2 * p.t.n_alt; i += blockDim.x) c.alt[i] = p.t.alt[i];
    for (uint32_t i = threadIdx.x; i < 2 * c.sum_len; i += blockDim.x) c.sum_lo[i] = 0u;
}

// ------------------------------------------------------------------------------------------
// the 48-byte record and the summary counters of one read (align.go:41-55, stats.go:140-187, handlers.go:203-215)
// ------------------------------------------------------------------------------------------
template <typename cnt_t>
__device__ __forceinline__ void emit_result(const AlignParams &p, const CtaSmem<cnt_t> &cs, int lane, uint64_t r, int n, int alen,
                                            uint32_t status, int score, int edit_stop, int junc_start, int junc_end, int junc_len,
                                            int has_mutation, int alt_editing, int mism, bool indel, uint32_t js_off,
                                            uint32_t js_len, unsigned long long &ctr, uint32_t raw_len_known = 0xffffffffu,
                                            uint32_t read_count_known = 0u)
{
    const int m = p.t.m;
    // read_count_known = 0 means "not loaded yet" (a true ReadCount of 0, header "x-0", just takes the load below)
    const uint32_t read_count = read_count_known ? read_count_known : p.read_count ? p.read_count[r] : 1u;

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
        o.score = score;
        o.junc_seq_off = js_off;
        o.junc_seq_len = js_len;
        o.aln_len = (uint16_t)alen;
        o.n_bases = (uint16_t)n;
        o.status = (uint8_t)status;
        o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
        o.raw_len = raw_len_known != 0xffffffffu ? raw_len_known : p.raw_len[r];  // the fused kernel has it in a register
        const uint4 *src = reinterpret_cast<const uint4 *>(&o);
        uint4 *dst = reinterpret_cast<uint4 *>(p.rec + r);
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
    }

    if (cs.sum_len) {
        // cmd/treat/stats.go:140-187: lane L accumulates counter L of the summary header in a register
        const unsigned long long rcw = read_count;
        const int cls = has_mutation ? 2 : 1;
        const int mm8 = mism & 0xff;
        const int mcls = indel ? 3 : mm8 == 1 ? 5 : mm8 == 2 ? 6 : mm8 > 2 ? 4 : 0;
        if (lane == 0 || lane == cls || (mcls && lane == mcls)) ctr += rcw;
        if (lane == 7 || lane == 7 + cls || (mcls && lane == 7 + mcls) || lane == 14) ctr += 1ull;
        if (lane == 15) ctr += (unsigned long long)m * (unsigned long long)n;
        // cmd/treat/handlers.go:203-215: three histograms, one lane each
        if (lane < 3 && !(edit_stop + p.t.edit_offset == p.t.tmpl_edit_stop && junc_len == 0)) {
            const uint32_t bin = lane == 0 ? (uint32_t)(edit_stop + 2) : lane == 1 ? (uint32_t)junc_len : (uint32_t)(junc_end + 2);
            if (bin < p.bins) {
                const uint32_t idx = TREAT_B200_SUMMARY_HEADER + (uint32_t)lane * p.bins + bin;
                const uint32_t old = atomicAdd(&cs.sum_lo[idx], read_count);
                if (old + read_count < old) atomicAdd(&cs.sum_hi[idx], 1u);
            }
        }
    }

}

// ------------------------------------------------------------------------------------------
// everything after the traceback for one read: computeT, JSS/ESS/JES, alt editing, JuncSeq window,
// record, summary counters, packed ops.  rbf(fi) returns the read base (code or byte) at index fi.
// ------------------------------------------------------------------------------------------
template <typename cnt_t, typename RbFn>
__device__ __forceinline__ void finish_read(const AlignParams &p, const CtaSmem<cnt_t> &cs, const WarpSmem &ws, int lane,
                                            uint64_t r, int n, uint32_t rflags, int score, int pos, int ngaps,
                                            const cnt_t *rcount, RbFn rbf, unsigned long long &ctr, int mism_known = -1,
                                            uint32_t raw_len_known = 0xffffffffu)
{
    const int m = p.t.m, len = p.t.len, len_pad = p.t.len_pad, K = p.t.K, n_alt = p.t.n_alt;
    const int tw = len_pad >> 5;
    const int alen = m + n - pos;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t *ops2 = ws.ops2;
    TB_CHECK(p, pos >= 0 && alen >= max(m, n) && alen <= m + n, 2);
    TB_CHECK(p, ((uint32_t)(m + n + 15) / 16 + 2) * 4 <= p.sm_ops, 3);
    TB_CHECK(p, n <= (int)p.ncap, 4);

    // ---- computeT part 1 (align.go:131-171): map template sites to read bases, count mismatches ----
    int mism = 0;
    const bool indel = ngaps != 0;
    if (!indel && mism_known >= 0) {
        mism = mism_known;  // the caller counted them (k_pack's one-mismatch mark)
    } else if (!indel) {
        // no gap in the alignment: column c pairs template base c with read base c
        if (score != m) {  // score == m == n: every base matches, nothing to count
            for (int c0 = 0; c0 < m; c0 += 32) {
                const int c = c0 + lane;
                const bool mm = c < m && rbf(c) != (int)cs.tb[c];
                mism += __popc(__ballot_sync(FULL_MASK, mm));
            }
        }
    } else {
        int ti_base = 0, fi_base = 0;
        for (int c0 = 0; c0 < alen; c0 += 32) {
            const int c = c0 + lane, pp = pos + c;
            const int op = c < alen ? (int)((ops2[pp >> 4] >> (2 * (pp & 15))) & 3u) : 3;
            const bool cT = op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_DEL;
            const bool cR = op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_INS;
            const uint32_t mT = __ballot_sync(FULL_MASK, cT), mR = __ballot_sync(FULL_MASK, cR);
            const int ti = ti_base + __popc(mT & lt), fi = fi_base + __popc(mR & lt);
            TB_CHECK(p, !cT || (ti >= 0 && ti < m), 5);
            TB_CHECK(p, op != TREAT_B200_OP_MATCH || (fi >= 0 && fi < n), 6);
            if (cT) ws.t2f[ti] = op == TREAT_B200_OP_MATCH ? (uint16_t)fi : (uint16_t)0xffff;
            const bool mm = op == TREAT_B200_OP_MATCH && rbf(fi) != (int)cs.tb[ti];
            mism += __popc(__ballot_sync(FULL_MASK, mm));
            ti_base += __popc(mT);
            fi_base += __popc(mR);
        }
        if (lane == 0) ws.t2f[m] = (uint16_t)n;  // last edit site (align.go:173-178)
        __syncwarp();
    }

    // ---- computeT part 2: edit-base count of the read at every template site vs the templates.
    // Only the FE and PE rows are always needed; an alternative template's row is built when the
    // junction start sits on the first site of its region (align.go:186-205).
    auto site_count = [&](int ti) -> uint32_t {  // the read's count aligned to template site ti (0 for a deletion)
        const uint32_t f = indel ? (uint32_t)ws.t2f[ti] : (uint32_t)ti;
        return f == 0xffffu ? 0u : (uint32_t)rcount[f];
    };
    auto build_row = [&](int k) -> uint32_t {  // lane w keeps bits ti = 32w..32w+31 of template k's match row
        uint32_t row = 0;
        for (int wv = 0; wv < tw; wv++) {
            const int ti = wv * 32 + lane;
            const bool match = ti < len && (uint32_t)cs.tcount[k * len_pad + ti] == site_count(ti);
            const uint32_t word = __ballot_sync(FULL_MASK, match);
            if (lane == wv) row = word;
        }
        return row;
    };
    uint32_t t0 = 0, t1 = 0;
    for (int wv = 0; wv < tw; wv++) {
        const int ti = wv * 32 + lane;
        const bool valid = ti < len;
        const uint32_t count = valid ? site_count(ti) : 0u;
        const uint32_t w0 = __ballot_sync(FULL_MASK, valid && (uint32_t)cs.tcount[ti] == count);
        const uint32_t w1 = __ballot_sync(FULL_MASK, valid && (uint32_t)cs.tcount[len_pad + ti] == count);
        if (lane == wv) {
            t0 = w0;
            t1 = w1;
        }
    }

    // ---- findJSS / computeAltEditing / findJES / NewAlignment (align.go:95-120,183-279) -----------
    const RowGeom G(lane, len);
    int has_mutation = 0;
    if (indel) has_mutation = 1;
    if (((p.flags & TREAT_B200_EXCLUDE_SNPS) && mism >= 1) || mism >= 3) has_mutation = 1;

    const int hc0 = G.highest_clear_le(t0, len - 1);
    const bool t0_all = hc0 < 0;
    int junc_start = t0_all ? len - 1 : (!G.any_set(t0) ? -1 : len - 1 - hc0);
    int alt_editing = 0;
    if (K > 2 && junc_start >= 0 && junc_start < len) {  // bitset.Test is false out of range
        int shift = junc_start, alt = -1;
        const int ti_s = len - 1 - shift;
        const uint32_t cstar = site_count(ti_s);
        for (int k = 2; k < K; k++) {
            if ((uint32_t)cs.tcount[k * len_pad + ti_s] == cstar) {
                alt = k - 2;
                break;
            }
        }
        if (alt >= 0 && shift == cs.alt[alt]) {
            const uint32_t ta = build_row(alt + 2);
            const int hc = G.highest_clear_le(ta, len - 1 - shift);
            const bool full_match = hc < 0;
            if (!full_match) shift = len - 1 - hc;
            if (full_match || shift >= cs.alt[n_alt + alt]) {
                alt_editing = alt + 1;
                if (full_match) {
                    junc_start = len - 1;
                } else {
                    const int h0 = G.highest_clear_le(t0, len - 1 - shift);
                    if (h0 >= 0) junc_start = len - 1 - h0;
                }
            }
        }
    }
    const int lc1 = G.lowest_clear(t1);
    int junc_end = lc1 < 0 ? -1 : len - 1 - lc1;
    int edit_stop = junc_start - 1;
    int junc_len = 0;
    uint32_t js_off = 0, js_len = 0;
    uint32_t status = rflags & ~kFlagsInternal;
    if (junc_end > edit_stop) {
        junc_len = junc_end - edit_stop;
        if (has_mutation == 0 && !t0_all) {
            const int from = (len - 1) - junc_end, to = (len - 1) - edit_stop;
            if (to > n + 1) {
                status |= TREAT_B200_ST_REF_PANIC;  // Go: index out of range on frag.EditSite
            } else {
                // JuncSeq = EditBase x EditSite[i] + Bases[i] for i in [from, to): a window of the
                // upper-cased raw read starting at the edit-base run in front of base `from`
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
    emit_result<cnt_t>(p, cs, lane, r, n, alen, status, score, edit_stop, junc_start, junc_end, junc_len, has_mutation, alt_editing,
                       mism, indel, js_off, js_len, ctr, raw_len_known);

    // ---- traceback ops, 2 bits per column, in alignment order: a funnel shift of ops2 by pos ------
    if (p.ops && (indel || !p.ops_gapped_only)) {
        uint64_t row = r;
        if (p.ops_count && indel) {  // count the gapped reads; with ops_idx their rows are compacted: claim the next one
            uint32_t k = 0;
            if (lane == 0) {
                k = atomicAdd(p.ops_count, 1u);
                if (p.ops_idx) p.ops_idx[k] = (uint32_t)r;
            }
            if (p.ops_idx) row = __shfl_sync(FULL_MASK, k, 0);
        }
        uint32_t *go = reinterpret_cast<uint32_t *>(p.ops + row * (uint64_t)p.ops_stride);
        const int nwords = (int)(p.ops_stride / 4);
        const int w0 = pos >> 4, sh = 2 * (pos & 15);
        const int avail = (int)(p.sm_ops / 4) - 1;  // ops2 words that exist (all zero past the alignment)
        for (int wi = lane; wi < nwords; wi += 32) {
            uint32_t w = 0;
            if (wi * 16 < alen) {
                const int si = w0 + wi;
                const uint32_t lo = ops2[si], hi = si + 1 <= avail ? ops2[si + 1] : 0u;
                w = __funnelshift_r(lo, hi, sh);
            }
            go[wi] = w;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// The same assembly for a read whose bases equal the template's (score m, no gap, no mismatch: read site ti is
// template site ti), on bytes: lane l holds the 16 edit-site counts [16 l, 16 l + 16) of the read and of the FE / PE
// rows in registers; "highest site where FE differs" / "lowest site where PE differs" are a clz / ffs on the XOR
// words plus one REDUX each.  Returns false (nothing written) when the junction start sits on the first site of an
// alternative template's region (align.go:186-205): those few reads take finish_read.
// ------------------------------------------------------------------------------------------
struct SameLane {
    uint32_t fe[4], pe[4], valid[4];  // template rows 0 / 1 and the byte mask of real sites, this lane's 16 sites
};

__device__ __forceinline__ SameLane load_same_lane(const CtaSmem<uint8_t> &cs, const AlignParams &p, int lane)
{
    SameLane sl;
    const int base = 16 * lane;
    uint4 f = make_uint4(0, 0, 0, 0), q = f;
    if (base < p.t.len_pad) {
        f = *reinterpret_cast<const uint4 *>(cs.tcount + base);
        q = *reinterpret_cast<const uint4 *>(cs.tcount + p.t.len_pad + base);
    }
    sl.fe[0] = f.x, sl.fe[1] = f.y, sl.fe[2] = f.z, sl.fe[3] = f.w;
    sl.pe[0] = q.x, sl.pe[1] = q.y, sl.pe[2] = q.z, sl.pe[3] = q.w;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int live = min(max(p.t.len - base - 4 * i, 0), 4);
        sl.valid[i] = live >= 4 ? 0xffffffffu : (1u << (8 * live)) - 1u;
    }
    return sl;
}

// G lanes work on one read (G = 16: two reads per warp, templates of up to 255 bases); gl = lane within the group,
// gmask = the group's lanes, lead = its first lane.  s_mask[t] = 16 bytes with the t lowest set (t = 0..16).
template <int G>
__device__ __forceinline__ bool finish_same(const AlignParams &p, const CtaSmem<uint8_t> &cs, const uint4 *s_mask, int gl,
                                            unsigned gmask, int lead, uint64_t r, uint32_t rflags, const uint4 &c,
                                            const SameLane &sl, unsigned long long &ctr, uint32_t raw_len_known = 0xffffffffu,
                                            int mism_known = -1, uint32_t read_count_known = 0u)
{
    // no gap; no mismatch, or exactly one (kFlagOneMismatch: the diagonal still is the unique optimum, score m - 2), or
    // as many as the caller's traceback found on the diagonal (mism_known, k_band_finish)
    const int mism = mism_known >= 0 ? mism_known : (rflags & kFlagOneMismatch) ? 1 : 0;
    const int has_mutation = (((p.flags & TREAT_B200_EXCLUDE_SNPS) && mism >= 1) || mism >= 3) ? 1 : 0;  // align.go:240-241
    const int m = p.t.m, n = m, len = p.t.len, K = p.t.K, base = 16 * gl;
    const uint32_t cw[4] = {c.x, c.y, c.z, c.w};
    uint32_t df[4], dp[4];  // non-zero byte: the read's count differs from FE / PE at that site
#pragma unroll
    for (int i = 0; i < 4; i++) {
        df[i] = (cw[i] ^ sl.fe[i]) & sl.valid[i];
        dp[i] = (cw[i] ^ sl.pe[i]) & sl.valid[i];
    }
    // findJSS (align.go:183-205): the highest array index whose FE bit is clear
    int hl = -1, ll = 0x7fffffff;
    bool fe_match = false;  // some real site where the read agrees with FE
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (df[i]) hl = base + 4 * i + ((31 - __clz(df[i])) >> 3);
        const uint32_t x = df[i] | ~sl.valid[i];
        fe_match |= ((x - 0x01010101u) & ~x & 0x80808080u) != 0u;
    }
#pragma unroll
    for (int i = 3; i >= 0; i--)
        if (dp[i]) ll = base + 4 * i + ((__ffs(dp[i]) - 1) >> 3);
    const int hc0 = __reduce_max_sync(gmask, hl);
    const bool t0_all = hc0 < 0;
    int junc_start = t0_all ? len - 1 : (!__any_sync(gmask, fe_match) ? -1 : len - 1 - hc0);
    if (K > 2 && junc_start >= 0 && junc_start < len) {
        // does the junction start sit on the first site of an alternative template's region that matches there?
        const int ti_s = len - 1 - junc_start, q = (ti_s >> 2) & 3;
        const uint32_t wsel = q == 0 ? cw[0] : q == 1 ? cw[1] : q == 2 ? cw[2] : cw[3];
        const uint32_t cstar = (__shfl_sync(gmask, wsel, lead + (ti_s >> 4)) >> (8 * (ti_s & 3))) & 0xffu;
        for (int k = 2; k < K; k++) {
            if ((uint32_t)cs.tcount[k * p.t.len_pad + ti_s] == cstar) {
                if (junc_start == cs.alt[k - 2]) return false;
                break;
            }
        }
    }
    const int lc1r = __reduce_min_sync(gmask, ll);
    int junc_end = lc1r == 0x7fffffff ? -1 : len - 1 - lc1r;
    int edit_stop = junc_start - 1, junc_len = 0;
    uint32_t js_off = 0, js_len = 0, status = rflags & ~kFlagsInternal;
    if (junc_end > edit_stop) {
        junc_len = junc_end - edit_stop;
        if (has_mutation == 0 && !t0_all) {
            const int from = (len - 1) - junc_end, to = (len - 1) - edit_stop;
            if (to > n + 1) {
                status |= TREAT_B200_ST_REF_PANIC;
            } else {
                // edit bases in front of base `from` / `to`: byte sums over the sites below them
                const uint4 mf = s_mask[min(max(from - base, 0), 16)], mt = s_mask[min(max(to - base, 0), 16)];
                uint32_t sf = __dp4a(cw[0] & mf.x, 0x01010101u, 0u), st = __dp4a(cw[0] & mt.x, 0x01010101u, 0u);
                sf = __dp4a(cw[1] & mf.y, 0x01010101u, sf), st = __dp4a(cw[1] & mt.y, 0x01010101u, st);
                sf = __dp4a(cw[2] & mf.z, 0x01010101u, sf), st = __dp4a(cw[2] & mt.z, 0x01010101u, st);
                sf = __dp4a(cw[3] & mf.w, 0x01010101u, sf), st = __dp4a(cw[3] & mt.w, 0x01010101u, st);
                sf = __reduce_add_sync(gmask, sf);
                st = __reduce_add_sync(gmask, st);
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
    emit_result<uint8_t>(p, cs, gl, r, n, /*alen*/ m, status, /*score*/ m - 2 * mism, edit_stop, junc_start, junc_end, junc_len,
                         has_mutation, /*alt_editing*/ 0, mism, /*indel*/ false, js_off, js_len, ctr, raw_len_known, read_count_known);
    if (p.ops && !p.ops_gapped_only) {  // m match columns: an all-zero row
        uint32_t *go = reinterpret_cast<uint32_t *>(p.ops + r * (uint64_t)p.ops_stride);
        for (int wi = gl; wi < (int)(p.ops_stride / 4); wi += G) go[wi] = 0u;
    }
    return true;
}

__device__ __forceinline__ void flush_summary(const AlignParams &p, uint32_t sum_len, const uint32_t *lo, const uint32_t *hi,
                                              int lane, unsigned long long ctr)
{
    if (!sum_len) return;
    unsigned long long *g = reinterpret_cast<unsigned long long *>(p.summary);
    if (lane < TREAT_B200_SUMMARY_HEADER && ctr) atomicAdd(&g[lane], ctr);
    __syncthreads();
    for (uint32_t i = TREAT_B200_SUMMARY_HEADER + threadIdx.x; i < sum_len; i += blockDim.x) {
        const unsigned long long v = (unsigned long long)lo[i] | ((unsigned long long)hi[i] << 32);
        if (v) atomicAdd(&g[i], v);
    }
}


// the general assembly for the few same-bases reads that touch an alternative template's region (kept out of line:
// it would otherwise set the register count of the whole kernel)
static __device__ __noinline__ void finish_same_general(const AlignParams &p, const CtaSmem<uint8_t> &cs, const WarpSmem &ws, int lane,
                                                 uint64_t r, uint32_t rflags, const uint8_t *rcount, unsigned long long &ctr,
                                                 uint32_t raw_len_known = 0xffffffffu, int mism_known = -1)
{
    const int m = p.t.m, mism = mism_known >= 0 ? mism_known : (rflags & kFlagOneMismatch) ? 1 : 0;  // the diagonal: score m - 2 per mismatch, no gap
    finish_read<uint8_t>(p, cs, ws, lane, r, m, rflags, /*score*/ m - 2 * mism, /*pos*/ m, /*ngaps*/ 0, rcount,
                         [](int) { return 0; }, ctr, mism, raw_len_known);
}

}  // namespace treat_b200
