// text_kernel.cu -- Alignment.WriteTo (align.go:388-520) for a batch on the device (sm_100a): the text block `treat align`
// prints per read -- header (name, JSS / ESS / JES / Junc Len / Alt) and the gapped FE / PE / A* / RD rows wrapped at
// tw - 4 columns -- from the raw read, the 48-byte record and the 2-bit traceback ops.  No second DP (the reference
// re-runs nwalgo.Align inside WriteTo, align.go:393): the ops are the alignment.
//
// One warp per read, two passes over the same code:
//   pass 1 (WRITE = false): NewFragment on the warp (bases + edit-base run lengths into shared memory), the width of
//       every alignment column (1 + max edit-base count over the rows that print one), hence the row length G and the
//       text length; an exclusive scan over the reads turns the lengths into offsets.
//   pass 2: the same walk, lane l owns column c0 + l and a warp scan of the widths gives its position x in the rows.
//       Character p of a column is '-' while p < max - count, the edit base while p < max, then the base, and the byte
//       address of (row k, position x) is arithmetic: chunk q = x / cols, all chunks but the last are full, rows inside a
//       chunk are label + ": " + chars + "\n".
//       MODE 2 (the normal case): the walk only records every column's position and (ti, fi, op) in shared memory plus
//       an inverse map position -> column; then lane l takes row position xi + l, looks its column up once and writes
//       that position of every row: 32 consecutive bytes of a line per store instruction.
//       MODE 1 (reads too long for the shared-memory tables): each lane writes its own column of every row directly.
// Integer / byte work, HBM (in practice PCIe D2H) bound: about 2.8 KB of text per RPS12 read.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

constexpr int kTextWarps = 8;
constexpr int kTextMaxRows = kMaxTemplates + 1;  // K template rows + RD

// decimal digits of v (with sign)
__device__ __forceinline__ uint32_t dec_len(int v)
{
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v, d = 1;
    while (u >= 10u) {
        u /= 10u;
        d++;
    }
    return d + (v < 0 ? 1u : 0u);
}
__device__ __forceinline__ uint8_t *put_dec(uint8_t *o, int v)
{
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    if (v < 0) *o++ = '-';
    char tmp[10];
    int d = 0;
    do {
        tmp[d++] = (char)('0' + u % 10u);
        u /= 10u;
    } while (u);
    while (d) *o++ = (uint8_t)tmp[--d];
    return o;
}
__device__ __forceinline__ uint8_t *put_str(uint8_t *o, const char *s)
{
    while (*s) *o++ = (uint8_t)*s++;
    return o;
}

// MODE 0: sizes only; 1: write, column-wise; 2: write, position-wise through shared-memory column tables
template <int MODE>
__global__ void __launch_bounds__(kTextWarps * 32) k_text(const __grid_constant__ TextParams p)
{
    constexpr bool WRITE = MODE != 0;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = p.m, len = p.len, K = p.K, R = K + 1;
    const uint32_t lt = (1u << lane) - 1u;
    // CTA-wide: label geometry, template bases, Max(i) (template.go:186), and the EditSite rows when they fit
    uint32_t *s_labsum = reinterpret_cast<uint32_t *>(smem);             // [R + 1] sum over rows j < k of (label_j + 3)
    uint32_t *s_pre = s_labsum + (kTextMaxRows + 1);                     // [R] where row k's characters start inside its line block
    uint32_t *s_tmax = s_pre + kTextMaxRows;                             // [len]
    uint32_t *s_tc = s_tmax + p.len_pad;                                 // [K][len_pad] (tmpl_in_smem)
    uint8_t *s_lab = reinterpret_cast<uint8_t *>(s_tc + (p.tmpl_in_smem ? (size_t)K * p.len_pad : 0));  // [R] label length
    uint8_t *s_tb = s_lab + 48;                                          // [m]
    uint8_t *wbase = s_tb + align16((uint32_t)m) + (size_t)warp * p.warp_bytes;
    uint8_t *sb = wbase;                                                  // [cap] read bases, upper-cased
    uint32_t *sc = reinterpret_cast<uint32_t *>(wbase + p.stage_cap);     // [cap] Fragment.EditSite
    // MODE 2: per alignment column its position in the rows and (ti | op << 11 | fi << 13); per row position its column
    uint32_t *cx = sc + p.stage_cap, *ci = cx + p.col_cap;
    uint16_t *cmap = reinterpret_cast<uint16_t *>(ci + p.col_cap);
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < R; k++) {
            const uint32_t l = (k >= 2 && k < K && k - 1 >= 10) ? 3u : 2u;  // FE, PE, A1.., RD
            s_lab[k] = (uint8_t)l;
            s_labsum[k] = acc;
            s_pre[k] = acc + l + 2u;
            acc += l + 3u;
        }
        s_labsum[R] = acc;
    }
    for (int i = threadIdx.x; i < len; i += blockDim.x) s_tmax[i] = p.tmax[i];
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_tb[i] = p.tb[i];
    if (p.tmpl_in_smem)
        for (int i = threadIdx.x; i < K * p.len_pad; i += blockDim.x) s_tc[i] = p.tc32[i];
    __syncthreads();
    const uint32_t *tc = p.tmpl_in_smem ? s_tc : p.tc32;
    const uint32_t cols = (uint32_t)(p.tw - 4), tw = (uint32_t)p.tw;
    const uint32_t block_full = s_labsum[R] + (uint32_t)R * cols + 1u;
    const uint8_t eb = p.edit_base;

    const uint64_t nwarps = blockDim.x >> 5;  // kTextWarps, fewer when very long reads need the shared memory
    const uint64_t gw = (uint64_t)blockIdx.x * nwarps + warp, nw = (uint64_t)gridDim.x * nwarps;
    uint32_t gmax = 0;  // MODE 0: longest row of this warp's reads
    for (uint64_t r = gw; r < p.N; r += nw) {
        // ---- NewFragment (fragment.go:91-121) on the warp: 32 characters per step ------------------------
        const uint64_t b = p.begin ? p.begin[r] : p.off[r] - p.off_bias;
        const uint32_t L = p.len_ ? p.len_[r] : (uint32_t)(p.off[r + 1] - p.off[r]);
        const uint8_t *seq = p.seqs + b;
        uint32_t nb = 0;
        int lastpos = -1;
        for (uint32_t c0 = 0; c0 < L; c0 += 32) {
            const uint32_t j = c0 + lane;
            uint32_t v = j < L ? seq[j] : eb;
            if (v >= 'a' && v <= 'z') v -= 32u;
            const bool isb = j < L && v != eb;
            const uint32_t mask = __ballot_sync(FULL_MASK, isb);
            if (isb) {
                const uint32_t below = mask & lt;
                const uint32_t idx = nb + __popc(below);
                const int prev = below ? (int)(c0 + 31 - __clz(below)) : lastpos;
                sb[idx] = (uint8_t)v;
                sc[idx] = (uint32_t)((int)j - prev - 1);
            }
            if (mask) lastpos = (int)(c0 + 31 - __clz(mask));
            nb += __popc(mask);
        }
        if (lane == 0) sc[nb] = (uint32_t)((int)L - 1 - lastpos);
        __syncwarp();
        const int n = (int)nb;

        // ---- the record's header fields ----------------------------------------------------------------
        const treat_b200_record *rc = p.rec + r;
        const int alen = (int)rc->aln_len;
        const int jss = rc->junc_start, ess = rc->edit_stop, jes = rc->junc_end, jl = rc->junc_len, alt = rc->alt_editing;
        const uint32_t nlen = p.name_len ? p.name_len[r] : (uint32_t)(p.name_off[r + 1] - p.name_off[r]);
        const uint32_t numlen = 5u + dec_len(jss) + 1u + 5u + dec_len(ess) + 1u + 5u + dec_len(jes) + 1u + 10u + dec_len(jl) + 1u +
                                (alt > 0 ? 5u + dec_len(alt) + 1u : 0u);
        const uint32_t hdr = (tw + 1u) + (nlen + 1u) + (tw + 1u) + numlen + (tw + 2u);

        uint32_t G = 0, nch = 0, ll = 0;
        uint8_t *out = nullptr;
        if (WRITE) {
            G = p.gtot[r];
            nch = (G + cols - 1u) / cols;
            ll = G - (nch - 1u) * cols;
            out = p.text + p.text_off[r];
        }
        const uint32_t *orow = reinterpret_cast<const uint32_t *>(p.ops + r * (uint64_t)p.ops_stride);

        // ---- walk the alignment columns (align.go:402-448); column alen is the last edit site (pads only) --
        uint32_t xbase = 0;
        int ti_base = 0, fi_base = 0;
        for (int c0 = 0; c0 <= alen; c0 += 32) {
            const int c = c0 + lane;
            const bool real = c < alen, valid = c <= alen;
            const uint32_t op = real ? (orow[c >> 4] >> (2 * (c & 15))) & 3u : 3u;
            const bool cT = real && (op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_DEL);
            const bool cR = real && (op == TREAT_B200_OP_MATCH || op == TREAT_B200_OP_INS);
            const uint32_t mT = __ballot_sync(FULL_MASK, cT), mR = __ballot_sync(FULL_MASK, cR);
            int ti = ti_base + __popc(mT & lt), fi = fi_base + __popc(mR & lt);
            ti_base += __popc(mT);
            fi_base += __popc(mR);
            ti = min(ti, m);  // (a malformed ops row must not index out of range)
            fi = min(fi, n);
            const bool ins = real && op == TREAT_B200_OP_INS, del = real && op == TREAT_B200_OP_DEL;
            const uint32_t cnt = (valid && !del) ? sc[fi] : 0u;
            const uint32_t tmx = (valid && !ins) ? s_tmax[ti] : 0u;
            const uint32_t mx = ins ? cnt : del ? tmx : max(tmx, cnt);
            const uint32_t width = !valid ? 0u : real ? mx + 1u : mx;
            uint32_t incl = width;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t xc = xbase + incl - width;
            xbase += __shfl_sync(FULL_MASK, incl, 31);
            if (MODE == 2 && valid) {
                cx[c] = xc;
                ci[c] = (uint32_t)ti | (op << 11) | ((uint32_t)fi << 13);
                for (uint32_t pp = 0; pp < width; pp++) cmap[xc + pp] = (uint16_t)c;
            }
            if (MODE == 1 && width) {
                const uint32_t q0 = xc / cols, r0 = xc - q0 * cols;
                const uint8_t tbase = ti < m ? s_tb[ti] : (uint8_t)'?', rbase = fi < n ? sb[fi] : (uint8_t)'?';
                for (int k = 0; k < R; k++) {
                    const bool rd = k == K;
                    const bool dashes = rd ? del : ins;
                    const uint32_t e = rd ? cnt : tc[(size_t)k * p.len_pad + ti];
                    const uint32_t ndash = dashes ? width : mx - min(e, mx);
                    const uint8_t bch = rd ? rbase : tbase;
                    const uint32_t pre = s_labsum[k] + s_lab[k] + 2u;
                    uint32_t q = q0, rr = r0;
                    uint8_t *line = out + hdr + (size_t)q * block_full + pre + (size_t)k * (q == nch - 1u ? ll : cols);
                    for (uint32_t pp = 0; pp < width; pp++) {
                        line[rr] = pp < ndash ? (uint8_t)'-' : pp < mx ? eb : bch;
                        if (++rr == cols) {
                            rr = 0;
                            q++;
                            line = out + hdr + (size_t)q * block_full + pre + (size_t)k * (q == nch - 1u ? ll : cols);
                        }
                    }
                }
            }
        }
        if (!WRITE) {
            if (lane == 0) {
                const uint32_t g = xbase;
                const uint32_t q = (g + cols - 1u) / cols, l2 = g - (q - 1u) * cols;
                p.gtot[r] = g;
                p.tlen[r] = hdr + (q - 1u) * block_full + s_labsum[R] + (uint32_t)R * l2 + 1u;
            }
            gmax = max(gmax, xbase);
            __syncwarp();
            continue;
        }
        if (MODE == 2) {
            // ---- 32 row positions per step: one column look-up per lane, then that position of every row ----
            __syncwarp();
            for (uint32_t xi = 0; xi < G; xi += 32) {
                const uint32_t x = xi + lane;
                if (x < G) {
                    const uint32_t c = cmap[x], pp = x - cx[c], info = ci[c];
                    const int ti = (int)(info & 0x7ffu), fi = (int)(info >> 13);
                    const uint32_t op = (info >> 11) & 3u;
                    const bool ins = op == TREAT_B200_OP_INS, del = op == TREAT_B200_OP_DEL;  // op 3: the last edit site
                    const uint32_t cnt = del ? 0u : sc[fi];
                    const uint32_t tmx = ins ? 0u : s_tmax[ti];
                    const uint32_t mx = ins ? cnt : del ? tmx : max(tmx, cnt);
                    const uint8_t tbase = ti < m ? s_tb[ti] : (uint8_t)'?', rbase = fi < n ? sb[fi] : (uint8_t)'?';
                    const uint32_t q = x / cols, rr = x - q * cols, lq = q == nch - 1u ? ll : cols;
                    uint8_t *at = out + hdr + (size_t)q * block_full + rr;
                    // template rows: '-' x (max - count), edit base x count, base; all '-' for an insertion
                    const uint32_t *tcol = tc + ti;
                    const uint32_t lim = ins ? 0xffffffffu : mx;  // pp < lim: pad or edit base; pp == mx: the base
                    for (int k = 0; k < K; k++) {
                        const uint32_t e = *tcol;
                        tcol += p.len_pad;
                        const uint32_t ndash = ins ? 0xffffffffu : mx - min(e, mx);
                        at[s_pre[k]] = pp < ndash ? (uint8_t)'-' : pp < lim ? eb : tbase;
                        at += lq;
                    }
                    // RD row: all '-' for a deletion
                    at[s_pre[K]] = (del || pp < mx - cnt) ? (uint8_t)'-' : pp < mx ? eb : rbase;
                }
            }
        }
        // ---- header (align.go:465-491) -------------------------------------------------------------------
        {
            const uint32_t o2 = tw + 1u + nlen + 1u, o3 = o2 + tw + 1u + numlen;
            for (uint32_t i = lane; i < tw; i += 32) {
                out[i] = '=';
                out[o2 + i] = '=';
                out[o3 + i] = '=';
            }
            const uint8_t *nm = p.names + p.name_off[r];
            for (uint32_t i = lane; i < nlen; i += 32) out[tw + 1u + i] = nm[i];
            if (lane == 0) {
                out[tw] = '\n';
                out[tw + 1u + nlen] = '\n';
                out[o2 + tw] = '\n';
                out[o3 + tw] = '\n';
                out[o3 + tw + 1u] = '\n';
            }
            if (lane == 1) {
                uint8_t *w = out + o2 + tw + 1u;
                w = put_dec(put_str(w, "JSS: "), jss);
                *w++ = '\n';
                w = put_dec(put_str(w, "ESS: "), ess);
                *w++ = '\n';
                w = put_dec(put_str(w, "JES: "), jes);
                *w++ = '\n';
                w = put_dec(put_str(w, "Junc Len: "), jl);
                *w++ = '\n';
                if (alt > 0) {
                    w = put_dec(put_str(w, "Alt: "), alt);
                    *w++ = '\n';
                }
            }
        }
        // ---- labels, line ends and the blank line behind every chunk (align.go:493-517) ------------------
        for (uint32_t i = lane; i < nch * (uint32_t)R; i += 32) {
            const uint32_t q = i / (uint32_t)R, k = i - q * (uint32_t)R;
            const uint32_t lq = q == nch - 1u ? ll : cols;
            uint8_t *line = out + hdr + (size_t)q * block_full + s_labsum[k] + (size_t)k * lq;
            if (k == 0) {
                line[0] = 'F';
                line[1] = 'E';
            } else if (k == 1) {
                line[0] = 'P';
                line[1] = 'E';
            } else if (k == (uint32_t)K) {
                line[0] = 'R';
                line[1] = 'D';
            } else {
                const uint32_t a = k - 1u;
                line[0] = 'A';
                if (a >= 10u) {
                    line[1] = (uint8_t)('0' + a / 10u);
                    line[2] = (uint8_t)('0' + a % 10u);
                } else {
                    line[1] = (uint8_t)('0' + a);
                }
            }
            const uint32_t l = s_lab[k];
            line[l] = ':';
            line[l + 1u] = ' ';
            line[l + 2u + lq] = '\n';
            if (k == (uint32_t)K) line[l + 2u + lq + 1u] = '\n';
        }
        __syncwarp();
    }
    if (MODE == 0 && lane == 0 && gmax && p.gmax) atomicMax(p.gmax, gmax);
}

// ------------------------------------------------------------------------------------------
// Alignment.MarshalBinary (align.go:316-335) for a batch: the blob `treat load` stores per read (cmd/treat/storage.go:
// 758-770) -- 36-byte big-endian header (EditStop, JuncStart, JuncEnd, JuncLen, ReadCount, HasMutation, AltEditing,
// Mismatches, Indel, Norm = 0.0, len(JuncSeq)) + JuncSeq, the upper-cased window of the raw read the record points at.
// Sizes pass (36 + junc_seq_len), exclusive scan, write pass: eight lanes per read, one big-endian word each for the
// header, then the JuncSeq bytes in steps of eight.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_marshal_sizes(const treat_b200_record *rec, uint64_t N, uint32_t *blen)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x)
        blen[i] = 36u + rec[i].junc_seq_len;
}

__global__ void __launch_bounds__(256) k_marshal_write(const MarshalParams p)
{
    const uint32_t sub = threadIdx.x & 7u;  // lane within the group of eight that owns a read
    const uint64_t groups = ((uint64_t)gridDim.x * blockDim.x) >> 3;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < p.N; r += groups) {
        const treat_b200_record *rc = p.rec + r;
        uint8_t *out = p.blob + p.blob_off[r];
        const uint32_t jl = rc->junc_seq_len;
        // header words 0..8: four int32, ReadCount, the four flag bytes (already in MarshalBinary order), Norm (two zero
        // words), len(JuncSeq); lane `sub` writes word `sub`, lane 0 also word 8
        const uint32_t *w = reinterpret_cast<const uint32_t *>(rc);  // record bytes 0-23 = MarshalBinary field order
        auto put_be = [&](uint32_t idx, uint32_t v) {
            uint8_t *o = out + 4u * idx;
            o[0] = (uint8_t)(v >> 24);
            o[1] = (uint8_t)(v >> 16);
            o[2] = (uint8_t)(v >> 8);
            o[3] = (uint8_t)v;
        };
        if (sub < 5u) put_be(sub, w[sub]);
        else if (sub == 5u) {  // HasMutation, AltEditing, Mismatches, Indel: single bytes, stored as they are
            const uint32_t f = w[5];
            out[20] = (uint8_t)f;
            out[21] = (uint8_t)(f >> 8);
            out[22] = (uint8_t)(f >> 16);
            out[23] = (uint8_t)(f >> 24);
        } else put_be(sub, 0u);  // words 6, 7: math.Float64bits(0.0)
        if (sub == 0u) put_be(8u, jl);
        const uint64_t b = p.begin ? p.begin[r] : p.off[r] - p.off_bias;
        const uint8_t *src = p.seqs + b + rc->junc_seq_off;
        for (uint32_t k = sub; k < jl; k += 8u) {
            uint32_t v = src[k];
            if (v >= 'a' && v <= 'z') v -= 32u;
            out[36u + k] = (uint8_t)v;
        }
    }
}

int launch_marshal_sizes(const treat_b200_record *d_rec, uint64_t N, uint32_t *d_blen, int num_sms, void *stream)
{
    if (!N) return 0;
    uint64_t blocks = (N + 255) / 256;
    if (blocks > (uint64_t)num_sms * 8) blocks = (uint64_t)num_sms * 8;
    k_marshal_sizes<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_rec, N, d_blen);
    return (int)cudaGetLastError();
}

int launch_marshal_write(const MarshalParams &p, int num_sms, void *stream)
{
    if (!p.N) return 0;
    uint64_t blocks = (p.N * 8 + 255) / 256;
    if (blocks > (uint64_t)num_sms * 8) blocks = (uint64_t)num_sms * 8;
    k_marshal_write<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

size_t text_smem_bytes(const TextParams &p, int warps)
{
    return (size_t)(2 * kTextMaxRows + 1) * 4 + (size_t)p.len_pad * 4 + (p.tmpl_in_smem ? (size_t)p.K * p.len_pad * 4 : 0) + 48 +
           align16((uint32_t)p.m) + (size_t)warps * p.warp_bytes;
}

// p.stage_cap, p.warp_bytes, p.col_cap and p.tmpl_in_smem are set here.  write: p.g_longest is the longest row (pass 1's
// gmax): the position-wise writer is used when its shared-memory tables fit, else the column-wise one.
int launch_text(TextParams &p, bool write, int num_sms, int max_smem_optin, void *stream)
{
    if (p.N == 0) return 0;
    p.stage_cap = align16(p.max_len + 1u);
    p.tmpl_in_smem = (size_t)p.K * p.len_pad * 4 <= 64u * 1024u ? 1u : 0u;
    p.col_cap = 0;
    p.warp_bytes = 5u * p.stage_cap;
    int mode = write ? 1 : 0;
    if (write && p.g_longest) {
        TextParams q = p;
        q.col_cap = align16((uint32_t)p.m + p.max_len + 2u);
        q.warp_bytes = 5u * q.stage_cap + 8u * q.col_cap + align16(2u * (p.g_longest + 32u));
        if (q.col_cap < 65536u && text_smem_bytes(q, kTextWarps) <= (size_t)max_smem_optin / 2) {  // two CTAs per SM at least
            p = q;
            mode = 2;
        }
    }
    int warps = kTextWarps;  // fewer warps per CTA when the reads are so long that eight staging areas do not fit
    while (warps > 1 && text_smem_bytes(p, warps) > (size_t)max_smem_optin) warps >>= 1;
    const size_t smem = text_smem_bytes(p, warps);
    if (smem > (size_t)max_smem_optin) return -1000;
    auto kern = mode == 2 ? k_text<2> : mode == 1 ? k_text<1> : k_text<0>;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    uint64_t blocks = (p.N + warps - 1) / warps;
    if (blocks > (uint64_t)num_sms * per_sm) blocks = (uint64_t)num_sms * per_sm;
    kern<<<(unsigned)blocks, warps * 32, smem, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int configure_text_kernels(int max_smem_optin)
{
    cudaError_t e = cudaFuncSetAttribute(k_text<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_text<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaFuncSetAttribute(k_text<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
}

}  // namespace treat_b200
