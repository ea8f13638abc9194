// fasta_kernel.cu -- FASTA ingest on the device: the record loop of
// gofasta.SimpleParser (cmd/treat/align.go:114, storage.go:738) + parseMergeCount (fragment.go:75-89)
// over the raw file bytes in HBM, with the record semantics of the host reader treat::ReadFasta:
//   * everything before the first '>' of the file is skipped; that '>' opens the first record wherever it is,
//     every later record starts at a '>' that is the first byte of a line;
//   * Id = header line after '>', strings.TrimSpace'd; sequence = the bytes of the following lines without
//     '\n' and without one '\r' in front of a '\n' (or at the end of the file);
//   * ReadCount = `.+[_\-](\d+)$` on the first space-separated token of Id, else 1.
//
//   k_fasta_mark     16 bytes per thread: record starts per 4 KB tile, position of the first '>'
//   scan kernels     exclusive prefix sums (tile counts -> record numbers, sequence lengths -> offsets)
//   k_fasta_starts   record start positions
//   k_fasta_records  one warp per record: header end, trimmed Id, ReadCount, sequence begin / length
//   k_fasta_compact  one warp per record: sequence bytes without line breaks -> contiguous reads
//                    (skipped when every record has a single sequence line: k_pack16 then reads the file in place)
#include <cuda_runtime.h>

#include <cstdint>

#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

constexpr int kTileThreads = 256;
constexpr uint32_t kTileBytes = kTileThreads * 16;

// 4-bit mask of the bytes of w equal to the byte replicated in c4
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t c4)
{
    const uint32_t z = w ^ c4;                                                  // zero byte <=> equal
    const uint32_t nz = (((z & 0x7f7f7f7fu) + 0x7f7f7f7fu) | z) & 0x80808080u;   // bit 7 of every non-zero byte
    return (((nz ^ 0x80808080u) >> 7) * 0x01020408u) >> 24;                      // bits 0,8,16,24 -> bits 0..3
}

// record-start flags of the 16 bytes at offset x (16-aligned): bit q <=> file[x+q] opens a record that is the first
// byte of a line; *gt = mask of all '>' bytes (for the first-'>' search)
__device__ __forceinline__ uint32_t start_flags16(const uint8_t *file, uint64_t len, uint64_t x, uint32_t *gt)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(file + x);  // the buffer is padded to a multiple of 16
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t g = 0, nl = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        g |= eq_mask4(w[i], 0x3e3e3e3eu) << (4 * i);
        nl |= eq_mask4(w[i], 0x0a0a0a0au) << (4 * i);
    }
    const uint32_t valid = len - x >= 16 ? 0xffffu : (1u << (uint32_t)(len - x)) - 1u;
    const uint32_t prev_nl = x == 0 ? 1u : (file[x - 1] == '\n' ? 1u : 0u);
    g &= valid;
    *gt = g;
    return g & ((nl << 1) | prev_nl) & 0xffffu;
}

__global__ void __launch_bounds__(kTileThreads) k_fasta_mark(const uint8_t *file, uint64_t len, uint64_t tile0, uint64_t ntiles,
                                                             uint32_t *tile_count, unsigned long long *first_gt)
{
    __shared__ uint32_t s_cnt;
    for (uint64_t tile = tile0 + blockIdx.x; tile < tile0 + ntiles; tile += gridDim.x) {
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        const uint64_t x = tile * kTileBytes + threadIdx.x * 16ull;
        if (x < len) {
            uint32_t gt;
            const uint32_t f = start_flags16(file, len, x, &gt);
            if (f) atomicAdd(&s_cnt, (uint32_t)__popc(f));
            if (gt) {
                const unsigned long long pos = x + (uint32_t)(__ffs(gt) - 1);
                if (pos < *reinterpret_cast<volatile unsigned long long *>(first_gt)) atomicMin(first_gt, pos);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) tile_count[tile] = s_cnt;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// exclusive scan of n uint32 values into uint64 (out[n] = total): block sums, one-block scan of the sums, apply
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256, kScanPer = 8, kScanTile = kScanThreads * kScanPer;

__device__ __forceinline__ unsigned long long block_exclusive(unsigned long long v, unsigned long long *s_warp, unsigned long long *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0ull, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(FULL_MASK, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < (int)(blockDim.x >> 5)) s_warp[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    const unsigned long long r = s_warp[warp] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(const uint32_t *in, uint64_t n, unsigned long long *sums)
{
    __shared__ unsigned long long s_warp[32], s_total;
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; k++)
        if (base + k < n) v += in[base + k];
    block_exclusive(v, s_warp, &s_total);
    if (threadIdx.x == 0) sums[blockIdx.x] = s_total;
}

__global__ void __launch_bounds__(1024) k_scan_top(unsigned long long *sums, uint64_t nblocks, unsigned long long *total)
{
    __shared__ unsigned long long s_warp[32], s_total;
    unsigned long long carry = 0;
    for (uint64_t b0 = 0; b0 < nblocks; b0 += blockDim.x) {
        const uint64_t i = b0 + threadIdx.x;
        const unsigned long long v = i < nblocks ? sums[i] : 0ull;
        const unsigned long long ex = block_exclusive(v, s_warp, &s_total);
        if (i < nblocks) sums[i] = carry + ex;
        carry += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const uint32_t *in, uint64_t n, const unsigned long long *sums,
                                                             unsigned long long *out)
{
    __shared__ unsigned long long s_warp[32], s_total;
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    uint32_t x[kScanPer];
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        x[k] = base + k < n ? in[base + k] : 0u;
        v += x[k];
    }
    unsigned long long ex = sums[blockIdx.x] + block_exclusive(v, s_warp, &s_total);
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += x[k];
    }
    if (base <= n && n < base + kScanPer) out[n] = ex;  // the thread whose range holds index n (its x[k >= n - base] are 0)
}

// out[0..n] = exclusive prefix sums of in[0..n) (uint64); sums: scratch of ceil(n / 2048) + 1 words
int launch_scan_u32(const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *sums, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t nblocks = n / kScanTile + 1;  // one more than needed when n is a multiple: that block writes out[n]
    k_scan_sums<<<(unsigned)nblocks, kScanThreads, 0, st>>>(in, n, reinterpret_cast<unsigned long long *>(sums));
    k_scan_top<<<1, 1024, 0, st>>>(reinterpret_cast<unsigned long long *>(sums), nblocks,
                                   reinterpret_cast<unsigned long long *>(sums) + nblocks);
    k_scan_apply<<<(unsigned)nblocks, kScanThreads, 0, st>>>(in, n, reinterpret_cast<const unsigned long long *>(sums),
                                                            reinterpret_cast<unsigned long long *>(out));
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// record starts.  tile_off = exclusive scan of the tile counts; extra = 1 when the first '>' of the file is not
// the first byte of a line (it then opens record 0 and shifts the others by one).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads) k_fasta_starts(const uint8_t *file, uint64_t len, uint64_t ntiles,
                                                               const unsigned long long *tile_off, const unsigned long long *first_gt,
                                                               unsigned long long *start)
{
    __shared__ unsigned long long s_warp[32], s_total;
    const unsigned long long first = *first_gt;
    const bool extra = first < len && !(first == 0 || file[first - 1] == '\n');
    if (extra && blockIdx.x == 0 && threadIdx.x == 0) start[0] = first;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t x = tile * kTileBytes + threadIdx.x * 16ull;
        uint32_t f = 0, gt;
        if (x < len) f = start_flags16(file, len, x, &gt);
        unsigned long long idx = tile_off[tile] + (extra ? 1ull : 0ull) + block_exclusive((unsigned long long)__popc(f), s_warp, &s_total);
        while (f) {
            const int q = __ffs(f) - 1;
            f &= f - 1;
            start[idx++] = x + q;
        }
    }
}

// ------------------------------------------------------------------------------------------
// one warp per record
// ------------------------------------------------------------------------------------------
struct FastaRecParams {
    const uint8_t *file;
    uint64_t len;
    const unsigned long long *start;  // [N]
    uint64_t N;
    unsigned long long *id_off;       // [N] Id inside the file
    uint32_t *id_len;
    uint32_t *read_count;
    unsigned long long *seq_begin;    // [N] first byte behind the header line
    uint32_t *seq_len;                // [N] sequence bytes (line breaks removed)
    uint32_t *scalars;                // [0] longest sequence, [1] some record has more than one sequence line, [2] too long
};

__device__ __forceinline__ bool is_space(uint32_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// parseMergeCount on the trimmed Id file[b, b + n)  (host mirror: treat::parseMergeCount)
__device__ uint32_t merge_count(const uint8_t *s, uint32_t n)
{
    uint32_t head = n;
    for (uint32_t i = 0; i < n; i++)
        if (s[i] == ' ') {
            head = i;
            break;
        }
    int sep = -1;
    for (int i = (int)head - 1; i >= 0; i--)
        if (s[i] == '_' || s[i] == '-') {
            sep = i;
            break;
        }
    if (sep <= 0 || (uint32_t)sep + 1 >= head) return 1u;
    unsigned long long v = 0;
    for (uint32_t i = (uint32_t)sep + 1; i < head; i++) {
        const uint32_t c = s[i];
        if (c < '0' || c > '9') return 1u;
        const unsigned long long x = c - '0';
        if (v > (0x7fffffffffffffffull - x) / 10ull) return 1u;  // strconv.Atoi range error
        v = v * 10ull + x;
    }
    return (uint32_t)v;
}

// 16-bit mask of the bytes of v equal to the byte replicated in c4
__device__ __forceinline__ uint32_t eq_mask16(const uint4 &v, uint32_t c4)
{
    return eq_mask4(v.x, c4) | (eq_mask4(v.y, c4) << 4) | (eq_mask4(v.z, c4) << 8) | (eq_mask4(v.w, c4) << 12);
}
__device__ __forceinline__ bool has_byte(const uint4 &v, uint32_t c4)
{
    uint32_t r = 0;
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t z = w[i] ^ c4;
        r |= (z - 0x01010101u) & ~z & 0x80808080u;
    }
    return r != 0u;
}
// bits [lo, hi) of a 16-bit mask, both clamped to 0..16
__device__ __forceinline__ uint32_t range16(long long lo, long long hi)
{
    const int l = (int)min(max(lo, 0ll), 16ll), h = (int)min(max(hi, 0ll), 16ll);
    return h > l ? ((1u << h) - 1u) & ~((1u << l) - 1u) : 0u;
}

// a record of any length, 32 bytes per step (the general form of the one-pass code in k_fasta_records)
__device__ __noinline__ void fasta_record_slow(const FastaRecParams &p, uint64_t r, int lane, uint32_t *wmax, uint32_t *wmulti)
{
    const uint64_t b = p.start[r], e = r + 1 < p.N ? p.start[r + 1] : p.len;
    // header line: [b + 1, eol)
    uint64_t eol = e;
    for (uint64_t x = b; x < e; x += 32) {
        const uint64_t pos = x + lane;
        const uint32_t m = __ballot_sync(FULL_MASK, pos < e && p.file[pos] == '\n');
        if (m) {
            eol = x + (uint32_t)(__ffs(m) - 1);
            break;
        }
    }
    // strings.TrimSpace of the header
    uint64_t ib = eol, il = 0;  // first non-space, one past the last non-space
    for (uint64_t x = b + 1; x < eol; x += 32) {
        const uint64_t pos = x + lane;
        const uint32_t m = __ballot_sync(FULL_MASK, pos < eol && !is_space(p.file[pos]));
        if (m) {
            if (ib == eol) ib = x + (uint32_t)(__ffs(m) - 1);
            il = x + (uint32_t)(32 - __clz(m));
        }
    }
    const uint32_t idn = il > ib ? (uint32_t)(il - ib) : 0u;
    // sequence bytes: everything behind the header line that is not a line break
    const uint64_t sb = eol < e ? eol + 1 : e;
    uint32_t kept = 0;
    uint64_t first_drop = e;
    for (uint64_t x = sb; x < e; x += 32) {
        const uint64_t pos = x + lane;
        bool drop = false, in = pos < e;
        if (in) {
            const uint32_t c = p.file[pos];
            drop = c == '\n' || (c == '\r' && (pos + 1 == p.len || p.file[pos + 1] == '\n'));
        }
        const uint32_t md = __ballot_sync(FULL_MASK, in && drop), mk = __ballot_sync(FULL_MASK, in && !drop);
        if (md && first_drop == e) first_drop = x + (uint32_t)(__ffs(md) - 1);
        kept += __popc(mk);
    }
    if (kept != (uint32_t)(first_drop - sb)) *wmulti = 1;  // a kept byte behind a dropped one: more than one line
    if (lane == 0) {
        p.id_off[r] = idn ? ib : (b + 1 < p.len ? b + 1 : p.len);
        p.id_len[r] = idn;
        p.read_count[r] = idn ? merge_count(p.file + ib, idn) : 1u;
        p.seq_begin[r] = sb;
        p.seq_len[r] = kept;
    }
    *wmax = max(*wmax, kept);
}

// One warp per record.  A record that fits 512 bytes from the 16-byte boundary below its '>' is handled in one pass:
// lane l holds bytes [16 l, 16 l + 16) as a uint4; newline / CR positions are 16-bit masks from byte-SIMD compares,
// the header end is a ballot + ffs, the sequence length a popcount + REDUX.
__global__ void __launch_bounds__(256) k_fasta_records(const __grid_constant__ FastaRecParams p)
{
    const int lane = threadIdx.x & 31;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t wmax = 0, wmulti = 0;
    for (uint64_t r = gw; r < p.N; r += nw) {
        const uint64_t b = p.start[r], e = r + 1 < p.N ? p.start[r + 1] : p.len;
        const uint64_t ab = b & ~15ull;
        if (e - ab > 512) {
            fasta_record_slow(p, r, lane, &wmax, &wmulti);
            continue;
        }
        const uint64_t x = ab + 16ull * lane;  // this lane's first byte
        uint4 v = make_uint4(0, 0, 0, 0);
        if (x < e) v = *reinterpret_cast<const uint4 *>(p.file + x);  // the buffer is padded past len
        const uint32_t valid = range16((long long)b - (long long)x, (long long)e - (long long)x);
        const uint32_t nl = eq_mask16(v, 0x0a0a0a0au) & valid;
        // header line [b + 1, eol): the first newline of the record
        const uint32_t fl = __ballot_sync(FULL_MASK, nl != 0u);
        uint64_t eol = e;
        if (fl) {
            const int f = __ffs(fl) - 1;
            eol = ab + 16ull * f + (uint32_t)(__shfl_sync(FULL_MASK, __ffs(nl) - 1, f));
        }
        const uint64_t sb = eol < e ? eol + 1 : e;
        // sequence bytes: behind the header line, not '\n', not a '\r' in front of a '\n' or at the end of the file
        const uint32_t smask = valid & range16((long long)sb - (long long)x, 16);
        uint32_t drop = nl & smask;
        if (__any_sync(FULL_MASK, has_byte(v, 0x0d0d0d0du))) {
            const uint32_t virt = (p.len >= x && p.len - x < 16) ? 1u << (uint32_t)(p.len - x) : 0u;  // the end of the file acts as '\n'
            const uint32_t nlv = (eq_mask16(v, 0x0a0a0a0au) & range16(0, (long long)p.len - (long long)x)) | virt;
            uint32_t nxt = __shfl_down_sync(FULL_MASK, nlv, 1) & 1u;
            if (lane == 31) nxt = (x + 16 == p.len) ? 1u : 0u;  // nothing of this record lies beyond lane 31
            drop |= eq_mask16(v, 0x0d0d0d0du) & ((nlv >> 1) | (nxt << 15)) & smask;
        }
        const uint32_t kept = __reduce_add_sync(FULL_MASK, (uint32_t)__popc(smask & ~drop));
        const uint32_t dl = __ballot_sync(FULL_MASK, drop != 0u);
        uint64_t first_drop = e;
        if (dl) {
            const int f = __ffs(dl) - 1;
            first_drop = ab + 16ull * f + (uint32_t)(__shfl_sync(FULL_MASK, __ffs(drop) - 1, f));
        }
        if (kept != (uint32_t)(first_drop - sb)) wmulti = 1;  // a kept byte behind a dropped one: more than one line
        // header: strings.TrimSpace and parseMergeCount.  Up to 32 header bytes are handled by the warp, one byte per
        // lane (the bytes were just loaded: L1 hits); longer headers, and counts of ten or more digits, by lane 0.
        const uint32_t hl = (uint32_t)(eol - (b + 1));
        uint64_t ib = b + 1;
        uint32_t idn = 0, rcount = 1u;
        if (hl <= 32) {
            const uint32_t c = (uint32_t)lane < hl ? (uint32_t)p.file[b + 1 + lane] : (uint32_t)' ';
            const uint32_t nsp = __ballot_sync(FULL_MASK, !is_space(c));
            const uint32_t spm = __ballot_sync(FULL_MASK, c == ' ');
            const uint32_t sepm = __ballot_sync(FULL_MASK, c == '_' || c == '-');
            const uint32_t dig = __ballot_sync(FULL_MASK, c >= '0' && c <= '9');
            if (nsp) {
                const int first = __ffs(nsp) - 1, last = 31 - __clz(nsp);
                idn = (uint32_t)(last - first + 1);
                ib = b + 1 + first;
                const uint32_t idm = (last == 31 ? 0xffffffffu : (2u << last) - 1u) & ~((1u << first) - 1u);
                const uint32_t sp = spm & idm;
                const int head_end = sp ? __ffs(sp) - 1 : last + 1;                 // first space-separated token: [first, head_end)
                const uint32_t hm = idm & (head_end == 32 ? 0xffffffffu : (1u << head_end) - 1u);
                const uint32_t sm = sepm & hm;
                if (sm) {
                    const int sep = 31 - __clz(sm);                                  // `.+[_\-](\d+)$`: the last '_' or '-' of the token
                    const uint32_t dm = hm & ~((2u << sep) - 1u);                    // what follows it
                    if (sep > first && dm != 0u && (dm & ~dig) == 0u) {
                        const int nd = head_end - sep - 1;
                        if (nd <= 9) {
                            uint32_t term = 0;
                            if ((dm >> lane) & 1u) {
                                const int ex = head_end - 1 - lane;
                                uint32_t pw = 1u;
                                if (ex & 1) pw *= 10u;
                                if (ex & 2) pw *= 100u;
                                if (ex & 4) pw *= 10000u;
                                if (ex & 8) pw *= 100000000u;
                                term = (c - '0') * pw;
                            }
                            rcount = __reduce_add_sync(FULL_MASK, term);
                        } else if (lane == 0) {
                            rcount = merge_count(p.file + ib, idn);  // int64 range check and the cut to 32 bits
                        }
                    }
                }
            }
        } else if (lane == 0) {
            uint64_t il = eol;
            while (ib < il && is_space(p.file[ib])) ib++;
            while (il > ib && is_space(p.file[il - 1])) il--;
            idn = il > ib ? (uint32_t)(il - ib) : 0u;
            rcount = idn ? merge_count(p.file + ib, idn) : 1u;
        }
        if (lane == 0) {
            p.id_off[r] = idn ? ib : (b + 1 < p.len ? b + 1 : p.len);
            p.id_len[r] = idn;
            p.read_count[r] = rcount;
            p.seq_begin[r] = sb;
            p.seq_len[r] = kept;
        }
        wmax = max(wmax, kept);
    }
    if (lane == 0) {
        if (wmax) atomicMax(&p.scalars[0], wmax);
        if (wmulti) atomicOr(&p.scalars[1], 1u);
    }
}

__global__ void __launch_bounds__(256) k_fasta_compact(const uint8_t *file, uint64_t len, const unsigned long long *start, uint64_t N,
                                                       const unsigned long long *seq_begin, const unsigned long long *seq_off,
                                                       uint8_t *seqs)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t r = gw; r < N; r += nw) {
        const uint64_t sb = seq_begin[r], e = r + 1 < N ? start[r + 1] : len;
        uint8_t *dst = seqs + seq_off[r];
        uint32_t done = 0;
        for (uint64_t x = sb; x < e; x += 32) {
            const uint64_t pos = x + lane;
            bool keep = false;
            uint32_t c = 0;
            if (pos < e) {
                c = file[pos];
                keep = !(c == '\n' || (c == '\r' && (pos + 1 == len || file[pos + 1] == '\n')));
            }
            const uint32_t mk = __ballot_sync(FULL_MASK, keep);
            if (keep) dst[done + __popc(mk & lt)] = (uint8_t)c;
            done += __popc(mk);
        }
    }
}

// ---- launchers --------------------------------------------------------------------------------
int launch_fasta_mark(const uint8_t *file, uint64_t len, uint64_t tile0, uint64_t ntiles, uint32_t *tile_count,
                      uint64_t *first_gt, int num_sms, void *stream)
{
    if (!ntiles) return 0;
    const uint64_t blocks = ntiles < (uint64_t)num_sms * 8 ? ntiles : (uint64_t)num_sms * 8;
    k_fasta_mark<<<(unsigned)blocks, kTileThreads, 0, (cudaStream_t)stream>>>(file, len, tile0, ntiles, tile_count,
                                                                              reinterpret_cast<unsigned long long *>(first_gt));
    return (int)cudaGetLastError();
}

int launch_fasta_starts(const uint8_t *file, uint64_t len, uint64_t ntiles, const uint64_t *tile_off, const uint64_t *first_gt,
                        uint64_t *start, int num_sms, void *stream)
{
    if (!ntiles) return 0;
    const uint64_t blocks = ntiles < (uint64_t)num_sms * 8 ? ntiles : (uint64_t)num_sms * 8;
    k_fasta_starts<<<(unsigned)blocks, kTileThreads, 0, (cudaStream_t)stream>>>(
        file, len, ntiles, reinterpret_cast<const unsigned long long *>(tile_off), reinterpret_cast<const unsigned long long *>(first_gt),
        reinterpret_cast<unsigned long long *>(start));
    return (int)cudaGetLastError();
}

int launch_fasta_records(const uint8_t *file, uint64_t len, const uint64_t *start, uint64_t N, uint64_t *id_off, uint32_t *id_len,
                         uint32_t *read_count, uint64_t *seq_begin, uint32_t *seq_len, uint32_t *scalars, int num_sms, void *stream)
{
    if (!N) return 0;
    FastaRecParams p;
    p.file = file;
    p.len = len;
    p.start = reinterpret_cast<const unsigned long long *>(start);
    p.N = N;
    p.id_off = reinterpret_cast<unsigned long long *>(id_off);
    p.id_len = id_len;
    p.read_count = read_count;
    p.seq_begin = reinterpret_cast<unsigned long long *>(seq_begin);
    p.seq_len = seq_len;
    p.scalars = scalars;
    uint64_t blocks = (N + 7) / 8;
    if (blocks > (uint64_t)num_sms * 8) blocks = (uint64_t)num_sms * 8;
    k_fasta_records<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return (int)cudaGetLastError();
}

int launch_fasta_compact(const uint8_t *file, uint64_t len, const uint64_t *start, uint64_t N, const uint64_t *seq_begin,
                         const uint64_t *seq_off, uint8_t *seqs, int num_sms, void *stream)
{
    if (!N) return 0;
    uint64_t blocks = (N + 7) / 8;
    if (blocks > (uint64_t)num_sms * 8) blocks = (uint64_t)num_sms * 8;
    k_fasta_compact<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        file, len, reinterpret_cast<const unsigned long long *>(start), N, reinterpret_cast<const unsigned long long *>(seq_begin),
        reinterpret_cast<const unsigned long long *>(seq_off), seqs);
    return (int)cudaGetLastError();
}

}  // namespace treat_b200
