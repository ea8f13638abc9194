// collapse_kernel.cu -- exact-duplicate collapse of a batch of reads on the device: the step the TREAT
// workflow runs upstream of `treat load` (fastx_collapser: identical sequences become one record whose id carries the
// count that parseMergeCount reads back, fragment.go:43-55,75-89).  Two reads with equal upper-cased sequences give the
// same Fragment (fragment.go:91-121: ToUpper, then the strip), hence the same Alignment, so only one of them has to be
// aligned and its ReadCount is the sum.
//
//   k_collapse_hash    one warp per read: a position-keyed 64-bit hash of the upper-cased characters (+ the length)
//   k_collapse_insert  one thread per read: open addressing (linear probing) on the 64-bit key with atomicCAS;
//                      the read remembers its slot
//   k_collapse_head    the lowest read index of every slot (atomicMin): the representative, independent of timing
//   k_collapse_verify  one warp per read: its characters against the representative's, byte for byte (case folded).
//                      Equal: it joins (its ReadCount is added to the representative's).  Different -- two sequences with
//                      the same 64-bit key -- is counted; the host then repeats the whole collapse with another seed, so
//                      the result never depends on the hash.
//   k_collapse_compact after a scan of the "I am a representative" flags: the unique reads in input order with their
//                      summed counts, as (begin, length) spans the pack kernels read in place; every read's unique number.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_format.h"
#include "ptx_helpers.cuh"

namespace treat_b200 {

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ uint32_t upper(uint32_t c) { return (c >= 'a' && c <= 'z') ? c - 32u : c; }

constexpr int kCollapseThreads = 256;

}  // namespace

__global__ void __launch_bounds__(kCollapseThreads) k_collapse_hash(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias, uint32_t N,
                                                                    uint32_t seed, unsigned long long *key)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = gw; r < N; r += nw) {
        const uint64_t b = off[r] - off_bias;
        const uint32_t L = (uint32_t)(off[r + 1] - off[r]);
        uint32_t h1 = 0, h2 = 0;
        for (uint32_t i = lane; i < L; i += 32) {
            const uint32_t c = upper(seqs[b + i]);
            const uint32_t x = (i << 8 | c) ^ seed;
            h1 += mix32(x * 0x9e3779b1u + 0x85ebca77u);
            h2 += mix32(x * 0xc2b2ae3du + 0x27d4eb2fu);
        }
        h1 = __reduce_add_sync(FULL_MASK, h1);
        h2 = __reduce_add_sync(FULL_MASK, h2);
        if (lane == 0) {
            unsigned long long k = ((unsigned long long)mix32(h2 ^ L) << 32) | mix32(h1 + 0x9e3779b9u * L);
            if (k == 0ull) k = 1ull;  // zero marks an empty slot
            key[r] = k;
        }
    }
}

__global__ void __launch_bounds__(kCollapseThreads) k_collapse_insert(const unsigned long long *key, uint32_t N, unsigned long long *table,
                                                                      uint32_t cap_mask, uint32_t *slot)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        const unsigned long long k = key[r];
        uint32_t s = (uint32_t)(k ^ (k >> 32)) & cap_mask;
        for (;;) {  // the table has at least twice as many slots as reads: an empty one is always found
            const unsigned long long prev = atomicCAS(&table[s], 0ull, k);
            if (prev == 0ull || prev == k) break;
            s = (s + 1u) & cap_mask;
        }
        slot[r] = s;
    }
}

__global__ void __launch_bounds__(kCollapseThreads) k_collapse_head(const uint32_t *slot, uint32_t N, uint32_t *head)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) atomicMin(&head[slot[r]], r);
}

__global__ void __launch_bounds__(kCollapseThreads) k_collapse_verify(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias,
                                                                      const uint32_t *read_count, uint32_t N, const uint32_t *slot,
                                                                      const uint32_t *head, uint32_t *rep, uint32_t *is_rep, uint32_t *sum,
                                                                      uint32_t *collisions, uint32_t *max_len)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    uint32_t longest = 0;
    for (uint32_t r = gw; r < N; r += nw) {
        const uint32_t h = head[slot[r]];
        const uint64_t b = off[r] - off_bias, bh = off[h] - off_bias;
        const uint32_t L = (uint32_t)(off[r + 1] - off[r]), Lh = (uint32_t)(off[h + 1] - off[h]);
        bool diff = L != Lh;
        if (!diff && h != r)
            for (uint32_t i = lane; i < L; i += 32) diff |= upper(seqs[b + i]) != upper(seqs[bh + i]);
        diff = __any_sync(FULL_MASK, diff);
        if (lane == 0) {
            if (diff) {
                atomicAdd(collisions, 1u);
                rep[r] = r;
                is_rep[r] = 1u;
            } else {
                rep[r] = h;
                is_rep[r] = h == r ? 1u : 0u;
                atomicAdd(&sum[h], read_count ? read_count[r] : 1u);  // uint32 like Fragment.ReadCount (fragment.go:63)
            }
        }
        if (h == r) longest = max(longest, L);
    }
    if (lane == 0 && longest) atomicMax(max_len, longest);
}

__global__ void __launch_bounds__(kCollapseThreads) k_collapse_compact(const uint64_t *off, uint64_t off_bias, uint32_t N, const uint32_t *rep,
                                                                       const uint32_t *is_rep, const unsigned long long *uidx,
                                                                       const uint32_t *sum, uint32_t *first, uint64_t *begin, uint32_t *len,
                                                                       uint32_t *count, uint32_t *member)
{
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < N; r += gridDim.x * blockDim.x) {
        if (is_rep[r]) {
            const uint32_t u = (uint32_t)uidx[r];
            first[u] = r;
            begin[u] = off[r] - off_bias;
            len[u] = (uint32_t)(off[r + 1] - off[r]);
            count[u] = sum[r];
        }
        if (member) member[r] = (uint32_t)uidx[rep[r]];
    }
}

// the distinct reads copied back to back (the wide-path kernels read contiguous reads): one warp per read
__global__ void __launch_bounds__(kCollapseThreads) k_collapse_gather(const uint8_t *seqs, const uint64_t *begin, const uint32_t *len,
                                                                      const unsigned long long *dst_off, uint32_t N, uint8_t *dst)
{
    const int lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = gw; r < N; r += nw) {
        const uint8_t *src = seqs + begin[r];
        uint8_t *out = dst + dst_off[r];
        for (uint32_t i = lane; i < len[r]; i += 32) out[i] = src[i];
    }
}

static unsigned grid_for(uint64_t threads, int num_sms)
{
    uint64_t b = (threads + kCollapseThreads - 1) / kCollapseThreads;
    const uint64_t cap = (uint64_t)num_sms * 8;
    if (b > cap) b = cap;
    return (unsigned)(b ? b : 1);
}

int launch_collapse_hash(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias, uint32_t N, uint32_t seed, uint64_t *key,
                         int num_sms, void *stream)
{
    k_collapse_hash<<<grid_for((uint64_t)N * 32, num_sms), kCollapseThreads, 0, (cudaStream_t)stream>>>(
        seqs, off, off_bias, N, seed, reinterpret_cast<unsigned long long *>(key));
    return (int)cudaGetLastError();
}

int launch_collapse_table(const uint64_t *key, uint32_t N, uint64_t *table, uint32_t cap, uint32_t *slot, uint32_t *head, int num_sms,
                          void *stream)
{
    const unsigned g = grid_for(N, num_sms);
    k_collapse_insert<<<g, kCollapseThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long *>(key), N,
                                                                        reinterpret_cast<unsigned long long *>(table), cap - 1u, slot);
    k_collapse_head<<<g, kCollapseThreads, 0, (cudaStream_t)stream>>>(slot, N, head);
    return (int)cudaGetLastError();
}

int launch_collapse_verify(const uint8_t *seqs, const uint64_t *off, uint64_t off_bias, const uint32_t *read_count, uint32_t N,
                           const uint32_t *slot, const uint32_t *head, uint32_t *rep, uint32_t *is_rep, uint32_t *sum, uint32_t *collisions,
                           uint32_t *max_len, int num_sms, void *stream)
{
    k_collapse_verify<<<grid_for((uint64_t)N * 32, num_sms), kCollapseThreads, 0, (cudaStream_t)stream>>>(
        seqs, off, off_bias, read_count, N, slot, head, rep, is_rep, sum, collisions, max_len);
    return (int)cudaGetLastError();
}

int launch_collapse_compact(const uint64_t *off, uint64_t off_bias, uint32_t N, const uint32_t *rep, const uint32_t *is_rep,
                            const uint64_t *uidx, const uint32_t *sum, uint32_t *first, uint64_t *begin, uint32_t *len, uint32_t *count,
                            uint32_t *member, int num_sms, void *stream)
{
    k_collapse_compact<<<grid_for(N, num_sms), kCollapseThreads, 0, (cudaStream_t)stream>>>(
        off, off_bias, N, rep, is_rep, reinterpret_cast<const unsigned long long *>(uidx), sum, first, begin, len, count, member);
    return (int)cudaGetLastError();
}

int launch_collapse_gather(const uint8_t *seqs, const uint64_t *begin, const uint32_t *len, const uint64_t *dst_off, uint32_t N, uint8_t *dst,
                           int num_sms, void *stream)
{
    k_collapse_gather<<<grid_for((uint64_t)N * 32, num_sms), kCollapseThreads, 0, (cudaStream_t)stream>>>(
        seqs, begin, len, reinterpret_cast<const unsigned long long *>(dst_off), N, dst);
    return (int)cudaGetLastError();
}

}  // namespace treat_b200
