// mbgc_b200/csrc/rcmatch.cu -- row f3 of SURVEY.md 8: the reverse-complement redundancy pass of mode 3 on the device (sm_100a).
//
// Replaces the per-base work of SimpleSequenceMatcher::rcMatchSequence
// (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176; called once per archive on the literal stream,
// mbgccoder/MBGC_Encoder.cpp:637-638) over CopMEMMatcher (matching/copmem/CopMEMMatcher.cpp):
//
//   index    genCumm / processRef (:139-227): every k1-th position of the sequence hashed over K bytes
//            (maRushPrime1HashSparsified, utils/Hashes.h:43-66) into a CSR table, at most 13 positions per bucket -- the 13
//            SMALLEST, because the single-threaded build takes them in position order.  Here: count (atomics), scan, fill in
//            arrival order, then every bucket sorts itself and keeps its 13 smallest (k_cm_sort_small / k_cm_sort_big).
//   query    processExactMatchQueryTight (:332-484): every k2-th position of the reverse-complemented sequence probes its
//            bucket; a candidate survives the self-match rule (:377-379), the K-mer compare and the length test after a
//            left / right extension.  The reference decides sequentially which candidates it even looks at (skip-ahead,
//            "inside the last match"); those decisions need only, per query position and in bucket order, the REAL seeds
//            (K-mer equal) with their match extents.  k_cm_query produces exactly that list, every position on its own;
//            extensions are cut at RC_EXT_CAP bytes per side (a seed inside a megabase repeat would otherwise walk the
//            whole repeat from every one of its thousands of seeds) and flagged open.  The host replays the walk over the
//            list -- linear in the number of seeds -- and finishes the open ends of the few matches it actually pushes.
#include "rcmatch.h"

namespace mbgc_b200 {

namespace {

constexpr int RC_THREADS = 256;
__constant__ uint8_t c_rcLut[256];

__device__ __forceinline__ uint32_t ld32(const uint8_t* p) {  // unaligned
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}

__device__ __forceinline__ uint32_t hash_sparsified(const uint8_t* s, int K) {  // utils/Hashes.h:43-66 (low 32 bits only matter)
    uint32_t h = static_cast<uint32_t>(K);
    const int words = K / 4;
    for (int j = 0; j < words; j++) {
        uint32_t k = ld32(s + 4 * j);
        k &= j < 3 ? 0x00FFFFFFu : 0x0000FFFFu;
        k += static_cast<uint32_t>(j);
        h ^= k;
        h *= 171717u;
    }
    return h;
}

__global__ void __launch_bounds__(RC_THREADS) k_rc_revcomp(const uint8_t* __restrict__ src, uint64_t n, uint8_t* __restrict__ dst) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = c_rcLut[src[n - 1 - i]];
}

// sampled position s (byte s * k1): its bucket, counted
__global__ void __launch_bounds__(RC_THREADS) k_cm_count(const uint8_t* __restrict__ seq, uint64_t nSamples, int K, int k1, uint32_t mask,
                                                         uint32_t* __restrict__ bucketOf, uint32_t* __restrict__ count) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t s = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < nSamples; s += stride) {
        const uint32_t h = hash_sparsified(seq + s * k1, K) & mask;
        bucketOf[s] = h;
        atomicAdd(&count[h], 1u);
    }
}

__global__ void __launch_bounds__(RC_THREADS) k_cm_fill(uint64_t nSamples, int k1, const uint32_t* __restrict__ bucketOf,
                                                        const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                                        uint32_t* __restrict__ entries) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t s = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < nSamples; s += stride) {
        const uint32_t h = bucketOf[s];
        entries[start[h] + atomicAdd(&cursor[h], 1u)] = static_cast<uint32_t>(s * k1);
    }
}

// buckets of 2 .. RC_SMALL entries: insertion sort by one thread; larger ones go to the list of big buckets
constexpr uint32_t RC_SMALL = 24;
__global__ void __launch_bounds__(RC_THREADS) k_cm_sort_small(uint32_t hashSize, const uint32_t* __restrict__ start, uint32_t* __restrict__ entries,
                                                              uint32_t* __restrict__ bigList, uint32_t* __restrict__ bigCount, uint32_t bigCap) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t h = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; h < hashSize; h += stride) {
        const uint32_t b = start[h], c = start[h + 1] - b;
        if (c < 2) continue;
        if (c > RC_SMALL) {
            const uint32_t at = atomicAdd(bigCount, 1u);
            if (at < bigCap) bigList[at] = static_cast<uint32_t>(h);
            continue;
        }
        uint32_t v[RC_SMALL];
        for (uint32_t i = 0; i < c; i++) v[i] = entries[b + i];
        for (uint32_t i = 1; i < c; i++) {
            const uint32_t x = v[i];
            uint32_t j = i;
            while (j > 0 && v[j - 1] > x) {
                v[j] = v[j - 1];
                j--;
            }
            v[j] = x;
        }
        for (uint32_t i = 0; i < c; i++) entries[b + i] = v[i];
    }
}

// a big bucket (a low-complexity k-mer: thousands of positions): a warp extracts its 13 smallest, ascending, into the
// bucket's first 13 slots (the rest of the bucket is never read again)
__global__ void __launch_bounds__(RC_THREADS) k_cm_sort_big(const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ bigCount,
                                                            uint32_t bigCap, const uint32_t* __restrict__ start, uint32_t* __restrict__ entries) {
    const int lane = threadIdx.x & 31;
    const uint32_t nBig = min(*bigCount, bigCap);
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nBig; w += warpsTotal) {
        const uint32_t h = bigList[w];
        const uint32_t b = start[h], c = start[h + 1] - b;
        uint32_t keep[RC_MAX_PER_BUCKET];
        uint64_t last = 0;  // smallest value not yet taken must be >= last (values are distinct positions; +1 after each take)
        for (int r = 0; r < RC_MAX_PER_BUCKET; r++) {
            uint32_t best = 0xFFFFFFFFu;
            for (uint32_t i = lane; i < c; i += 32) {
                const uint32_t x = entries[b + i];
                if (x >= last && x < best) best = x;
            }
            for (int d = 16; d; d >>= 1) best = min(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
            keep[r] = best;
            last = static_cast<uint64_t>(best) + 1;
        }
        __syncwarp();
        if (lane == 0)
            for (int r = 0; r < RC_MAX_PER_BUCKET; r++) entries[b + r] = keep[r];
        __syncwarp();
    }
}

// number of equal bytes of a[0..) and b[0..) going right, at most `limit`
__device__ __forceinline__ uint32_t equal_right(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint32_t limit) {
    uint32_t k = 0;
    while (k + 4 <= limit) {
        const uint32_t x = ld32(a + k) ^ ld32(b + k);
        if (x) return k + ((__ffs(x) - 1) >> 3);
        k += 4;
    }
    while (k < limit && a[k] == b[k]) k++;
    return k;
}
// number of equal bytes a[-1], a[-2], .. vs b[-1], b[-2], .. going left, at most `limit`
__device__ __forceinline__ uint32_t equal_left(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint32_t limit) {
    uint32_t k = 0;
    while (k + 4 <= limit) {
        const uint32_t x = ld32(a - k - 4) ^ ld32(b - k - 4);
        if (x) return k + (__clz(x) >> 3);
        k += 4;
    }
    while (k < limit && a[-static_cast<int64_t>(k) - 1] == b[-static_cast<int64_t>(k) - 1]) k++;
    return k;
}

// One query position per thread.  For every candidate of its bucket, in bucket order, that passes the self-match rule and
// whose K-mer equals the query's: a record (query position, order in the bucket, reference position, extents).
// The extents follow the reference's two loops literally (:396-403): right from the byte behind the K-mer while both texts
// last; left from the seed's first byte *itself* down to, but never including, byte 0 of either text.
__global__ void __launch_bounds__(RC_THREADS) k_cm_query(const uint8_t* __restrict__ seq, const uint8_t* __restrict__ q, uint64_t N, int K,
                                                         int k2, uint32_t mask, const uint32_t* __restrict__ start,
                                                         const uint32_t* __restrict__ entries, RcSeed* __restrict__ out,
                                                         uint32_t* __restrict__ outCount, uint32_t outCap) {
    const uint64_t nQuery = N >= static_cast<uint64_t>(K) ? (N - K) / k2 + 1 : 0;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t t = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < nQuery; t += stride) {
        const uint64_t pos = t * k2;
        const uint32_t h = hash_sparsified(q + pos, K) & mask;
        const uint32_t b = start[h];
        uint32_t c = start[h + 1] - b;
        if (c == 0) continue;
        if (c > RC_MAX_PER_BUCKET) c = RC_MAX_PER_BUCKET;
        for (uint32_t j = 0; j < c; j++) {
            const uint64_t src = entries[b + j];
            if (N - src < pos) continue;                                    // :377-379 (destIsSrc && revComplMatching)
            if (equal_right(seq + src, q + pos, static_cast<uint32_t>(K)) < static_cast<uint32_t>(K)) continue;   // memcmp(curr1, curr2, K), :405
            // right: first index r >= K with src + r == N or pos + r == N or a mismatch
            const uint64_t roomR = (N - src < N - pos ? N - src : N - pos) - K;
            const uint32_t capR = static_cast<uint32_t>(roomR < RC_EXT_CAP ? roomR : RC_EXT_CAP);
            const uint32_t er = equal_right(seq + src + K, q + pos + K, capR);
            // left: the loop starts ON the seed's first byte (equal: the K-mer matched) and stops before byte 0 of either text
            const uint64_t roomL = src < pos ? src : pos;                   // steps the loop can take
            uint32_t el = 0;                                                // bytes left of the seed that are equal AND reachable
            uint32_t openL = 0;
            if (roomL > 0) {
                // step 1 moves from the seed byte to src-1 / pos-1; it continues while those are equal and not at index 0
                const uint32_t capL = static_cast<uint32_t>(roomL - 1 < RC_EXT_CAP ? roomL - 1 : RC_EXT_CAP);
                el = equal_left(seq + src, q + pos, capL);
                openL = (el == capL && capL == RC_EXT_CAP) ? 1u : 0u;
            }
            const uint32_t at = atomicAdd(outCount, 1u);
            if (at < outCap) {
                RcSeed r;
                r.pos = static_cast<uint32_t>(pos);
                r.src = static_cast<uint32_t>(src);
                r.left = el;
                r.right = er;
                r.order = static_cast<uint16_t>(j);
                r.flags = static_cast<uint16_t>(openL | ((er == capR && capR == RC_EXT_CAP) ? 2u : 0u) | (roomL == 0 ? 4u : 0u));
                out[at] = r;
            }
        }
    }
}

inline unsigned grid_for(uint64_t items, unsigned cap) {
    const uint64_t b = (items + RC_THREADS - 1) / RC_THREADS;
    return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void rc_upload_lut(const uint8_t lut[256]) { cudaMemcpyToSymbol(c_rcLut, lut, 256); }

int launch_rc_revcomp(const uint8_t* src, uint64_t n, uint8_t* dst, cudaStream_t st) {
    if (n == 0) return 0;
    k_rc_revcomp<<<grid_for(n, 148 * 16), RC_THREADS, 0, st>>>(src, n, dst);
    return 1;
}

int launch_cm_count(const uint8_t* seq, uint64_t nSamples, int K, int k1, uint32_t mask, uint32_t* bucketOf, uint32_t* count, cudaStream_t st) {
    if (nSamples == 0) return 0;
    k_cm_count<<<grid_for(nSamples, 148 * 16), RC_THREADS, 0, st>>>(seq, nSamples, K, k1, mask, bucketOf, count);
    return 1;
}

int launch_cm_fill(uint64_t nSamples, int k1, const uint32_t* bucketOf, const uint32_t* start, uint32_t* cursor, uint32_t* entries, cudaStream_t st) {
    if (nSamples == 0) return 0;
    k_cm_fill<<<grid_for(nSamples, 148 * 16), RC_THREADS, 0, st>>>(nSamples, k1, bucketOf, start, cursor, entries);
    return 1;
}

int launch_cm_sort(uint32_t hashSize, const uint32_t* start, uint32_t* entries, uint32_t* bigList, uint32_t* bigCount, uint32_t bigCap, cudaStream_t st) {
    k_cm_sort_small<<<148 * 16, RC_THREADS, 0, st>>>(hashSize, start, entries, bigList, bigCount, bigCap);
    k_cm_sort_big<<<148 * 4, RC_THREADS, 0, st>>>(bigList, bigCount, bigCap, start, entries);
    return 2;
}

int launch_cm_query(const uint8_t* seq, const uint8_t* q, uint64_t N, int K, int k2, uint32_t mask, const uint32_t* start,
                    const uint32_t* entries, RcSeed* out, uint32_t* outCount, uint32_t outCap, cudaStream_t st) {
    if (N < static_cast<uint64_t>(K)) return 0;
    k_cm_query<<<grid_for((N - K) / k2 + 1, 148 * 16), RC_THREADS, 0, st>>>(seq, q, N, K, k2, mask, start, entries, out, outCount, outCap);
    return 1;
}

}  // namespace mbgc_b200
