// mbgc_b200/csrc/kernels.cu -- hand-written sm_100a kernels of the MBGC match finder.
//
// Everything here is integer / byte work bounded by HBM and L2 bandwidth (no tensor cores: nothing on the
// path is a contraction).  Results are bit-exact with the reference's single-threaded run; the design note
// for each kernel says which reference loop it replaces (paths relative to /root/reference, ".cpp" =
// matching/SlidingWindowSparseEMMatcher.cpp).
//
// Data flow of one match pass over a batch of target texts ("contigs") against a frozen index:
//
//   K3 k_probe        every target position: hash (utils/Hashes.h:28-40) -> shared-memory filter -> occupancy
//                     bit (L2) -> table slot -> window test (.cpp:211-222) -> K-mer verify (.cpp:298)
//                     => position-indexed seed staging per 512-position chunk
//   scan + k_compact  chunk counts -> dense, position-ordered seed arrays (pos, enc)
//   k_heads/k_runs    consecutive seeds on one diagonal with spacing k1 collapse into runs (their k-mers
//                     overlap, so the whole span is known equal without comparing it again)
//   K4 k_extend       per run: raw left / right reach beyond its first / last seed (.cpp:227-246, 257-259),
//                     warp-cooperative 8-byte compares; adjacent runs on one diagonal are linked, not re-walked
//   K5 k_resolve      per contig: the reference's sequential accept / overlap / skip logic (.cpp:250-315)
//                     replayed over the run list as pure integer arithmetic (no text access)
//   k_gather          per-contig match lists -> one dense list for the device->host copy
#include "kernels.h"

#include <atomic>

#include <cstdio>

namespace mbgc_b200 {

// ------------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------------

// 4 bytes at an arbitrary address from two aligned words (buffers carry >= 64 bytes of padding)
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    return __funnelshift_r(w[0], w[1], static_cast<uint32_t>(a & 3) * 8);
}

// same through the read-only path (global memory that no kernel of the pass writes)
__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    return __funnelshift_r(__ldg(w), __ldg(w + 1), static_cast<uint32_t>(a & 3) * 8);
}

__device__ __forceinline__ uint64_t ld_u64_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    return static_cast<uint64_t>(__funnelshift_r(w0, w1, sh)) |
           (static_cast<uint64_t>(__funnelshift_r(w1, w2, sh)) << 32);
}

// ---- mbarrier / bulk-copy (TMA) primitives: Blackwell async proxy, inline PTX ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// utils/Hashes.h:28-40 (maRushPrime1HashSimplified<K>): only the low 32 bits are ever used
__device__ __forceinline__ uint32_t hash_step(uint32_t h, uint32_t w, uint32_t j) { return (h ^ (w + j)) * 171717u; }

__device__ __forceinline__ uint64_t decode_pos(uint32_t enc, int k1ord) {
    return k1ord >= 0 ? (static_cast<uint64_t>(enc) << k1ord) : static_cast<uint64_t>(enc);
}

// .cpp:212-220: is a candidate at reference position s usable under the sliding window?
__device__ __forceinline__ bool window_ok(uint64_t s, int K, uint64_t pos1, uint64_t swStop, int& cls) {
    const bool endsBefore = s + K < pos1;
    const bool startsBefore = s < swStop;
    cls = (endsBefore ? 1 : 0) | (startsBefore ? 2 : 0);
    if (pos1 <= swStop) return !(!endsBefore && startsBefore);
    return !(!endsBefore || startsBefore);
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// largest c in [0, n) with arr[c] <= x   (arr ascending, arr[0] <= x)
__device__ __forceinline__ int upper_contig(const uint32_t* arr, int n, uint32_t x) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (arr[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------------
// K2: index insert  (.cpp:146-171)
//
// One thread per sampled reference position.  The reference stores ht[hash] = encodedPos with "last writer
// in program order wins".  Here every sample carries its global program-order sequence number in the high
// half of a 64-bit slot, so a single atomicMax realises exactly that order across threads, launches and
// buffer wrap-arounds.  Lanes of a warp that hit the same slot are merged with __match_any_sync first (the
// highest lane has the highest sequence number), so each distinct slot costs one atomic per warp.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_index_insert(const uint8_t* __restrict__ ref, unsigned long long* __restrict__ ht,
                                                      uint32_t* __restrict__ occ, uint32_t* __restrict__ filter,
                                                      uint32_t filterBits, uint32_t hashMask, int KW, int k1,
                                                      int k1ord, InsertRange r0, InsertRange r1) {
    const int64_t total = r0.count + r1.count;
    const int64_t totalPadded = (total + 31) & ~static_cast<int64_t>(31);
    const int lane = threadIdx.x & 31;
    for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < totalPadded;
         idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const bool valid = idx < total;
        unsigned long long slotKey = (1ull << 40) + lane;  // unique per lane: never merges
        unsigned long long key = 0;
        uint32_t slot = 0, hfull = 0;
        if (valid) {
            const bool second = idx >= r0.count;
            const int64_t k = second ? idx - r0.count : idx;
            const int64_t pos = (second ? r1.begin : r0.begin) + k * k1;
            const uint64_t seq = (second ? r1.seqBase : r0.seqBase) + static_cast<uint64_t>(k);
            uint32_t h = static_cast<uint32_t>(KW * 4);
            for (int j = 0; j < KW; j++) h = hash_step(h, ld_u32_unaligned(ref + pos + 4 * j), j);
            slot = h & hashMask;
            hfull = h;
            const uint32_t enc = static_cast<uint32_t>(k1ord >= 0 ? (static_cast<uint64_t>(pos) >> k1ord)
                                                                   : static_cast<uint64_t>(pos));
            key = (seq << 32) | enc;
            slotKey = slot;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, slotKey);
        if (valid && lane == 31 - __clz(peers)) {
            atomicMax(&ht[slot], key);
            atomicOr(&occ[slot >> 5], 0x80000000u >> (slot & 31));  // bit i lives at 31 - (i & 31), see k_probe
        }
        if (valid) {  // first-level (shared-memory) filter of the probe kernel; bits are never cleared
            const uint32_t fi = __umulhi(hfull, filterBits);
            atomicOr(&filter[fi >> 5], 0x80000000u >> (fi & 31));
        }
    }
}

void launch_index_insert(const uint8_t* ref, unsigned long long* ht, uint32_t* occ, uint32_t* filter, uint32_t filterBits,
                         uint32_t hashMask, int K, int k1, int k1ord, InsertRange r0, InsertRange r1, cudaStream_t st) {
    const int64_t total = r0.count + r1.count;
    if (total <= 0) return;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    k_index_insert<<<static_cast<unsigned>(blocks), 256, 0, st>>>(ref, ht, occ, filter, filterBits, hashMask, K / 4, k1,
                                                                  k1ord, r0, r1);
}

__global__ void k_table_low32(const unsigned long long* __restrict__ ht, uint32_t* __restrict__ out, uint64_t n) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<uint32_t>(ht[i]);
}

// epoch rebase of the insert sequence numbers: every slot keeps its entry (low half) and forgets its number
__global__ void k_table_rebase(unsigned long long* __restrict__ ht, uint64_t n) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        ht[i] &= 0xFFFFFFFFull;
}

void launch_table_rebase(unsigned long long* ht, uint64_t n, cudaStream_t st) {
    k_table_rebase<<<device_sm_count() * 8, 256, 0, st>>>(ht, n);
}

void launch_table_low32(const unsigned long long* ht, uint32_t* out, uint64_t n, cudaStream_t st) {
    k_table_low32<<<device_sm_count() * 8, 256, 0, st>>>(ht, out, n);
}

// ------------------------------------------------------------------------------------------------------
// reference segments from the staged text of a loadRef call (.cpp:420-424): a straight copy, or the upper-casing
// reverse complement of utils/helper.cpp:405-410 through its 256-entry table (:312-358)
// ------------------------------------------------------------------------------------------------------
__constant__ uint8_t c_complement[256];

__global__ void __launch_bounds__(256) k_ref_segment(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n,
                                                     int reverseComplement) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = reverseComplement ? c_complement[src[n - 1 - i]] : src[i];
}

void upload_complement_table(const uint8_t table[256]) { cudaMemcpyToSymbol(c_complement, table, 256); }

void launch_ref_segment(const uint8_t* src, uint8_t* dst, uint64_t n, bool reverseComplement, cudaStream_t st) {
    if (n == 0) return;
    uint64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const uint64_t cap = static_cast<uint64_t>(device_sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    k_ref_segment<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, dst, n, reverseComplement ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------------
// exclusive scan of a uint32 array whose length may live in device memory (no host sync between stages)
// ------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sWarp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (blockDim.x >> 5) ? sWarp[lane] : 0;
        const uint32_t wi = warp_incl_scan(w, lane);
        sWarp[lane] = wi - w;
        if (lane == 31) sWarp[32] = wi;
    }
    __syncthreads();
    const uint32_t res = inc - v + sWarp[warp];
    *total = sWarp[32];
    __syncthreads();
    return res;
}

// n = *nPtr (device-resident length) or nConst; clamped to nMax so an overflowing pass cannot write out of bounds
__device__ __forceinline__ uint32_t scan_len(const uint32_t* nPtr, uint32_t nConst, uint32_t nMax) {
    const uint32_t n = nPtr ? *nPtr : nConst;
    return n < nMax ? n : nMax;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, const uint32_t* nPtr,
                                                              uint32_t nConst, uint32_t nMax,
                                                              uint32_t* __restrict__ blockSums) {
    __shared__ uint32_t sWarp[33];
    const uint32_t n = scan_len(nPtr, nConst, nMax);
    const uint32_t nBlocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const uint32_t i = b * SCAN_BLOCK + k * SCAN_THREADS + threadIdx.x;
            if (i < n) s += in[i];
        }
        uint32_t total;
        block_excl_scan(s, &total, sWarp);
        if (threadIdx.x == 0) blockSums[b] = total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(uint32_t* __restrict__ blockSums, const uint32_t* nPtr,
                                                                uint32_t nConst, uint32_t nMax) {
    __shared__ uint32_t sWarp[33];
    const uint32_t n = scan_len(nPtr, nConst, nMax);
    const uint32_t nBlocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nBlocks; base += SCAN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nBlocks ? blockSums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan(v, &total, sWarp);
        if (i < nBlocks) blockSums[i] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) blockSums[nBlocks] = carry;
}

// out may alias in.  out[n] receives the grand total.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const uint32_t* in, uint32_t* out, const uint32_t* nPtr,
                                                             uint32_t nConst, uint32_t nMax,
                                                             const uint32_t* __restrict__ blockSums,
                                                             uint32_t* totalOut) {
    __shared__ uint32_t sWarp[33];
    const uint32_t n = scan_len(nPtr, nConst, nMax);
    const uint32_t nBlocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        uint32_t v[SCAN_ITEMS];
        uint32_t s = 0;
        const uint32_t i0 = b * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            s += v[k];
        }
        uint32_t total;
        uint32_t ex = block_excl_scan(s, &total, sWarp) + blockSums[b];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint32_t t = blockSums[nBlocks];
        out[n] = t;
        if (totalOut) *totalOut = t;
    }
}

// Single-pass exclusive scan (decoupled look-back): one launch instead of three.  Every 4096-element tile publishes its
// aggregate, then its inclusive prefix, in one 64-bit word of `state` ((flag << 32) | value; flag 1 = aggregate, 2 =
// inclusive prefix, 0 = nothing yet -- `state` is zeroed together with the pass counters; its last word is the tile
// ticket).  Tiles are handed out by ticket, so every predecessor of a tile belongs to a CTA that is already running: no
// CTA can wait for one that has not been scheduled.  The first warp of a tile looks back 32 predecessors at a time
// until it meets an inclusive prefix.  out may alias in.  out[n] receives the grand total.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const uint32_t* in, uint32_t* out, const uint32_t* nPtr,
                                                                uint32_t nConst, uint32_t nMax,
                                                                unsigned long long* state, uint32_t stateWords,
                                                                uint32_t* totalOut) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sWarp[33];
    __shared__ uint32_t sPrefix, sTile;
    const uint32_t n = scan_len(nPtr, nConst, nMax);
    const uint32_t nBlocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nBlocks == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            out[0] = 0;
            if (totalOut) *totalOut = 0;
        }
        return;
    }
    volatile unsigned long long* vs = state;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(state + stateWords - 1);
    const int lane = threadIdx.x & 31;
    for (;;) {
        if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t b = sTile;
        if (b >= nBlocks) break;
        uint32_t v[SCAN_ITEMS];
        uint32_t s = 0;
        const uint32_t i0 = b * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0;
            s += v[k];
        }
        uint32_t total;
        uint32_t ex = block_excl_scan(s, &total, sWarp);
        if (threadIdx.x < 32) {
            uint32_t prefix = 0;
            if (b == 0) {
                if (lane == 0) vs[0] = (2ull << 32) | total;
            } else {
                if (lane == 0) vs[b] = (1ull << 32) | total;
                for (int64_t base = static_cast<int64_t>(b) - 1;; base -= 32) {
                    const int64_t j = base - lane;
                    unsigned long long w = 2ull << 32;  // in front of tile 0: an inclusive prefix of 0
                    if (j >= 0) {
                        do {
                            w = vs[j];
                        } while ((w >> 32) == 0);
                    }
                    const uint32_t incl = __ballot_sync(0xFFFFFFFFu, (w >> 32) == 2);
                    const int first = incl ? __ffs(incl) - 1 : 31;  // nearest predecessor that knows its inclusive prefix
                    prefix += __reduce_add_sync(0xFFFFFFFFu, lane <= first ? static_cast<uint32_t>(w) : 0u);
                    if (incl) break;
                }
                if (lane == 0) vs[b] = (2ull << 32) | static_cast<uint32_t>(prefix + total);
            }
            if (lane == 0) {
                sPrefix = prefix;
                if (b == nBlocks - 1) {
                    out[n] = prefix + total;
                    if (totalOut) *totalOut = prefix + total;
                }
            }
        }
        __syncthreads();
        ex += sPrefix;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < n) out[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();  // sTile / sPrefix are rewritten in the next round
    }
}

// Single-pass exclusive scan over contiguous RANGES: the n elements are cut into at most gridDim.x ranges of whole
// 4096-element tiles, handed out by ticket.  A CTA reduces its range, publishes the range total ((1 << 32) | total in
// `state`, zeroed with the pass counters; last word = ticket), adds up the totals of all ranges before it -- their CTAs
// hold earlier tickets, so they are running, and they publish without waiting for anybody: no chain of dependent
// look-backs, one global round trip -- and then scans its range RANGE_GROUP tiles at a time (the warp scans of the tiles
// of a group share two barriers).  A range of at most RANGE_GROUP tiles is read once and stays in registers.
// in / out 16-byte aligned; out may alias in.  out[n] receives the grand total.
constexpr int RANGE_GROUP = 4;

__device__ __forceinline__ uint4 scan_load4(const uint32_t* in, uint32_t i0, uint32_t n) {
    if (i0 + 3 < n) return *reinterpret_cast<const uint4*>(in + i0);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i0 < n) v.x = in[i0];
    if (i0 + 1 < n) v.y = in[i0 + 1];
    if (i0 + 2 < n) v.z = in[i0 + 2];
    return v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_ranges(const uint32_t* in, uint32_t* out, const uint32_t* nPtr,
                                                              uint32_t nConst, uint32_t nMax, unsigned long long* state,
                                                              uint32_t stateWords, uint32_t* totalOut) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sW[RANGE_GROUP][33];
    __shared__ uint32_t sTileTot[RANGE_GROUP];
    __shared__ uint32_t sRed[33];
    __shared__ uint32_t sTicket, sPrefix;
    const uint32_t n = scan_len(nPtr, nConst, nMax);
    const uint32_t nTiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nTiles == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            out[0] = 0;
            if (totalOut) *totalOut = 0;
        }
        return;
    }
    const uint32_t tpr = (nTiles + gridDim.x - 1) / gridDim.x;  // tiles per range
    const uint32_t nRanges = (nTiles + tpr - 1) / tpr;
    const bool keep = tpr <= RANGE_GROUP;
    volatile unsigned long long* vs = state;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(state + stateWords - 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) sTicket = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t t = sTicket;
        if (t >= nRanges) break;
        const uint32_t tile0 = t * tpr, tile1 = min(nTiles, tile0 + tpr);
        uint4 v[RANGE_GROUP];
        uint32_t s[RANGE_GROUP], inc[RANGE_GROUP];
        uint32_t rangeTotal = 0;
        // ---- the range total
        if (keep) {
#pragma unroll
            for (int k = 0; k < RANGE_GROUP; k++) {
                v[k] = make_uint4(0, 0, 0, 0);
                if (tile0 + k < tile1) v[k] = scan_load4(in, (tile0 + k) * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS, n);
            }
#pragma unroll
            for (int k = 0; k < RANGE_GROUP; k++) {
                s[k] = v[k].x + v[k].y + v[k].z + v[k].w;
                inc[k] = warp_incl_scan(s[k], lane);
                if (lane == 31) sW[k][warp] = inc[k];
            }
            __syncthreads();
            if (warp < RANGE_GROUP) {
                const uint32_t w = sW[warp][lane];
                const uint32_t wi = warp_incl_scan(w, lane);
                sW[warp][lane] = wi - w;
                if (lane == 31) sTileTot[warp] = wi;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RANGE_GROUP; k++) rangeTotal += sTileTot[k];
        } else {
            uint32_t acc = 0;
            for (uint32_t tile = tile0; tile < tile1; tile++) {
                const uint4 x = scan_load4(in, tile * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS, n);
                acc += x.x + x.y + x.z + x.w;
            }
            acc = __reduce_add_sync(0xFFFFFFFFu, acc);
            if (lane == 0) sRed[warp] = acc;
            __syncthreads();
            if (warp == 0) {
                const uint32_t r = __reduce_add_sync(0xFFFFFFFFu, sRed[lane]);
                if (lane == 0) sRed[32] = r;
            }
            __syncthreads();
            rangeTotal = sRed[32];
        }
        // ---- publish it, add up the ranges in front
        if (warp == 0) {
            if (lane == 0) vs[t] = (1ull << 32) | rangeTotal;
            uint32_t p = 0;
            for (uint32_t j = lane; j < t; j += 32) {
                unsigned long long w;
                do {
                    w = vs[j];
                } while ((w >> 32) == 0);
                p += static_cast<uint32_t>(w);
            }
            p = __reduce_add_sync(0xFFFFFFFFu, p);
            if (lane == 0) {
                sPrefix = p;
                if (t == nRanges - 1) {
                    out[n] = p + rangeTotal;
                    if (totalOut) *totalOut = p + rangeTotal;
                }
            }
        }
        __syncthreads();
        uint32_t carry = sPrefix;
        // ---- the scan of the range
        for (uint32_t g0 = tile0; g0 < tile1; g0 += RANGE_GROUP) {
            if (!keep) {
#pragma unroll
                for (int k = 0; k < RANGE_GROUP; k++) {
                    v[k] = make_uint4(0, 0, 0, 0);
                    if (g0 + k < tile1) v[k] = scan_load4(in, (g0 + k) * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS, n);
                }
#pragma unroll
                for (int k = 0; k < RANGE_GROUP; k++) {
                    s[k] = v[k].x + v[k].y + v[k].z + v[k].w;
                    inc[k] = warp_incl_scan(s[k], lane);
                    if (lane == 31) sW[k][warp] = inc[k];
                }
                __syncthreads();
                if (warp < RANGE_GROUP) {
                    const uint32_t w = sW[warp][lane];
                    const uint32_t wi = warp_incl_scan(w, lane);
                    sW[warp][lane] = wi - w;
                    if (lane == 31) sTileTot[warp] = wi;
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < RANGE_GROUP; k++) {
                if (g0 + k < tile1) {
                    const uint32_t i0 = (g0 + k) * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
                    uint4 e;
                    e.x = carry + sW[k][warp] + inc[k] - s[k];
                    e.y = e.x + v[k].x;
                    e.z = e.y + v[k].y;
                    e.w = e.z + v[k].z;
                    if (i0 + 3 < n) {
                        *reinterpret_cast<uint4*>(out + i0) = e;
                    } else {
                        if (i0 < n) out[i0] = e.x;
                        if (i0 + 1 < n) out[i0 + 1] = e.y;
                        if (i0 + 2 < n) out[i0 + 2] = e.z;
                    }
                }
                carry += sTileTot[k];
            }
            __syncthreads();  // sW / sTileTot are rewritten by the next group, sTicket / sPrefix by the next range
        }
    }
}

static int scan_exclusive(const uint32_t* in, uint32_t* out, const uint32_t* nPtr, uint32_t nConst, uint32_t nMax,
                          unsigned long long* state, uint32_t stateWords, uint32_t* totalOut, cudaStream_t st) {
    uint32_t blocks = (nMax + SCAN_BLOCK - 1) / SCAN_BLOCK;
    const uint32_t cap = static_cast<uint32_t>(device_sm_count()) * 2;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    static const bool lookback = []() { const char* e = getenv("MBGC_B200_SCAN_LOOKBACK"); return e && e[0] == '1'; }();
    const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (lookback || !aligned) launch_pdl(k_scan_lookback, blocks, SCAN_THREADS, 0, st, in, out, nPtr, nConst, nMax, state, stateWords, totalOut);
    else launch_pdl(k_scan_ranges, blocks, SCAN_THREADS, 0, st, in, out, nPtr, nConst, nMax, state, stateWords, totalOut);
    return 1;
}

// exclusive scan for other translation units (rcmatch.cpp): in[0 .. n) -> out[0 .. n], out[n] = total; `state` = zeroed
// scratch of n / 4096 + 8 words
int launch_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* state, cudaStream_t st) {
    return scan_exclusive(in, out, nullptr, n, n, state, n / SCAN_BLOCK + 8, nullptr, st);
}

// ------------------------------------------------------------------------------------------------------
// K3: probe  (.cpp:200-226 and the K-mer verify of :298, for every target position)
//
// Warp-autonomous: a warp owns a chunk of 512 consecutive positions of one contig (lane = 16 positions) and
// never synchronises with other warps, so the SM hides latency with 24-32 resident warps instead of idling
// at block barriers.  Per chunk:
//   1. each lane pulls its 16+K-1 bytes as coalesced 128-bit loads (the 2nd/3rd hit L1: neighbour lanes just
//      fetched them), forms the unaligned 4-byte words with one funnel shift each and runs the K/4 hash rounds
//      of its 16 positions as independent chains;
//   2. first-level filter in SHARED memory: a ~1.5 Mbit Bloom-style bitmap (bit = mulhi(hash, bits)) that every
//      CTA copies from global once per launch.  A divergent global load costs one L1TEX wavefront per lane
//      (~1 lookup/clk/SM, the measured wall of the first version); a shared-memory lookup costs ~3 wavefronts
//      per warp.  With a 10 Mbp reference (625 k sampled k-mers) ~30 % of the positions survive;
//   3. survivors test the exact occupancy bitmap in global memory (L2-resident, predicated loads: only the
//      surviving lanes generate wavefronts);
//   4. set bits (seeds + load-factor false positives) read the 64-bit table slot, apply the window test and
//      compare the K-mer against the reference;
//   5. seeds are staged position-indexed (enc at scratchEnc[position], a 16-bit mask per lane): order is
//      implicit, no scan or barrier in the hot kernel.  k_compact turns the staging into dense arrays.
// ------------------------------------------------------------------------------------------------------
struct ProbeArgs {
    IndexState ix;
    BatchDesc b;
    const uint4* chunkDesc;  // per chunk: text byte offset (lo, hi), contig-local base position, nValid of the contig
    const uint32_t* chunkContig;
    uint32_t* chunkCount;
    uint16_t* seedMask;
    uint32_t* scratchEnc;
    const uint32_t* laneList;    // nullptr: every lane of every chunk; else the (chunk * 32 + lane) items k_clean left over
    const uint32_t* laneCount;   // device-resident length of laneList
    unsigned long long* anchor;  // k_anchor output: per chunk, the proposed reference position of its start (0 = none)
};

__host__ __device__ constexpr int probe_words(int KW) { return ((PROBE_NPT / 4 + KW + 3) / 4) * 4; }  // aligned words a lane holds

// 16 consecutive positions of one lane: hash (utils/Hashes.h:28-40) -> slot, first-level filter, occupancy bit.
// A = the lane's aligned text words.  Returns the mask of positions whose occupancy bit is set (position p at
// bit 15 - p).  Bit bookkeeping: filter and occupancy words store bit i at position 31 - (i & 31), so
// "w << (i & 31)" (one wrapping shift) brings the tested bit to bit 31 and one funnel shift appends it to a mask.
template <int KW, bool FILTER>
__device__ __forceinline__ uint32_t probe_lane16(const IndexState& ix, const uint32_t* sFilter,
                                                 const uint32_t (&A)[probe_words(KW)], uint32_t (&slots)[PROBE_NPT]) {
    constexpr int NPT = PROBE_NPT;
    uint32_t pass = 0;  // appended in (r, q) order
#pragma unroll
    for (int r = 0; r < 4; r++) {
        uint32_t V[NPT / 4 + KW - 1];
#pragma unroll
        for (int m = 0; m < NPT / 4 + KW - 1; m++) V[m] = r == 0 ? A[m] : __funnelshift_r(A[m], A[m + 1], 8 * r);
#pragma unroll
        for (int q = 0; q < NPT / 4; q++) {
            uint32_t h = static_cast<uint32_t>(KW * 4);
#pragma unroll
            for (int j = 0; j < KW; j++) h = hash_step(h, V[q + j], j);
            slots[4 * q + r] = h & ix.hashMask;
            if (FILTER) {
                const uint32_t fi = __umulhi(h, ix.filterBits);
                pass = __funnelshift_l(sFilter[fi >> 5] << (fi & 31), pass, 1);
            }
        }
    }
    if (!FILTER) pass = 0xffffu;
    uint32_t occw[NPT];
#pragma unroll
    for (int p = 0; p < NPT; p++) {
        occw[p] = 0;
        if (pass & (1u << (15 - ((p & 3) * 4 + (p >> 2))))) occw[p] = __ldg(&ix.occ[slots[p] >> 5]);
    }
    uint32_t hit = 0;  // appended in p order
#pragma unroll
    for (int p = 0; p < NPT; p++) hit = __funnelshift_l(occw[p] << (slots[p] & 31), hit, 1);
    return hit & 0xffffu;
}

// slots[p] for a run-time p: a 4-level select tree keeps the array in registers
__device__ __forceinline__ uint32_t select16(const uint32_t (&slots)[PROBE_NPT], int p) {
    uint32_t s8[8], s4[4], s2[2];
#pragma unroll
    for (int k = 0; k < 8; k++) s8[k] = (p & 1) ? slots[2 * k + 1] : slots[2 * k];
#pragma unroll
    for (int k = 0; k < 4; k++) s4[k] = (p & 2) ? s8[2 * k + 1] : s8[2 * k];
#pragma unroll
    for (int k = 0; k < 2; k++) s2[k] = (p & 4) ? s4[2 * k + 1] : s4[2 * k];
    return (p & 8) ? s2[1] : s2[0];
}

// candidate at lane position p: table slot -> window test (.cpp:212-220) -> K-mer compare (.cpp:298).
// Returns the table entry of a verified seed, else 0.
template <int KW>
__device__ __forceinline__ uint32_t verify_candidate(const IndexState& ix, const uint32_t (&A)[probe_words(KW)], int p,
                                                     uint32_t slot, uint64_t swStop) {
    const uint32_t enc = static_cast<uint32_t>(__ldg(&ix.ht[slot]));
    if (enc == 0) return 0;
    const uint64_t s = decode_pos(enc, ix.k1ord);
    int cls;
    if (!window_ok(s, KW * 4, ix.pos1, swStop, cls)) return 0;
    // target K-mer from the lane's own registers: words A[q .. q+KW] for q = p / 4 through a 2-level select, then
    // one funnel shift per word (no loads)
    const int q = p >> 2;
    const uint32_t sh = static_cast<uint32_t>(p & 3) * 8;
    uint32_t B[KW + 1];
#pragma unroll
    for (int m = 0; m <= KW; m++) {
        const uint32_t lo = (q & 1) ? A[m + 1] : A[m];
        const uint32_t hi = (q & 1) ? A[m + 3] : A[m + 2];
        B[m] = (q & 2) ? hi : lo;
    }
    const uint8_t* r = ix.ref + s;
    uint32_t diff = 0;
    if ((s & 15) == 0 && KW <= 8) {  // Exp variant with k1 >= 16: two aligned 128-bit loads
        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(r));
        const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(r) + 1);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int j = 0; j < KW; j++) diff |= __funnelshift_r(B[j], B[j + 1], sh) ^ rw[j < 8 ? j : 0];
    } else {
#pragma unroll
        for (int j = 0; j < KW; j++) diff |= __funnelshift_r(B[j], B[j + 1], sh) ^ ldg_u32_unaligned(r + 4 * j);
    }
    return diff == 0 ? enc : 0;
}

// The unit of work is a LANE ITEM: 16 consecutive positions of one chunk (item = chunk * 32 + lane-in-chunk).
// Without the fast path the items are simply 0 .. 32 nChunks - 1, so a warp still owns one chunk and its text
// loads coalesce; with it, k_clean hands over only the items it could not prove identical to the reference.
template <int KW, bool FILTER>
__global__ void __launch_bounds__(FILTER ? PROBE_THREADS_FILTER : PROBE_THREADS_PLAIN, 1) k_probe(const ProbeArgs a) {
    griddep_launch_dependents();
    constexpr int NPT = PROBE_NPT;
    constexpr int NWA = probe_words(KW);
    extern __shared__ __align__(128) uint32_t sFilter[];  // [filter words][mbarrier]
    const int lane = threadIdx.x & 31;
    // The first-level filter comes in by bulk copy (TMA): one thread queues the 192 KiB in 32 KiB pieces on an
    // mbarrier and the CTA goes straight on to its first items' descriptor and text loads; the wait sits in front of
    // the first filter lookup.
    const uint32_t barAddr = smem_addr(sFilter + a.ix.filterWords);
    bool filterPending = FILTER;
    if (FILTER) {
        if (threadIdx.x == 0) mbar_init(barAddr, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = a.ix.filterWords * 4;
            mbar_arrive_expect_tx(barAddr, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768)
                tma_load_1d(smem_addr(sFilter) + off, reinterpret_cast<const uint8_t*>(a.ix.filter) + off,
                            min(32768u, bytes - off), barAddr);
        }
    }
    // everything above is independent of the pass (the filter belongs to the index): the kernels in front may still run
    griddep_wait();
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    const uint32_t nItems = a.laneList ? *a.laneCount : a.b.nTiles * 32u;
    for (uint32_t kb = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; kb < nItems; kb += warpsTotal * 32) {
        const bool valid = kb + lane < nItems;
        const uint32_t item = valid ? (a.laneList ? a.laneList[kb + lane] : kb + lane) : 0;
        const uint32_t g = item >> 5, cl = item & 31u;  // chunk, lane within the chunk
        const uint4 desc = __ldg(&a.chunkDesc[g]);
        const uint8_t* ctext = a.b.text + ((static_cast<uint64_t>(desc.y) << 32) | desc.x);  // text of the chunk start
        const uint32_t base = desc.z;    // contig-local position of the chunk start
        const uint32_t nValid = desc.w;  // positions of the contig that can start a K-mer
        const uint32_t p0 = base + cl * NPT;

        uint32_t A[NWA];
        {
            const uint4* gp = reinterpret_cast<const uint4*>(ctext + cl * NPT);
#pragma unroll
            for (int v = 0; v < NWA / 4; v++) {
                const uint4 x = __ldg(gp + v);
                A[4 * v] = x.x; A[4 * v + 1] = x.y; A[4 * v + 2] = x.z; A[4 * v + 3] = x.w;
            }
        }
        if (filterPending) {
            mbar_wait(barAddr, 0);
            filterPending = false;
        }
        uint32_t slots[NPT];
        uint32_t hit = probe_lane16<KW, FILTER>(a.ix, sFilter, A, slots);
        // positions past the last valid one never probe (position p = bit 15 - p)
        if (p0 + NPT > nValid) hit &= p0 >= nValid ? 0u : (0xffffu << (p0 + NPT - nValid)) & 0xffffu;
        if (!valid) hit = 0;

        uint32_t mask = 0;
        if (hit) {
            const uint64_t lock = a.b.contigLock ? a.b.contigLock[a.chunkContig[g]] : NO_LOCK;
            const uint64_t swStop = lock != NO_LOCK ? lock : a.ix.pos1;
            do {
                const int bit = 31 - __clz(hit);
                hit &= ~(1u << bit);
                const int p = 15 - bit;
                const uint32_t enc = verify_candidate<KW>(a.ix, A, p, select16(slots, p), swStop);
                if (enc) {
                    a.scratchEnc[static_cast<uint64_t>(g) * PROBE_CHUNK + cl * NPT + p] = enc;
                    mask |= 1u << p;
                }
            } while (hit);
        }
        if (valid) a.seedMask[static_cast<uint64_t>(g) * 32 + cl] = static_cast<uint16_t>(mask);
        // seeds per chunk: one atomic per distinct chunk in the warp (all 32 lanes share it without the fast path)
        const uint32_t cnt = __popc(mask);
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? g : 0xffffffffu);
        uint32_t sum = cnt;
        if (peers == 0xffffffffu) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        } else {
            sum = 0;
            for (unsigned rest = peers; rest; rest &= rest - 1) sum += __shfl_sync(peers, cnt, __ffs(rest) - 1);
        }
        if (valid && sum && lane == __ffs(peers) - 1) atomicAdd(&a.chunkCount[g], sum);
    }
    // a warp without items never reached the wait in the loop: no thread may leave the CTA while the bulk copy into its
    // shared memory is still in flight (the async proxy would write into memory of a retired CTA)
    if (FILTER && filterPending) mbar_wait(barAddr, 0);
}

// ------------------------------------------------------------------------------------------------------
// Fast path for near-identical texts (the README's 1k E. coli regime): most 512-position chunks of a target
// are byte-identical to a stretch of the reference, and probing identical bytes gives identical answers.
//
//   memo    (per index state) the reference buffer is probed once as if it were a text, window test off
//           ("self-probe").  memoX has one bit per 16 reference positions: 0 = probing those 16 positions
//           gives exactly the default answer -- one seed, at the sampled (16-aligned) position, pointing at
//           itself; 1 = anything else (a repeat found elsewhere, an overwritten or unsampled slot).
//   k_anchor  one LANE per anchored chunk (every 8th chunk and the first chunk of a contig) probes the chunk's
//           first 16 positions in full; the first verified seed proposes a diagonal d (reference position -
//           target position) for that chunk and the chunks up to the next anchored one.  A proposal is only a
//           guess -- k_clean verifies every byte -- so a wrong one (an indel in between) merely sends chunks
//           to the full probe.
//   k_clean   one WARP per chunk: if the chunk's 512 + K-1 bytes equal the reference on diagonal d, all
//           covered memo bits are 0 and the window class is uniform, then the chunk's seeds are exactly the 32
//           positions with (i + d) % 16 == 0 and entries (i + d) >> 4: it leaves as ONE 32-seed segment.
//           Everything else goes to the work list of k_probe.  A chunk is only ever declared clean when the
//           full probe would provably return the same seeds, so the seed-run list downstream is identical.
//
// Needs k1 == 16 (Exp variant, MBGC's only setting) and a reference buffer that has not wrapped; the host
// falls back to the full probe of every chunk otherwise.
// ------------------------------------------------------------------------------------------------------
template <int KW>
__global__ void __launch_bounds__(256) k_anchor(const ProbeArgs a) {
    griddep_launch_dependents();
    griddep_wait();
    constexpr int NPT = PROBE_NPT;
    constexpr int NWA = probe_words(KW);
    // anchored chunks: every ANCHOR_STRIDE-th chunk, and the first chunk of every contig
    const uint32_t nStride = (a.b.nTiles + ANCHOR_STRIDE - 1) / ANCHOR_STRIDE;
    const uint32_t nItems = nStride + static_cast<uint32_t>(a.b.nContigs);
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < nItems; it += gridDim.x * blockDim.x) {
        uint32_t g;
        if (it < nStride) {
            g = it * ANCHOR_STRIDE;
        } else {
            const uint32_t c = it - nStride;
            g = a.b.tileStart[c];
            if (g == a.b.tileStart[c + 1] || g % ANCHOR_STRIDE == 0) continue;  // no chunk, or covered above
        }
        const uint4 desc = __ldg(&a.chunkDesc[g]);
        const uint8_t* ctext = a.b.text + ((static_cast<uint64_t>(desc.y) << 32) | desc.x);
        const uint32_t base = desc.z, nValid = desc.w;
        uint32_t A[NWA];
        {
            const uint4* gp = reinterpret_cast<const uint4*>(ctext);
#pragma unroll
            for (int v = 0; v < NWA / 4; v++) {
                const uint4 x = __ldg(gp + v);
                A[4 * v] = x.x; A[4 * v + 1] = x.y; A[4 * v + 2] = x.z; A[4 * v + 3] = x.w;
            }
        }
        uint32_t slots[NPT];
        uint32_t hit = probe_lane16<KW, false>(a.ix, nullptr, A, slots);
        if (base + NPT > nValid) hit &= base >= nValid ? 0u : (0xffffu << (base + NPT - nValid)) & 0xffffu;
        unsigned long long an = 0;
        if (hit) {
            const uint64_t lock = a.b.contigLock ? a.b.contigLock[a.chunkContig[g]] : NO_LOCK;
            const uint64_t swStop = lock != NO_LOCK ? lock : a.ix.pos1;
            do {
                const int bit = 31 - __clz(hit);
                hit &= ~(1u << bit);
                const int p = 15 - bit;
                const uint32_t enc = verify_candidate<KW>(a.ix, A, p, select16(slots, p), swStop);
                if (enc) {
                    an = (static_cast<unsigned long long>(p) << 32) | enc;
                    break;
                }
            } while (hit);
        }
        // the proposal for this chunk and the following ones up to the next anchored chunk (same contig): reference
        // position of the chunk start on the anchor's diagonal, 0 = none
        const int64_t r0a = an ? static_cast<int64_t>(decode_pos(static_cast<uint32_t>(an), a.ix.k1ord)) -
                                     static_cast<int64_t>(an >> 32)
                               : 0;
        const uint32_t cEnd = a.b.tileStart[a.chunkContig[g] + 1];
        const uint32_t hi = min(cEnd, (g & ~(ANCHOR_STRIDE - 1)) + ANCHOR_STRIDE);
        for (uint32_t j = g; j < hi; j++) {
            const int64_t r = r0a + static_cast<int64_t>(j - g) * PROBE_CHUNK;
            a.anchor[j] = an && r >= 1 ? static_cast<unsigned long long>(r) : 0ull;
        }
    }
}

struct CleanArgs {
    IndexState ix;
    BatchDesc b;
    const uint4* chunkDesc;
    const uint32_t* chunkContig;
    const unsigned long long* anchor;
    const uint32_t* memoX;
    const uint32_t* memoE;
    uint64_t memoLen;
    uint32_t* chunkCount;
    uint16_t* seedMask;
    uint32_t* scratchEnc;
    uint8_t* chunkSegN;
    uint2* chunkSeg;
    uint32_t* dirtyMask;
    uint32_t* workList;
    uint32_t* laneList;
    uint32_t* counters;
};

// LPC lanes per chunk, 32 / LPC chunks per warp: the per-chunk scalar work (descriptor, diagonal, window class, memo
// window, the stores) is ~250 instructions that every lane of the chunk repeats, so fewer lanes per chunk means fewer
// issue slots per chunk; each lane compares 32 / LPC pieces of 16 bytes, interleaved so that the lanes of a chunk read
// 16 LPC contiguous bytes per load instruction (whole sectors as long as LPC >= 2).
template <int LPC, int OCC>
__global__ void __launch_bounds__(256, OCC) k_clean(const CleanArgs a) {
    griddep_launch_dependents();
    griddep_wait();
    constexpr int CPW = 32 / LPC;  // chunks per warp
    constexpr int PPL = 32 / LPC;  // 16-byte pieces per lane
    constexpr uint32_t GMASK = LPC == 32 ? 0xffffffffu : (1u << (LPC & 31)) - 1u;
    const int lane = threadIdx.x & 31, sub = lane & (LPC - 1), cg = lane / LPC;
    const int K = a.ix.K;
    const int tailWords = (K + 2) / 4;  // the K-1 bytes the chunk's last k-mers reach past it, rounded up to words
    const int span = (K + 30) / 16;     // 16-byte pieces the 16 k-mers of a lane touch
    const unsigned long long tailPieces = ((1ull << ((K + 14) / 16)) - 1ull) << 32;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    const uint32_t nTiles = a.b.nTiles;
    for (uint32_t gb = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * CPW; gb < nTiles; gb += warpsTotal * CPW) {
        const bool valid = gb + cg < nTiles;
        const uint32_t g = valid ? gb + cg : nTiles - 1;
        const uint4 desc = __ldg(&a.chunkDesc[g]);
        const uint32_t base = desc.z, nValid = desc.w;
        // diagonal proposal (k_anchor): reference position of the chunk start, 0 = none
        const int64_t r0 = static_cast<int64_t>(a.anchor[g]);
        uint64_t lock = NO_LOCK;
        if (a.b.contigLock) lock = __ldg(&a.b.contigLock[__ldg(&a.chunkContig[g])]);
        bool ok = valid && r0 >= 1 && base + PROBE_CHUNK <= nValid;
        ok = ok && static_cast<uint64_t>(r0) + PROBE_CHUNK + K - 1 <= a.memoLen;
        const uint32_t pin = (16u - (static_cast<uint32_t>(r0) & 15u)) & 15u;  // first position with (i + d) % 16 == 0
        if (ok) {
            const uint64_t swStop = lock != NO_LOCK ? lock : a.ix.pos1;
            const uint64_t sFirst = static_cast<uint64_t>(r0) + pin, sLast = sFirst + 31 * 16;
            int c0, c1;
            const bool w0 = window_ok(sFirst, K, a.ix.pos1, swStop, c0);
            const bool w1 = window_ok(sLast, K, a.ix.pos1, swStop, c1);
            ok = w0 && w1 && c0 == c1;  // both window predicates are monotone in s: uniform class in between
        }
        const uint8_t* T = a.b.text + ((static_cast<uint64_t>(desc.y) << 32) | desc.x);
        const uint8_t* R = ok ? a.ix.ref + r0 : T;  // chunks that already failed compare the text with itself
        // memo window: bit 0 = the partial group before the first grid position, bits 1..32 = the groups of the
        // chunk's 32 grid positions
        const uint64_t grp0 = ok ? (static_cast<uint64_t>(r0) + pin) / 16 : 1;
        const uint64_t wi = (grp0 - 1) >> 5;
        const uint32_t ws = static_cast<uint32_t>(grp0 - 1) & 31u;
        const uint64_t xw = (static_cast<uint64_t>(__ldg(&a.memoX[wi + 1])) << 32 | __ldg(&a.memoX[wi])) >> ws;
        const uint64_t ew = (static_cast<uint64_t>(__ldg(&a.memoE[wi + 1])) << 32 | __ldg(&a.memoE[wi])) >> ws;

        // 512 + K-1 bytes of text against the reference on the diagonal.  The reference side is unaligned: three
        // 8-byte loads cover a 16-byte piece (24-byte window from the 8-byte boundary below it), a select picks the
        // 5 words the piece touches and a funnel shift per word removes the byte offset.  (Five 4-byte loads would
        // make every lane group request the same sectors five times; one aligned 16-byte load per piece with the
        // neighbouring piece taken from the next lane by shuffle halves the sector count again but costs more issue
        // slots than it saves: measured slower, profiles/README.md r02c.)
        const uintptr_t ra = reinterpret_cast<uintptr_t>(R + 16 * sub);
        const uint2* rp = reinterpret_cast<const uint2*>(ra & ~static_cast<uintptr_t>(7));
        const bool hiHalf = (ra & 4) != 0;
        const uint32_t sh = static_cast<uint32_t>(ra & 3) * 8;
        const uint4* tp = reinterpret_cast<const uint4*>(T) + sub;
        // which of the chunk's 32 + 2 sixteen-byte pieces differ: lane sub compares pieces sub, sub + LPC, sub + 2 LPC, ...
        uint32_t pflags = 0, pt = 0;  // bit k of pflags: this lane's k-th piece differs
#pragma unroll
        for (int k = 0; k < PPL; k++) {
            const uint4 tw = __ldg(tp + LPC * k);
            const uint2 a0 = __ldg(rp + 2 * LPC * k), a1 = __ldg(rp + 2 * LPC * k + 1), a2 = __ldg(rp + 2 * LPC * k + 2);
            const uint32_t w0 = hiHalf ? a0.y : a0.x, w1 = hiHalf ? a1.x : a0.y, w2 = hiHalf ? a1.y : a1.x,
                           w3 = hiHalf ? a2.x : a1.y, w4 = hiHalf ? a2.y : a2.x;
            const uint32_t pd = (tw.x ^ __funnelshift_r(w0, w1, sh)) | (tw.y ^ __funnelshift_r(w1, w2, sh)) |
                                (tw.z ^ __funnelshift_r(w2, w3, sh)) | (tw.w ^ __funnelshift_r(w3, w4, sh));
            pflags |= (pd != 0 ? 1u : 0u) << k;
        }
        {
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(R) & ~static_cast<uintptr_t>(3));
#pragma unroll
            for (int w = 0; w < 16 / LPC; w++) {  // tail: words sub, sub + LPC, ... (at most 14: K <= 56)
                const int tl = sub + LPC * w;
                if (tl < tailWords) {
                    const uint32_t* rt = rw + PROBE_CHUNK / 4 + tl;  // R + PROBE_CHUNK + 4 tl, rounded down to a word
                    pt |= __ldg(reinterpret_cast<const uint32_t*>(T + PROBE_CHUNK) + tl) ^ __funnelshift_r(__ldg(rt), __ldg(rt + 1), sh);
                }
            }
        }
        // Most warps see no difference at all: one vote settles that, and everything that deals with dirty lanes is
        // behind warp-uniform branches.
        const uint32_t X1 = static_cast<uint32_t>(xw >> 1), X0 = static_cast<uint32_t>(xw);
        uint32_t L = ok ? (X1 | (pin ? X0 : 0u)) : 0xffffffffu;  // dirty lanes: memo first (lane l overlaps the reference
                                                                 // groups of bits l, before its grid position, and l + 1)
        if (!valid) L = 0;
        if (__any_sync(0xffffffffu, (pflags | pt) != 0)) {
            unsigned long long pieces = 0;  // bit q = piece q differs (pieces 32.. = the K-1 bytes past the chunk)
#pragma unroll
            for (int k = 0; k < PPL; k++) {
                const unsigned bal = (__ballot_sync(0xffffffffu, (pflags >> k) & 1u) >> (LPC * cg)) & GMASK;
                pieces |= static_cast<unsigned long long>(bal) << (LPC * k);
            }
            const unsigned balT = (__ballot_sync(0xffffffffu, pt != 0) >> (LPC * cg)) & GMASK;
            if (balT) pieces |= tailPieces;  // any difference in the tail dirties all tail pieces (conservative)
            // lane l of the chunk (positions 16 l .. 16 l + 15) reads bytes [16 l, 16 l + 15 + K): pieces l .. l + span - 1
            unsigned long long dm = 0;
            for (int j = 0; j < span; j++) dm |= pieces >> j;
            if (valid) L |= static_cast<uint32_t>(dm);
        }
        if (__any_sync(0xffffffffu, L != 0)) {
            // dirty lanes -> work items.  One 64-bit atomic per warp reserves the slots of its chunks in both lists
            // (counters 6 = chunks with dirty lanes, 7 = dirty lanes), then the lanes of a chunk write its items
            // (PPL bits of L each).
            uint32_t laneBase = 0, chunkBase = 0, laneTot = 0, chunkTot = 0;
#pragma unroll
            for (int j = 0; j < CPW; j++) {
                const uint32_t Lj = __shfl_sync(0xffffffffu, L, LPC * j);
                if (j == cg) {
                    laneBase = laneTot;
                    chunkBase = chunkTot;
                }
                laneTot += __popc(Lj);
                chunkTot += Lj ? 1u : 0u;
            }
            unsigned long long old = 0;
            if (lane == 0)
                old = atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 6),
                                (static_cast<unsigned long long>(laneTot) << 32) | chunkTot);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (L) {
                if (sub == 0) a.workList[static_cast<uint32_t>(old) + chunkBase] = g;
                constexpr uint32_t PMASK = PPL == 32 ? 0xffffffffu : (1u << (PPL & 31)) - 1u;
                uint32_t mine = (L >> (PPL * sub)) & PMASK;
                uint32_t k = static_cast<uint32_t>(old >> 32) + laneBase + __popc(L & ((1u << (PPL * sub)) - 1u));
                while (mine) {
                    a.laneList[k++] = g * 32u + PPL * sub + (__ffs(mine) - 1);
                    mine &= mine - 1;
                }
            }
        }
        if (sub == 0 && valid) {
            // seeds of the clean lanes = their grid position, where the memo group is in the default state
            const uint32_t D = ~static_cast<uint32_t>(ew >> 1) & ~L;
            a.dirtyMask[g] = L;
            a.chunkSegN[g] = L == 0xffffffffu ? 0 : static_cast<uint8_t>(0x80u | pin);
            a.chunkSeg[g] = make_uint2(static_cast<uint32_t>(grp0), D);
            // clean chunk: maximal stretches of D leave as segments; mixed chunk: every seed on its own (k_probe adds
            // the seeds of the dirty lanes)
            a.chunkCount[g] = L == 0 ? __popc(D & ~(D << 1)) : __popc(D);
        }
    }
}

// memo bits from a finished self-probe pass: the "text" is the reference buffer from position 0, so chunk g,
// lane l is the group of 16 reference positions number 32 g + l.  A group is
//   default  one seed, at the grid position, pointing at itself            (X = 0, E = 0)
//   empty    no seed at all: its sampled k-mer lost the table slot          (X = 0, E = 1)
//   other    anything else, or not completely probed (end of the reference) (X = 1)
__global__ void __launch_bounds__(256) k_memo(uint32_t nTiles, uint32_t nValid, const uint16_t* __restrict__ seedMask,
                                              const uint32_t* __restrict__ scratchEnc, uint32_t* __restrict__ memoX,
                                              uint32_t* __restrict__ memoE) {
    griddep_launch_dependents();
    griddep_wait();
    const int lane = threadIdx.x & 31;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < nTiles; g += warpsTotal) {
        const uint32_t m = seedMask[static_cast<uint64_t>(g) * 32 + lane];
        const uint32_t grp = g * 32 + lane;
        const bool probed = static_cast<uint64_t>(grp) * 16 + 16 <= nValid;
        bool def = false;
        if (m == 1u) def = scratchEnc[static_cast<uint64_t>(g) * PROBE_CHUNK + lane * PROBE_NPT] == grp;
        const bool empty = m == 0u;
        const unsigned x = __ballot_sync(0xffffffffu, !probed || !(def || empty));
        const unsigned e = __ballot_sync(0xffffffffu, empty);
        if (lane == 0) {
            memoX[g] = x;
            memoE[g] = e;
        }
    }
}

bool probe_supports_K(int K) {
    switch (K) {
        case 12: case 16: case 20: case 24: case 28: case 32: case 36: case 40: case 44: case 56: return true;
        default: return false;
    }
}

// per device: one process may hold matchers on several GPUs (mbgcg_create_multi), each driven by its own host thread
static std::atomic<int> g_smCount[64];
int device_sm_count() {
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev & 63;
    int n = g_smCount[slot].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        g_smCount[slot].store(n, std::memory_order_relaxed);
    }
    return n;
}

// ------------------------------------------------------------------------------------------------------
// Text transfer: blocks the host packed 4 bases per byte (hostpack.cpp) expand back to the exact ASCII bytes.
// One thread per 16 output bytes = one packed 32-bit word; a byte permute maps four 2-bit codes to "ACTG"[code]
// at once.  Bytes that are not A/C/G/T were sent as exception runs and are patched in afterwards.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_blocks(const PackedBlock* __restrict__ blocks, uint32_t nBlocks,
                                                       const uint8_t* __restrict__ packed, uint8_t* __restrict__ text) {
    for (uint32_t b = blockIdx.x; b < nBlocks; b += gridDim.x) {
        const PackedBlock pb = blocks[b];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(packed + pb.srcOff);
        uint8_t* dst = text + pb.dstOff;
        const uint32_t groups = (pb.n + 15) / 16;
        for (uint32_t g = threadIdx.x; g < groups; g += blockDim.x) {
            const uint32_t w = __ldg(src + g);
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t q = (w >> (8 * j)) & 0xffu;  // 4 codes, first base in the lowest bits
                const uint32_t sel = (q & 3u) | ((q & 0xcu) << 2) | ((q & 0x30u) << 4) | ((q & 0xc0u) << 6);
                o[j] = __byte_perm(0x47544341u, 0u, sel);   // code 0 'A', 1 'C', 2 'T', 3 'G'
            }
            if (g * 16 + 16 <= pb.n) {
                reinterpret_cast<uint4*>(dst)[g] = make_uint4(o[0], o[1], o[2], o[3]);
            } else {
                for (uint32_t k = g * 16; k < pb.n; k++) dst[k] = static_cast<uint8_t>(o[(k & 15) >> 2] >> (8 * (k & 3)));
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_patch_exceptions(const ExcRunDev* __restrict__ runs, uint32_t nRuns,
                                                          uint8_t* __restrict__ text) {
    const int lane = threadIdx.x & 31;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nRuns; r += warpsTotal) {
        const ExcRunDev e = runs[r];
        const uint64_t step = (e.byte >> 8) ? (e.byte >> 8) : 1u;  // bits 8-31: distance between the bytes of the run (line ends)
        for (uint32_t k = lane; k < e.len; k += 32) text[e.off + k * step] = static_cast<uint8_t>(e.byte);
    }
}

// a whole file image that travelled 4 bases per byte (hostpack.cpp: pack_fasta_image): 16 bytes of text per thread
__global__ void __launch_bounds__(256) k_unpack_image(const uint32_t* __restrict__ packed, uint64_t n, uint8_t* __restrict__ text) {
    const uint64_t groups = (n + 15) / 16;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups;
         g += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t w = __ldg(packed + g);
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t q = (w >> (8 * j)) & 0xffu;
            const uint32_t sel = (q & 3u) | ((q & 0xcu) << 2) | ((q & 0x30u) << 4) | ((q & 0xc0u) << 6);
            o[j] = __byte_perm(0x47544341u, 0u, sel);
        }
        if (g * 16 + 16 <= n) {
            reinterpret_cast<uint4*>(text)[g] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (uint64_t k = g * 16; k < n; k++) text[k] = static_cast<uint8_t>(o[(k & 15) >> 2] >> (8 * (k & 3)));
        }
    }
}

int launch_unpack_image(const uint8_t* packed, uint64_t n, const ExcRunDev* runs, uint32_t nRuns, uint8_t* text, cudaStream_t st) {
    int launches = 0;
    if (n) {
        const uint64_t blocks = ((n + 15) / 16 + 255) / 256;
        const uint64_t cap = static_cast<uint64_t>(device_sm_count()) * 16;
        k_unpack_image<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(packed), n, text);
        launches++;
    }
    if (nRuns) {
        uint32_t blocksN = (nRuns + 7) / 8;
        const uint32_t cap = static_cast<uint32_t>(device_sm_count()) * 8;
        k_patch_exceptions<<<blocksN < cap ? blocksN : cap, 256, 0, st>>>(runs, nRuns, text);
        launches++;
    }
    return launches;
}

int launch_unpack(const PackedBlock* blocks, uint32_t nBlocks, const uint8_t* packed, const ExcRunDev* runs, uint32_t nRuns,
                  uint8_t* text, cudaStream_t st) {
    int launches = 0;
    if (nBlocks) {
        const uint32_t cap = static_cast<uint32_t>(device_sm_count()) * 8;
        k_unpack_blocks<<<nBlocks < cap ? nBlocks : cap, 256, 0, st>>>(blocks, nBlocks, packed, text);
        launches++;
    }
    if (nRuns) {
        uint32_t blocksN = (nRuns + 7) / 8;
        const uint32_t cap = static_cast<uint32_t>(device_sm_count()) * 8;
        k_patch_exceptions<<<blocksN < cap ? blocksN : cap, 256, 0, st>>>(runs, nRuns, text);
        launches++;
    }
    return launches;
}

// chunk -> contig table and self-contained chunk descriptors (one binary search per chunk, once per pass)
__global__ void __launch_bounds__(256) k_chunk_contig(BatchDesc b, int K, uint32_t* __restrict__ chunkContig,
                                                      uint4* __restrict__ chunkDesc, uint32_t* __restrict__ chunkCount) {
    griddep_launch_dependents();
    griddep_wait();
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < b.nTiles; g += gridDim.x * blockDim.x) {
        const uint32_t c = static_cast<uint32_t>(upper_contig(b.tileStart, b.nContigs, g));
        chunkContig[g] = c;
        chunkCount[g] = 0;  // k_clean stores, k_probe adds
        const uint32_t base = (g - b.tileStart[c]) * PROBE_CHUNK;
        const uint64_t off = b.contigOff[c] + base;
        const uint32_t N2 = b.contigLen[c];
        chunkDesc[g] = make_uint4(static_cast<uint32_t>(off), static_cast<uint32_t>(off >> 32), base,
                                  N2 >= static_cast<uint32_t>(K) ? N2 - K + 1 : 0);
    }
}

template <int KW>
static void launch_probe_k(const ProbeArgs& a, bool useFilter, int mode, cudaStream_t st) {
    const int sm = device_sm_count();
    if (mode == 1) {  // anchors: one lane per anchored chunk
        uint32_t blocks = (a.b.nTiles / ANCHOR_STRIDE + a.b.nContigs + 256) / 256;
        const uint32_t cap = static_cast<uint32_t>(sm) * 8;
        if (blocks > cap) blocks = cap;
        launch_pdl(k_anchor<KW>, blocks, 256, 0, st, a);
        return;
    }
    if (useFilter) {
        const size_t smem = static_cast<size_t>(a.ix.filterWords) * 4 + 16;  // filter + mbarrier
        // the opt-in applies to the device that is current when it is made: once per (device, instantiation)
        static std::atomic<uint64_t> attrSet{0};
        int dev = 0;
        cudaGetDevice(&dev);
        const uint64_t bit = 1ull << (dev & 63);
        if (!(attrSet.load(std::memory_order_acquire) & bit)) {
            cudaFuncSetAttribute(k_probe<KW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(PROBE_FILTER_BYTES + 16));
            attrSet.fetch_or(bit, std::memory_order_release);
        }
        launch_pdl(k_probe<KW, true>, sm, PROBE_THREADS_FILTER, smem, st, a);
    } else {
        uint32_t blocks = (a.b.nTiles + PROBE_THREADS_PLAIN / 32 - 1) / (PROBE_THREADS_PLAIN / 32);
        const uint32_t cap = static_cast<uint32_t>(sm) * 2;
        if (blocks > cap) blocks = cap;
        launch_pdl(k_probe<KW, false>, blocks, PROBE_THREADS_PLAIN, 0, st, a);
    }
}

static int dispatch_probe(const ProbeArgs& a, bool useFilter, int mode, cudaStream_t st) {
#define MBGC_PROBE_CASE(KK) \
    case KK: launch_probe_k<KK / 4>(a, useFilter, mode, st); break;
    switch (a.ix.K) {
        MBGC_PROBE_CASE(12) MBGC_PROBE_CASE(16) MBGC_PROBE_CASE(20) MBGC_PROBE_CASE(24) MBGC_PROBE_CASE(28)
        MBGC_PROBE_CASE(32) MBGC_PROBE_CASE(36) MBGC_PROBE_CASE(40) MBGC_PROBE_CASE(44) MBGC_PROBE_CASE(56)
        default: return -1;
    }
#undef MBGC_PROBE_CASE
    return 1;
}

int launch_chunk_desc(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st) {
    if (b.nTiles == 0) {
        cudaMemsetAsync(w.tileCount, 0, sizeof(uint32_t), st);
        return 0;
    }
    launch_pdl(k_chunk_contig, device_sm_count() * 4, 256, 0, st, b, ix.K, w.chunkContig, w.chunkDesc, w.tileCount);
    return 1;
}

int launch_probe(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, bool useFilter, bool useWorkList,
                 cudaStream_t st) {
    if (b.nTiles == 0) return 0;
    ProbeArgs a{ix, b, w.chunkDesc, w.chunkContig, w.tileCount, w.seedMask, w.scratchEnc,
                useWorkList ? w.laneList : nullptr, useWorkList ? w.counters + 7 : nullptr, w.anchor};
    return dispatch_probe(a, useFilter, 0, st);
}

int launch_anchor(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st) {
    if (b.nTiles == 0) return 0;
    ProbeArgs a{ix, b, w.chunkDesc, w.chunkContig, w.tileCount, w.seedMask, w.scratchEnc, nullptr, nullptr, w.anchor};
    return dispatch_probe(a, false, 1, st);
}

int launch_clean(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, const MemoState& memo, cudaStream_t st) {
    if (b.nTiles == 0) return 0;
    CleanArgs a{ix, b, w.chunkDesc, w.chunkContig, w.anchor, memo.memoX, memo.memoE, memo.memoLen, w.tileCount,
                w.seedMask, w.scratchEnc, w.chunkSegN, w.chunkSeg, w.dirtyMask, w.workList, w.laneList, w.counters};
    // 8 lanes per chunk, 8 CTAs per SM: 4 or 2 lanes per chunk execute 30-50 % fewer instructions per chunk and were measured
    // 25 % / 2.2x slower, 6 CTAs per SM with 40 registers 6 % slower (profiles/README.md, r02c experiments)
    launch_pdl(k_clean<8, 8>, device_sm_count() * 8, 256, 0, st, a);
    return 1;
}

int launch_memo(const BatchDesc& b, const WorkBuffers& w, uint32_t nValid, uint32_t* memoX, uint32_t* memoE,
                cudaStream_t st) {
    if (b.nTiles == 0) return 0;
    launch_pdl(k_memo, device_sm_count() * 4, 256, 0, st, b.nTiles, nValid, w.seedMask, w.scratchEnc, memoX, memoE);
    return 1;
}

// ------------------------------------------------------------------------------------------------------
// seed compaction + run detection
// ------------------------------------------------------------------------------------------------------

// position-indexed seed staging -> dense, position-ordered segment arrays.  chunkOff = exclusive scan of the
// chunk counts.  One warp per chunk that has probed lanes (all chunks, or the fast path's list of chunks with at
// least one dirty lane).  In such a "mixed" chunk the clean lanes contribute their memo seed (k_clean: mask D, one
// grid position per lane) and the dirty lanes what k_probe found; every seed leaves as its own segment.
__global__ void __launch_bounds__(256) k_compact(BatchDesc b, const uint32_t* __restrict__ chunkContig,
                                                 const uint32_t* __restrict__ chunkOff,
                                                 const uint16_t* __restrict__ seedMask,
                                                 const uint32_t* __restrict__ scratchEnc,
                                                 const uint32_t* __restrict__ workList, const uint32_t* workCount,
                                                 const uint32_t* __restrict__ dirtyMask,
                                                 const uint8_t* __restrict__ chunkSegN, const uint2* __restrict__ chunkSeg,
                                                 SeedRec* __restrict__ seeds, uint32_t seedCap, uint32_t* counters) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t nSeeds = chunkOff[b.nTiles];
    if (nSeeds > seedCap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) counters[3] = 1;  // overflow: host retries with more room
        return;
    }
    const int lane = threadIdx.x & 31;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    const uint32_t nWork = workList ? *workCount : b.nTiles;
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nWork; w += warpsTotal) {
        const uint32_t g = workList ? workList[w] : w;
        const uint32_t o = chunkOff[g];
        if (chunkOff[g + 1] == o) continue;
        const bool dirty = workList ? ((dirtyMask[g] >> lane) & 1u) != 0 : true;
        uint32_t m, memoEnc = 0;
        if (dirty) {
            m = seedMask[static_cast<uint64_t>(g) * 32 + lane];
        } else {
            const uint2 sg = chunkSeg[g];  // .x = entry of lane 0's grid position, .y = lanes whose grid position is a seed
            m = ((sg.y >> lane) & 1u) << (chunkSegN[g] & 15u);
            memoEnc = sg.x + lane;
        }
        const uint32_t c = chunkContig[g];
        const uint32_t base = (g - b.tileStart[c]) * PROBE_CHUNK + lane * PROBE_NPT;
        const uint64_t src = static_cast<uint64_t>(g) * PROBE_CHUNK + lane * PROBE_NPT;
        const uint32_t cnt = __popc(m);
        uint32_t k = o + warp_incl_scan(cnt, lane) - cnt;
        while (m) {
            const int p = __ffs(m) - 1;
            m &= m - 1;
            SeedRec rec;
            rec.pos = base + p;
            rec.enc = dirty ? scratchEnc[src + p] : memoEnc;
            rec.n = 1;
            rec.contig = c;
            seeds[k++] = rec;
        }
    }
}

// clean chunks (k_clean, no dirty lane): one THREAD per chunk.  The chunk's seeds sit on one diagonal, k1 apart, at
// most one per lane-sized group of 16 positions; D has a bit for every group with a seed, and every maximal stretch
// of set bits leaves as one segment.
__global__ void __launch_bounds__(256) k_compact_clean(BatchDesc b, const uint32_t* __restrict__ chunkContig,
                                                       const uint32_t* __restrict__ chunkOff,
                                                       const uint32_t* __restrict__ dirtyMask,
                                                       const uint8_t* __restrict__ chunkSegN,
                                                       const uint2* __restrict__ chunkSeg, SeedRec* __restrict__ seeds,
                                                       uint32_t seedCap) {
    griddep_launch_dependents();
    griddep_wait();
    if (chunkOff[b.nTiles] > seedCap) return;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < b.nTiles; g += gridDim.x * blockDim.x) {
        if (dirtyMask[g]) continue;
        const uint32_t pin = chunkSegN[g] & 15u;
        const uint2 sg = chunkSeg[g];  // .x = entry of group 0's seed, .y = D
        const uint32_t c = chunkContig[g];
        const uint32_t base = (g - b.tileStart[c]) * PROBE_CHUNK + pin;
        uint32_t k = chunkOff[g];
        unsigned D = sg.y;
        while (D) {
            const int l = __ffs(D) - 1;
            // consecutive groups with a seed from l on (64-bit so that 32 ones give 32)
            const int n = __ffsll(static_cast<long long>(~static_cast<unsigned long long>(D >> l))) - 1;
            SeedRec rec;
            rec.pos = base + 16u * l;
            rec.enc = sg.x + l;
            rec.n = static_cast<uint32_t>(n);
            rec.contig = c;
            seeds[k++] = rec;
            D = l + n >= 32 ? 0u : D & (~0u << (l + n));
        }
    }
}

// head[j] = 1 iff segment j starts a run: first of its contig, or not (previous pos + k1, previous enc + step),
// or a different window class (the class fixes the extension bounds tmpStart1/tmpEnd1, .cpp:221-222).
// Seeds only collapse when their k-mers touch or overlap (k1 <= K); otherwise (protein profile: K = 12 < k1 = 16)
// the bases between two seeds are unverified and every seed is its own run.  One thread per segment.
__global__ void __launch_bounds__(256) k_heads(IndexState ix, BatchDesc b, const uint32_t* __restrict__ chunkOff,
                                               const SeedRec* __restrict__ seeds, uint32_t* __restrict__ head,
                                               const uint32_t* counters) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    const uint32_t nSeeds = chunkOff[b.nTiles];
    const uint32_t encStep = ix.k1ord >= 0 ? static_cast<uint32_t>(ix.k1 >> ix.k1ord) : static_cast<uint32_t>(ix.k1);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nSeeds; j += gridDim.x * blockDim.x) {
        uint32_t h = 1;
        const SeedRec cur = seeds[j];
        const SeedRec prev = j ? seeds[j - 1] : cur;
        const uint32_t c = cur.contig;
        if (j != 0 && prev.contig == c) {
            // last seed of the previous segment (a segment = n seeds, spacing k1, entries ascending)
            const uint32_t nm1 = prev.n - 1u;
            const uint32_t e = cur.enc, ep = prev.enc + nm1 * encStep;
            const uint32_t pp = prev.pos + nm1 * static_cast<uint32_t>(ix.k1);
            if (ix.k1 <= ix.K && cur.pos - pp == static_cast<uint32_t>(ix.k1) && e - ep == encStep) {
                const uint64_t lock = b.contigLock ? b.contigLock[c] : NO_LOCK;
                const uint64_t swStop = lock != NO_LOCK ? lock : ix.pos1;
                int c0, c1;
                window_ok(decode_pos(e, ix.k1ord), ix.K, ix.pos1, swStop, c0);
                window_ok(decode_pos(ep, ix.k1ord), ix.K, ix.pos1, swStop, c1);
                h = c0 != c1;
            }
        }
        head[j] = h;
    }
}

// runFirst[run] = first seed of the run; runStart[c] = first run of contig c (runStart[nContigs] = nRuns)
__global__ void __launch_bounds__(256) k_runs(BatchDesc b, const uint32_t* __restrict__ tileOff,
                                              const uint32_t* __restrict__ headScan, uint32_t* __restrict__ runFirst,
                                              uint32_t* __restrict__ runStart, uint32_t* counters) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    const uint32_t nSeeds = tileOff[b.nTiles];
    const uint32_t nRuns = headScan[nSeeds];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (uint32_t j = gtid; j < nSeeds; j += gsz)
        if (headScan[j + 1] != headScan[j]) runFirst[headScan[j]] = j;
    for (uint32_t c = gtid; c <= static_cast<uint32_t>(b.nContigs); c += gsz) {
        const uint32_t s = c < static_cast<uint32_t>(b.nContigs) ? tileOff[b.tileStart[c]] : nSeeds;
        runStart[c] = headScan[s];  // headScan[nSeeds] = nRuns
    }
    if (gtid == 0) {
        runFirst[nRuns] = nSeeds;
        counters[0] = nSeeds;
        counters[1] = nRuns;
    }
}

int launch_compact_and_runs(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, bool segments,
                            cudaStream_t st) {
    int launches = 0;
    const int sm = device_sm_count();
    // tileCount -> tileOff (in place), tileOff[nTiles] = nSeeds
    launches += scan_exclusive(w.tileCount, w.tileCount, nullptr, b.nTiles, b.nTiles, w.scanState, w.scanStride, nullptr, st);
    launch_pdl(k_compact, sm * 8, 256, 0, st, b, w.chunkContig, w.tileCount, w.seedMask, w.scratchEnc,
                                      segments ? w.workList : nullptr, segments ? w.counters + 6 : nullptr,
                                      w.dirtyMask, w.chunkSegN, w.chunkSeg, w.seeds, w.seedCap, w.counters);
    if (segments) {
        launch_pdl(k_compact_clean, sm * 8, 256, 0, st, b, w.chunkContig, w.tileCount, w.dirtyMask, w.chunkSegN, w.chunkSeg,
                                                w.seeds, w.seedCap);
        launches += 1;
    }
    launch_pdl(k_heads, sm * 8, 256, 0, st, ix, b, w.tileCount, w.seeds, w.headScan, w.counters);
    launches += 2;
    // head flags -> exclusive scan in place; length lives on the device (tileOff[nTiles])
    launches += scan_exclusive(w.headScan, w.headScan, w.tileCount + b.nTiles, 0, w.seedCap, w.scanState + w.scanStride, w.scanStride, nullptr, st);
    launch_pdl(k_runs, sm * 8, 256, 0, st, b, w.tileCount, w.headScan, w.runFirst, w.runStart, w.counters);
    launches += 1;
    return launches;
}

// ------------------------------------------------------------------------------------------------------
// K4: extend  (.cpp:227-246 right, :257-259 / :278-280 raw left)
//
// Eight lanes per seed run (four runs per warp).  Inside a run every base is already known equal (overlapping
// verified k-mers), so only the stretch left of the first seed and right of the last seed's k-mer is compared:
// 8 lanes x 8 bytes per step, XOR + __clzll/__ffsll for the first differing byte, __ballot_sync for the first
// lane that stops.  These stretches are short by construction (a run's seeds reach to within k1 + K bytes of the
// mismatch that ends it), so one 64-byte step almost always finishes a side.
// If the neighbouring run lies on the same diagonal (a sampled k-mer was overwritten in the table, so the seed
// chain broke although the bases still match) the walk stops at the neighbour's verified span and the run is
// marked linked; k_links resolves link chains, so no stretch of text is ever compared twice.
// ------------------------------------------------------------------------------------------------------
struct ExtendArgs {
    IndexState ix;
    BatchDesc b;
    const SeedRec* seeds;
    const uint32_t* runFirst;
    const uint32_t* runStart;
    RunRec* runs;
    uint8_t* runLink;
    const uint32_t* counters;
};

__global__ void __launch_bounds__(256) k_extend(const ExtendArgs a) {
    griddep_launch_dependents();
    griddep_wait();
    if (a.counters[3]) return;
    const uint32_t nRuns = a.counters[1];
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    const int K = a.ix.K, k1 = a.ix.k1;
    const uint32_t encStep = a.ix.k1ord >= 0 ? static_cast<uint32_t>(k1 >> a.ix.k1ord) : static_cast<uint32_t>(k1);
    const uint8_t* R = a.ix.ref;
    for (uint32_t rb = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; rb < nRuns; rb += warpsTotal * 4) {
        const bool valid = rb + grp < nRuns;
        const uint32_t r = valid ? rb + grp : nRuns - 1;
        const uint32_t j0 = a.runFirst[r], j1 = a.runFirst[r + 1];
        // first / last segment of the run and their neighbours: four 128-bit loads, issued together
        const SeedRec sFirst = a.seeds[j0], sLast = a.seeds[j1 - 1];
        const SeedRec sPrev = a.seeds[j0 ? j0 - 1 : 0], sNext = a.seeds[j1];  // seeds[] has two spare entries at the end
        const uint32_t c = sFirst.contig;
        const uint32_t firstPos = sFirst.pos, enc0 = sFirst.enc;
        const uint32_t lastPos = sLast.pos + (sLast.n - 1u) * static_cast<uint32_t>(k1);
        const uint32_t rs = a.runStart[c], re = a.runStart[c + 1];
        const uint32_t N2 = a.b.contigLen[c];
        const uint8_t* T = a.b.text + a.b.contigOff[c];
        const uint64_t lock = a.b.contigLock ? a.b.contigLock[c] : NO_LOCK;
        const uint64_t swStop = lock != NO_LOCK ? lock : a.ix.pos1;
        const int64_t s0 = static_cast<int64_t>(decode_pos(enc0, a.ix.k1ord));
        const int64_t d = s0 - static_cast<int64_t>(firstPos);  // reference position = target position + d
        int cls;
        window_ok(static_cast<uint64_t>(s0), K, a.ix.pos1, swStop, cls);

        // ---- bounds: the verified span of the nearest run (within LINK_WINDOW) that continues this diagonal.  When the
        // reference holds several similar copies of a region, seeds of their diagonals interleave in position order and
        // the neighbour on the same diagonal is a few runs away, not adjacent; without it every run would walk the whole
        // equal interval again.
        bool leftCand = false;
        uint32_t linkBack = 0, linkAhead = 0;
        int64_t stopL = d < 0 ? -d : 0;  // target position below which nothing is compared (ref index >= 0)
        for (uint32_t k = 1; k <= LINK_WINDOW && r >= rs + k; k++) {
            const SeedRec sp_ = k == 1 ? sPrev : a.seeds[a.runFirst[r - k + 1] - 1];  // last segment of run r - k
            const uint32_t nm1 = sp_.n - 1u;  // its last seed
            const uint32_t pp = sp_.pos + nm1 * static_cast<uint32_t>(k1);
            // interleaved diagonals sit within k1 positions of each other; a run further back than 2 k1 (and all before
            // it: the list is in position order) belongs to another stretch of the text -- stop looking
            if (k > 1 && static_cast<int64_t>(pp) + 2 * k1 < static_cast<int64_t>(firstPos)) break;
            const int64_t sp = static_cast<int64_t>(decode_pos(sp_.enc + nm1 * encStep, a.ix.k1ord));
            if (sp - static_cast<int64_t>(pp) != d) continue;
            int clsP;
            window_ok(static_cast<uint64_t>(sp), K, a.ix.pos1, swStop, clsP);
            if (clsP == cls) {
                leftCand = true;
                linkBack = k;
                stopL = static_cast<int64_t>(pp) + K;  // end of that run's verified span
                break;
            }
        }
        const int64_t refCap = static_cast<int64_t>(a.ix.refLen);
        int64_t stopR = N2;  // exclusive target bound
        if (refCap - d < stopR) stopR = refCap - d;
        bool rightCand = false;
        for (uint32_t k = 1; k <= LINK_WINDOW && r + k < re; k++) {
            const SeedRec sn_ = k == 1 ? sNext : a.seeds[a.runFirst[r + k]];  // first segment of run r + k
            const uint32_t np = sn_.pos;
            if (k > 1 && static_cast<int64_t>(np) > static_cast<int64_t>(lastPos) + 2 * k1) break;
            const int64_t sn = static_cast<int64_t>(decode_pos(sn_.enc, a.ix.k1ord));
            if (sn - static_cast<int64_t>(np) != d) continue;
            int clsN;
            window_ok(static_cast<uint64_t>(sn), K, a.ix.pos1, swStop, clsN);
            if (clsN == cls) {
                if (static_cast<int64_t>(np) <= stopR) {
                    rightCand = true;
                    linkAhead = k;
                    stopR = np;
                }
                break;
            }
        }

        // ---- left: bytes [stopL, firstPos), walking down from firstPos
        int64_t hiEnd = firstPos;
        int64_t mism = stopL - 1;  // position of the first mismatch (or the stop bound - 1)
        bool dataMismatch = false;
        bool act = valid && hiEnd > stopL;
        while (__any_sync(0xffffffffu, act)) {
            int64_t m = INT64_MIN;
            int isData = 0;
            if (act) {
                const int64_t hi_l = hiEnd - 8 * sub, lo_l = hi_l - 8;
                if (hi_l > stopL) {
                    uint64_t diff = ld_u64_unaligned(T + lo_l) ^ ld_u64_unaligned(R + lo_l + d);
                    // bytes below stopL are outside the comparison: drop them from the diff
                    if (lo_l < stopL) diff &= ~0ull << (8 * (stopL - lo_l));
                    if (diff) {
                        m = lo_l + (63 - __clzll(diff)) / 8;
                        isData = 1;
                    } else if (lo_l <= stopL) {
                        m = stopL - 1;
                    }
                } else {
                    m = stopL - 1;
                }
            }
            const unsigned bal = (__ballot_sync(0xffffffffu, m != INT64_MIN) >> (8 * grp)) & 0xffu;
            const int src = bal ? 8 * grp + __ffs(bal) - 1 : lane;
            const int64_t mS = __shfl_sync(0xffffffffu, m, src);
            const int dS = __shfl_sync(0xffffffffu, isData, src);
            if (act) {
                if (bal) {
                    mism = mS;
                    dataMismatch = dS != 0;
                    act = false;
                } else {
                    hiEnd -= 64;
                    act = hiEnd > stopL;
                }
            }
        }
        const bool leftLinked = leftCand && !dataMismatch;
        const uint32_t aFree = static_cast<uint32_t>(mism + 1);

        // ---- right: bytes [lastPos + K, stopR)
        int64_t loBeg = static_cast<int64_t>(lastPos) + K;
        int64_t mismR = stopR;
        bool dataMismatchR = false;
        act = valid && loBeg < stopR;
        while (__any_sync(0xffffffffu, act)) {
            int64_t m = INT64_MIN;
            int isData = 0;
            if (act) {
                const int64_t lo_l = loBeg + 8 * sub, hi_l = lo_l + 8;
                if (lo_l < stopR) {
                    uint64_t diff = ld_u64_unaligned(T + lo_l) ^ ld_u64_unaligned(R + lo_l + d);
                    if (hi_l > stopR) diff &= ~0ull >> (8 * (hi_l - stopR));
                    if (diff) {
                        m = lo_l + (__ffsll(static_cast<long long>(diff)) - 1) / 8;
                        isData = 1;
                    } else if (hi_l >= stopR) {
                        m = stopR;
                    }
                } else {
                    m = stopR;
                }
            }
            const unsigned bal = (__ballot_sync(0xffffffffu, m != INT64_MIN) >> (8 * grp)) & 0xffu;
            const int src = bal ? 8 * grp + __ffs(bal) - 1 : lane;
            const int64_t mS = __shfl_sync(0xffffffffu, m, src);
            const int dS = __shfl_sync(0xffffffffu, isData, src);
            if (act) {
                if (bal) {
                    mismR = mS;
                    dataMismatchR = dS != 0;
                    act = false;
                } else {
                    loBeg += 64;
                    act = loBeg < stopR;
                }
            }
        }
        const bool rightLinked = rightCand && !dataMismatchR;
        if (valid && sub == 0) {
            RunRec rec;
            rec.firstPos = firstPos;
            rec.nSeeds = (lastPos - firstPos) / static_cast<uint32_t>(k1) + 1;
            rec.enc0 = enc0;
            rec.aFree = aFree;
            rec.bFree = static_cast<uint32_t>(mismR);
            rec.flags = (leftLinked ? 1u : 0u) | (rightLinked ? 2u : 0u) | (c << 2);
            a.runs[r] = rec;
            a.runLink[r] = static_cast<uint8_t>((leftLinked ? linkBack : 0u) | ((rightLinked ? linkAhead : 0u) << 4));
        }
    }
}

// Link chains (adjacent runs on one diagonal whose gap compared equal) share one maximal interval: copy the
// chain's left end forward and its right end backward so that K5 reads final values.  Every run belongs to
// exactly one chain and only the thread of the chain end (start) walks it: no races.
__global__ void __launch_bounds__(256) k_links(RunRec* __restrict__ runs, const uint8_t* __restrict__ runLink,
                                               const uint32_t* counters) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    const uint32_t nRuns = counters[1];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < nRuns; r += gridDim.x * blockDim.x) {
        const uint32_t lk = runLink[r];
        const uint32_t back = lk & 15u, ahead = lk >> 4;
        if (!ahead && back) {  // chain end (chains have >= 2 runs): push bFree backward
            const uint32_t b = runs[r].bFree;
            uint32_t q = r, step = back;
            while (step) {  // a left-linked run always has its predecessor `step` runs back
                q -= step;
                runs[q].bFree = b;
                step = runLink[q] & 15u;
            }
        }
        if (!back && ahead) {  // chain start: push aFree forward
            const uint32_t a = runs[r].aFree;
            uint32_t q = r, step = ahead;
            while (step) {
                q += step;
                runs[q].aFree = a;
                step = runLink[q] >> 4;
            }
        }
    }
}

int launch_extend(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st) {
    ExtendArgs a{ix, b, w.seeds, w.runFirst, w.runStart, w.runs, w.runLink, w.counters};
    launch_pdl(k_extend, device_sm_count() * 8, 256, 0, st, a);
    launch_pdl(k_links, device_sm_count() * 4, 256, 0, st, w.runs, w.runLink, w.counters);
    return 2;
}

// ------------------------------------------------------------------------------------------------------
// K5: resolve  (.cpp:250-315)
//
// The reference scans a contig left to right; an accepted match makes the cursor jump to matchEnd -
// skipMargin, and its left edge depends on the stack of matches accepted so far (trim to the previous end,
// or swallow the previous match when the new diagonal also covers it).  Given the run list this is pure
// integer arithmetic: for the first seed at or after the cursor, the maximal equal interval [aFree, bFree)
// of its run replaces every byte comparison of the reference loop.
//
// Parallelism: ISLANDS.  Let P be a target position that no run's equal interval contains in its interior
// (bFree <= P for every earlier run, aFree >= P for every later one).  Then nothing crosses P: a match of an
// earlier run ends at or before P, so the cursor (matchEnd - skipMargin, or seed + 1) is still at or before the
// first seed after P; and a later run can neither be trimmed by nor swallow a match from before P, because its
// left walk stops at aFree >= P >= every earlier match end (the reference's comparisons at .cpp:257-266 evaluate
// exactly as with an empty stack).  So the runs between two such cuts -- an island -- replay independently with
// an empty stack and cursor 0, and the contig's list is the concatenation of the islands' lists.  Cuts are where
// prefix-max(bFree) <= suffix-min(aFree): two scans over the run list.  Every mismatch between a target and the
// reference is a cut unless a repeat bridges it, so islands are a handful of runs; one thread replays one island.
// ------------------------------------------------------------------------------------------------------
constexpr int ISL_THREADS = 256;
constexpr int ISL_ITEMS = 4;
constexpr int ISL_TILE = ISL_THREADS * ISL_ITEMS;

template <bool IS_MAX>
__device__ __forceinline__ unsigned long long op64(unsigned long long x, unsigned long long y) {
    return IS_MAX ? (x > y ? x : y) : (x < y ? x : y);
}
template <bool IS_MAX>
__device__ __forceinline__ unsigned long long ident64() { return IS_MAX ? 0ull : ~0ull; }

// inclusive scan of one value per thread over the block; *total = reduction of the whole block
template <bool IS_MAX>
__device__ __forceinline__ unsigned long long block_scan64(unsigned long long v, unsigned long long* sWarp,
                                                           unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v = op64<IS_MAX>(v, t);
    }
    if (lane == 31) sWarp[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < nWarps ? sWarp[lane] : ident64<IS_MAX>();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = op64<IS_MAX>(w, t);
        }
        sWarp[32 + lane] = w;  // inclusive over warps
    }
    __syncthreads();
    if (warp > 0) v = op64<IS_MAX>(v, sWarp[32 + warp - 1]);
    *total = sWarp[32 + nWarps - 1];
    __syncthreads();
    return v;
}

// scan keys: the contig index in the high half makes one global scan behave as a per-contig scan
__device__ __forceinline__ unsigned long long key_b(const RunRec& r) {
    return (static_cast<unsigned long long>(r.flags >> 2) << 32) | r.bFree;
}
__device__ __forceinline__ unsigned long long key_a(const RunRec& r) {
    return (static_cast<unsigned long long>(r.flags >> 2) << 32) | r.aFree;
}

// per tile: max of key_b, min of key_a
__global__ void __launch_bounds__(ISL_THREADS) k_island_reduce(const RunRec* __restrict__ runs, const uint32_t* counters,
                                                               unsigned long long* __restrict__ tileFwd,
                                                               unsigned long long* __restrict__ tileBwd) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    __shared__ unsigned long long sWarp[64];
    const uint32_t nRuns = counters[1];
    const uint32_t nTiles = (nRuns + ISL_TILE - 1) / ISL_TILE;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        unsigned long long mx = 0ull, mn = ~0ull;
#pragma unroll
        for (int k = 0; k < ISL_ITEMS; k++) {
            const uint32_t r = t * ISL_TILE + k * ISL_THREADS + threadIdx.x;
            if (r < nRuns) {
                const RunRec rr = runs[r];
                mx = op64<true>(mx, key_b(rr));
                mn = op64<false>(mn, key_a(rr));
            }
        }
        unsigned long long tot;
        block_scan64<true>(mx, sWarp, &tot);
        if (threadIdx.x == 0) tileFwd[t] = tot;
        block_scan64<false>(mn, sWarp, &tot);
        if (threadIdx.x == 0) tileBwd[t] = tot;
    }
}

// tileFwd -> exclusive prefix max (carry into each tile); tileBwd -> exclusive suffix min
__global__ void __launch_bounds__(1024) k_island_partials(const uint32_t* counters, unsigned long long* __restrict__ tileFwd,
                                                          unsigned long long* __restrict__ tileBwd) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    __shared__ unsigned long long sWarp[64];
    __shared__ unsigned long long sInc[1024];
    const uint32_t nRuns = counters[1];
    const uint32_t nTiles = (nRuns + ISL_TILE - 1) / ISL_TILE;
    unsigned long long carry = 0ull;
    for (uint32_t base = 0; base < nTiles; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        unsigned long long tot;
        sInc[threadIdx.x] = block_scan64<true>(i < nTiles ? tileFwd[i] : 0ull, sWarp, &tot);
        __syncthreads();
        const unsigned long long ex = threadIdx.x ? sInc[threadIdx.x - 1] : 0ull;
        __syncthreads();
        if (i < nTiles) tileFwd[i] = op64<true>(ex, carry);
        carry = op64<true>(carry, tot);
    }
    carry = ~0ull;
    for (uint32_t base = 0; base < nTiles; base += blockDim.x) {
        const uint32_t q = base + threadIdx.x;  // tiles in reverse order
        const uint32_t i = nTiles - 1 - q;
        unsigned long long tot;
        sInc[threadIdx.x] = block_scan64<false>(q < nTiles ? tileBwd[i] : ~0ull, sWarp, &tot);
        __syncthreads();
        const unsigned long long ex = threadIdx.x ? sInc[threadIdx.x - 1] : ~0ull;
        __syncthreads();
        if (q < nTiles) tileBwd[i] = op64<false>(ex, carry);
        carry = op64<false>(carry, tot);
    }
}

// isHead[r] = 1 iff run r starts an island: first run of its contig, or prefix-max(bFree) of the contig's earlier
// runs <= suffix-min(aFree) of run r and the later ones
__global__ void __launch_bounds__(ISL_THREADS) k_island_heads(const RunRec* __restrict__ runs, const uint32_t* counters,
                                                              const unsigned long long* __restrict__ tileFwd,
                                                              const unsigned long long* __restrict__ tileBwd,
                                                              uint32_t* __restrict__ isHead) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    __shared__ unsigned long long sWarp[64];
    __shared__ unsigned long long sFwdEx[ISL_THREADS];
    const uint32_t nRuns = counters[1];
    const uint32_t nTiles = (nRuns + ISL_TILE - 1) / ISL_TILE;
    for (uint32_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        const uint32_t r0 = t * ISL_TILE + threadIdx.x * ISL_ITEMS;  // this thread's ISL_ITEMS consecutive runs
        unsigned long long kb[ISL_ITEMS], ka[ISL_ITEMS];
        uint32_t contig[ISL_ITEMS];
#pragma unroll
        for (int k = 0; k < ISL_ITEMS; k++) {
            kb[k] = 0ull; ka[k] = ~0ull; contig[k] = 0;
            if (r0 + k < nRuns) {
                const RunRec rr = runs[r0 + k];
                kb[k] = key_b(rr); ka[k] = key_a(rr); contig[k] = rr.flags >> 2;
            }
        }
        // forward: exclusive prefix max
        unsigned long long loc = 0ull;
#pragma unroll
        for (int k = 0; k < ISL_ITEMS; k++) loc = op64<true>(loc, kb[k]);
        unsigned long long tot;
        const unsigned long long inc = block_scan64<true>(loc, sWarp, &tot);
        sFwdEx[threadIdx.x] = inc;
        __syncthreads();
        unsigned long long run = op64<true>(threadIdx.x ? sFwdEx[threadIdx.x - 1] : 0ull, tileFwd[t]);
        __syncthreads();
        unsigned long long fwdEx[ISL_ITEMS];
#pragma unroll
        for (int k = 0; k < ISL_ITEMS; k++) {
            fwdEx[k] = run;
            run = op64<true>(run, kb[k]);
        }
        // backward: inclusive suffix min (threads scanned in reverse order)
        loc = ~0ull;
#pragma unroll
        for (int k = 0; k < ISL_ITEMS; k++) loc = op64<false>(loc, ka[k]);
        // reverse the thread order through shared memory, scan, reverse back
        sFwdEx[ISL_THREADS - 1 - threadIdx.x] = loc;
        __syncthreads();
        const unsigned long long rev = sFwdEx[threadIdx.x];
        __syncthreads();
        const unsigned long long incR = block_scan64<false>(rev, sWarp, &tot);
        sFwdEx[threadIdx.x] = incR;  // inclusive over reversed threads 0..x
        __syncthreads();
        // this thread (reversed index q = ISL_THREADS-1-threadIdx.x): min over the threads after it = inclusive at q-1
        const int q = ISL_THREADS - 1 - threadIdx.x;
        unsigned long long after = op64<false>(q ? sFwdEx[q - 1] : ~0ull, tileBwd[t]);
        __syncthreads();
#pragma unroll
        for (int k = ISL_ITEMS - 1; k >= 0; k--) {
            after = op64<false>(after, ka[k]);  // inclusive suffix min at run r0 + k
            if (r0 + k < nRuns) {
                const bool head = static_cast<uint32_t>(fwdEx[k] >> 32) != contig[k] || (fwdEx[k] >> 32) != (after >> 32) ||
                                  static_cast<uint32_t>(fwdEx[k]) <= static_cast<uint32_t>(after);
                isHead[r0 + k] = head ? 1u : 0u;
            }
        }
    }
}

struct ResolveArgs {
    IndexState ix;
    BatchDesc b;
    const RunRec* runs;
    const uint32_t* runStart;
    const uint32_t* isHead;
    MatchRec* matches;
    uint32_t* matchCountRun;  // [nRuns+1] matches of the island starting at this run (0 for other runs)
    uint32_t* counters;          // 1 nRuns, 3 overflow (read); 4 longest island, 5 / 8 long islands (diagnostics, work list)
    uint32_t minMatchLength;
    uint32_t* bigList;           // first runs of the long islands (k_resolve -> k_resolve_big)
};

__device__ __forceinline__ int64_t max3(int64_t a, int64_t b, int64_t c) {
    const int64_t m = a > b ? a : b;
    return m > c ? m : c;
}

// State of one island replay: the cursor and the stack of accepted matches (top kept in registers).
struct ReplayState {
    int64_t cur = 0;
    uint32_t nOut = 0;
    MatchRec top{0, 0, 0};
};

struct ReplayConst {
    int K, k1, M, k1Shift;
    bool k1Pow2;
    int64_t pos1, refLen, swStop;
    uint32_t cap;  // runs of the island = capacity of its output region
};

// Everything the reference does with the seeds of ONE run that lie at or after the cursor (.cpp:250-315).
__device__ __forceinline__ void replay_run(const ResolveArgs& a, const ReplayConst& rc, const RunRec& run, MatchRec* out,
                                           ReplayState& st) {
    const int K = rc.K, k1 = rc.k1, M = rc.M, k1Shift = rc.k1Shift;
    const bool k1Pow2 = rc.k1Pow2;
    const int64_t OVERLAP_MATCH_MAX_LENGTH = 1 << 13;  // SlidingWindowSparseEMMatcher.h:18
    const int64_t pos1 = rc.pos1, refLen = rc.refLen, swStop = rc.swStop;
    const uint32_t rs = 0, re = rc.cap;  // (the bound of the output region is written as re - rs below)
    int64_t& cur = st.cur;
    uint32_t& nOut = st.nOut;
    MatchRec& top = st.top;
    const int64_t aRes = run.aFree;  // link chains already resolved by k_links
    const int64_t firstPos = run.firstPos;
    const int64_t lastPos = firstPos + static_cast<int64_t>(run.nSeeds - 1) * k1;
    while (lastPos >= cur) {
        int64_t i = firstPos;
        if (cur > i) {  // first seed of the run at or after the cursor (32-bit math: targets are < 4 GiB)
            const uint32_t gap = static_cast<uint32_t>(cur - i) + static_cast<uint32_t>(k1) - 1;
            i += k1Pow2 ? ((gap >> k1Shift) << k1Shift) : (gap / static_cast<uint32_t>(k1)) * static_cast<uint32_t>(k1);
        }
        const int64_t s = static_cast<int64_t>(decode_pos(run.enc0, a.ix.k1ord)) + (i - firstPos);
        const int64_t d = s - i;
        const bool endsBefore = s + K < pos1;
        const bool startsBefore = s < swStop;
        const int64_t hi = endsBefore ? pos1 : refLen;  // tmpEnd1
        const int64_t lo = startsBefore ? 0 : swStop;   // tmpStart1
        int64_t right1 = static_cast<int64_t>(run.bFree) + d;
        if (hi < right1) right1 = hi;

        // left edge with overlap resolution (.cpp:251-291); p2 = inclusive matched cursor,
        // e = exclusive cursor (the reference's final p2)
        int64_t p2 = i, e = 0;
        uint32_t idx = nOut;
        bool broke = false;
        while (idx > 0) {
            const MatchRec last = idx == nOut ? top : out[idx - 1];
            const int64_t lastDst = last.dst, lastEnd = lastDst + last.len;
            if (lastEnd < p2) {
                const int64_t stop = max3(aRes - 1, lo - d, lastEnd - 1);
                if (stop > lastEnd - 1) {
                    e = stop;
                    broke = true;
                    break;
                }
                p2 = lastEnd;
            }
            const int64_t lastDelta = p2 - lastDst;
            if ((p2 + d) - lo < lastDelta || static_cast<int64_t>(last.len) > OVERLAP_MATCH_MAX_LENGTH ||
                aRes > lastDst) {
                e = p2 - 1;
                broke = true;
                break;
            }
            p2 = lastDst;
            idx--;
        }
        if (!broke) {
            e = max3(aRes - 1, lo - d, -1);
        } else {
            const MatchRec keep = idx == nOut ? top : out[idx - 1];
            const int64_t overlap = static_cast<int64_t>(keep.dst) + keep.len - (e + 1);
            if (overlap > 0) e += overlap;
        }
        const int64_t len = right1 - (e + d) - 1;
        if (len >= static_cast<int64_t>(a.minMatchLength) && idx < re - rs) {
            nOut = idx;
            MatchRec mr;
            mr.src = static_cast<uint64_t>(e + d + 1);
            mr.dst = static_cast<uint32_t>(e + 1);
            mr.len = static_cast<uint32_t>(len);
            out[nOut++] = mr;
            top = mr;
            int64_t skip = (e + 1 + len) - i;
            skip -= skip > M ? M : skip;
            cur = i + (skip ? skip : 1);
        } else {
            cur = i + 1;
        }
    }
}

constexpr uint32_t BIG_ISLAND = 256;  // islands longer than this go to the warp-cooperative kernel

__global__ void __launch_bounds__(128) k_resolve(const ResolveArgs a) {
    griddep_launch_dependents();
    griddep_wait();
    if (a.counters[3]) return;
    const uint32_t nRuns = a.counters[1];
    ReplayConst rc;
    rc.K = a.ix.K;
    rc.k1 = a.ix.k1;
    rc.M = a.ix.skipMargin;
    rc.k1Pow2 = (rc.k1 & (rc.k1 - 1)) == 0;
    rc.k1Shift = 31 - __clz(rc.k1);
    rc.pos1 = static_cast<int64_t>(a.ix.pos1);
    rc.refLen = static_cast<int64_t>(a.ix.refLen);
    const int k1 = rc.k1;
    for (uint32_t rs = blockIdx.x * blockDim.x + threadIdx.x; rs < nRuns; rs += gridDim.x * blockDim.x) {
        if (!a.isHead[rs]) {
            a.matchCountRun[rs] = 0;
            continue;
        }
        uint32_t re = rs + 1;  // island = runs [rs, re): up to the next head
        bool big = false;
        // most islands are one or two runs: look at the next flag alone first, then sixteen per step (loads in flight together)
        if (re < nRuns && !a.isHead[re]) re++;
        while (re > rs + 1 && re < nRuns) {
            uint32_t heads = 0;
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (re + k < nRuns && a.isHead[re + k]) heads |= 1u << k;
            if (heads) {
                re += __ffs(heads) - 1;
                break;
            }
            re += 16;
            if (re - rs > BIG_ISLAND) {  // a long island (several similar copies in the reference): a whole warp takes it
                big = true;
                break;
            }
        }
        if (big) {
            a.bigList[atomicAdd(&a.counters[8], 1u)] = rs;
            continue;
        }
        if (re > nRuns) re = nRuns;
        const uint32_t c = a.runs[rs].flags >> 2;
        const uint64_t lock = a.b.contigLock ? a.b.contigLock[c] : NO_LOCK;
        rc.swStop = lock != NO_LOCK ? static_cast<int64_t>(lock) : rc.pos1;
        rc.cap = re - rs;
        MatchRec* out = a.matches + rs;
        ReplayState st;
        for (uint32_t r = rs; r < re; r++) {
            const RunRec run = a.runs[r];
            const int64_t lastPos = static_cast<int64_t>(run.firstPos) + static_cast<int64_t>(run.nSeeds - 1) * k1;
            if (lastPos < st.cur) continue;  // behind the cursor (inside the match just accepted)
            replay_run(a, rc, run, out, st);
        }
        a.matchCountRun[rs] = st.nOut;
    }
}

// Long islands, a warp each: lane 0 replays, all 32 lanes find the island's end and the next run that is not behind
// the cursor (a ballot over 32 runs per step instead of one dependent load per run).
__global__ void __launch_bounds__(128) k_resolve_big(const ResolveArgs a) {
    griddep_launch_dependents();
    griddep_wait();
    if (a.counters[3]) return;
    const uint32_t nRuns = a.counters[1], nBig = a.counters[8];
    const int lane = threadIdx.x & 31;
    ReplayConst rc;
    rc.K = a.ix.K;
    rc.k1 = a.ix.k1;
    rc.M = a.ix.skipMargin;
    rc.k1Pow2 = (rc.k1 & (rc.k1 - 1)) == 0;
    rc.k1Shift = 31 - __clz(rc.k1);
    rc.pos1 = static_cast<int64_t>(a.ix.pos1);
    rc.refLen = static_cast<int64_t>(a.ix.refLen);
    const int k1 = rc.k1;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nBig; w += warps) {
        const uint32_t rs = a.bigList[w];
        uint32_t re = rs + 1;
        while (re < nRuns) {
            const uint32_t q = re + lane;
            const unsigned heads = __ballot_sync(0xffffffffu, q < nRuns && a.isHead[q] != 0);
            if (heads) {
                re += __ffs(heads) - 1;
                break;
            }
            re += 32;
        }
        if (re > nRuns) re = nRuns;
        const uint32_t c = a.runs[rs].flags >> 2;
        const uint64_t lock = a.b.contigLock ? a.b.contigLock[c] : NO_LOCK;
        rc.swStop = lock != NO_LOCK ? static_cast<int64_t>(lock) : rc.pos1;
        rc.cap = re - rs;
        MatchRec* out = a.matches + rs;
        ReplayState st;
        uint32_t r = rs;
        while (r < re) {
            const uint32_t q = r + lane;
            bool ahead = false;
            if (q < re) {
                const RunRec& rr = a.runs[q];
                ahead = static_cast<int64_t>(rr.firstPos) + static_cast<int64_t>(rr.nSeeds - 1) * k1 >= st.cur;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ahead);
            if (!bal) {
                r += 32;
                continue;
            }
            r += __ffs(bal) - 1;
            if (lane == 0) replay_run(a, rc, a.runs[r], out, st);
            st.cur = __shfl_sync(0xffffffffu, st.cur, 0);
            r++;
        }
        if (lane == 0) {
            a.matchCountRun[rs] = st.nOut;
            atomicAdd(&a.counters[5], 1u);
            atomicMax(&a.counters[4], re - rs);
        }
    }
}

// island lists -> one dense list; matchOffRun = exclusive scan of matchCountRun (matchOffRun[nRuns] = total)
__global__ void __launch_bounds__(256) k_gather(const MatchRec* __restrict__ matches, const uint32_t* __restrict__ isHead,
                                                const uint32_t* __restrict__ matchOffRun,
                                                const uint32_t* __restrict__ runStart, TextMatchRec* __restrict__ dense,
                                                uint32_t* __restrict__ matchOff, int nContigs, uint32_t* counters) {
    griddep_launch_dependents();
    griddep_wait();
    if (counters[3]) return;
    const uint32_t nRuns = counters[1];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (uint32_t r = gtid; r < nRuns; r += gsz) {
        if (!isHead[r]) continue;
        const uint32_t o = matchOffRun[r], n = matchOffRun[r + 1] - o;
        for (uint32_t k = 0; k < n; k++) {
            const MatchRec m = matches[r + k];
            TextMatchRec t;
            t.posSrcText = m.src;
            t.length = m.len;
            t.posDestText = m.dst;
            t.nextSrcRegionLoadingPos = 0;
            dense[o + k] = t;
        }
    }
    for (uint32_t c = gtid; c <= static_cast<uint32_t>(nContigs); c += gsz) matchOff[c] = matchOffRun[runStart[c]];
    if (gtid == 0) counters[2] = matchOffRun[nRuns];
}

int launch_resolve(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, uint32_t minMatchLength,
                   cudaStream_t st) {
    const int sm = device_sm_count();
    int launches = 0;
    // islands: headScan (free after k_runs) takes the island-head flags, blockSums-sized u64 temporaries the tile carries
    launch_pdl(k_island_reduce, sm * 4, ISL_THREADS, 0, st, w.runs, w.counters, w.islandFwd, w.islandBwd);
    launch_pdl(k_island_partials, 1, 1024, 0, st, w.counters, w.islandFwd, w.islandBwd);
    launch_pdl(k_island_heads, sm * 4, ISL_THREADS, 0, st, w.runs, w.counters, w.islandFwd, w.islandBwd, w.headScan);
    launches += 3;
    // runFirst (free after k_extend) takes the per-island match counts, then their exclusive scan
    // scratchEnc (position-indexed seed staging, free since the compaction) holds the list of long islands
    ResolveArgs a{ix, b, w.runs, w.runStart, w.headScan, w.matches, w.runFirst, w.counters, minMatchLength, w.scratchEnc};
    launch_pdl(k_resolve, sm * 16, 128, 0, st, a);
    launch_pdl(k_resolve_big, sm * 4, 128, 0, st, a);
    launches += 2;
    launches += scan_exclusive(w.runFirst, w.runFirst, w.counters + 1, 0, w.seedCap, w.scanState + 2 * w.scanStride, w.scanStride, nullptr, st);
    launch_pdl(k_gather, sm * 8, 256, 0, st, w.matches, w.headScan, w.runFirst, w.runStart, w.matchesDense, w.matchOff,
                                     b.nContigs, w.counters);
    launches += 1;
    return launches;
}

}  // namespace mbgc_b200
