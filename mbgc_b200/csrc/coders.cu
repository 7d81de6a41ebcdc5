// mbgc_b200/csrc/coders.cu -- row a14 of SURVEY.md 8: the two symbol packers the north star names next to the stream
// packing, on the device (sm_100a).  Both are legacy code paths of the reference (PgRC heritage; SURVEY.md 0.9: no call
// site / coder id never requested), restated here from their published behaviour and pinned against the reference's own
// classes compiled from where they lie (oracle/ref_coders_driver.cpp, tests/test_coders_oracle.py).
//
//   SymbolsPackingFacility::packSequence / reverseSequence   /root/reference/coders/SymbolsPackingFacility.cpp:123-165, 190-215
//       symbols of an alphabet of N letters, as many per byte as fit (N = 4: four, N = 5: three), first symbol most
//       significant, a short tail padded with the alphabet's first letter.  One thread per output byte; pure streaming.
//   VarLenDNACoder::encode / decode                          /root/reference/coders/VarLenDNACoder.cpp:55-128
//       greedy longest-first parse with a code book of up to 256 words of 1..4 symbols, looked up through the
//       reference's own 2^24-entry table over the low 6 bits of 4 bytes (VarLenDNACoder.h:40-43, cpp:30-36).  The parse
//       is sequential as written; its state is only "which of the next 4 positions starts a code word", so chunks of 64
//       positions become functions {entry offset 0..3} -> (exit offset, words), composed by a scan, and a last pass
//       writes every chunk's words at its known output offset.  decode is an expansion: lengths, scan, scatter.
#include "coders.h"

namespace mbgc_b200 {

namespace {

constexpr int CD_THREADS = 256;

// ---------------------------------------------------------------------------------------------------------------------
// SymbolsPackingFacility
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CD_THREADS) k_sym_pack(const uint8_t* __restrict__ src, uint64_t n, uint8_t* __restrict__ dst,
                                                         uint64_t nOut, const int8_t* __restrict__ order, int symbolsCount,
                                                         int perElement, uint32_t* __restrict__ bad) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t o = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < nOut; o += stride) {
        uint32_t v = 0;
        const uint64_t base = o * perElement;
        for (int j = 0; j < perElement; j++) {
            v *= symbolsCount;
            if (base + j < n) {
                const int k = order[src[base + j]];
                if (k < 0) atomicOr(bad, 1u);  // the reference's validateSymbol ends the program here
                else v += k;
            }
        }
        dst[o] = static_cast<uint8_t>(v);
    }
}

__global__ void __launch_bounds__(CD_THREADS) k_sym_unpack(const uint8_t* __restrict__ packed, uint64_t pos, uint64_t length,
                                                           uint8_t* __restrict__ dst, const uint8_t* __restrict__ reverse,
                                                           int perElement) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t k = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < length; k += stride) {
        const uint64_t p = pos + k;
        dst[k] = reverse[packed[p / perElement] * perElement + p % perElement];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// VarLenDNACoder
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t VL_LUT_SIZE = 1u << 24;
constexpr int VL_CHUNK = 64;          // positions per chunk map
constexpr int VL_BLOCK = 256;         // chunks per block map

__device__ __forceinline__ uint32_t vl_idx(uint32_t v) {  // getLUTidx, VarLenDNACoder.h:40-43
    return (v & 0x3Fu) | ((v & 0x3F00u) >> 2) | ((v & 0x3F0000u) >> 4) | ((v & 0x3F000000u) >> 6);
}

// the code book into the table, in book order (a later word that aliases an earlier one wins, as in initUsing :30-36)
__global__ void k_vl_build_lut(const uint8_t* __restrict__ book /* 256 x 5 */, int nCodes, uint8_t* __restrict__ lut) {
    if (blockIdx.x || threadIdx.x) return;
    for (int i = 0; i < nCodes; i++) {
        uint32_t t = 0;
        for (int j = 0; j < 4; j++) t |= static_cast<uint32_t>(book[i * 5 + j]) << (8 * j);
        lut[vl_idx(t)] = static_cast<uint8_t>(i);
    }
}

// the word the greedy parse emits when it stands at position p: (code, length); length 0 = unexpected symbol
__device__ __forceinline__ uint32_t vl_word(const uint8_t* __restrict__ src, uint64_t n, uint64_t p, const uint8_t* __restrict__ lut) {
    uint32_t t = 0;
    const uint64_t left = n - p;
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (static_cast<uint64_t>(j) < left) t |= static_cast<uint32_t>(src[p + j]) << (8 * j);
    t = vl_idx(t);
    uint32_t c;
    if (left >= 4 && (c = lut[t]) != 0) return (c << 8) | 4u;                  // :64-68
    t &= 0x0003FFFFu;
    if (left >= 3 && (c = lut[t]) != 0) return (c << 8) | 3u;                  // :69-74 / :94-99
    t &= 0x00000FFFu;
    if (left >= 2 && (c = lut[t]) != 0) return (c << 8) | 2u;                  // :75-80 / :100-105
    t &= 0x0000003Fu;
    c = lut[t];
    if (c == 0 && t != (65u & 0x3Fu)) return 0;                                // :81-85: "unexpected symbol"
    return (c << 8) | 1u;
}

// a chunk as a function of its entry offset e (the parse arrives at chunk start + e, e = 0..3): words emitted inside the
// chunk and the exit offset into the next chunk.  Packed: 4 x (count << 2 | exit).
__global__ void __launch_bounds__(CD_THREADS) k_vl_map(const uint8_t* __restrict__ src, uint64_t n, const uint8_t* __restrict__ lut,
                                                       uint4* __restrict__ chunkMap, uint64_t nChunks, uint32_t* __restrict__ bad) {
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= nChunks) return;
    const uint64_t lo = c * VL_CHUNK, hi = lo + VL_CHUNK < n ? lo + VL_CHUNK : n;
    uint32_t m[4];
    for (int e = 0; e < 4; e++) {
        uint64_t p = lo + e;
        uint32_t cnt = 0;
        while (p < hi) {
            const uint32_t w = vl_word(src, n, p, lut);
            if ((w & 0xFF) == 0) {
                atomicOr(bad, 1u);  // informational: whether the position lies on the parse path is known after the scan (k_vl_emit)
                p += 1;
            } else
                p += w & 0xFF;
            cnt++;
        }
        m[e] = (cnt << 2) | static_cast<uint32_t>(p >= hi ? (p - hi > 3 ? 3 : p - hi) : 0);
    }
    chunkMap[c] = make_uint4(m[0], m[1], m[2], m[3]);
}

__device__ __forceinline__ uint4 vl_compose(const uint4 a, const uint4 b) {  // a first, then b
    const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const uint32_t x = B[A[e] & 3];
        r[e] = (((A[e] >> 2) + (x >> 2)) << 2) | (x & 3);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}

// block map = composition of its VL_BLOCK chunk maps (one thread per block of chunks: 256 steps)
__global__ void __launch_bounds__(CD_THREADS) k_vl_blockmap(const uint4* __restrict__ chunkMap, uint64_t nChunks,
                                                            uint4* __restrict__ blockMap, uint64_t nBlocks) {
    const uint64_t b = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    uint4 acc = make_uint4(0, 1, 2, 3);  // identity
    const uint64_t lo = b * VL_BLOCK, hi = lo + VL_BLOCK < nChunks ? lo + VL_BLOCK : nChunks;
    for (uint64_t c = lo; c < hi; c++) acc = vl_compose(acc, chunkMap[c]);
    blockMap[b] = acc;
}

// entry state (offset, words so far) of every block: one thread walks the block maps (a few thousand steps per GB / 16 KiB)
__global__ void k_vl_blockscan(const uint4* __restrict__ blockMap, uint64_t nBlocks, uint2* __restrict__ blockEntry,
                               uint64_t* __restrict__ total) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t off = 0;
    uint64_t cnt = 0;
    for (uint64_t b = 0; b < nBlocks; b++) {
        reinterpret_cast<uint64_t*>(blockEntry)[b] = (cnt << 2) | off;
        const uint4 m = blockMap[b];
        const uint32_t M[4] = {m.x, m.y, m.z, m.w};
        cnt += M[off] >> 2;
        off = M[off] & 3;
    }
    *total = cnt;
}

// per block: thread 0 turns the block's entry state into the entry state of each of its chunks (shared memory), then every
// thread emits the words of its chunk
__global__ void __launch_bounds__(VL_BLOCK) k_vl_emit(const uint8_t* __restrict__ src, uint64_t n, const uint8_t* __restrict__ lut,
                                                      const uint4* __restrict__ chunkMap, uint64_t nChunks,
                                                      const uint64_t* __restrict__ blockEntry, uint8_t* __restrict__ dst,
                                                      uint64_t dstCap, uint32_t* __restrict__ bad) {
    __shared__ uint64_t sEntry[VL_BLOCK];
    const uint64_t b = blockIdx.x;
    const uint64_t lo = b * VL_BLOCK, hi = lo + VL_BLOCK < nChunks ? lo + VL_BLOCK : nChunks;
    if (threadIdx.x == 0) {
        uint64_t st = blockEntry[b];
        for (uint64_t c = lo; c < hi; c++) {
            sEntry[c - lo] = st;
            const uint4 m = chunkMap[c];
            const uint32_t M[4] = {m.x, m.y, m.z, m.w};
            const uint32_t x = M[st & 3];
            st = (((st >> 2) + (x >> 2)) << 2) | (x & 3);
        }
    }
    __syncthreads();
    const uint64_t c = lo + threadIdx.x;
    if (c >= hi) return;
    const uint64_t st = sEntry[threadIdx.x];
    uint64_t out = st >> 2;
    const uint64_t cend = c * VL_CHUNK + VL_CHUNK < n ? c * VL_CHUNK + VL_CHUNK : n;
    uint64_t p = c * VL_CHUNK + (st & 3);
    while (p < cend) {
        const uint32_t w = vl_word(src, n, p, lut);
        if ((w & 0xFF) == 0) {
            atomicOr(bad, 2u);  // an unexpected symbol on the parse path
            return;
        }
        if (out < dstCap) dst[out] = static_cast<uint8_t>(w >> 8);
        out++;
        p += w & 0xFF;
    }
}

// decode: word lengths of the codes, scanned by blocks; then the scatter
__global__ void __launch_bounds__(CD_THREADS) k_vl_declen(const uint8_t* __restrict__ codes, uint64_t n, const uint8_t* __restrict__ book,
                                                          uint32_t* __restrict__ blockLen) {
    __shared__ uint32_t sSum;
    if (threadIdx.x == 0) sSum = 0;
    __syncthreads();
    const uint64_t i0 = static_cast<uint64_t>(blockIdx.x) * (CD_THREADS * 16);
    uint32_t s = 0;
    for (int k = 0; k < 16; k++) {
        const uint64_t i = i0 + static_cast<uint64_t>(k) * CD_THREADS + threadIdx.x;
        if (i < n) {
            const uint8_t* w = book + codes[i] * 5;
            s += w[0] == 0 ? 0 : (w[1] == 0 ? 1 : (w[2] == 0 ? 2 : (w[3] == 0 ? 3 : 4)));
        }
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sSum, s);
    __syncthreads();
    if (threadIdx.x == 0) blockLen[blockIdx.x] = sSum;
}

__global__ void k_vl_decscan(uint32_t* __restrict__ blockLen, uint64_t nBlocks, uint64_t* __restrict__ blockOff, uint64_t* __restrict__ total) {
    if (blockIdx.x || threadIdx.x) return;
    uint64_t acc = 0;
    for (uint64_t b = 0; b < nBlocks; b++) {
        blockOff[b] = acc;
        acc += blockLen[b];
    }
    *total = acc;
}

// one warp per 32 x 16 codes: lane l expands codes [16 l, 16 l + 16) of the warp's slice behind a warp scan of the lengths
__global__ void __launch_bounds__(CD_THREADS) k_vl_decode(const uint8_t* __restrict__ codes, uint64_t n, const uint8_t* __restrict__ book,
                                                          const uint64_t* __restrict__ blockOff, uint8_t* __restrict__ dst, uint64_t dstCap) {
    __shared__ uint32_t sWarp[CD_THREADS / 32 + 1];
    const uint64_t i0 = static_cast<uint64_t>(blockIdx.x) * (CD_THREADS * 16) + static_cast<uint64_t>(threadIdx.x) * 16;
    uint32_t len[16], s = 0;
    for (int k = 0; k < 16; k++) {
        len[k] = 0;
        if (i0 + k < n) {
            const uint8_t* w = book + codes[i0 + k] * 5;
            len[k] = w[0] == 0 ? 0 : (w[1] == 0 ? 1 : (w[2] == 0 ? 2 : (w[3] == 0 ? 3 : 4)));
        }
        s += len[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = s;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int w = 0; w < CD_THREADS / 32; w++) {
            const uint32_t t = sWarp[w];
            sWarp[w] = acc;
            acc += t;
        }
    }
    __syncthreads();
    uint64_t out = blockOff[blockIdx.x] + sWarp[warp] + inc - s;
    for (int k = 0; k < 16; k++) {
        if (i0 + k >= n) break;
        const uint8_t* w = book + codes[i0 + k] * 5;
        for (uint32_t j = 0; j < len[k]; j++)
            if (out + j < dstCap) dst[out + j] = w[j];
        out += len[k];
    }
}

inline unsigned grid_for(uint64_t items, int threads, unsigned cap) {
    const uint64_t b = (items + threads - 1) / threads;
    return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int launch_sym_pack(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t nOut, const int8_t* order, int symbolsCount,
                    int perElement, uint32_t* bad, cudaStream_t st) {
    if (nOut == 0) return 0;
    k_sym_pack<<<grid_for(nOut, CD_THREADS, 148 * 16), CD_THREADS, 0, st>>>(src, n, dst, nOut, order, symbolsCount, perElement, bad);
    return 1;
}

int launch_sym_unpack(const uint8_t* packed, uint64_t pos, uint64_t length, uint8_t* dst, const uint8_t* reverse, int perElement,
                      cudaStream_t st) {
    if (length == 0) return 0;
    k_sym_unpack<<<grid_for(length, CD_THREADS, 148 * 16), CD_THREADS, 0, st>>>(packed, pos, length, dst, reverse, perElement);
    return 1;
}

int launch_vl_build_lut(const uint8_t* book, int nCodes, uint8_t* lut, cudaStream_t st) {
    cudaMemsetAsync(lut, 0, VL_LUT_SIZE, st);
    k_vl_build_lut<<<1, 32, 0, st>>>(book, nCodes, lut);
    return 1;
}

uint64_t vl_chunks(uint64_t n) { return (n + VL_CHUNK - 1) / VL_CHUNK; }
uint64_t vl_blocks(uint64_t n) { return (vl_chunks(n) + VL_BLOCK - 1) / VL_BLOCK; }

int launch_vl_encode(const uint8_t* src, uint64_t n, const uint8_t* lut, uint4* chunkMap, uint4* blockMap, uint64_t* blockEntry,
                     uint64_t* total, uint8_t* dst, uint64_t dstCap, uint32_t* bad, cudaStream_t st) {
    const uint64_t nChunks = vl_chunks(n), nBlocks = vl_blocks(n);
    if (nChunks == 0) {
        cudaMemsetAsync(total, 0, 8, st);
        return 0;
    }
    k_vl_map<<<static_cast<unsigned>((nChunks + CD_THREADS - 1) / CD_THREADS), CD_THREADS, 0, st>>>(src, n, lut, chunkMap, nChunks, bad + 1);
    k_vl_blockmap<<<static_cast<unsigned>((nBlocks + CD_THREADS - 1) / CD_THREADS), CD_THREADS, 0, st>>>(chunkMap, nChunks, blockMap, nBlocks);
    k_vl_blockscan<<<1, 32, 0, st>>>(blockMap, nBlocks, reinterpret_cast<uint2*>(blockEntry), total);
    k_vl_emit<<<static_cast<unsigned>(nBlocks), VL_BLOCK, 0, st>>>(src, n, lut, chunkMap, nChunks, blockEntry, dst, dstCap, bad);
    return 4;
}

uint64_t vl_dec_blocks(uint64_t n) { return (n + CD_THREADS * 16 - 1) / (CD_THREADS * 16); }

int launch_vl_decode(const uint8_t* codes, uint64_t n, const uint8_t* book, uint32_t* blockLen, uint64_t* blockOff, uint64_t* total,
                     uint8_t* dst, uint64_t dstCap, cudaStream_t st) {
    const uint64_t nb = vl_dec_blocks(n);
    if (nb == 0) {
        cudaMemsetAsync(total, 0, 8, st);
        return 0;
    }
    k_vl_declen<<<static_cast<unsigned>(nb), CD_THREADS, 0, st>>>(codes, n, book, blockLen);
    k_vl_decscan<<<1, 32, 0, st>>>(blockLen, nb, blockOff, total);
    k_vl_decode<<<static_cast<unsigned>(nb), CD_THREADS, 0, st>>>(codes, n, book, blockOff, dst, dstCap);
    return 3;
}

}  // namespace mbgc_b200
