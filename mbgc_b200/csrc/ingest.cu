// mbgc_b200/csrc/ingest.cu -- FASTA ingest on the device (sm_100a): a file image as it is on disk -> contiguous texts.
//
// Replaces the per-character loop of the reference's reader, kseq_read_lossless_fasta
// (/root/reference/utils/kseq.h:233-274), which MBGC runs on one thread per file
// (matching/MultipleGenomeMatchingProcessor.cpp:9-14, 126-139, 400-406).  The reader is sequential only in how it is
// written; what it computes is local:
//
//   records        a record starts at every '>' that is the first byte of a line (:250: the loop looks at the first
//                  character of every line); its header is the rest of that line (:244, KS_SEP_LINE, '\r' is data)
//   sequence       the lines up to the next record start, line ends removed (:250-259)
//   line length    every line but a record's last must have the file's line length (set by the first such line, :253-256),
//                  no line may be empty (:251), no last line may be longer (:260-264)  -- otherwise "DNA not well-formed"
//
// So: k_fa_count / k_fa_starts find the record starts in file order (block counts, scan, ranks); k_fa_records (a warp
// per record) finds the end of the header and of the first line; k_fa_check (a thread per 16 input bytes) verifies
// that every line end of a record sits exactly where its first line's length puts it -- which makes the sequence a pure
// index transformation, out[k] = in[seqStart + k + k / lineLen], done by k_fa_gather with 16-byte stores.  The one
// genuinely sequential piece, the file-wide line-length state across records, is a loop over the per-record
// summaries on the host (ingest.cpp).  All byte work, HBM-bound: the input is read three times, the output written once.
#include "ingest.h"

namespace mbgc_b200 {

namespace {

__device__ __forceinline__ uint32_t fa_warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

template <int THREADS>
__device__ __forceinline__ uint32_t fa_block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sWarp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = fa_warp_incl_scan(v, lane);
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (THREADS >> 5) ? sWarp[lane] : 0;
        const uint32_t wi = fa_warp_incl_scan(w, lane);
        sWarp[lane] = wi - w;
        if (lane == 31) sWarp[32] = wi;
    }
    __syncthreads();
    const uint32_t res = inc - v + sWarp[warp];
    *total = sWarp[32];
    __syncthreads();
    return res;
}

// bit k set where byte k of the four equals c (SIMD-in-register compare, then one bit per byte)
__device__ __forceinline__ uint32_t eq_mask4(uint32_t x, uint32_t c4) {
    const uint32_t r = __vcmpeq4(x, c4);  // 0xFF per equal byte
    return (r & 1u) | ((r >> 7) & 2u) | ((r >> 14) & 4u) | ((r >> 21) & 8u);
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4& v, uint32_t c4) {
    return eq_mask4(v.x, c4) | (eq_mask4(v.y, c4) << 4) | (eq_mask4(v.z, c4) << 8) | (eq_mask4(v.w, c4) << 12);
}

// bit 7 of every byte of x that equals the byte replicated in c4 (exact: no carries between bytes)
__device__ __forceinline__ uint32_t eq_bytes(uint32_t x, uint32_t c4) {
    const uint32_t y = x ^ c4;
    return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y) & 0x80808080u;
}

// The 16 bytes at i0 (a multiple of 16; the buffer is 16-byte aligned with a '\n' in front of byte 0 and zeros behind
// byte n-1) -> bit k set where byte k starts a record: a '>' whose predecessor is a '\n'.
__device__ __forceinline__ uint32_t start_mask(const uint8_t* __restrict__ in, uint64_t n, uint64_t i0) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i0);
    // sequence lines hold no '>' at all: one cheap test per word rejects nearly every chunk
    if ((eq_bytes(v.x, 0x3E3E3E3Eu) | eq_bytes(v.y, 0x3E3E3E3Eu) | eq_bytes(v.z, 0x3E3E3E3Eu) | eq_bytes(v.w, 0x3E3E3E3Eu)) == 0)
        return 0;
    const uint32_t gt = eq_mask16(v, 0x3E3E3E3Eu);
    if (gt == 0) return 0;
    const uint32_t nl = eq_mask16(v, 0x0A0A0A0Au);
    const uint32_t prevNl = in[static_cast<int64_t>(i0) - 1] == '\n' ? 1u : 0u;
    return gt & ((nl << 1) | prevNl) & 0xFFFFu;  // (bytes behind n-1 are zeros, never '>')
}

__global__ void __launch_bounds__(FA_THREADS) k_fa_count(const uint8_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ blockCnt) {
    __shared__ uint32_t sTot;
    if (threadIdx.x == 0) sTot = 0;
    __syncthreads();
    const uint64_t b0 = static_cast<uint64_t>(blockIdx.x) * FA_BLOCK_BYTES;
    uint32_t c = 0;
#pragma unroll
    for (uint32_t k = 0; k < FA_PER_THREAD; k++) {
        const uint64_t i0 = b0 + (k * FA_THREADS + threadIdx.x) * FA_CHUNK;
        if (i0 < n) c += __popc(start_mask(in, n, i0));
    }
    const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, c);
    if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sTot, t);
    __syncthreads();
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = sTot;
}

// exclusive scan in place, one CTA of 1024 threads, 4 values per thread and round; total at [nBlocks]
__global__ void __launch_bounds__(1024) k_fa_scan(uint32_t* __restrict__ blockCnt, uint32_t nBlocks) {
    __shared__ uint32_t sWarp[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nBlocks; base += 4096) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v[q] = i0 + q < nBlocks ? blockCnt[i0 + q] : 0;
            sum += v[q];
        }
        uint32_t total;
        uint32_t ex = fa_block_excl_scan<1024>(sum, &total, sWarp) + carry;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (i0 + q < nBlocks) blockCnt[i0 + q] = ex;
            ex += v[q];
        }
        carry += total;
    }
    if (threadIdx.x == 0) blockCnt[nBlocks] = carry;
}

__global__ void __launch_bounds__(FA_THREADS) k_fa_starts(const uint8_t* __restrict__ in, uint64_t n,
                                                          const uint32_t* __restrict__ blockOff, uint64_t* __restrict__ starts) {
    __shared__ uint32_t sWarp[33];
    uint32_t at0 = blockOff[blockIdx.x];
    if (blockOff[blockIdx.x + 1] == at0) return;  // no record starts in this 16 KiB: nearly every CTA ends here
    const uint64_t b0 = static_cast<uint64_t>(blockIdx.x) * FA_BLOCK_BYTES;
    for (uint32_t k = 0; k < FA_PER_THREAD; k++) {  // file order = round k, then thread
        const uint64_t i0 = b0 + (k * FA_THREADS + threadIdx.x) * FA_CHUNK;
        uint32_t m = i0 < n ? start_mask(in, n, i0) : 0;
        uint32_t total;
        uint32_t at = fa_block_excl_scan<FA_THREADS>(__popc(m), &total, sWarp) + at0;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            starts[at++] = i0 + b;
        }
        at0 += total;
    }
}

// first '\n' in [from, to) found by the whole CTA (16 bytes per thread and round, 4 KiB per round), or `to`
__device__ __forceinline__ uint64_t block_find_nl(const uint8_t* __restrict__ in, uint64_t from, uint64_t to,
                                                  unsigned long long* sPos) {
    if (threadIdx.x == 0) *sPos = to;
    __syncthreads();
    for (uint64_t base = from & ~static_cast<uint64_t>(15); base < to; base += FA_THREADS * FA_CHUNK) {
        const uint64_t i0 = base + threadIdx.x * FA_CHUNK;
        uint32_t m = 0;
        if (i0 < to) {
            m = eq_mask16(*reinterpret_cast<const uint4*>(in + i0), 0x0A0A0A0Au);
            if (i0 < from) m &= 0xFFFFu << (from - i0);
            if (i0 + 16 > to) m &= (1u << (to - i0)) - 1;
        }
        if (m) atomicMin(sPos, static_cast<unsigned long long>(i0 + (__ffs(m) - 1)));
        if (__syncthreads_or(m != 0)) break;
    }
    __syncthreads();
    const uint64_t r = *sPos;
    __syncthreads();
    return r;
}

// a CTA per record (a header or a first line can be a whole unwrapped chromosome)
// fileEnds (batched ingest, several file images in one buffer): ascending end offsets of the images; a record ends where
// its file ends even if the next record start lies further on (the bytes between two images are filler)
__global__ void __launch_bounds__(FA_THREADS) k_fa_records(const uint8_t* __restrict__ in, uint64_t n,
                                                           const uint64_t* __restrict__ starts, uint32_t nRecs,
                                                           FaRecord* __restrict__ recs,
                                                           const uint64_t* __restrict__ fileEnds, uint32_t nFiles) {
    __shared__ unsigned long long sPos;
    const uint32_t r = blockIdx.x;
    const uint64_t start = starts[r];
    uint64_t regionEnd = r + 1 < nRecs ? starts[r + 1] : n;
    if (fileEnds) {  // first file end behind `start`
        uint32_t lo = 0, hi = nFiles;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (fileEnds[mid] > start) hi = mid; else lo = mid + 1;
        }
        if (lo < nFiles && fileEnds[lo] < regionEnd) regionEnd = fileEnds[lo];
    }
    const uint64_t hdrEnd = block_find_nl(in, start + 1, regionEnd, &sPos);  // the header line cannot reach the next record
    const uint64_t seqStart = hdrEnd + 1 < regionEnd ? hdrEnd + 1 : regionEnd;
    const uint64_t firstNl = block_find_nl(in, seqStart, regionEnd, &sPos);
    if (threadIdx.x == 0) {
        FaRecord rec;
        rec.start = start;
        rec.hdrEnd = hdrEnd;
        rec.regionEnd = regionEnd;
        rec.firstLine = firstNl - seqStart;
        rec.endsWithNl = regionEnd > seqStart && in[regionEnd - 1] == '\n';
        rec.bad = regionEnd > seqStart && firstNl == seqStart;  // the record begins with an empty line (:251)
        recs[r] = rec;
    }
}

// index r with key[r] <= x < key[r + 1] (key ascending, key[0] <= x), keys read through `get`
template <typename F>
__device__ __forceinline__ uint32_t fa_segment(uint32_t nRecs, uint64_t x, F get) {
    uint32_t lo = 0, hi = nRecs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (get(mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// Every byte of a record's sequence region against the layout its first line dictates: a '\n' exactly after every
// `firstLine` bytes; the last line may be shorter (not empty) and may lack its '\n'.
__device__ __forceinline__ void check_chunk(const uint8_t* __restrict__ in, uint64_t n, uint64_t i0, uint32_t r,
                                            FaRecord* __restrict__ recs) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i0);
    {   // the whole chunk inside one record's sequence, lines of 16 bytes or more, the region's last byte not in it:
        // at most one line end is due in these 16 bytes, compare the masks
        const FaRecord rec = recs[r];
        const uint64_t seqStart = rec.hdrEnd + 1;
        if (i0 >= seqStart && i0 + 16 < rec.regionEnd && rec.firstLine >= 16) {
            const uint64_t o = i0 - seqStart, period = rec.firstLine + 1;
            const uint64_t m0 = (o | period) >> 32 ? o % period : static_cast<uint32_t>(o) % static_cast<uint32_t>(period);
            const uint32_t due = static_cast<uint32_t>(rec.firstLine - m0 < 16 ? rec.firstLine - m0 : 16);  // 16: none
            // line ends as bit 7 per byte, word by word, against the one that is due
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            bool ok = true;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t want = (due >> 2) == static_cast<uint32_t>(q) ? 0x80u << (8 * (due & 3)) : 0u;
                ok &= eq_bytes(w[q], 0x0A0A0A0Au) == want;
            }
            if (!ok) recs[r].bad = 1;
            return;
        }
    }
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint64_t seqStart = 0, len = 0, line = 0, next = 0, m = 0;
    bool loaded = false, bad = false;
    for (int k = 0; k < 16 && i0 + k < n; k++) {
        const uint64_t i = i0 + k;
        if (loaded && i >= next) {
            if (bad) recs[r].bad = 1;
            bad = false;
            loaded = false;
            r++;
        }
        if (!loaded) {
            FaRecord rec = recs[r];
            // batched ingest: the chunk starts in the filler behind record r's file image -- the next record owns the
            // rest of it (i < n = end of the last image, so there is one)
            if (i >= rec.regionEnd) rec = recs[++r];
            seqStart = rec.hdrEnd + 1;
            next = rec.regionEnd;
            len = next > seqStart ? next - seqStart : 0;
            line = rec.firstLine;
            loaded = true;
            m = i >= seqStart ? (i - seqStart) % (line + 1) : 0;
        }
        if (i < seqStart) continue;
        const bool nl = static_cast<uint8_t>(w[k >> 2] >> (8 * (k & 3))) == '\n';
        const bool expect = m == line;
        const uint64_t o = i - seqStart;
        if (nl != expect && !(nl && o == len - 1 && m != 0)) bad = true;
        m = expect ? 0 : m + 1;
    }
    if (bad) recs[r].bad = 1;
}

__global__ void __launch_bounds__(FA_THREADS) k_fa_check(const uint8_t* __restrict__ in, uint64_t n,
                                                         const uint64_t* __restrict__ starts, uint32_t nRecs,
                                                         FaRecord* __restrict__ recs) {
    const uint64_t b0 = static_cast<uint64_t>(blockIdx.x) * FA_BLOCK_BYTES;
    if (b0 >= n) return;
    // one search for the CTA's 16 KiB; records are far longer than that, so usually all of it is one record
    const uint32_t rb = fa_segment(nRecs, b0, [&](uint32_t k) { return starts[k]; });
    const bool oneRecord = rb + 1 >= nRecs || starts[rb + 1] >= b0 + FA_BLOCK_BYTES;
#pragma unroll
    for (uint32_t k = 0; k < FA_PER_THREAD; k++) {
        const uint64_t i0 = b0 + (k * FA_THREADS + threadIdx.x) * FA_CHUNK;
        if (i0 >= n) break;
        const uint32_t r = oneRecord ? rb : fa_segment(nRecs, i0, [&](uint32_t q) { return starts[q]; });
        check_chunk(in, n, i0, r, recs);
    }
}

// 16 output bytes of a validated record: out[k] = in[seqStart + k + k / lineLen] for k = k0 .. k0 + 15
__device__ __forceinline__ uint4 gather_chunk(const uint8_t* __restrict__ in, const FaGather& rec, uint32_t k0) {
    // sequences are shorter than 4 GiB (checked on the host): 32-bit index arithmetic
    const uint32_t lineLen = static_cast<uint32_t>(rec.lineLen), seqLen = static_cast<uint32_t>(rec.seqLen);
    const uint32_t q = lineLen ? k0 / lineLen : 0, rem = lineLen ? k0 - q * lineLen : 0;
    const uint8_t* src = in + rec.seqStart + k0 + q;
    const uint32_t valid = k0 >= seqLen ? 0 : (seqLen - k0 < 16 ? seqLen - k0 : 16);
    uint32_t w[4] = {0, 0, 0, 0};
    if (valid == 0) return make_uint4(0, 0, 0, 0);  // the zero gap behind a text (its chunks belong to the record too)
    if (lineLen >= 16) {
        // at most one line end falls inside the chunk: after `e` output bytes the source skips one byte.  24 source bytes
        // from six aligned words, F = bytes 0.., G = bytes 1..; every output word is F, G or a byte-wise mix of the two.
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uint32_t* p = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
        const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
        uint32_t x[6];
#pragma unroll
        for (int i = 0; i < 6; i++) x[i] = __ldg(p + i);
        uint32_t F[5];
#pragma unroll
        for (int i = 0; i < 5; i++) F[i] = __funnelshift_r(x[i], x[i + 1], sh);
        const uint32_t e = lineLen - rem;  // output bytes before the line end (>= 16: none inside)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t G = __funnelshift_r(F[i], F[i + 1], 8);
            const uint32_t lo = 4u * i;
            if (e >= lo + 4) w[i] = F[i];
            else if (e <= lo) w[i] = G;
            else {
                const uint32_t keep = (1u << (8 * (e - lo))) - 1;
                w[i] = (F[i] & keep) | (G & ~keep);
            }
        }
        if (valid < 16) {  // the record's last chunk: zeros behind the sequence
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t lo = 4u * i;
                if (valid <= lo) w[i] = 0;
                else if (valid < lo + 4) w[i] &= (1u << (8 * (valid - lo))) - 1;
            }
        }
    } else {  // short lines: byte by byte
        uint32_t r2 = rem, skip = 0;
        for (uint32_t j = 0; j < valid; j++) {
            w[j >> 2] |= static_cast<uint32_t>(src[j + skip]) << (8 * (j & 3));
            if (++r2 == lineLen) {
                r2 = 0;
                skip++;
            }
        }
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// a thread per 4 x 16 output bytes, 16-byte stores into the batch layout
__global__ void __launch_bounds__(FA_THREADS) k_fa_gather(const uint8_t* __restrict__ in, const FaGather* __restrict__ g,
                                                          uint32_t nRecs, uint64_t nChunks, uint8_t* __restrict__ out) {
    const uint64_t c0 = static_cast<uint64_t>(blockIdx.x) * (FA_THREADS * FA_PER_THREAD);
    if (c0 >= nChunks) return;
    const uint32_t rb = fa_segment(nRecs, c0, [&](uint32_t k) { return g[k].chunkStart; });
    const bool oneRecord = rb + 1 >= nRecs || g[rb + 1].chunkStart >= c0 + FA_THREADS * FA_PER_THREAD;
    FaGather rec = g[rb];
#pragma unroll
    for (uint32_t k = 0; k < FA_PER_THREAD; k++) {
        const uint64_t c = c0 + k * FA_THREADS + threadIdx.x;
        if (c >= nChunks) break;
        if (!oneRecord) rec = g[fa_segment(nRecs, c, [&](uint32_t q) { return g[q].chunkStart; })];
        const uint32_t k0 = static_cast<uint32_t>(c - rec.chunkStart) * 16u;
        *reinterpret_cast<uint4*>(out + rec.outOff + k0) = gather_chunk(in, rec, k0);
    }
}

inline uint32_t fa_blocks(uint64_t n) { return static_cast<uint32_t>((n + FA_BLOCK_BYTES - 1) / FA_BLOCK_BYTES); }

}  // namespace

int launch_fa_count(const uint8_t* in, uint64_t n, uint32_t* blockCnt, cudaStream_t st) {
    const uint32_t blocks = fa_blocks(n);
    k_fa_count<<<blocks, FA_THREADS, 0, st>>>(in, n, blockCnt);
    k_fa_scan<<<1, 1024, 0, st>>>(blockCnt, blocks);
    return 2;
}

int launch_fa_starts(const uint8_t* in, uint64_t n, const uint32_t* blockOff, uint64_t* starts, cudaStream_t st) {
    k_fa_starts<<<fa_blocks(n), FA_THREADS, 0, st>>>(in, n, blockOff, starts);
    return 1;
}

// the filler between the images of a batched ingest: 16 x '\n' in front of every image (so that its first byte starts a
// line), zeros from an image's end to the next filler.  spans[2f], spans[2f+1] = start and end offset of image f.
__global__ void k_fa_filler(uint8_t* __restrict__ in, const uint64_t* __restrict__ spans, uint32_t nFiles, uint64_t bufEnd) {
    const uint32_t f = blockIdx.x;
    const uint64_t end = spans[2 * f + 1];
    const uint64_t nextStart = f + 1 < nFiles ? spans[2 * f + 2] : bufEnd + 16;
    for (uint64_t i = end + threadIdx.x; i < nextStart - 16; i += blockDim.x) in[i] = 0;
    if (f + 1 < nFiles && threadIdx.x < 16) in[nextStart - 16 + threadIdx.x] = '\n';
    if (f == 0 && threadIdx.x < 16) in[static_cast<int64_t>(spans[0]) - 16 + threadIdx.x] = '\n';
}

int launch_fa_filler(uint8_t* in, const uint64_t* spans, uint32_t nFiles, uint64_t bufEnd, cudaStream_t st) {
    if (nFiles == 0) return 0;
    k_fa_filler<<<nFiles, 64, 0, st>>>(in, spans, nFiles, bufEnd);
    return 1;
}

int launch_fa_records(const uint8_t* in, uint64_t n, const uint64_t* starts, uint32_t nRecs, FaRecord* recs, cudaStream_t st,
                      const uint64_t* fileEnds, uint32_t nFiles) {
    k_fa_records<<<nRecs, FA_THREADS, 0, st>>>(in, n, starts, nRecs, recs, fileEnds, nFiles);
    k_fa_check<<<fa_blocks(n), FA_THREADS, 0, st>>>(in, n, starts, nRecs, recs);
    return 2;
}

// header lines of the accepted records -> one contiguous buffer (a warp per header; they go back to the host in one copy)
__global__ void __launch_bounds__(256) k_fa_headers(const uint8_t* __restrict__ in, const uint64_t* __restrict__ hdr, uint32_t nRecs,
                                                    uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nRecs; r += warpsTotal) {
        const uint64_t src = hdr[3 * r], len = hdr[3 * r + 1], dst = hdr[3 * r + 2];
        for (uint64_t k = lane; k < len; k += 32) out[dst + k] = in[src + k];
    }
}

int launch_fa_headers(const uint8_t* in, const uint64_t* hdr, uint32_t nRecs, uint8_t* out, cudaStream_t st) {
    if (nRecs == 0) return 0;
    const uint32_t blocks = (nRecs + 7) / 8;
    k_fa_headers<<<blocks < 1024 ? blocks : 1024, 256, 0, st>>>(in, hdr, nRecs, out);
    return 1;
}

int launch_fa_gather(const uint8_t* in, const FaGather* g, uint32_t nRecs, uint64_t nChunks, uint8_t* out, cudaStream_t st) {
    if (nChunks == 0) return 0;
    const uint64_t perBlock = static_cast<uint64_t>(FA_THREADS) * FA_PER_THREAD;
    k_fa_gather<<<static_cast<uint32_t>((nChunks + perBlock - 1) / perBlock), FA_THREADS, 0, st>>>(in, g, nRecs, nChunks, out);
    return 1;
}

}  // namespace mbgc_b200
