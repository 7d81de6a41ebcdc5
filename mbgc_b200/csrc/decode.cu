// mbgc_b200/csrc/decode.cu -- row f4 of SURVEY.md 8: the decoder's match-copy path on the device
// (/root/reference/mbgccoder/MBGC_Decoder.cpp: decodeSequenceAndReturnUnmatchedChars :319-432, extendMatchRight :434-462,
// extendMatchLeft :464-527).
//
// The reference decodes one contig at a time: it walks the literal stream from MATCH_MARK to MATCH_MARK, reads an offset (or
// derives it from an earlier paired match), a length, a gap delta, runs the mismatch-score automaton over the flag stream for
// the extensions, and appends literals / reference bytes to the contig.  Two observations split that into parallel work:
//
//  * the CONTROL FLOW (how many bytes every piece has, where it reads from) depends on the six streams only, never on
//    the reference's content.  k_dc_parse replays it -- one thread per target file, all targets at once -- and writes a
//    list of segments (literal run | reference copy | right extension | left extension) with contig-relative positions;
//  * given the segments, every OUTPUT BYTE is independent.  k_dc_copy writes the DNA: one CTA per tile of output (8 to 64 KiB) walks
//    the segments that intersect its tile; copies move 16 bytes per thread (aligned stores, two aligned loads and a byte
//    shift for the source), extensions are done by a warp each behind a ballot prefix of their mismatch flags.
//
// Where every target starts in every stream (the reference finds that by decoding everything in front of it) comes from marker
// passes over the streams: k_dc_count / k_dc_scatter (SEQ_SEPARATOR_MARK and MATCH_MARK in the literals, FILE_SEPARATOR_MARK
// in the flags), k_dc_contig_gd / k_dc_contig_nz (gap deltas and omitted offsets per contig) and scans; the frugal length
// stream is expanded with composable chunk maps (k_dl_*: which 16-bit word starts a record is the only state).
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode.h"

namespace mbgc_b200 {

namespace {

constexpr int DC_THREADS = 256;
constexpr int MAX_GAP_DEPTH = 128;                                // mbgccoder/MBGC_Params.h:50
constexpr int64_t MAX_EXTEND_LEFT = 1 << 24;                      // MAX_EXTEND_MATCH_LEFT_LENGTH, mbgccoder/MBGC_Params.h:55
constexpr int64_t REF_SHIFT = 1;                                  // mbgccoder/MBGC_Decoder.h:35

// ContextAwareMismatchesCoder::code2mis (coders/ContextAwareMismatchesCoder.cpp:8-17 from mis2code, .h:13-17): row = the
// reference's symbol (A C G T N), column = code 0..3
__constant__ uint8_t c_code2mis[5][4];
__constant__ uint8_t c_sym2val[256];

__device__ __forceinline__ uint8_t code2mismatch(uint8_t actual, uint8_t code) {  // .cpp:29-34
    if (code >= 4) return code;  // codes 4..31 never leave the encoder; >= 32 is a plain symbol
    const uint8_t row = c_sym2val[actual];
    return row > 4 ? code : c_code2mis[row][code];
}

// ---- marker passes ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DC_THREADS) k_dc_count(const uint8_t* __restrict__ src, uint64_t n, int b0, int b1,
                                                         uint32_t* __restrict__ cnt0, uint32_t* __restrict__ cnt1) {
    __shared__ uint32_t sSum;
    if (threadIdx.x == 0) sSum = 0;
    __syncthreads();
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * DC_COUNT_BLOCK + threadIdx.x * 16;
    uint32_t c0 = 0, c1 = 0;
    if (base < n) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + base);  // the buffers are padded
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const int lim = n - base < 16 ? static_cast<int>(n - base) : 16;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int v = (w[k >> 2] >> ((k & 3) * 8)) & 0xFF;
            if (k < lim) {
                c0 += v == b0;
                c1 += v == b1;
            }
        }
    }
    uint32_t packed = (c0 << 16) | c1;
    packed = __reduce_add_sync(0xFFFFFFFFu, packed);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sSum, packed);
    __syncthreads();
    if (threadIdx.x == 0) {
        cnt0[blockIdx.x] = sSum >> 16;
        if (cnt1) cnt1[blockIdx.x] = sSum & 0xFFFF;
    }
}

__global__ void __launch_bounds__(DC_THREADS) k_dc_scatter(const uint8_t* __restrict__ src, uint64_t n, int b0, int b1,
                                                           const uint32_t* __restrict__ off0, const uint32_t* __restrict__ off1,
                                                           uint32_t* __restrict__ pos0, uint32_t* __restrict__ rank1of0,
                                                           uint32_t* __restrict__ pos1) {
    __shared__ uint32_t sWarp[DC_THREADS / 32];
    const uint64_t base = static_cast<uint64_t>(blockIdx.x) * DC_COUNT_BLOCK + threadIdx.x * 16;
    uint32_t w[4] = {0, 0, 0, 0};
    int lim = 0;
    if (base < n) {
        const uint4 q = *reinterpret_cast<const uint4*>(src + base);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        lim = n - base < 16 ? static_cast<int>(n - base) : 16;
    }
    uint32_t c0 = 0, c1 = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int v = (w[k >> 2] >> ((k & 3) * 8)) & 0xFF;
        if (k < lim) {
            c0 += v == b0;
            c1 += v == b1;
        }
    }
    // exclusive scan of (c0, c1) over the block, packed into one word (both stay below 4096)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mine = (c0 << 16) | c1;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    uint32_t before = inc - mine;
    for (int k = 0; k < warp; k++) before += sWarp[k];
    uint32_t r0 = off0[blockIdx.x] + (before >> 16);
    uint32_t r1 = (off1 ? off1[blockIdx.x] : 0) + (before & 0xFFFF);
    if (c0 | c1) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int v = (w[k >> 2] >> ((k & 3) * 8)) & 0xFF;
            if (k < lim) {
                if (v == b0) {
                    pos0[r0] = static_cast<uint32_t>(base + k);
                    if (rank1of0) rank1of0[r0] = r1;
                    r0++;
                } else if (v == b1) {
                    if (pos1) pos1[r1] = static_cast<uint32_t>(base + k);
                    r1++;
                }
            }
        }
    }
}

// gap-delta bytes a contig consumes: one per match except its last (MBGC_Decoder.cpp:387: read only while another mark follows)
__global__ void __launch_bounds__(DC_THREADS) k_dc_contig_gd(uint32_t nContigs, const uint32_t* __restrict__ sepMarks, int gapDepthOffset,
                                                             uint32_t* __restrict__ gdCount) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nContigs) return;
    const uint32_t m = sepMarks[c] - (c ? sepMarks[c - 1] : 0);
    gdCount[c] = gapDepthOffset && m ? m - 1 : 0;
}

// non-zero gap deltas of a contig = its matches that omit their offset (warp per contig)
__global__ void __launch_bounds__(DC_THREADS) k_dc_contig_nz(uint32_t nContigs, const uint32_t* __restrict__ gdStart,
                                                             const uint8_t* __restrict__ gd, uint64_t nGd, uint32_t* __restrict__ nzCount) {
    const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= nContigs) return;
    uint64_t lo = gdStart[c], hi = gdStart[c + 1];
    if (hi > nGd) hi = nGd;
    uint32_t s = 0;
    for (uint64_t i = lo + lane; i < hi; i += 32) s += gd[i] != 0;
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if (lane == 0) nzCount[c] = s;
}

// ---- frugal length stream (utils/helper.h:256-272: u16, or 0xFFFF + u32, or 0xFFFF + 0xFFFFFFFF + u64) --------------------------
// A chunk map answers, for each of the 7 word offsets a record can start at: how many records start in the chunk and at which
// offset of the NEXT chunk the parse continues.  Maps compose associatively.
struct LenMap {
    uint32_t m[8];  // (count << 3) | exit offset; m[7] unused
};

__device__ __forceinline__ uint32_t len_word(const uint8_t* s, uint64_t w) { return s[2 * w] | (s[2 * w + 1] << 8); }

// parses the record at word p (p < nWords); returns its size in words, *v = value
__device__ __forceinline__ int len_record(const uint8_t* s, uint64_t nWords, uint64_t p, uint32_t* v) {
    const uint32_t w = len_word(s, p);
    if (w != 0xFFFF) {
        *v = w;
        return 1;
    }
    if (p + 2 >= nWords) {  // a truncated stream ends the parse
        *v = 0;
        return 7;
    }
    const uint32_t y = len_word(s, p + 1) | (len_word(s, p + 2) << 16);
    if (y != 0xFFFFFFFFu) {
        *v = y;
        return 3;
    }
    *v = p + 4 < nWords ? (len_word(s, p + 3) | (len_word(s, p + 4) << 16)) : 0;  // low half of the u64 (readUInt64Frugal<uint32_t>)
    return 7;
}

__global__ void __launch_bounds__(DC_THREADS) k_dl_map(const uint8_t* __restrict__ src, uint64_t nWords, LenMap* __restrict__ chunkMap,
                                                       uint64_t nChunks) {
    const uint64_t c = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= nChunks) return;
    const uint64_t lo = c * DC_LEN_CHUNK, hi = lo + DC_LEN_CHUNK;
    LenMap r;
    r.m[7] = 0;
    for (int e = 0; e < 7; e++) {
        uint64_t p = lo + e;
        uint32_t cnt = 0, v;
        while (p < hi && p < nWords) {
            p += len_record(src, nWords, p, &v);
            cnt++;
        }
        const uint32_t exitOff = p >= hi ? static_cast<uint32_t>(p - hi) : 0;
        r.m[e] = (cnt << 3) | (exitOff > 6 ? 6 : exitOff);
    }
    chunkMap[c] = r;
}

__device__ __forceinline__ LenMap len_compose(const LenMap& a, const LenMap& b) {  // a first, then b
    LenMap r;
#pragma unroll
    for (int e = 0; e < 7; e++) {
        const uint32_t x = b.m[a.m[e] & 7];
        r.m[e] = (((a.m[e] >> 3) + (x >> 3)) << 3) | (x & 7);
    }
    r.m[7] = 0;
    return r;
}

__global__ void __launch_bounds__(DC_THREADS) k_dl_blockmap(const LenMap* __restrict__ chunkMap, uint64_t nChunks, LenMap* __restrict__ blockMap,
                                                            uint64_t nBlocks) {
    const uint64_t b = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (b >= nBlocks) return;
    LenMap acc;
    for (int e = 0; e < 8; e++) acc.m[e] = e < 7 ? e : 0;  // identity
    const uint64_t lo = b * DC_LEN_BLOCK, hi = lo + DC_LEN_BLOCK < nChunks ? lo + DC_LEN_BLOCK : nChunks;
    for (uint64_t c = lo; c < hi; c++) acc = len_compose(acc, chunkMap[c]);
    blockMap[b] = acc;
}

// entry state of every block: (records so far << 3) | offset; one thread (the length stream is a few bytes per match)
__global__ void k_dl_blockscan(const LenMap* __restrict__ blockMap, uint64_t nBlocks, uint64_t* __restrict__ blockEntry,
                               uint64_t* __restrict__ total) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t off = 0;
    uint64_t cnt = 0;
    for (uint64_t b = 0; b < nBlocks; b++) {
        blockEntry[b] = (cnt << 3) | off;
        const uint32_t x = blockMap[b].m[off];
        cnt += x >> 3;
        off = x & 7;
    }
    *total = cnt;
}

__global__ void __launch_bounds__(DC_LEN_BLOCK) k_dl_emit(const uint8_t* __restrict__ src, uint64_t nWords, const LenMap* __restrict__ chunkMap,
                                                          uint64_t nChunks, const uint64_t* __restrict__ blockEntry,
                                                          uint32_t* __restrict__ out, uint64_t outCap) {
    __shared__ uint64_t sEntry[DC_LEN_BLOCK];
    const uint64_t b = blockIdx.x;
    const uint64_t lo = b * DC_LEN_BLOCK, hi = lo + DC_LEN_BLOCK < nChunks ? lo + DC_LEN_BLOCK : nChunks;
    if (threadIdx.x == 0) {
        uint64_t st = blockEntry[b];
        for (uint64_t c = lo; c < hi; c++) {
            sEntry[c - lo] = st;
            const uint32_t x = chunkMap[c].m[st & 7];
            st = (((st >> 3) + (x >> 3)) << 3) | (x & 7);
        }
    }
    __syncthreads();
    const uint64_t c = lo + threadIdx.x;
    if (c >= hi) return;
    const uint64_t st = sEntry[threadIdx.x];
    uint64_t k = st >> 3;
    const uint64_t end = c * DC_LEN_CHUNK + DC_LEN_CHUNK;
    uint64_t p = c * DC_LEN_CHUNK + (st & 7);
    while (p < end && p < nWords) {
        uint32_t v;
        p += len_record(src, nWords, p, &v);
        if (k < outCap) out[k] = v;
        k++;
    }
}

// ---- the parse: MBGC_Decoder.cpp:319-527 without the bytes ---------------------------------------------------------------------
struct ParseCtx {
    const DcParams& p;
    const DcStreams& s;
    uint64_t litPos, flagsPos;
    bool bad;
    int lane;
};

// The mismatch-score automaton of extendMatchRight / extendMatchLeft over a stretch of the flag stream, 32 flags per step of
// the WARP (every lane calls this with the same arguments).  One step of the reference's loops is
//     stop if the step limit is reached, the literal cursor sits at the guard (stopAtGuard), or the score reached the threshold;
//     read a flag; stop (flag consumed) if it is a mismatch and literal cursor + litBias == guard;
//     score = mismatch ? score + penalty : max(0, score - bonus); a mismatch moves the literal cursor.
// The cursor before step i is l0 + (mismatches before i): a ballot.  The score before step i is a Lindley recursion,
// W_i = P_i + max(W_0, -min_{1<=j<=i} P_j) with P the prefix sums of the +penalty / -bonus steps: two warp scans.
struct Walk {
    uint64_t steps;   // complete steps (= flags consumed by them)
    uint64_t mism;    // mismatches among them
    bool atGuard;     // stopped by a mismatch flag at the guard (that flag is consumed as well)
    bool limit;       // stopped by the step limit
};

__device__ Walk warp_walk(ParseCtx& x, uint64_t fp, uint64_t l0, uint64_t guard, uint32_t litBias, uint64_t maxSteps, bool useScore,
                          bool stopAtGuard) {
    const int lane = x.lane;
    const uint32_t lt = (1u << lane) - 1u;
    Walk w{0, 0, false, false};
    int s0 = x.p.initialScore;
    for (;;) {
        const uint64_t i = w.steps + lane;
        const bool inRange = fp + i < x.s.nFlags;
        const bool f = inRange ? x.s.flags[fp + i] != 0 : true;
        const uint32_t ball = __ballot_sync(0xFFFFFFFFu, f);
        const uint32_t c = __popc(ball & lt);
        const uint64_t l = l0 + w.mism + c;
        int sc = 0;
        if (useScore) {
            // one scan of (sum, least non-empty prefix sum): (a then b) = (a.sum + b.sum, min(a.min, a.sum + b.min))
            const int d = f ? x.p.mismatchPenalty : -x.p.matchBonus;
            int P = d, M = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tp = __shfl_up_sync(0xFFFFFFFFu, P, o);
                const int tm = __shfl_up_sync(0xFFFFFFFFu, M, o);
                if (lane >= o) {
                    const int m2 = tp + M;   // the earlier block comes first: its sum shifts this block's prefixes
                    M = tm < m2 ? tm : m2;
                    P += tp;
                }
            }
            // this step starts from what the steps before it left: the inclusive values of the lane below
            const int Pe = __shfl_up_sync(0xFFFFFFFFu, P, 1);
            const int Me = __shfl_up_sync(0xFFFFFFFFu, M, 1);
            sc = lane == 0 ? s0 : Pe + (s0 > -Me ? s0 : -Me);
        }
        const bool lim = i >= maxSteps;
        const bool A = lim || (stopAtGuard && l == guard) || (useScore && sc >= x.p.scoreThreshold);
        const bool B = f && (l + litBias == guard);
        const uint32_t stop = __ballot_sync(0xFFFFFFFFu, A || !inRange || B);
        if (stop) {
            const int k = __ffs(stop) - 1;
            const uint32_t packed = (A ? 1u : 0u) | (inRange ? 2u : 0u) | (lim ? 4u : 0u) | (c << 8);
            const uint32_t pk = __shfl_sync(0xFFFFFFFFu, packed, k);
            w.steps += k;
            w.mism += pk >> 8;
            if (pk & 1u) {
                w.limit = (pk & 4u) != 0;
            } else if (!(pk & 2u)) {
                x.bad = true;  // the automaton wants a flag behind the stream's end
            } else {
                w.atGuard = true;
            }
            return w;
        }
        w.steps += 32;
        w.mism += __popc(ball);
        if (useScore) {
            const int after = f ? sc + x.p.mismatchPenalty : (sc - x.p.matchBonus < 0 ? 0 : sc - x.p.matchBonus);
            s0 = __shfl_sync(0xFFFFFFFFu, after, 31);
        }
    }
}

// extendMatchRight (:434-462): returns the extension's length; the segment describes what the copy pass has to write
__device__ uint64_t parse_extend_right(ParseCtx& x, uint64_t destSize, int64_t offsetDelta, bool isGap, bool gapStart, bool gapMiddle,
                                       bool gapEnd, uint64_t guardLitPos, DcSeg& seg) {
    seg.len = 0;
    if (x.litPos == guardLitPos && !gapMiddle) return 0;
    const bool implicitFirst = gapStart || !isGap;
    seg.src = static_cast<uint64_t>(static_cast<int64_t>(destSize) + offsetDelta);
    seg.litPos = static_cast<uint32_t>(x.litPos);
    uint64_t n = 0;
    if (implicitFirst) {
        x.litPos++;
        n++;
    }
    seg.flagsPos = static_cast<uint32_t>(x.flagsPos);
    // while ((!gapEnd || litPos != guard) && (isGap || score < threshold)) { flag; if (flag && litPos == guard) break; ... }
    const Walk w = warp_walk(x, x.flagsPos, x.litPos, guardLitPos, 0, UINT64_MAX, !isGap, gapEnd);
    x.flagsPos += w.steps + (w.atGuard ? 1 : 0);
    x.litPos += w.mism;
    n += w.steps;
    seg.len = static_cast<uint32_t>(n);
    seg.meta = DC_EXT_RIGHT | (implicitFirst ? 4u : 0u);
    // the reference's buffer has maxConsecutiveMismatches (a uint8) spare bytes behind refTotalLength (:1083); ours has 256
    if (static_cast<int64_t>(seg.src) < 0 || seg.src + n > x.p.refTotalLength + 255 || n >= (1ull << 32)) x.bad = true;
    return n;
}

// extendMatchLeft (:464-527): may move matchSrcPos (an omitted offset points at the extension's left end)
__device__ uint64_t parse_extend_left(ParseCtx& x, uint64_t& matchSrcPos, bool skipOffset, uint64_t refLockPos, uint64_t markPos, DcSeg& seg) {
    seg.len = 0;
    int64_t srcMatch = static_cast<int64_t>(matchSrcPos);
    int64_t srcGuard = srcMatch - MAX_EXTEND_LEFT;
    if (!skipOffset) {
        if (srcGuard < REF_SHIFT) srcGuard = REF_SHIFT;
        const int64_t srcLock = static_cast<int64_t>(refLockPos);
        if (srcGuard < srcLock && srcLock <= srcMatch) srcGuard = srcLock;
    }
    if (srcGuard == srcMatch) return 0;
    if (skipOffset) {
        // while (--srcPtr > srcGuard && score < threshold) { flag (peeked); if (flag && litPos + ++mismatches == markPos) break; ... }
        const int64_t room = srcMatch - 2 - srcGuard;
        const Walk w = warp_walk(x, x.flagsPos, x.litPos, markPos, 1, room > 0 ? static_cast<uint64_t>(room) : 0, true, false);
        if (x.bad) return 0;
        uint64_t length = w.steps + (w.atGuard ? 1 : 0), mismatches = w.mism + (w.atGuard ? 1 : 0);
        if (w.limit && !w.atGuard) {  // srcPtr == srcGuard && isExtLenKnown
            mismatches++;
            length++;
        }
        const uint64_t matchingChars = length - mismatches;
        matchSrcPos += matchingChars;
        srcGuard += static_cast<int64_t>(matchingChars);
        srcMatch += static_cast<int64_t>(matchingChars);
    }
    seg.src = static_cast<uint64_t>(srcMatch);  // the copy pass reads src - 1 - w for the w-th byte it walks
    seg.litPos = static_cast<uint32_t>(x.litPos);
    seg.flagsPos = static_cast<uint32_t>(x.flagsPos);
    x.litPos++;
    // while (--srcPtr >= srcGuard && score < threshold) { flag; if (flag && litPos == markPos) break; ... }
    const int64_t room = srcMatch - 1 - srcGuard;
    const Walk w = warp_walk(x, x.flagsPos, x.litPos, markPos, 0, room > 0 ? static_cast<uint64_t>(room) : 0, true, false);
    x.flagsPos += w.steps + (w.atGuard ? 1 : 0);
    x.litPos += w.mism;
    const uint64_t n = 1 + w.steps;
    seg.len = static_cast<uint32_t>(n);
    seg.meta = DC_EXT_LEFT | 4u;
    if (srcMatch < static_cast<int64_t>(n) || static_cast<uint64_t>(srcMatch) > x.p.refTotalLength) x.bad = true;
    return n;
}

__device__ __forceinline__ void put_seg(DcSeg* segs, uint64_t at, const DcSeg& g) {
    if (threadIdx.x) return;  // the lanes hold identical values
    uint4* d = reinterpret_cast<uint4*>(segs + at);
    const uint4* q = reinterpret_cast<const uint4*>(&g);
    d[0] = q[0];
    d[1] = q[1];
}

// One WARP per target.  All 32 lanes run the (scalar) replay with identical values -- for a warp that costs what one lane
// would -- so that the flag walks of the extensions, where the time goes, can use the whole warp (warp_walk).  32 targets in one
// warp would instead run their completely divergent parses one after the other.
__global__ void __launch_bounds__(32) k_dc_parse(DcParams p, DcStreams s, DcIndex ix, const DcTarget* __restrict__ targets, uint32_t nTargets,
                                                  DcSeg* __restrict__ segs, uint64_t* __restrict__ contigLen,
                                                  uint64_t* __restrict__ contigUnmatched, uint32_t* __restrict__ error) {
    const uint32_t t = blockIdx.x;
    if (t >= nTargets) return;
    const int lane = threadIdx.x;
    const DcTarget tg = targets[t];
    __shared__ int64_t pairedGapOffsetDelta[MAX_GAP_DEPTH];  // every lane reads and writes the same entries
    ParseCtx x{p, s, 0, 0, false, lane};
    const uint32_t c0 = tg.firstContig;
    x.litPos = c0 ? static_cast<uint64_t>(ix.sepPos[c0 - 1]) + 1 : 0;
    x.flagsPos = (tg.flagIdx == UINT32_MAX || tg.flagIdx == 0) ? 0 : static_cast<uint64_t>(ix.flagSep[tg.flagIdx - 1]) + 1;
    if (tg.flagIdx != UINT32_MAX && tg.flagIdx > ix.nFlagSep) x.bad = true;
    const uint64_t refLockPos = tg.lockPos;
    const int64_t NO_GAP = -1;
    for (uint32_t c = c0; c < c0 + tg.nContigs && !x.bad; c++) {
        const uint64_t marksLo = c ? ix.sepMarks[c - 1] : 0, marksHi = ix.sepMarks[c];
        const uint64_t seqEnd = ix.sepPos[c];
        uint64_t gdPos = ix.gdStart[c];
        uint64_t offIdx = marksLo - ix.nzStart[c];
        uint64_t destSize = 0, unmatched = 0;
        __syncwarp();
        for (int i = lane; i < MAX_GAP_DEPTH; i += 32) pairedGapOffsetDelta[i] = INT64_MAX;
        __syncwarp();
        int64_t gapStartIdx = NO_GAP, gapEndIdx = NO_GAP;
        int gapCurIdx = 0;
        int64_t offsetDelta = -1;
        uint64_t extRightLen = 0;
        bool isGap = false;
        int64_t j = 0;
        DcSeg g;
        for (uint64_t mk = marksLo; mk < marksHi; mk++, j++) {
            const uint64_t segAt = 4 * mk + c;
            const uint64_t markPos = ix.markPos[mk];
            const uint64_t literalsLeft = markPos - x.litPos;
            uint64_t matchSrcPos;
            const int64_t pairedDelta = pairedGapOffsetDelta[gapCurIdx];
            const bool skipOffset = pairedDelta != INT64_MAX;
            __syncwarp();  // every lane has read the entry before lane 0 clears it
            if (skipOffset) {
                matchSrcPos = static_cast<uint64_t>(pairedDelta + static_cast<int64_t>(destSize) + static_cast<int64_t>(literalsLeft));
                if (lane == 0) pairedGapOffsetDelta[gapCurIdx] = INT64_MAX;
                __syncwarp();
            } else {
                if (offIdx >= s.nOff) { x.bad = true; break; }
                const uint8_t* o = s.off + 4 * offIdx;
                matchSrcPos = o[0] | (o[1] << 8) | (o[2] << 16) | (static_cast<uint32_t>(o[3]) << 24);
                if (p.refTotalLength > UINT32_MAX) {
                    if (offIdx >= s.nOff5) { x.bad = true; break; }
                    matchSrcPos += static_cast<uint64_t>(s.off5[offIdx]) << 32;
                }
                offIdx++;
            }
            uint64_t extLeftLen = 0;
            DcSeg left;
            left.len = 0;
            left.src = left.dst = 0;
            left.flagsPos = left.litPos = 0;
            left.meta = DC_EXT_LEFT;
            if (p.enableExt) {
                if (!isGap && literalsLeft) extLeftLen = parse_extend_left(x, matchSrcPos, skipOffset, refLockPos, markPos, left);
                if (gapEndIdx == j) {
                    gapStartIdx = NO_GAP;
                    gapEndIdx = NO_GAP;
                }
            }
            if (x.bad) break;
            const uint64_t plain = markPos - x.litPos;
            unmatched += plain + extLeftLen + extRightLen;
            g.dst = destSize; g.src = x.litPos; g.len = static_cast<uint32_t>(plain); g.flagsPos = g.litPos = 0; g.meta = DC_LITERAL | (c << 4);
            put_seg(segs, segAt, g);
            destSize += plain;
            left.dst = destSize;
            left.meta |= c << 4;
            put_seg(segs, segAt + 1, left);
            destSize += extLeftLen;
            x.litPos = markPos + 1;

            if (mk >= s.nLen) { x.bad = true; break; }
            const uint32_t matchLength = s.len[mk];
            if (matchSrcPos + matchLength > p.refTotalLength) { x.bad = true; break; }  // the copy pass must stay inside the reference
            const uint64_t prevMatchDestPos = destSize;
            g.dst = destSize; g.src = matchSrcPos; g.len = matchLength; g.meta = DC_MATCH | (c << 4);
            put_seg(segs, segAt + 2, g);
            destSize += matchLength;

            const bool more = mk + 1 < marksHi;
            const uint64_t nextMark = more ? ix.markPos[mk + 1] : seqEnd;
            if (p.gapDepthOffset && more) {
                if (gdPos >= s.nGd) { x.bad = true; break; }
                uint8_t gapDelta = s.gd[gdPos++];
                if (gapDelta) {
                    int gapIdx = gapCurIdx;
                    int gg = gapDelta;
                    if (!p.lazy && gapStartIdx == NO_GAP && nextMark - x.litPos == 0) {
                        gapIdx = (gapIdx + 1) % MAX_GAP_DEPTH;
                        gg++;
                    }
                    int guard = 0;
                    while (gapDelta) {
                        gapIdx = (gapIdx + 1) % MAX_GAP_DEPTH;
                        if (pairedGapOffsetDelta[gapIdx] == INT64_MAX)
                            gapDelta--;
                        else
                            gg++;
                        if (++guard > 4 * MAX_GAP_DEPTH) { x.bad = true; break; }  // a full ring: corrupt stream
                    }
                    __syncwarp();
                    if (lane == 0) pairedGapOffsetDelta[gapIdx] = static_cast<int64_t>(matchSrcPos) - static_cast<int64_t>(prevMatchDestPos);
                    __syncwarp();
                    if (p.enableExt && gapEndIdx <= j + gg && gg <= p.gapDepthMismatches) {
                        gapStartIdx = j;
                        gapEndIdx = j + gg;
                    }
                }
            }
            gapCurIdx = (gapCurIdx + 1) % MAX_GAP_DEPTH;
            const bool gapStart = gapStartIdx == j;
            const bool gapEnd = gapEndIdx == j + 1;
            const bool gapMiddle = gapStartIdx < j && j + 1 < gapEndIdx;
            isGap = gapStart || gapMiddle || gapEnd;
            extRightLen = 0;
            g.len = 0; g.dst = destSize; g.src = 0; g.flagsPos = g.litPos = 0; g.meta = DC_EXT_RIGHT;
            if (p.enableExt) {
                if (!isGap || gapStart) offsetDelta = static_cast<int64_t>(matchSrcPos) + matchLength - static_cast<int64_t>(destSize);
                extRightLen = parse_extend_right(x, destSize, offsetDelta, isGap, gapStart, gapMiddle, gapEnd, nextMark, g);
            }
            g.dst = destSize;
            g.meta |= c << 4;
            put_seg(segs, segAt + 3, g);
            destSize += extRightLen;
            if (x.bad) break;
        }
        if (x.bad) break;
        const uint64_t plain = seqEnd - x.litPos;
        unmatched += plain + extRightLen;
        g.dst = destSize; g.src = x.litPos; g.len = static_cast<uint32_t>(plain); g.flagsPos = g.litPos = 0; g.meta = DC_LITERAL | (c << 4);
        put_seg(segs, 4 * marksHi + c, g);
        destSize += plain;
        x.litPos = seqEnd + 1;
        if (lane == 0) {
            contigLen[c] = destSize;
            contigUnmatched[c] = unmatched & 0xFFFFFFFFull;  // uint32_t unmatchedChars (:329)
        }
    }
    if (x.bad && lane == 0) atomicOr(error, 1u);
}

// ---- the copy pass --------------------------------------------------------------------------------------------------------------
// which segment covers the first byte of every output tile (each tile start lies in exactly one segment of non-zero length)
__global__ void __launch_bounds__(DC_THREADS) k_dc_tiles(const DcSeg* __restrict__ segs, uint64_t segBegin, uint64_t segEnd,
                                                         const uint64_t* __restrict__ contigBase, uint64_t dstBegin, int tileShift,
                                                         uint32_t* __restrict__ tileSeg) {
    const uint64_t s = segBegin + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= segEnd) return;
    const uint32_t len = segs[s].len;
    if (!len) return;
    const uint64_t a = contigBase[segs[s].meta >> 4] + segs[s].dst;
    const uint64_t tileBytes = 1ull << tileShift;
    if (a == dstBegin) tileSeg[a >> tileShift] = static_cast<uint32_t>(s);
    for (uint64_t t = (a + tileBytes - 1) >> tileShift; (t << tileShift) < a + len; t++) tileSeg[t] = static_cast<uint32_t>(s);
}

// dst[0 .. n) = src[0 .. n) by `nthr` threads (idx = this thread's rank), n at most a tile: 16-byte aligned stores, the source
// through two aligned 16-byte loads and a byte shift (sources are padded: up to 31 bytes behind src + n are read, never used).
// The word offset of the shift is the same for the whole call: one loop per offset, so the words are picked at compile time.
// Every lane loads ONE aligned 16-byte piece; the piece behind it comes from the next lane by shuffle (lane 31 loads it): the
// kernel was bound by L1 sector throughput with two loads per store.  The calling warp is converged (idx = lane or threadIdx.x).
template <int J>
__device__ __forceinline__ void copy_vectors(uint4* __restrict__ d, const uint4* __restrict__ q, uint32_t nvec, uint32_t bits, int idx, int nthr) {
    const int lane = idx & 31;
    for (uint32_t v0 = idx - lane; v0 < nvec; v0 += nthr) {  // warp-uniform
        const uint32_t v = v0 + lane;
        const uint4 lo = q[v < nvec ? v : nvec];   // q[nvec] is readable (the buffers are padded)
        uint4 hi;
        hi.x = __shfl_down_sync(0xFFFFFFFFu, lo.x, 1);
        hi.y = __shfl_down_sync(0xFFFFFFFFu, lo.y, 1);
        hi.z = __shfl_down_sync(0xFFFFFFFFu, lo.z, 1);
        hi.w = __shfl_down_sync(0xFFFFFFFFu, lo.w, 1);
        if (lane == 31 && v < nvec) hi = q[v + 1];
        if (v >= nvec) continue;
        const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        d[v] = make_uint4(__funnelshift_r(W[J], W[J + 1], bits), __funnelshift_r(W[J + 1], W[J + 2], bits),
                          __funnelshift_r(W[J + 2], W[J + 3], bits), __funnelshift_r(W[J + 3], W[J + 4], bits));
    }
}

__device__ __forceinline__ void copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint64_t n64, int idx, int nthr) {
    const uint32_t n = static_cast<uint32_t>(n64);
    uint32_t head = (16u - (static_cast<uint32_t>(reinterpret_cast<uintptr_t>(dst)) & 15u)) & 15u;
    if (head > n) head = n;
    for (uint32_t i = idx; i < head; i += nthr) dst[i] = src[i];
    const uint32_t nvec = (n - head) >> 4;
    const uint8_t* sp = src + head;
    const uint32_t sh = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(sp)) & 15u;
    const uint4* q = reinterpret_cast<const uint4*>(sp - sh);
    uint4* d = reinterpret_cast<uint4*>(dst + head);
    const uint32_t bits = (sh & 3u) * 8u;
    if (sh == 0) {
        for (uint32_t v = idx; v < nvec; v += nthr) d[v] = q[v];
    } else {
        switch (sh >> 2) {  // uniform over the call
            case 0: copy_vectors<0>(d, q, nvec, bits, idx, nthr); break;
            case 1: copy_vectors<1>(d, q, nvec, bits, idx, nthr); break;
            case 2: copy_vectors<2>(d, q, nvec, bits, idx, nthr); break;
            default: copy_vectors<3>(d, q, nvec, bits, idx, nthr); break;
        }
    }
    for (uint32_t i = head + (nvec << 4) + idx; i < n; i += nthr) dst[i] = src[i];
}

// one extension by one warp: byte w of the walk reads the reference at src + w (right) or src - 1 - w (left), is a mismatch
// if it is the implicit first byte or its flag is set, and then takes the next mismatch code of the literal stream
__device__ __forceinline__ void copy_extension(const DcSeg& g, uint8_t* __restrict__ out, const uint8_t* __restrict__ ref,
                                               const DcStreams& s, int lane) {
    const bool left = (g.meta & 3) == DC_EXT_LEFT;
    const bool implicitFirst = (g.meta & 4) != 0;
    uint32_t used = 0;  // mismatch codes consumed so far
    for (uint32_t w0 = 0; w0 < g.len; w0 += 32) {
        const uint32_t w = w0 + lane;
        bool mm = false;
        if (w < g.len) {
            if (implicitFirst) mm = w == 0 ? true : s.flags[g.flagsPos + w - 1] != 0;
            else mm = s.flags[g.flagsPos + w] != 0;
        }
        const uint32_t ball = __ballot_sync(0xFFFFFFFFu, mm);
        if (w < g.len) {
            const uint8_t r = ref[left ? g.src - 1 - w : g.src + w];
            const uint8_t b = mm ? code2mismatch(r, s.lit[g.litPos + used + __popc(ball & ((1u << lane) - 1))]) : r;
            out[left ? g.len - 1 - w : w] = b;
        }
        used += __popc(ball);
    }
}

// One CTA per tile of output (8 to 64 KiB, chosen per batch from the mean match length).  The segments that may touch the tile are staged in shared memory a chunk at a time (one
// coalesced load instead of a chain of dependent ones); long copies are done by the whole CTA, everything else by one warp per
// segment.  Destinations are disjoint, so the warps never wait for each other.
constexpr int DC_STAGE = 256;

__global__ void __launch_bounds__(DC_THREADS) k_dc_copy(const DcSeg* __restrict__ segs, uint64_t segEnd, const uint64_t* __restrict__ contigBase,
                                                        const uint32_t* __restrict__ tileSeg, uint64_t tileBegin, int tileShift,
                                                        uint64_t dstBegin, uint64_t dstEnd, const uint8_t* __restrict__ ref, DcStreams s,
                                                        uint8_t* __restrict__ dna) {
    __shared__ uint4 sSeg[DC_STAGE][2];
    __shared__ uint64_t sAbs[DC_STAGE];
    const uint64_t tile = tileBegin + blockIdx.x;
    uint64_t lo = tile << tileShift, hi = lo + (1ull << tileShift);
    if (lo < dstBegin) lo = dstBegin;
    if (hi > dstEnd) hi = dstEnd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int chunk = 32;  // most tiles of a near-identical collection lie inside one long match: look at few segments first
    for (uint64_t base = tileSeg[tile]; base < segEnd; base += chunk, chunk = DC_STAGE) {
        __syncthreads();
        if (threadIdx.x < chunk) {
            const uint64_t si = base + threadIdx.x;
            uint64_t a = UINT64_MAX;  // behind the list: ends the walk
            uint4 q0 = make_uint4(0, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
            if (si < segEnd) {
                const uint4* q = reinterpret_cast<const uint4*>(segs + si);
                q0 = q[0];
                q1 = q[1];
                // empty segments carry their position too: the walk over a tile ends at the first segment behind it, empty or
                // not (most slots of a near-identical collection are empty: no literals, no extensions between two matches)
                a = contigBase[q1.w >> 4] + (q0.x | (static_cast<uint64_t>(q0.y) << 32));
            }
            sSeg[threadIdx.x][0] = q0;
            sSeg[threadIdx.x][1] = q1;
            sAbs[threadIdx.x] = a;
        }
        __syncthreads();
        bool done = false;
        // long copies: every thread
        for (int k = 0; k < chunk; k++) {
            const uint4 q1 = sSeg[k][1];
            const uint64_t a = sAbs[k];
            if (a >= hi) {  // (UINT64_MAX: behind the list)
                done = true;
                break;
            }
            if (!q1.x || (q1.w & 3) > DC_MATCH) continue;
            const uint64_t i0 = a > lo ? a : lo;
            const uint64_t i1 = a + q1.x < hi ? a + q1.x : hi;
            if (i1 <= i0 || i1 - i0 < 1024) continue;
            const uint4 q0 = sSeg[k][0];
            const uint64_t src = q0.z | (static_cast<uint64_t>(q0.w) << 32);
            copy_bytes(dna + i0, ((q1.w & 3) == DC_MATCH ? ref : s.lit) + src + (i0 - a), i1 - i0, threadIdx.x, DC_THREADS);
        }
        // short copies and extensions: a warp each
        for (int k = warp; k < chunk; k += DC_THREADS / 32) {
            const uint4 q1 = sSeg[k][1];
            const uint64_t a = sAbs[k];
            if (a >= hi) break;
            if (!q1.x) continue;
            const uint4 q0 = sSeg[k][0];
            const uint32_t kind = q1.w & 3;
            const uint64_t src = q0.z | (static_cast<uint64_t>(q0.w) << 32);
            if (kind <= DC_MATCH) {
                const uint64_t i0 = a > lo ? a : lo;
                const uint64_t i1 = a + q1.x < hi ? a + q1.x : hi;
                if (i1 <= i0 || i1 - i0 >= 1024) continue;
                copy_bytes(dna + i0, (kind == DC_MATCH ? ref : s.lit) + src + (i0 - a), i1 - i0, lane, 32);
            } else if (a >= lo) {  // an extension is written whole by the tile it starts in
                DcSeg g;
                g.dst = 0; g.src = src; g.len = q1.x; g.flagsPos = q1.y; g.litPos = q1.z; g.meta = q1.w;
                copy_extension(g, dna + a, ref, s, lane);
            }
        }
        if (done) break;
    }
}

}  // namespace

void dc_upload_tables(cudaStream_t st) {
    // mis2code (coders/ContextAwareMismatchesCoder.h:13-17) inverted as the reference's static initializer does (.cpp:8-17)
    static const uint8_t mis2code[5][5] = {{'A', 2, 0, 1, 3}, {1, 'C', 2, 0, 3}, {0, 2, 'G', 1, 3}, {1, 0, 2, 'T', 3}, {1, 2, 3, 0, 'N'}};
    static const char sym[] = "ACGTN";
    uint8_t c2m[5][4], s2v[256];
    for (int i = 0; i < 256; i++) s2v[i] = 255;
    for (int i = 0; i < 5; i++) {
        s2v[static_cast<uint8_t>(sym[i])] = static_cast<uint8_t>(i);
        for (int k = 0; k < 4; k++) c2m[i][k] = static_cast<uint8_t>(k);
        for (int j = 0; j < 5; j++)
            if (mis2code[i][j] < 4) c2m[i][mis2code[i][j]] = static_cast<uint8_t>(sym[j]);
    }
    cudaMemcpyToSymbolAsync(c_code2mis, c2m, sizeof(c2m), 0, cudaMemcpyHostToDevice, st);
    cudaMemcpyToSymbolAsync(c_sym2val, s2v, sizeof(s2v), 0, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);  // the sources are on this stack
}

static unsigned blocks_for(uint64_t items, uint64_t per) { return static_cast<unsigned>((items + per - 1) / per); }

int launch_dc_count(const uint8_t* src, uint64_t n, int b0, int b1, uint32_t* cnt0, uint32_t* cnt1, cudaStream_t st) {
    if (!n) return 0;
    k_dc_count<<<blocks_for(n, DC_COUNT_BLOCK), DC_THREADS, 0, st>>>(src, n, b0, b1, cnt0, cnt1);
    return 1;
}

int launch_dc_scatter(const uint8_t* src, uint64_t n, int b0, int b1, const uint32_t* off0, const uint32_t* off1, uint32_t* pos0,
                      uint32_t* rank1of0, uint32_t* pos1, cudaStream_t st) {
    if (!n) return 0;
    k_dc_scatter<<<blocks_for(n, DC_COUNT_BLOCK), DC_THREADS, 0, st>>>(src, n, b0, b1, off0, off1, pos0, rank1of0, pos1);
    return 1;
}

int launch_dc_contig_gd(uint32_t nContigs, const uint32_t* sepMarks, int gapDepthOffset, uint32_t* gdCount, cudaStream_t st) {
    if (!nContigs) return 0;
    k_dc_contig_gd<<<blocks_for(nContigs, DC_THREADS), DC_THREADS, 0, st>>>(nContigs, sepMarks, gapDepthOffset, gdCount);
    return 1;
}

int launch_dc_contig_nz(uint32_t nContigs, const uint32_t* gdStart, const uint8_t* gd, uint64_t nGd, uint32_t* nzCount, cudaStream_t st) {
    if (!nContigs) return 0;
    k_dc_contig_nz<<<blocks_for(static_cast<uint64_t>(nContigs) * 32, DC_THREADS), DC_THREADS, 0, st>>>(nContigs, gdStart, gd, nGd, nzCount);
    return 1;
}

uint64_t dc_len_chunks(uint64_t nBytes) { return (nBytes / 2 + DC_LEN_CHUNK - 1) / DC_LEN_CHUNK; }
uint64_t dc_len_blocks(uint64_t nBytes) { return (dc_len_chunks(nBytes) + DC_LEN_BLOCK - 1) / DC_LEN_BLOCK; }

int launch_dc_len_expand(const uint8_t* src, uint64_t nBytes, uint32_t* chunkMap, uint32_t* blockMap, uint64_t* blockEntry,
                         uint64_t* total, uint32_t* out, uint64_t outCap, cudaStream_t st) {
    const uint64_t nWords = nBytes / 2, nChunks = dc_len_chunks(nBytes), nBlocks = dc_len_blocks(nBytes);
    if (!nChunks) {
        cudaMemsetAsync(total, 0, 8, st);
        return 0;
    }
    k_dl_map<<<blocks_for(nChunks, DC_THREADS), DC_THREADS, 0, st>>>(src, nWords, reinterpret_cast<LenMap*>(chunkMap), nChunks);
    k_dl_blockmap<<<blocks_for(nBlocks, DC_THREADS), DC_THREADS, 0, st>>>(reinterpret_cast<const LenMap*>(chunkMap), nChunks,
                                                                         reinterpret_cast<LenMap*>(blockMap), nBlocks);
    k_dl_blockscan<<<1, 32, 0, st>>>(reinterpret_cast<const LenMap*>(blockMap), nBlocks, blockEntry, total);
    k_dl_emit<<<static_cast<unsigned>(nBlocks), DC_LEN_BLOCK, 0, st>>>(src, nWords, reinterpret_cast<const LenMap*>(chunkMap), nChunks, blockEntry,
                                                                      out, outCap);
    return 4;
}

int launch_dc_parse(const DcParams& p, const DcStreams& s, const DcIndex& ix, const DcTarget* targets, uint32_t nTargets, DcSeg* segs,
                    uint64_t* contigLen, uint64_t* contigUnmatched, uint32_t* error, cudaStream_t st) {
    if (!nTargets) return 0;
    k_dc_parse<<<nTargets, 32, 0, st>>>(p, s, ix, targets, nTargets, segs, contigLen, contigUnmatched, error);
    return 1;
}

int launch_dc_tiles(const DcSeg* segs, uint64_t segBegin, uint64_t segEnd, const uint64_t* contigBase, uint64_t dstBegin, int tileShift,
                    uint32_t* tileSeg, cudaStream_t st) {
    if (segEnd <= segBegin) return 0;
    k_dc_tiles<<<blocks_for(segEnd - segBegin, DC_THREADS), DC_THREADS, 0, st>>>(segs, segBegin, segEnd, contigBase, dstBegin, tileShift, tileSeg);
    return 1;
}

int launch_dc_copy(const DcSeg* segs, uint64_t segEnd, const uint64_t* contigBase, const uint32_t* tileSeg, int tileShift, uint64_t dstBegin,
                   uint64_t dstEnd, const uint8_t* ref, const DcStreams& s, uint8_t* dna, cudaStream_t st) {
    if (dstEnd <= dstBegin) return 0;
    const uint64_t tileBegin = dstBegin >> tileShift, tileEnd = ((dstEnd - 1) >> tileShift) + 1;
    k_dc_copy<<<static_cast<unsigned>(tileEnd - tileBegin), DC_THREADS, 0, st>>>(segs, segEnd, contigBase, tileSeg, tileBegin, tileShift, dstBegin,
                                                                                dstEnd, ref, s, dna);
    return 1;
}

}  // namespace mbgc_b200
