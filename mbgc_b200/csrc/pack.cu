// mbgc_b200/csrc/pack.cu -- K6: match/literal stream packing on the device (sm_100a).
//
// Replaces MBGC_Encoder::processMatches (/root/reference/mbgccoder/MBGC_Encoder.cpp:143-308) with
// extendMatchRight (:310-371), extendMatchLeft (:373-427) and ContextAwareMismatchesCoder::mismatch2code
// (coders/ContextAwareMismatchesCoder.cpp:22-27) for a whole batch of texts at once.  All byte / integer work.
//
// The reference walks the match list of one contig with a state machine.  Its state is small, and every piece
// of it is a function of a bounded window of earlier matches -- so each piece is recomputed in parallel:
//
//   pass 1 (:153-199)  "drop match j if j+1 continues the diagonal of the last kept match and j is short".
//                      removed(j) depends on removed(i < j) only (last kept match; j's length grows by a left
//                      walk when j-1 was dropped).  The recurrence has exactly one solution, found left to
//                      right; any fixed point of "recompute every removed(j) from the current array" satisfies
//                      it at every j and therefore IS that solution.  k_pk_removal is one such sweep over all
//                      matches of the batch; sweeps repeat until one changes nothing (two in practice: the
//                      second only confirms; the host checks the flag of the last sweep and adds sweeps if not).
//   pairing (:229-273) match j "claims" the first later match (within gapDepthOffsetEncoding) on its diagonal and
//                      on its side of the lock position that no earlier match claimed (pairedGap ring), subject
//                      to the lazy-decompression region rule (:252-261, nextSrcRegionLoadingPos inherited along
//                      the chain).  claim(j) depends on claims of i < j only: same fixed-point argument,
//                      k_pk_claim_* sweeps.  The claimed match omits its offset; the gapDelta byte is g - reduce.
//   gap state (:219-222, 263-279)  with gapDepthMismatchesEncoding <= 2 the (gapStartIdx, gapEndIdx) pair at step
//                      j is a closed form of claim(j) and claim(j-1) (derivation in gap_state()).
//   extensions + emission (:214-287)  everything the reference appends between two consecutive matches -- the
//                      right extension of the earlier one, the left extension of the later one, the remaining
//                      plain literals, MATCH_MARK, offset / length / gapDelta -- depends only on those two matches
//                      and their gap state.  One thread per such "gap" runs the score automaton twice: counting
//                      (k_pk_count), then, after an exclusive scan of the six per-gap byte counts, writing
//                      (k_pk_emit) straight into the final streams.  Plain literal runs are copied by the warp
//                      when longer than 16 bytes and by a separate grid (k_pk_copy_long) from 4 KiB on.
#include "pack.h"

namespace mbgc_b200 {

namespace {

constexpr uint8_t MATCH_MARK = static_cast<uint8_t>('%' + 128);   // mbgccoder/MBGC_Params.h:45
constexpr uint8_t REF_REGION_SEPARATOR = 0;                       // matching/MGMP_Params.h:14
constexpr int64_t MAX_EXTEND_MATCH_LEFT_LENGTH = 1 << 24;         // mbgccoder/MBGC_Params.h:55
constexpr uint64_t REF_SHIFT = 1;
constexpr uint64_t REF_READ_SLACK = 8191;                         // readable bytes behind the reference buffer

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan over the block (THREADS threads, at most 1024); *total = block sum
template <int THREADS = PK_THREADS>
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* total, uint32_t* sWarp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan_u32(v, lane);
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < (THREADS >> 5) ? sWarp[lane] : 0;
        const uint32_t wi = warp_incl_scan_u32(w, lane);
        sWarp[lane] = wi - w;
        if (lane == 31) sWarp[32] = wi;
    }
    __syncthreads();
    const uint32_t res = inc - v + sWarp[warp];
    *total = sWarp[32];
    __syncthreads();
    return res;
}

// index c with off[c] <= x < off[c+1] (off ascending, off[0] <= x < off[n])
__device__ __forceinline__ uint32_t segment_of(const uint32_t* __restrict__ off, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;  // invariant: off[lo] <= x, x < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// the left walk of MBGC_Encoder.cpp:178-181: equal bytes going left from (dst, src), both exclusive
__device__ __forceinline__ uint32_t left_walk(const uint8_t* __restrict__ T, const uint8_t* __restrict__ R, uint32_t dst,
                                              uint64_t src) {
    uint32_t n = 0;
    while (dst > n && src > n && __ldg(T + dst - 1 - n) == __ldg(R + src - 1 - n)) n++;
    return n;
}

// ------------------------------------------------------------------------------------------------------
// pass 1: gap-breaking matches (MBGC_Encoder.cpp:153-199)
// ------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(PK_THREADS) k_pk_init(PackBatch b, uint32_t* __restrict__ mContig,
                                                        uint8_t* __restrict__ removed) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t j = blockIdx.x * PK_THREADS + threadIdx.x;
    if (j >= b.nMatches) return;
    mContig[j] = segment_of(b.matchOff, static_cast<uint32_t>(b.nContigs), j);
    removed[j] = 0;
}

// One sweep of: removed(j) = j+1 exists && a kept predecessor P exists && diag(j+1) == diag(P) != diag(j)
//                            && len'(j) < gapBreakingMatchMinLength
// with len'(j) = len(j) + leftWalk(j) when j-1 was removed and ends where j starts (:174-182).
__global__ void __launch_bounds__(PK_THREADS) k_pk_removal(PackParams p, PackState s, PackBatch b,
                                                           const uint32_t* __restrict__ mContig,
                                                           uint8_t* removed, uint32_t* __restrict__ changed) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t j = blockIdx.x * PK_THREADS + threadIdx.x;
    if (j >= b.nMatches) return;
    const uint32_t c = mContig[j];
    const uint32_t lo = b.matchOff[c], hi = b.matchOff[c + 1];
    uint8_t rem = 0;
    if (p.enableExt && j + 1 < hi && j > lo) {
        int64_t q = static_cast<int64_t>(j) - 1;
        while (q >= static_cast<int64_t>(lo) && removed[q]) q--;
        if (q >= static_cast<int64_t>(lo)) {
            const TextMatchRec mp = b.matches[q], mj = b.matches[j], mn = b.matches[j + 1];
            const uint64_t dP = mp.posSrcText - mp.posDestText;
            if (mn.posSrcText - mn.posDestText == dP && mj.posSrcText - mj.posDestText != dP) {
                uint64_t len = mj.length;
                if (removed[j - 1]) {
                    const TextMatchRec m1 = b.matches[j - 1];
                    if (m1.posDestText + m1.length == mj.posDestText)
                        len += left_walk(b.text + b.contigOff[c], s.ref, static_cast<uint32_t>(mj.posDestText), mj.posSrcText);
                }
                if (len < p.gapBreakingMinLen) rem = 1;
            }
        }
    }
    if (rem != removed[j]) {
        removed[j] = rem;
        *changed = 1;
    }
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_kept_tot(uint32_t nM, const uint8_t* __restrict__ removed,
                                                            uint32_t* __restrict__ blockTot) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sWarp[33];
    const uint32_t j = blockIdx.x * PK_THREADS + threadIdx.x;
    const uint32_t v = (j < nM && !removed[j]) ? 1u : 0u;
    uint32_t total;
    block_excl_scan_u32(v, &total, sWarp);
    if (threadIdx.x == 0) blockTot[blockIdx.x] = total;
}

// exclusive scan of per-block totals, COLS interleaved columns, one CTA of 1024 threads per column (4 values per
// thread and round); totals land at row nBlocks
template <int COLS>
__global__ void __launch_bounds__(1024) k_pk_partials(uint32_t* __restrict__ blockTot, uint32_t nBlocks) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sWarp[33];
    const int k = blockIdx.x;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nBlocks; base += 4096) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            v[q] = i0 + q < nBlocks ? blockTot[static_cast<uint64_t>(i0 + q) * COLS + k] : 0;
            sum += v[q];
        }
        uint32_t total;
        uint32_t ex = block_excl_scan_u32<1024>(sum, &total, sWarp) + carry;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (i0 + q < nBlocks) blockTot[static_cast<uint64_t>(i0 + q) * COLS + k] = ex;
            ex += v[q];
        }
        carry += total;
    }
    if (threadIdx.x == 0) blockTot[static_cast<uint64_t>(nBlocks) * COLS + k] = carry;
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_compact(PackState s, PackBatch b, PackWork w,
                                                           const uint32_t* __restrict__ blockOff) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sWarp[33];
    const uint32_t j = blockIdx.x * PK_THREADS + threadIdx.x;
    const bool in = j < b.nMatches;
    const bool keep = in && !w.removed[j];
    uint32_t total;
    const uint32_t idx = block_excl_scan_u32(keep ? 1u : 0u, &total, sWarp) + blockOff[blockIdx.x];
    uint32_t c = PK_NONE;
    uint64_t len = 0;
    if (in) {
        w.keptScan[j] = idx;
        if (j == b.nMatches - 1) w.keptScan[b.nMatches] = idx + (keep ? 1u : 0u);
    }
    if (keep) {
        c = w.mContig[j];
        const TextMatchRec m = b.matches[j];
        KeptRec r{m.posSrcText, static_cast<uint32_t>(m.posDestText), static_cast<uint32_t>(m.length)};
        if (j > b.matchOff[c] && w.removed[j - 1]) {  // shiftStartPos(-leftExtension), :182
            const TextMatchRec m1 = b.matches[j - 1];
            if (m1.posDestText + m1.length == m.posDestText) {
                const uint32_t L = left_walk(b.text + b.contigOff[c], s.ref, r.dst, r.src);
                r.src -= L;
                r.dst -= L;
                r.len += L;
            }
        }
        w.kept[idx] = r;
        w.kContig[idx] = c;
        w.kDiag[idx] = static_cast<uint32_t>(r.src) - r.dst;
        w.claim[idx] = 0;
        w.delta[idx] = 0;
        len = r.len;
    }
    // sum of kept lengths per text: one atomic per warp when the warp sits in one text
    const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, c, 0);
    if (__all_sync(0xFFFFFFFFu, c == c0)) {
#pragma unroll
        for (int d = 16; d; d >>= 1) len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
        if ((threadIdx.x & 31) == 0 && c0 != PK_NONE) atomicAdd(&w.sumLen[c0], static_cast<unsigned long long>(len));
    } else if (c != PK_NONE) {
        atomicAdd(&w.sumLen[c], static_cast<unsigned long long>(len));
    }
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_koff(PackBatch b, const uint32_t* __restrict__ keptScan,
                                                        uint32_t* __restrict__ kOff) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t c = blockIdx.x * PK_THREADS + threadIdx.x;
    if (c <= static_cast<uint32_t>(b.nContigs)) kOff[c] = keptScan[b.matchOff[c]];
}

// ------------------------------------------------------------------------------------------------------
// pairing (MBGC_Encoder.cpp:229-273)
// ------------------------------------------------------------------------------------------------------

// MBGC_Encoder.cpp:137-141
__device__ __forceinline__ uint64_t loaded_pos(const PackState& s, uint64_t pos) {
    if (s.nLoaded == 0) return pos;
    const uint64_t back = s.loadedArr[s.nLoaded - 1];
    const uint64_t step = s.refLength - REF_SHIFT;
    if (step == 0) return pos;
    while (pos + step < back) pos += step;
    return pos;
}

// :253-256: *std::upper_bound(refExtLoadedPosArr, getMatchLoadedPos(src)), UINT64_MAX at the end
__device__ __forceinline__ uint64_t next_region_loading_pos(const PackState& s, uint64_t src) {
    const uint64_t key = loaded_pos(s, src);
    uint32_t lo = 0, hi = s.nLoaded;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s.loadedArr[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo == s.nLoaded ? UINT64_MAX : s.loadedArr[lo];
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_claim_scatter(int nContigs, const uint32_t* __restrict__ kOff,
                                                                 const uint8_t* __restrict__ claim,
                                                                 uint32_t* __restrict__ claimedBy) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t m = blockIdx.x * PK_THREADS + threadIdx.x;
    if (m >= kOff[nContigs]) return;
    const uint32_t g = claim[m];
    if (g) atomicMin(&claimedBy[m + g], m);
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_claim_sweep(PackParams p, PackState s, PackBatch b, PackWork w,
                                                               uint32_t* __restrict__ changed) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t m = blockIdx.x * PK_THREADS + threadIdx.x;
    if (m >= w.kOff[b.nContigs]) return;
    const uint32_t c = w.kContig[m];
    const uint32_t base = w.kOff[c], n = w.kOff[c + 1] - base, j = m - base;
    const KeptRec me = w.kept[m];
    const uint64_t lock = b.contigLock ? b.contigLock[c] : NO_LOCK;
    const int gCnt = min(static_cast<int>(n - j - 1), p.gapDepthOffset);
    uint8_t newClaim = 0, newDelta = 0;
    bool haveNl = false;
    uint64_t myNl = 0;
    if (gCnt > 0) {
        // gapStartIdx != NO_GAP at this step <=> the previous match paired at distance 2 (see gap_state)
        const bool active = p.enableExt && p.gapDepthMismatches >= 2 && j > 0 && w.claim[m - 1] == 2;
        const uint32_t end = me.dst + me.len;
        const uint32_t literalsLeft = (j + 1 < n ? w.kept[m + 1].dst : b.contigLen[c]) - end;
        const bool special = !p.lazy && !active && literalsLeft == 0;  // :247
        const uint64_t diag = me.src - me.dst;
        const uint32_t diag32 = static_cast<uint32_t>(diag);
        const int side = me.src > lock ? 1 : (me.src < lock ? -1 : 0);  // TextMatchers.h:46-50
        // Only a candidate on the same diagonal can end the scan, so the scan itself reads 4 bytes per candidate;
        // the claimed / skipped ones before the hit (`reduce`, :247-250) are counted once the hit is known.
        if (side != 0)
            for (int g = 1; g <= gCnt; g++) {
                const uint32_t k = m + static_cast<uint32_t>(g);
                if (w.kDiag[k] != diag32) continue;
                if (w.claimedBy[k] < m || (special && g == 1)) continue;
                const KeptRec kr = w.kept[k];
                if (kr.src - kr.dst != diag) continue;
                if ((kr.src > lock ? 1 : (kr.src < lock ? -1 : 0)) != side) continue;
                if (p.lazy) {  // :252-261
                    if (!haveNl) {
                        const uint32_t cb = w.claimedBy[m];
                        myNl = cb != PK_NONE ? w.nl[cb] : next_region_loading_pos(s, me.src);
                        haveNl = true;
                    }
                    if (loaded_pos(s, kr.src) >= myNl) continue;
                }
                int reduce = 0;
                for (int q = 1; q < g; q++)
                    if (w.claimedBy[m + static_cast<uint32_t>(q)] < m || (special && q == 1)) reduce++;
                newClaim = static_cast<uint8_t>(g);
                newDelta = static_cast<uint8_t>(g - reduce);
                break;
            }
    }
    bool ch = false;
    if (haveNl && w.nl[m] != myNl) {
        w.nl[m] = myNl;
        ch = true;
    }
    if (w.claim[m] != newClaim) {
        w.claim[m] = newClaim;
        ch = true;
    }
    w.delta[m] = newDelta;
    if (ch) *changed = 1;
}

// ------------------------------------------------------------------------------------------------------
// extensions with mismatches (MBGC_Encoder.cpp:310-427)
// ------------------------------------------------------------------------------------------------------

// coders/ContextAwareMismatchesCoder.h:13-17: rows = actual (A,C,G,T,N), columns = mismatch
__constant__ uint8_t c_mis2code[25] = {'A', 2, 0, 1, 3, 1, 'C', 2, 0, 3, 0, 2, 'G', 1, 3, 1, 0, 2, 'T', 3, 1, 2, 3, 0, 'N'};

__device__ __forceinline__ int sym2val5(uint8_t ch) {
    switch (ch) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'N': return 4;
        default: return 255;  // every other symbol is > 4 for the test of .cpp:23
    }
}

// coders/ContextAwareMismatchesCoder.cpp:22-27
__device__ __forceinline__ uint8_t mismatch2code(uint8_t actual, uint8_t mismatch) {
    const int a = sym2val5(actual), m = sym2val5(mismatch);
    if (actual == mismatch || a > 4 || m > 4) return mismatch;
    return c_mis2code[a * 5 + m];
}

template <bool W>
struct Sink {
    uint8_t* lit;
    uint8_t* flg;
    uint32_t nLit = 0, nFlg = 0;
    __device__ __forceinline__ void putLit(uint8_t v) {
        if (W) lit[nLit] = v;
        nLit++;
    }
    __device__ __forceinline__ void putFlag(uint8_t v) {
        if (W) flg[nFlg] = v;
        nFlg++;
    }
    // `n` flags at once: byte i of mm has bit 7 set where position i mismatches
    __device__ __forceinline__ void putFlags(uint64_t mm, int n) {
        if (W)
            for (int i = 0; i < n; i++) flg[nFlg + i] = static_cast<uint8_t>((mm >> (8 * i + 7)) & 1);
        nFlg += n;
    }
};

__device__ __forceinline__ uint8_t ref_at(const PackState& s, int64_t i) {
    const int64_t hi = static_cast<int64_t>(s.maxRefLength + REF_READ_SLACK);
    return __ldg(s.ref + (i < hi ? i : hi));  // memory safety only: the reference never reads there in a valid run
}

// 8 bytes at an arbitrary address (little endian: byte i = p[i]) from three aligned words
__device__ __forceinline__ uint64_t load8(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
    const uint32_t sh = static_cast<uint32_t>(a & 3) * 8;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    return static_cast<uint64_t>(__funnelshift_r(w0, w1, sh)) | (static_cast<uint64_t>(__funnelshift_r(w1, w2, sh)) << 32);
}
__device__ __forceinline__ uint64_t load8_ref(const PackState& s, int64_t i) {
    if (i + 12 <= static_cast<int64_t>(s.maxRefLength + REF_READ_SLACK)) return load8(s.ref + i);
    uint64_t v = 0;
    for (int k = 0; k < 8; k++) v |= static_cast<uint64_t>(ref_at(s, i + k)) << (8 * k);
    return v;
}
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    return (static_cast<uint64_t>(__byte_perm(static_cast<uint32_t>(x), 0, 0x0123)) << 32) |
           __byte_perm(static_cast<uint32_t>(x >> 32), 0, 0x0123);
}
// bit 7 of every byte that is non-zero
__device__ __forceinline__ uint64_t nonzero_bytes(uint64_t x) {
    return (((x & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | x) & 0x8080808080808080ull;
}
__device__ __forceinline__ uint64_t low_bytes(uint64_t x, int n) { return n >= 8 ? x : x & ((1ull << (8 * n)) - 1); }

// The score automaton of both extensions over up to 8 positions held in registers (rw = reference bytes, tw = text
// bytes, byte i = i-th position in walking order, `nb` of them in range).  Emits a flag per position walked and a
// literal per mismatch; returns how many positions it walked.  `limited`: the walk also ends where the score has
// reached the threshold (checked before every position, as the loop conditions of :346 and :397 do).
template <bool W>
__device__ __forceinline__ int walk_word(const PackParams& p, uint64_t rw, uint64_t tw, int nb, bool limited, int& score,
                                         Sink<W>& out) {
    uint64_t mm = low_bytes(nonzero_bytes(rw ^ tw), nb);
    int done = nb;
    if (limited) {
        if (mm == 0) {  // all equal: the score only falls (floor 0), no stop inside
            score -= nb * p.matchBonus;
            if (score < 0) score = 0;
        } else {
            for (done = 0; done < nb; done++) {
                if (score >= p.scoreThreshold) break;
                if ((mm >> (8 * done + 7)) & 1)
                    score += p.mismatchPenalty;
                else {
                    score -= p.matchBonus;
                    if (score < 0) score = 0;
                }
            }
            mm = low_bytes(mm, done);
        }
    }
    out.putFlags(mm, done);
    while (mm) {
        const int i = (__ffsll(static_cast<long long>(mm)) - 1) >> 3;
        mm &= mm - 1;
        const uint8_t r = static_cast<uint8_t>(rw >> (8 * i)), t = static_cast<uint8_t>(tw >> (8 * i));
        out.putLit(p.withExclusion ? mismatch2code(r, t) : t);
    }
    return done;
}

// MBGC_Encoder.cpp:310-371.  T = text of the contig, gapPos = first byte after `match`, srcStart = reference
// position aligned with it (along the core match's diagonal inside a gap).  Returns the bytes consumed.
template <bool W>
__device__ uint32_t extend_right(const PackParams& p, const PackState& s, const uint8_t* __restrict__ T, uint32_t gapPos,
                                 int64_t srcStart, uint64_t length, bool isGap, bool gapStart, bool gapEnd, Sink<W>& out) {
    if (length == 0) return 0;  // (gapMiddle cannot occur with gapDepthMismatchesEncoding <= 2)
    int64_t srcPtr = srcStart;
    uint32_t gp = gapPos;
    const int64_t loadingPos = static_cast<int64_t>(s.loadingPos);
    const int64_t srcGuard = srcPtr + static_cast<int64_t>(length);
    int64_t validGuard = srcGuard;
    if (srcPtr == loadingPos) validGuard = srcPtr;
    if (!isGap) {
        const int64_t srcEnd = static_cast<int64_t>(s.maxRefLength);
        if (validGuard > srcEnd) validGuard = srcEnd;
        if (srcPtr <= loadingPos && loadingPos < validGuard) validGuard = loadingPos;
    }
    if (gapStart || !isGap) {
        const uint8_t r = ref_at(s, srcPtr);
        if (p.lazy && r == REF_REGION_SEPARATOR) validGuard = srcPtr;
        const uint8_t t = __ldg(T + gp);
        out.putLit(p.withExclusion && srcPtr < validGuard ? mismatch2code(r, t) : t);
        gp++;
    } else
        srcPtr--;
    int score = p.initialScore;
    // :346-361, eight positions per step.  The loop leaves srcPtr at the first position it did not walk: the guard, a
    // region separator (lazy mode) or, outside a gap, where the score reached the threshold.
    ++srcPtr;
    for (;;) {
        const int64_t avail = validGuard - srcPtr;
        if (avail <= 0 || (!isGap && score >= p.scoreThreshold)) break;
        int nb = avail < 8 ? static_cast<int>(avail) : 8;
        const uint64_t rw = load8_ref(s, srcPtr), tw = load8(T + gp);
        bool sep = false;
        if (p.lazy) {
            const uint64_t z = ~nonzero_bytes(rw) & 0x8080808080808080ull;
            if (z) {
                const int zi = (__ffsll(static_cast<long long>(z)) - 1) >> 3;
                if (zi < nb) {
                    nb = zi;
                    sep = true;
                }
            }
        }
        const int done = walk_word<W>(p, rw, tw, nb, !isGap, score, out);
        srcPtr += done;
        gp += static_cast<uint32_t>(done);
        if (done < nb || sep) break;
    }
    while (srcPtr++ < srcGuard && (isGap || score < p.scoreThreshold)) {
        out.putFlag(1);
        out.putLit(__ldg(T + gp));
        gp++;
        score += p.mismatchPenalty;
    }
    if ((isGap && !gapEnd) || (!isGap && score < p.scoreThreshold)) out.putFlag(1);
    return gp - gapPos;
}

// MBGC_Encoder.cpp:373-427.  length = literals between the cursor and the match.  Returns the bytes consumed.
template <bool W>
__device__ uint32_t extend_left(const PackParams& p, const PackState& s, const uint8_t* __restrict__ T, uint64_t length,
                                const KeptRec& match, uint64_t lockPos, Sink<W>& out) {
    const int64_t srcMatch = static_cast<int64_t>(match.src);
    int64_t srcGuard = 1;
    if (srcGuard < srcMatch - MAX_EXTEND_MATCH_LEFT_LENGTH) srcGuard = srcMatch - MAX_EXTEND_MATCH_LEFT_LENGTH;
    const int64_t srcLock = static_cast<int64_t>(lockPos);  // SIZE_MAX compares as -1 (ref + SIZE_MAX wraps below ref)
    if (srcGuard < srcLock && srcLock <= srcMatch) srcGuard = srcLock;
    bool guardKnown = true;
    if (srcGuard < srcMatch - static_cast<int64_t>(length)) {
        srcGuard = srcMatch - static_cast<int64_t>(length);
        guardKnown = false;
    }
    if (srcGuard == srcMatch) return 0;
    int64_t srcPtr = srcMatch - 1;
    int64_t gp = static_cast<int64_t>(match.dst) - 1;
    uint8_t r = ref_at(s, srcPtr);
    bool validSrc = !p.lazy || r != REF_REGION_SEPARATOR;
    {
        const uint8_t t = __ldg(T + gp);
        out.putLit(p.withExclusion && validSrc ? mismatch2code(r, t) : t);
    }
    int score = p.initialScore;
    // :397-415, eight positions per step while both buffers have eight bytes to the left (walking order = descending
    // addresses, hence the byte swap); the byte-wise loop below finishes near the buffer starts
    while (validSrc && srcPtr - srcGuard >= 1 && score < p.scoreThreshold && gp >= 8 && srcPtr >= 8) {
        const int64_t avail = srcPtr - srcGuard;
        int nb = avail < 8 ? static_cast<int>(avail) : 8;
        const uint64_t rw = bswap64(load8_ref(s, srcPtr - 8)), tw = bswap64(load8(T + gp - 8));
        bool sep = false;
        if (p.lazy) {
            const uint64_t z = ~nonzero_bytes(rw) & 0x8080808080808080ull;
            if (z) {
                const int zi = (__ffsll(static_cast<long long>(z)) - 1) >> 3;
                if (zi < nb) {
                    nb = zi;
                    sep = true;
                }
            }
        }
        const int done = walk_word<W>(p, rw, tw, nb, true, score, out);
        srcPtr -= done;
        gp -= done;
        if (done == nb && sep) validSrc = false;  // the next position is a separator: :399-404
        if (done < nb || sep) break;
    }
    while (validSrc && srcPtr > srcGuard && score < p.scoreThreshold) {
        --gp;
        --srcPtr;
        r = ref_at(s, srcPtr);
        const uint8_t t = __ldg(T + gp);
        const bool mismatch = t != r;
        if (p.lazy && r == REF_REGION_SEPARATOR) {
            validSrc = false;
            srcPtr++;
            gp++;
            break;
        }
        out.putFlag(mismatch ? 1 : 0);
        if (mismatch) {
            score += p.mismatchPenalty;
            out.putLit(p.withExclusion ? mismatch2code(r, t) : t);
        } else {
            score -= p.matchBonus;
            if (score < 0) score = 0;
        }
    }
    while (!validSrc && srcPtr > srcGuard && score < p.scoreThreshold) {
        srcPtr--;
        out.putFlag(1);
        --gp;
        out.putLit(__ldg(T + gp));
        score += p.mismatchPenalty;
    }
    if ((srcPtr != srcGuard || !guardKnown) && score < p.scoreThreshold) out.putFlag(1);
    return static_cast<uint32_t>(srcMatch - srcPtr);
}

// Gap state of step j (0-based inside its contig) after its pairing loop, for gapDepthMismatchesEncoding G <= 2.
// The reference keeps (gapStartIdx, gapEndIdx); a pair found at distance g <= G sets it to (j, j+g) -- the test
// `gapEndIdx <= j + g` of :263 always holds, because the only state that survives the reset of :219-222 into step j is
// (j-1, j+1), set by step j-1 pairing at distance 2.  Hence:
//   gapStart(j) = claim(j) in [1, G]
//   gapEnd(j)   = (gapEndIdx == j+1) = claim(j) == 1 (own gap of width 1), or no own gap and claim(j-1) == 2
//   gapMiddle   = never;   the core match (textMatches[gapStartIdx]) is j itself, or j-1 in the second case.
struct GapState {
    bool isGap, gapStart, gapEnd, coreIsPrev;
};
__device__ __forceinline__ GapState gap_state(const PackParams& p, const uint8_t* __restrict__ claim, uint32_t m, uint32_t j) {
    GapState g{false, false, false, false};
    if (!p.enableExt) return g;
    const int cl = claim[m];
    const bool own = cl >= 1 && cl <= p.gapDepthMismatches;
    const bool active = p.gapDepthMismatches >= 2 && j > 0 && claim[m - 1] == 2;
    g.gapStart = own;
    g.gapEnd = (own && cl == 1) || (!own && active);
    g.isGap = g.gapStart || g.gapEnd;
    g.coreIsPrev = !own && active;
    return g;
}

// utils/helper.cpp:237-246
__device__ __forceinline__ uint32_t frugal_len_bytes(uint64_t v) { return v < 0xFFFFu ? 2u : (v < 0xFFFFFFFFull ? 6u : 14u); }
__device__ __forceinline__ void put_le(uint8_t* p, uint64_t v, int n) {
    for (int i = 0; i < n; i++) p[i] = static_cast<uint8_t>(v >> (8 * i));
}

struct GapOut {            // where one gap writes (emit) / how much it writes (count)
    uint32_t n[PACK_STREAMS];
};

// Everything the reference appends from the end of match j-1 (or the contig start) up to and including the
// bookkeeping of match j (or the contig end for the tail gap j == n).  W = false counts, W = true writes at `at`.
// The plain literal run is returned for the caller to copy (plainSrc / plainN / plainDst).
template <bool W>
__device__ __forceinline__ void process_gap(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w,
                                            uint32_t c, uint32_t j, uint32_t n, const GapOut& at, GapOut& cnt,
                                            const uint8_t*& plainSrc, uint8_t*& plainDst, uint32_t& plainN) {
    const uint32_t base = w.kOff[c];
    const uint8_t* T = b.text + b.contigOff[c];
    const uint32_t destLen = b.contigLen[c];
    const uint64_t lock = b.contigLock ? b.contigLock[c] : NO_LOCK;
    Sink<W> out;
    out.lit = W ? w.out[PK_LITERALS] + at.n[PK_LITERALS] : nullptr;
    out.flg = W ? w.out[PK_GAPFLAGS] + at.n[PK_GAPFLAGS] : nullptr;
    KeptRec cur{0, 0, 0};
    if (j < n) cur = w.kept[base + j];
    uint32_t pos = 0;
    bool prevIsGap = false;
    if (j > 0) {
        const KeptRec prev = w.kept[base + j - 1];
        pos = prev.dst + prev.len;
        if (p.enableExt) {  // :275-285, the tail of step j-1
            const GapState g = gap_state(p, w.claim, base + j - 1, j - 1);
            int64_t srcStart = static_cast<int64_t>(prev.src + prev.len);
            if (g.coreIsPrev) {
                const KeptRec core = w.kept[base + j - 2];
                srcStart = static_cast<int64_t>(core.src + pos - core.dst);
            }
            const uint64_t length = (j < n ? cur.dst : destLen) - pos;
            pos += extend_right<W>(p, s, T, pos, srcStart, length, g.isGap, g.gapStart, g.gapEnd, out);
            prevIsGap = g.isGap;
        }
    }
    uint32_t literalsLeft = (j < n ? cur.dst : destLen) - pos;
    if (j < n && p.enableExt && !prevIsGap && literalsLeft)  // :216-218
        literalsLeft -= extend_left<W>(p, s, T, literalsLeft, cur, lock, out);
    plainSrc = T + pos;
    plainN = literalsLeft;
    plainDst = W ? out.lit + out.nLit : nullptr;
    out.nLit += literalsLeft;
    cnt.n[PK_LITERALS] = out.nLit;
    cnt.n[PK_GAPFLAGS] = out.nFlg;
    cnt.n[PK_MAPOFF] = cnt.n[PK_MAPOFF5] = cnt.n[PK_MAPLEN] = cnt.n[PK_GAPDELTA] = 0;
    if (j < n) {
        const uint32_t m = base + j;
        if (W) out.lit[out.nLit] = MATCH_MARK;  // :239
        cnt.n[PK_LITERALS] = out.nLit + 1;
        if (w.claimedBy[m] == PK_NONE) {  // !pairedGap[gapCurIdx], :227-233
            cnt.n[PK_MAPOFF] = 4;
            if (W) put_le(w.out[PK_MAPOFF] + at.n[PK_MAPOFF], static_cast<uint32_t>(cur.src), 4);
            if (p.enable40bit) {
                cnt.n[PK_MAPOFF5] = 1;
                if (W) w.out[PK_MAPOFF5][at.n[PK_MAPOFF5]] = static_cast<uint8_t>(cur.src >> 32);
            }
        }
        if (p.frugalLen) {  // :235-238
            const uint32_t nb = frugal_len_bytes(cur.len);
            cnt.n[PK_MAPLEN] = nb;
            if (W) {
                uint8_t* d = w.out[PK_MAPLEN] + at.n[PK_MAPLEN];
                put_le(d, cur.len < 0xFFFFu ? cur.len : 0xFFFFu, 2);
                if (nb > 2) put_le(d + 2, cur.len, 4);
                if (nb > 6) put_le(d + 6, cur.len, 8);
            }
        } else {
            cnt.n[PK_MAPLEN] = 4;
            if (W) put_le(w.out[PK_MAPLEN] + at.n[PK_MAPLEN], cur.len, 4);
        }
        if (n - j - 1 > 0 && p.gapDepthOffset > 0) {  // gCnt, :270-271
            cnt.n[PK_GAPDELTA] = 1;
            if (W) w.out[PK_GAPDELTA][at.n[PK_GAPDELTA]] = w.delta[m];
        }
    }
}

// gap gi of the batch -> (contig, index inside the contig, kept matches of the contig).  Gaps of contig c start at
// kOff[c] + c: one per kept match plus the tail gap.
__device__ __forceinline__ bool locate_gap(const PackBatch& b, const uint32_t* __restrict__ kOff, uint32_t gi, uint32_t& c,
                                           uint32_t& j, uint32_t& n) {
    const uint32_t nC = static_cast<uint32_t>(b.nContigs);
    if (gi >= kOff[nC] + nC) return false;
    uint32_t lo = 0, hi = nC;  // kOff[lo] + lo <= gi < kOff[hi] + hi
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (kOff[mid] + mid <= gi) lo = mid; else hi = mid;
    }
    c = lo;
    j = gi - (kOff[c] + c);
    n = kOff[c + 1] - kOff[c];
    return true;
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_count(PackParams p, PackState s, PackBatch b, PackWork w) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sTot[PACK_STREAMS];
    const uint32_t gi = blockIdx.x * PK_THREADS + threadIdx.x;
    GapOut cnt;
#pragma unroll
    for (int k = 0; k < PACK_STREAMS; k++) cnt.n[k] = 0;
    uint32_t c, j, n;
    if (locate_gap(b, w.kOff, gi, c, j, n)) {
        GapOut at{};
        const uint8_t* ps;
        uint8_t* pd;
        uint32_t pn;
        process_gap<false>(p, s, b, w, c, j, n, at, cnt, ps, pd, pn);
#pragma unroll
        for (int k = 0; k < PACK_STREAMS; k++) w.gapCnt[static_cast<uint64_t>(gi) * PACK_STREAMS + k] = cnt.n[k];
    }
    // per-block totals of the six columns: one hardware warp reduction each, then 8 warps x 6 shared atomics
    if (threadIdx.x < PACK_STREAMS) sTot[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PACK_STREAMS; k++) {
        const uint32_t t = __reduce_add_sync(0xFFFFFFFFu, cnt.n[k]);
        if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sTot[k], t);
    }
    __syncthreads();
    if (threadIdx.x < PACK_STREAMS) w.blockTot[static_cast<uint64_t>(blockIdx.x) * PACK_STREAMS + threadIdx.x] = sTot[threadIdx.x];
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_emit(PackParams p, PackState s, PackBatch b, PackWork w) {
    griddep_launch_dependents();
    griddep_wait();
    __shared__ uint32_t sWarpTot[PK_THREADS / 32][PACK_STREAMS];
    const uint32_t gi = blockIdx.x * PK_THREADS + threadIdx.x;
    uint32_t c = 0, j = 0, n = 0;
    const bool live = locate_gap(b, w.kOff, gi, c, j, n);
    GapOut at;
    {   // exclusive scan of the six columns over the block: warp scans, then the (8 x 6) warp totals
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < PACK_STREAMS; k++) {
            const uint32_t v = live ? w.gapCnt[static_cast<uint64_t>(gi) * PACK_STREAMS + k] : 0;
            const uint32_t inc = warp_incl_scan_u32(v, lane);
            at.n[k] = inc - v;
            if (lane == 31) sWarpTot[warp][k] = inc;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PACK_STREAMS; k++) {
            uint32_t before = w.blockTot[static_cast<uint64_t>(blockIdx.x) * PACK_STREAMS + k];
            for (int q = 0; q < warp; q++) before += sWarpTot[q][k];
            at.n[k] += before;
        }
    }
    const uint8_t* plainSrc = nullptr;
    uint8_t* plainDst = nullptr;
    uint32_t plainN = 0;
    if (live) {
        if (j == 0) {  // first gap of its text: the text's streams start here
#pragma unroll
            for (int k = 0; k < PACK_STREAMS; k++) w.streamOff[static_cast<uint64_t>(k) * (b.nContigs + 1) + c] = at.n[k];
        }
        GapOut cnt;
        process_gap<true>(p, s, b, w, c, j, n, at, cnt, plainSrc, plainDst, plainN);
        // the (pos, literalsLeft) processLiteral sees at :228 / :290
        if (w.litSpan) w.litSpan[gi] = make_uint2(static_cast<uint32_t>(plainSrc - (b.text + b.contigOff[c])), plainN);
        if (plainN <= 16)
            for (uint32_t i = 0; i < plainN; i++) plainDst[i] = __ldg(plainSrc + i);
    }
    // plain literal runs: the warp copies the medium ones together, the long ones go to k_pk_copy_long
    const int lane = threadIdx.x & 31;
    unsigned pending = __ballot_sync(0xFFFFFFFFu, plainN > 16);
    while (pending) {
        const int l = __ffs(pending) - 1;
        pending &= pending - 1;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(plainSrc), l));
        uint8_t* dst = reinterpret_cast<uint8_t*>(__shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(plainDst), l));
        const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, plainN, l);
        if (cnt >= PK_LONG_RUN) {
            if (lane == 0) {
                const uint32_t at2 = atomicAdd(&w.flags[32], 1u);
                if (at2 < w.longCap) w.longRuns[at2] = LongRun{src, dst, cnt};
            }
        } else {
            for (uint32_t i = lane; i < cnt; i += 32) dst[i] = __ldg(src + i);
        }
    }
}

// long plain-literal runs: every CTA takes 16 KiB pieces of every run round-robin; 4-byte stores at aligned
// destinations, sources read through two aligned words and a funnel shift
__global__ void __launch_bounds__(PK_THREADS) k_pk_copy_long(const LongRun* __restrict__ runs, const uint32_t* __restrict__ flags,
                                                             uint32_t cap) {
    griddep_launch_dependents();
    griddep_wait();
    constexpr uint64_t PIECE = 16384;
    uint32_t nRuns = flags[32];
    if (nRuns > cap) nRuns = cap;
    for (uint32_t r = 0; r < nRuns; r++) {
        const LongRun run = runs[r];
        const uint64_t head = (4 - (reinterpret_cast<uintptr_t>(run.dst) & 3)) & 3;  // bytes before the first aligned word
        const uint64_t body = (run.n - head) / 4;                                      // aligned words
        const uint64_t nPieces = (body * 4 + PIECE - 1) / PIECE;
        for (uint64_t piece = blockIdx.x; piece < nPieces; piece += gridDim.x) {
            const uint64_t w0 = piece * (PIECE / 4);
            const uint64_t w1 = w0 + PIECE / 4 < body ? w0 + PIECE / 4 : body;
            uint32_t* d = reinterpret_cast<uint32_t*>(run.dst + head);
            const uint8_t* sp = run.src + head;
            for (uint64_t i = w0 + threadIdx.x; i < w1; i += PK_THREADS) {
                const uintptr_t a = reinterpret_cast<uintptr_t>(sp + i * 4);
                const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~static_cast<uintptr_t>(3));
                d[i] = __funnelshift_r(__ldg(q), __ldg(q + 1), static_cast<uint32_t>(a & 3) * 8);
            }
        }
        if (blockIdx.x == r % gridDim.x) {  // head and tail bytes
            const uint64_t tail0 = head + body * 4;
            for (uint64_t i = threadIdx.x; i < head; i += PK_THREADS) run.dst[i] = __ldg(run.src + i);
            for (uint64_t i = tail0 + threadIdx.x; i < run.n; i += PK_THREADS) run.dst[i] = __ldg(run.src + i);
        }
    }
}

// MBGC_Encoder::processLiteral (mbgccoder/MBGC_Encoder.cpp:128-135): which literal runs extend the reference.  A run counts
// when its contig is NOT proper for a whole-contig extension (loadLiteralsToRef, :211; MGMP_Params.h:179-187 on the
// unmatched count before the extensions = length - sum of the kept match lengths) and it is longer than the threshold.
__global__ void __launch_bounds__(PK_THREADS) k_pk_lit_flag(LitExtParams le, PackBatch b, PackWork w) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t gi = blockIdx.x * PK_THREADS + threadIdx.x;
    if (gi >= b.nMatches + static_cast<uint32_t>(b.nContigs)) return;
    uint32_t c, j, n;
    if (!locate_gap(b, w.kOff, gi, c, j, n)) {  // behind the last gap (pass 1 removed matches): the scan runs over all nM + nC slots
        w.litFlag[gi] = 0;
        return;
    }
    const uint64_t destLen = b.contigLen[c];
    const uint64_t unmatched = destLen - w.sumLen[c];
    const bool proper = unmatched * le.factor > destLen;
    w.litFlag[gi] = !proper && static_cast<uint64_t>(w.litSpan[gi].y) > le.minLen ? 1u : 0u;
}

__global__ void __launch_bounds__(PK_THREADS) k_pk_lit_ranges(LitExtParams le, PackBatch b, PackWork w) {
    griddep_launch_dependents();
    griddep_wait();
    const uint32_t gi = blockIdx.x * PK_THREADS + threadIdx.x;
    uint32_t c, j, n;
    if (!locate_gap(b, w.kOff, gi, c, j, n)) return;
    const uint32_t k = w.litFlag[gi];
    if (w.litFlag[gi + 1] == k) return;
    const uint64_t destLen = b.contigLen[c];
    const uint64_t pos = w.litSpan[gi].x, length = w.litSpan[gi].y;
    const uint64_t before = le.beforeAfter < pos ? le.beforeAfter : pos;                                       // :130
    const uint64_t after = pos + length + le.beforeAfter < destLen ? le.beforeAfter : destLen - pos - length;  // :131-132
    w.litRanges[k] = LitRange{c, static_cast<uint32_t>(pos - before), static_cast<uint32_t>(length + before + after), 0};
}

inline uint32_t blocks_for(uint64_t n) { return static_cast<uint32_t>((n + PK_THREADS - 1) / PK_THREADS); }

}  // namespace

int launch_pack_pass1(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, int sweeps,
                      cudaStream_t st) {
    const uint32_t nM = b.nMatches;
    if (nM == 0) return 0;
    const uint32_t blocks = blocks_for(nM);
    int launches = 0;
    launch_pdl(k_pk_init, blocks, PK_THREADS, 0, st, b, w.mContig, w.removed);
    launches++;
    if (p.enableExt)
        for (int i = 0; i < sweeps; i++) {
            launch_pdl(k_pk_removal, blocks, PK_THREADS, 0, st, p, s, b, w.mContig, w.removed, w.flags + (i == sweeps - 1 ? 15 : i % 15));  // the last sweep has a flag of its own
            launches++;
        }
    launch_pdl(k_pk_kept_tot, blocks, PK_THREADS, 0, st, nM, w.removed, w.blockTot);
    launch_pdl(k_pk_partials<1>, 1, 1024, 0, st, w.blockTot, blocks);
    launch_pdl(k_pk_compact, blocks, PK_THREADS, 0, st, s, b, w, w.blockTot);
    launch_pdl(k_pk_koff, blocks_for(static_cast<uint64_t>(b.nContigs) + 1), PK_THREADS, 0, st, b, w.keptScan, w.kOff);
    return launches + 4;
}

int launch_pack_claims(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, int sweeps,
                       cudaStream_t st) {
    const uint32_t nM = b.nMatches;
    if (nM == 0) return 0;
    const uint32_t blocks = blocks_for(nM);
    int launches = 0;
    if (p.gapDepthOffset == 0) {  // no pairing at all: every match writes its offset
        cudaMemsetAsync(w.claimedBy, 0xFF, static_cast<uint64_t>(nM) * 4, st);
        return 0;
    }
    for (int i = 0; i < sweeps; i++) {
        cudaMemsetAsync(w.claimedBy, 0xFF, static_cast<uint64_t>(nM) * 4, st);
        if (i > 0) {
            launch_pdl(k_pk_claim_scatter, blocks, PK_THREADS, 0, st, b.nContigs, w.kOff, w.claim, w.claimedBy);
            launches++;
        }
        launch_pdl(k_pk_claim_sweep, blocks, PK_THREADS, 0, st, p, s, b, w, w.flags + 16 + (i == sweeps - 1 ? 15 : i % 15));
        launches++;
    }
    // the claims of the last sweep decide which matches omit their offset
    cudaMemsetAsync(w.claimedBy, 0xFF, static_cast<uint64_t>(nM) * 4, st);
    launch_pdl(k_pk_claim_scatter, blocks, PK_THREADS, 0, st, b.nContigs, w.kOff, w.claim, w.claimedBy);
    return launches + 1;
}

int launch_pack_count(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, cudaStream_t st) {
    const uint32_t blocks = blocks_for(static_cast<uint64_t>(b.nMatches) + b.nContigs);
    launch_pdl(k_pk_count, blocks, PK_THREADS, 0, st, p, s, b, w);
    launch_pdl(k_pk_partials<PACK_STREAMS>, PACK_STREAMS, 1024, 0, st, w.blockTot, blocks);
    return 2;
}

int launch_pack_literal_ext(const LitExtParams& le, const PackBatch& b, const PackWork& w, cudaStream_t st) {
    const uint32_t nGaps = b.nMatches + static_cast<uint32_t>(b.nContigs);
    const uint32_t blocks = blocks_for(nGaps);
    launch_pdl(k_pk_lit_flag, blocks, PK_THREADS, 0, st, le, b, w);
    const int l = launch_scan_u32(w.litFlag, w.litFlag, nGaps, w.litScanState, st);
    launch_pdl(k_pk_lit_ranges, blocks, PK_THREADS, 0, st, le, b, w);
    return 2 + l;
}

int launch_pack_emit(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, cudaStream_t st) {
    const uint32_t blocks = blocks_for(static_cast<uint64_t>(b.nMatches) + b.nContigs);
    launch_pdl(k_pk_emit, blocks, PK_THREADS, 0, st, p, s, b, w);
    launch_pdl(k_pk_copy_long, static_cast<unsigned>(device_sm_count()) * 4, PK_THREADS, 0, st, w.longRuns, w.flags, w.longCap);
    return 2;
}

}  // namespace mbgc_b200
