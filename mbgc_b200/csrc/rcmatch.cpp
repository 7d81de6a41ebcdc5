// mbgc_b200/csrc/rcmatch.cpp -- host side of row f3: SimpleSequenceMatcher::rcMatchSequence
// (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176) with the index build and the probe / verify / extend work
// of CopMEMMatcher (matching/copmem/CopMEMMatcher.cpp) on the device (rcmatch.cu).  What stays here is what is
// sequential over MATCHES, not over bases: the replay of the query loop's skip / "inside the last match" decisions over
// the device's seed list, the collision fix-up, sort, overlap trimming and the rewrite of the sequence (:79-140, :150-163).
#include "rcmatch.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbgc_b200.h"
#include "matcher.hpp"

namespace mbgc_b200 {

namespace {

void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw MatcherError(MBGCG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct Dev {
    void* p = nullptr;
    explicit Dev(uint64_t n) { check(cudaMalloc(&p, n ? n : 1), "cudaMalloc(rc match)"); }
    ~Dev() { cudaFree(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

// PgHelpers::reverseComplement's table (utils/helper.cpp:312-358): case-preserving, identity elsewhere, 0x7F -> 0
void complement_lut(uint8_t lut[256]) {
    for (int i = 0; i < 256; i++) lut[i] = static_cast<uint8_t>(i);
    lut[0x7F] = 0;
    const char* from = "ACGTNYRKMBDHVWS";
    const char* to = "TGCANRYMKVHDBSW";
    for (int i = 0; from[i]; i++) {
        lut[static_cast<uint8_t>(from[i])] = static_cast<uint8_t>(to[i]);
        lut[static_cast<uint8_t>(from[i] + 32)] = static_cast<uint8_t>(to[i] + 32);
    }
    lut['U'] = 'U';
    lut['u'] = 'u';
}

// CopMEMMatcher::initParams / calcCoprimes (:63-93, :108-137)
void copmem_params(uint64_t N, uint32_t L, uint32_t minMatchLength, int& K, int& k1, int& k2, uint32_t& hashSize) {
    if (L > 110) K = 56;
    else if (L > 62) K = 44;
    else if (L > 53) K = 40;
    else if (L > 46) K = 36;
    else if (L > 42) K = 32;
    else if (L > 32) K = 28;
    else K = (static_cast<int>(L) / 4 - 1) * 4;
    if (minMatchLength < 24) throw MatcherError(MBGCG_E_PARAM, "Error: Minimal matching length too short!");
    const int kmm = (static_cast<int>(minMatchLength) / 4 - 1) * 4;
    if (kmm < K) K = kmm;
    const int t = static_cast<int>(L) - K + 1;
    if (t <= 0) throw MatcherError(MBGCG_E_PARAM, "L and K mismatch.");
    if (t >= 20) {
        k1 = static_cast<int>(std::pow(t, 0.5)) + 1;
        k2 = k1 - 1;
        if (k1 * k2 > t) { --k2; --k1; }
    } else if (t >= 15) { k1 = 5; k2 = 3; }
    else if (t >= 12) { k1 = 4; k2 = 3; }
    else if (t >= 10) { k1 = 5; k2 = 2; }
    else if (t >= 6) { k1 = 3; k2 = 2; }
    else { k1 = t; k2 = 1; }
    uint8_t i = 24;
    do {
        hashSize = static_cast<uint32_t>(1) << (i++);
    } while (i <= 31 && hashSize < N / static_cast<uint64_t>(k1));
}

struct Match { uint64_t src, len, dst; };

uint64_t frugal_bytes(uint8_t* dst, uint64_t at, uint64_t v) {  // writeUIntByteFrugal, utils/helper.cpp:217-225
    while (v >= 128) {
        dst[at++] = static_cast<uint8_t>(128 + v % 128);
        v /= 128;
    }
    dst[at++] = static_cast<uint8_t>(v);
    return at;
}

}  // namespace

uint64_t rc_match_sequence(B200SparseEMMatcher* m, uint8_t* seq, uint64_t N, uint32_t targetMatchLength, uint32_t minMatchLength,
                           uint8_t* mapOff, uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, float deviceMs[2],
                           uint64_t* launchesOut) {
    *mapOffLen = *mapLenLen = 0;
    if (deviceMs) deviceMs[0] = deviceMs[1] = 0;
    if (launchesOut) *launchesOut = 0;
    if (N < targetMatchLength) return N;  // no matcher: both maps stay empty (SimpleSequenceMatcher.cpp:17-19, 72-76)
    if (N >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "reverse-complement pass on a sequence of 4 GiB or more (64-bit CopMEM arrays)");
    uint32_t minLen = minMatchLength > targetMatchLength ? targetMatchLength : minMatchLength;  // CopMEMMatcher ctor :489-490
    int K, k1, k2;
    uint32_t hashSize;
    copmem_params(N, targetMatchLength, minLen, K, k1, k2, hashSize);
    if (minMatchLength == UINT32_MAX) minMatchLength = targetMatchLength;                        // :83-84
    if (minMatchLength < static_cast<uint32_t>(K))
        throw MatcherError(MBGCG_E_PARAM, "Minimal matching length cannot be smaller than K");   // CopMEMMatcher::matchTexts :513-516
    const uint32_t mask = hashSize - 1;

    check(cudaSetDevice(m->deviceIndex()), "cudaSetDevice");
    cudaStream_t st = static_cast<cudaStream_t>(m->streamHandle());
    uint8_t lut[256];
    complement_lut(lut);
    rc_upload_lut(lut);
    const uint64_t nSamples = (N - K) / k1 + 1;
    const uint64_t PADB = 64;
    Dev dSeq(N + 2 * PADB), dQ(N + 2 * PADB), dBucket(nSamples * 4), dCount((static_cast<uint64_t>(hashSize) + 2) * 4),
        dStart((static_cast<uint64_t>(hashSize) + 2) * 4), dEntries((nSamples + RC_MAX_PER_BUCKET + 1) * 4),
        dState((hashSize / 4096 + 16) * 8), dBig((1u << 20) * 4 + 8);
    uint32_t seedCap = static_cast<uint32_t>(std::min<uint64_t>(N / 8 + (1u << 16), 1u << 28));
    cudaEvent_t ev[3];
    for (auto& e : ev) check(cudaEventCreate(&e), "cudaEventCreate");
    std::vector<RcSeed> seeds;
    int launches = 0;
    try {
        check(cudaMemsetAsync(dSeq.p, 0, N + 2 * PADB, st), "memset");
        check(cudaMemsetAsync(dQ.p, 0, N + 2 * PADB, st), "memset");
        uint8_t* s = dSeq.as<uint8_t>() + PADB;
        uint8_t* q = dQ.as<uint8_t>() + PADB;
        check(cudaMemcpyAsync(s, seq, N, cudaMemcpyHostToDevice, st), "H2D(sequence)");
        check(cudaEventRecord(ev[0], st), "event");
        launches += launch_rc_revcomp(s, N, q, st);
        // ---- index: count, scan, fill, sort every bucket, keep its 13 smallest ----
        check(cudaMemsetAsync(dCount.p, 0, (static_cast<uint64_t>(hashSize) + 2) * 4, st), "memset(count)");
        check(cudaMemsetAsync(dState.p, 0, (hashSize / 4096 + 16) * 8, st), "memset(scan state)");
        check(cudaMemsetAsync(dBig.p, 0, 8, st), "memset(big bucket count)");
        launches += launch_cm_count(s, nSamples, K, k1, mask, dBucket.as<uint32_t>(), dCount.as<uint32_t>(), st);
        launches += launch_scan_u32(dCount.as<uint32_t>(), dStart.as<uint32_t>(), hashSize, dState.as<unsigned long long>(), st);
        check(cudaMemsetAsync(dCount.p, 0, (static_cast<uint64_t>(hashSize) + 2) * 4, st), "memset(cursor)");
        launches += launch_cm_fill(nSamples, k1, dBucket.as<uint32_t>(), dStart.as<uint32_t>(), dCount.as<uint32_t>(), dEntries.as<uint32_t>(), st);
        uint32_t* bigCount = dBig.as<uint32_t>();
        launches += launch_cm_sort(hashSize, dStart.as<uint32_t>(), dEntries.as<uint32_t>(), bigCount + 2, bigCount, 1u << 20, st);
        check(cudaEventRecord(ev[1], st), "event");
        // ---- query: the real seeds of every k2-th position of the reverse complement ----
        for (int attempt = 0;; attempt++) {
            Dev dSeeds(static_cast<uint64_t>(seedCap) * sizeof(RcSeed)), dN(8);
            check(cudaMemsetAsync(dN.p, 0, 8, st), "memset");
            launches += launch_cm_query(s, q, N, K, k2, mask, dStart.as<uint32_t>(), dEntries.as<uint32_t>(), dSeeds.as<RcSeed>(),
                                        dN.as<uint32_t>(), seedCap, st);
            check(cudaGetLastError(), "rc match launch");
            uint32_t n = 0, big = 0;
            check(cudaMemcpyAsync(&n, dN.p, 4, cudaMemcpyDeviceToHost, st), "D2H(seed count)");
            check(cudaMemcpyAsync(&big, bigCount, 4, cudaMemcpyDeviceToHost, st), "D2H(big buckets)");
            check(cudaEventRecord(ev[2], st), "event");
            check(cudaStreamSynchronize(st), "rc match sync");
            if (big > (1u << 20)) throw MatcherError(MBGCG_E_UNSUPPORTED, "more than 2^20 over-full hash buckets in the reverse-complement pass");
            if (n > seedCap) {  // rerun with room for all of them
                if (attempt) throw MatcherError(MBGCG_E_NOMEM, "seed list overflow after retry");
                seedCap = n + 1024;
                continue;
            }
            seeds.resize(n);
            if (n) {
                check(cudaMemcpyAsync(seeds.data(), dSeeds.p, static_cast<uint64_t>(n) * sizeof(RcSeed), cudaMemcpyDeviceToHost, st), "D2H(seeds)");
                check(cudaStreamSynchronize(st), "rc match sync");
            }
            break;
        }
        if (deviceMs) {
            check(cudaEventElapsedTime(&deviceMs[0], ev[0], ev[1]), "event time");
            check(cudaEventElapsedTime(&deviceMs[1], ev[1], ev[2]), "event time");
        }
    } catch (...) {
        for (auto& e : ev) cudaEventDestroy(e);
        throw;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    m->totalLaunches += static_cast<uint64_t>(launches);
    if (launchesOut) *launchesOut = static_cast<uint64_t>(launches);

    // ---- the query loop's sequential part, replayed over the seeds (:349-472) ----
    std::sort(seeds.begin(), seeds.end(), [](const RcSeed& a, const RcSeed& b) { return a.pos != b.pos ? a.pos < b.pos : a.order < b.order; });
    std::vector<uint8_t> qHost;  // the reverse complement on the host: only for matches whose ends the device left open
    auto needQ = [&] {
        if (qHost.empty()) {
            qHost.resize(N);
            for (uint64_t i = 0; i < N; i++) qHost[N - 1 - i] = lut[seq[i]];
        }
    };
    const uint64_t MULTI = 256, k2MULTI = static_cast<uint64_t>(k2) * MULTI;
    uint64_t nMainChunks = 0;  // chunks of 256 query positions the reference's main loop covers; the rest is its tail loop
    for (uint64_t i1 = 0; i1 + K + k2MULTI < N + 1; i1 += k2MULTI) nMainChunks++;
    const uint64_t tailFirstIdx = nMainChunks * MULTI;
    const uint64_t skip = static_cast<uint64_t>(K / k1 - 1);
    std::vector<Match> res;
    uint64_t skipUntil = 0;  // query indices below this are jumped over
    size_t i = 0;
    while (i < seeds.size()) {
        const uint64_t pos = seeds[i].pos, idx = pos / k2;
        size_t e = i;
        while (e < seeds.size() && seeds[e].pos == pos) e++;
        if (idx >= skipUntil) {
            for (size_t k = i; k < e; k++) {
                const RcSeed& sd = seeds[k];
                bool hit = false;
                if (!res.empty()) {  // a seed inside the match pushed last, on its diagonal (:380-386)
                    const Match& b = res.back();
                    if (pos - sd.src == b.dst - b.src && pos + K < b.dst + b.len) hit = true;
                }
                if (!hit) {
                    uint64_t left = sd.left, right = sd.right;
                    if (sd.flags & 3) {
                        needQ();
                        if (sd.flags & 1) {
                            const uint64_t lim = std::min<uint64_t>(sd.src, pos) - 1;
                            while (left < lim && seq[sd.src - left - 1] == qHost[pos - left - 1]) left++;
                        }
                        if (sd.flags & 2) {
                            const uint64_t lim = std::min<uint64_t>(N - sd.src, N - pos) - K;
                            while (right < lim && seq[sd.src + K + right] == qHost[pos + K + right]) right++;
                        }
                    }
                    Match mt;
                    if (sd.flags & 4) {  // the seed sits on byte 0 of a text: the left loop does not run, the match starts behind it
                        mt.src = static_cast<uint64_t>(sd.src) + 1;
                        mt.dst = pos + 1;
                        mt.len = K + right - 1;
                    } else {
                        mt.src = sd.src - left;
                        mt.dst = pos - left;
                        mt.len = left + K + right;
                    }
                    if (mt.len >= minMatchLength) {  // right - p1 > minMatchLength, :405
                        res.push_back(mt);
                        hit = true;
                    }
                }
                if (hit) {
                    skipUntil = idx + skip + 1;
                    if (idx < tailFirstIdx) {  // the main loop's skip ends with its chunk of 256 positions
                        const uint64_t chunkEnd = (idx / MULTI + 1) * MULTI;
                        if (skipUntil > chunkEnd) skipUntil = chunkEnd;
                    }
                    break;
                }
            }
        }
        i = e;
    }

    // ---- markAndRemoveExactMatches behind exactMatchSequence (:58-61, :150-163, :88-140) ----
    for (Match& t : res) {
        t.dst = N - (t.dst + t.len);
        if (t.src > t.dst) std::swap(t.src, t.dst);
        if (t.src + t.len > t.dst) {
            const uint64_t margin = (t.src + t.len - t.dst + 1) / 2;
            t.len -= margin;
            t.dst += margin;
        }
    }
    uint64_t lo = frugal_bytes(mapLen, 0, minMatchLength);
    std::sort(res.begin(), res.end(), [](const Match& a, const Match& b) {
        if (a.dst != b.dst) return a.dst < b.dst;
        if (a.src != b.src) return a.src < b.src;
        return a.len < b.len;
    });
    res.erase(std::unique(res.begin(), res.end(), [](const Match& a, const Match& b) { return a.src == b.src && a.len == b.len && a.dst == b.dst; }),
              res.end());
    uint64_t pos = 0, nPos = 0, oo = 0;
    for (Match& t : res) {
        if (t.dst < pos) {
            const uint64_t overflow = pos - t.dst;
            if (overflow >= t.len) {
                t.len = 0;
                continue;
            }
            t.len -= overflow;
            t.dst += overflow;
        }
        if (t.len < minMatchLength) continue;
        const uint64_t length = t.dst - pos;
        std::memmove(seq + nPos, seq + pos, length);
        nPos += length;
        seq[nPos++] = static_cast<uint8_t>('$' + 128);  // RC_MATCH_MARK, mbgccoder/MBGC_Params.h:48
        const uint32_t src32 = static_cast<uint32_t>(t.src);
        std::memcpy(mapOff + oo, &src32, 4);
        oo += 4;
        lo = frugal_bytes(mapLen, lo, t.len - minMatchLength);
        pos = t.dst + t.len;
    }
    std::memmove(seq + nPos, seq + pos, N - pos);
    nPos += N - pos;
    *mapOffLen = oo;
    *mapLenLen = lo;
    return nPos;
}

}  // namespace mbgc_b200
