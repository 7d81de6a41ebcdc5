// mbgc_b200/csrc/matcher.cpp -- host logic of B200SparseEMMatcher (see matcher.hpp).
//
// Reference citations: ".cpp"/".h" = /root/reference/matching/SlidingWindowSparseEMMatcher.{cpp,h}.
#include "matcher.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>

#include <sys/mman.h>

#include "../../include/mbgc_b200.h"

namespace mbgc_b200 {

namespace {

constexpr uint64_t PAD = 64;            // padding before/after device buffers (unaligned loads may over-read)
constexpr uint64_t TEXT_GAP = 64;       // bytes between texts of a batch
constexpr uint64_t TEXT_TAIL = 8192;    // readable bytes after the last text (tile staging over-reads)
constexpr uint64_t REF_TAIL = 8192;     // same after the reference buffer (the self-probe reads it as a text)

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// utils/helper.cpp:312-358: the upper-casing complement table.  Identity, except the listed IUPAC letters
// in either case; the reference fills the identity only for CHAR_MIN..CHAR_MAX-1, so 0x7F maps to 0.
struct ComplementLut {
    uint8_t t[256];
    ComplementLut() {
        for (int i = 0; i < 256; i++) t[i] = static_cast<uint8_t>(i);
        t[0x7F] = 0;
        const char* from = "ACGTNUYRKMBDHVWS";
        const char* to = "TGCANARYMKVHDBSW";
        for (int i = 0; from[i]; i++) {
            t[static_cast<uint8_t>(from[i])] = static_cast<uint8_t>(to[i]);
            t[static_cast<uint8_t>(from[i] + 32)] = static_cast<uint8_t>(to[i]);
        }
    }
};
const ComplementLut g_lut;

// .cpp:75-87
int kmer_from_L(int L, uint32_t minMatchLength) {
    int K;
    if (L > 110) K = 56;
    else if (L > 62) K = 44;
    else if (L > 53) K = 40;
    else if (L > 46) K = 36;
    else if (L > 42) K = 32;
    else if (L > 32) K = 28;
    else K = (L / 4 - 1) * 4;
    const int KmmL = (static_cast<int>(minMatchLength) / 4 - 1) * 4;
    if (KmmL < K) K = KmmL;
    return K;
}

// .cpp:118-144
void calc_coprimes(int L, int K, int& k1, int& k2) {
    const int t = L - K + 1;
    if (t >= 20) {
        k1 = static_cast<int>(std::pow(t, 0.5)) + 1;
        k2 = k1 - 1;
        if (k1 * k2 > t) { --k2; --k1; }
    } else if (t >= 15) { k1 = 5; k2 = 3; }
    else if (t >= 12) { k1 = 4; k2 = 3; }
    else if (t >= 10) { k1 = 5; k2 = 2; }
    else if (t >= 6) { k1 = 3; k2 = 2; }
    else { k1 = t; k2 = 1; }
}

}  // namespace

void B200SparseEMMatcher::selftestReverseComplement(const uint8_t* src, uint64_t n, uint8_t* dst) {
    reverse_complement(src, n, dst, g_lut.t);
}

void B200SparseEMMatcher::check(cudaError_t e, const char* what) {
    if (e != cudaSuccess)
        throw MatcherError(MBGCG_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

B200SparseEMMatcher::B200SparseEMMatcher(uint64_t refLengthLimit, uint32_t targetMatchLength, int _k1, int _k2,
                                         int _skipMargin, int dev)
    : maxRefLength(refLengthLimit), L(static_cast<int>(targetMatchLength)), skipMargin(_skipMargin), device(dev) {
    const bool trace = std::getenv("MBGC_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace)
            fprintf(stderr, "[mbgc_b200]   create: %-28s at %9.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        throw MatcherError(MBGCG_E_CUDA, "no CUDA device: the B200 matcher has no CPU fallback");
    check(cudaSetDevice(device), "cudaSetDevice");
    check(cudaFree(nullptr), "context");
    lap("CUDA context");
    if (targetMatchLength < 16)  // MIN_MATCH_LENGTH, .h:23 / .cpp:82-85
        throw MatcherError(MBGCG_E_PARAM, "Error: Minimal matching length too short!");
    if (maxRefLength < 2) throw MatcherError(MBGCG_E_PARAM, "reference length limit too small");

    pos1 = static_cast<int64_t>(REF_SHIFT);
    swEnd = maxRefLength;
    setSlidingWindowSize(16);  // SW_WIDTH_FACTOR, .h:47
    K = kmer_from_L(L, targetMatchLength);  // minMatchLength defaults to UINT32_MAX, clamped to L (.cpp:342-344)
    calc_coprimes(L, K, k1, k2);
    if (_k1 != -1) k1 = _k1;
    if (_k2 != -1) k2 = _k2;
    if (k2 != 1) throw MatcherError(MBGCG_E_UNSUPPORTED, "k2 != 1 is outside the accelerated scope (MBGC always uses k2 = 1)");
    if (!probe_supports_K(K)) throw MatcherError(MBGCG_E_UNSUPPORTED, "hashed prefix length K not in the reference's hash matrix");
    uint8_t i = 24;  // HASH_SIZE_MIN_ORDER .. MAX_ORDER, .cpp:99-103
    do {
        hashSize = static_cast<uint32_t>(1) << (i++);
    } while (i <= 31 && hashSize < maxRefLength / static_cast<uint64_t>(k1));
    if (__builtin_ctz(static_cast<uint32_t>(k1))) {
        k1ord = __builtin_ctz(static_cast<uint32_t>(k1));
        samplingPos = static_cast<uint64_t>(k1);  // .cpp:503
        if ((maxRefLength >> k1ord) >= (1ull << 32))
            throw MatcherError(MBGCG_E_UNSUPPORTED, "reference needs the 64-bit table (ref .cpp:505-508)");
    } else {
        k1ord = -1;
        samplingPos = REF_SHIFT;
        if (maxRefLength >= (1ull << 32))
            throw MatcherError(MBGCG_E_UNSUPPORTED, "reference needs the 64-bit table (ref .cpp:347-350)");
    }

    if (const char* e = std::getenv("MBGC_B200_SEQ_LIMIT")) {  // tests: rebase the insert sequence numbers early
        const unsigned long long v = std::strtoull(e, nullptr, 10);
        if (v >= 1024 && v < (1ull << 32)) seqLimit = v;
    }
    upload_complement_table(g_lut.t);
    // anonymous mapping (zero pages, like calloc) with transparent huge pages asked for: the mirror is written once,
    // front to back, and with 4 KiB pages every loadRef call would spend most of its time in first-touch page faults
    hostRefBytes = align_up(maxRefLength + PAD, 2ull << 20);
    void* mem = mmap(nullptr, hostRefBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) throw MatcherError(MBGCG_E_NOMEM, "host mirror allocation failed");
    madvise(mem, hostRefBytes, MADV_HUGEPAGE);
    hostRef = static_cast<char*>(mem);
    lap("host mirror");
    check(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : ev) check(cudaEventCreate(&e), "cudaEventCreate");
    check(cudaMalloc(&dRefAlloc, maxRefLength + PAD + REF_TAIL), "cudaMalloc(ref)");
    check(cudaMemsetAsync(dRefAlloc, 0, maxRefLength + PAD + REF_TAIL, stream), "memset(ref)");
    dRef = dRefAlloc + PAD;
    check(cudaMalloc(&dHt, static_cast<uint64_t>(hashSize) * sizeof(unsigned long long)), "cudaMalloc(table)");
    check(cudaMemsetAsync(dHt, 0, static_cast<uint64_t>(hashSize) * sizeof(unsigned long long), stream), "memset(table)");
    check(cudaMalloc(&dOcc, hashSize / 8), "cudaMalloc(occupancy)");
    check(cudaMemsetAsync(dOcc, 0, hashSize / 8, stream), "memset(occupancy)");
    check(cudaMalloc(&dFilter, PROBE_FILTER_BYTES), "cudaMalloc(filter)");
    check(cudaMemsetAsync(dFilter, 0, PROBE_FILTER_BYTES, stream), "memset(filter)");
    check(cudaStreamSynchronize(stream), "init sync");
    lap("device buffers allocated + zeroed");
}

B200SparseEMMatcher::~B200SparseEMMatcher() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(dRefAlloc);
    cudaFree(dHt);
    cudaFree(dOcc);
    cudaFree(dFilter);
    cudaFree(dMemoX);
    cudaFree(dStage);
    if (hStage) cudaFreeHost(hStage);
    delete pool;
    if (hPacked) cudaFreeHost(hPacked);
    if (hXferMeta) cudaFreeHost(hXferMeta);
    cudaFree(dPacked);
    cudaFree(dXferMeta);
    cudaFree(dText);
    cudaFree(dMeta);
    cudaFree(dWork);
    if (hMeta) cudaFreeHost(hMeta);
    if (hMatches) cudaFreeHost(hMatches);
    if (hCounts) cudaFreeHost(hCounts);
    releasePack();
    releaseIngest();
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (hostRef) munmap(hostRef, hostRefBytes);
}

// ---- index build ---------------------------------------------------------------------------------------

// .cpp:146-171: the chunked (OpenMP) loop and the sequential tail are two independent integer ranges; the
// kernel gets them with program-order sequence numbers so that "last writer wins" is reproduced exactly.
void B200SparseEMMatcher::indexPass() {
    const int64_t chunk = static_cast<int64_t>(k1) * 128;
    int64_t i1 = static_cast<int64_t>(samplingPos);
    const int64_t chunkedBegin = i1;
    if (i1 < (pos1 - K) - chunk) i1 += ((pos1 - K) - chunk - i1 + chunk - 1) / chunk * chunk;
    InsertRange r0{chunkedBegin, (i1 - chunkedBegin) / k1, insertSeq};
    int64_t t = k1 + ((pos1 - K - 1) / chunk) * chunk;
    const int64_t tailBegin = t;
    if (t < pos1 - K + 1) t += ((pos1 - K + 1) - t + k1 - 1) / k1 * k1;
    InsertRange r1{tailBegin, (t - tailBegin) / k1, insertSeq + static_cast<uint64_t>(r0.count)};
    samplingPos = static_cast<uint64_t>(t);
    const uint64_t n = static_cast<uint64_t>(r0.count + r1.count);
    if (n == 0) return;
    if (insertSeq + n >= seqLimit) {
        // Epoch rebase: the sequence numbers only order samples that race for a slot, and every sample still to come must win
        // over everything in the table.  So the table forgets its numbers (each slot keeps its entry, sequence 0) and the
        // count restarts at 1 -- the table's low halves, i.e. the reference's ht32bit, do not change.
        if (n + 1 >= seqLimit) throw MatcherError(MBGCG_E_UNSUPPORTED, "one loadRef call with 2^32 index samples or more");
        launch_table_rebase(dHt, hashSize, stream);
        totalLaunches++;
        memoSeq = memoSeq == insertSeq ? 1 : 0;
        insertSeq = 1;
        r0.seqBase = insertSeq;
        r1.seqBase = insertSeq + static_cast<uint64_t>(r0.count);
        seqRebases++;
    }
    insertSeq += n;
    samplesInserted += n;
    passesSinceInsert = 0;
    launch_index_insert(dRef, dHt, dOcc, dFilter, PROBE_FILTER_BITS, hashSize - 1, K, k1, k1ord, r0, r1, stream);
    totalLaunches++;
    check(cudaGetLastError(), "index_insert launch");
}

// .cpp:402-437 (tail recursion written as a loop)
// The bytes of `text` are on the device at dSrc (staged once per loadRef call, or the resident text of an ingested
// contig): the device reference gets its bytes from there (copy / reverse-complement kernel) while the host cores write
// the mirror -- unless the mirror is switched off (text may then be null).
void B200SparseEMMatcher::loadSegment(const char* text, const uint8_t* dSrc, uint64_t len, bool rc, bool addSep, char sep) {
    while (len != 0) {
        if (static_cast<uint64_t>(pos1) == maxRefLength && swEnd != maxRefLength) {
            reachedRefLengthCount++;
            pos1 = static_cast<int64_t>(REF_SHIFT);
            samplingPos = REF_SHIFT;
        }
        const uint64_t tmpEnd = swEnd;
        uint64_t tmpLen = len;
        const uint64_t tmpMax = tmpEnd < static_cast<uint64_t>(pos1) ? maxRefLength : tmpEnd;
        if (static_cast<uint64_t>(pos1) + tmpLen > tmpMax) tmpLen = tmpMax - static_cast<uint64_t>(pos1);
        const bool mirror = mirrorEnabled && text != nullptr;
        if (mirror) {
            char* dst = hostRef + pos1;
            if (rc) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(text) + len - tmpLen;
                reverse_complement(src, tmpLen, reinterpret_cast<uint8_t*>(dst), g_lut.t);
            } else
                std::memcpy(dst, text, tmpLen);
        }
        const bool sepHit = addSep && static_cast<uint64_t>(pos1) + tmpLen == swEnd;
        if (sepHit && mirrorEnabled) hostRef[swEnd - 1] = sep;
        if (tmpLen) {
            launch_ref_segment(dSrc + (rc ? len - tmpLen : 0), dRef + pos1, tmpLen, rc, stream);
            totalLaunches++;
        }
        if (sepHit)  // the separator replaces the segment's last byte
            check(cudaMemsetAsync(dRef + swEnd - 1, static_cast<unsigned char>(sep), 1, stream), "memset(sep)");
        pos1 += static_cast<int64_t>(tmpLen);
        indexPass();
        if (!rc) {
            if (text) text += tmpLen;
            dSrc += tmpLen;
        }
        len = (static_cast<uint64_t>(pos1) == tmpEnd) ? 0 : len - tmpLen;
    }
}

void B200SparseEMMatcher::loadRef(const char* refText, uint64_t refLength, bool loadRCRef, bool addRegionSeparators,
                                  char regionSeparator) {
    check(cudaSetDevice(device), "cudaSetDevice");
    static const bool trace = std::getenv("MBGC_B200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (trace)
            fprintf(stderr, "[mbgc_b200]   loadRef: %-22s at %8.3f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    const uint64_t loaded0 = getLoadedRefLength();
    check(cudaEventRecord(ev[10], stream), "event");
    if (refLength) {  // the text goes up once, through pinned staging (a pageable source would be staged by the driver)
        growPinned(reinterpret_cast<void**>(&hStage), &hStageCap, refLength, "cudaHostAlloc(ref staging)");
        growDevice(reinterpret_cast<void**>(&dStage), &dStageCap, refLength + 64, "cudaMalloc(ref staging)");
        std::memcpy(hStage, refText, refLength);
        check(cudaMemcpyAsync(dStage, hStage, refLength, cudaMemcpyHostToDevice, stream), "H2D(ref text)");
    }
    lap("text staged");
    loadSegment(refText, dStage, refLength, false, addRegionSeparators, regionSeparator);
    lap("forward segment");
    if (loadRCRef) loadSegment(refText, dStage, refLength, true, addRegionSeparators, regionSeparator);
    lap("reverse complement");
    // the caller may free refText right after the call; pageable copies are staged before returning, but
    // keep the ordering explicit
    check(cudaEventRecord(ev[11], stream), "event");
    check(cudaStreamSynchronize(stream), "loadRef sync");
    lap("device done");
    check(cudaEventElapsedTime(&lastLoadRefMs[0], ev[10], ev[11]), "event time");
    lastLoadRefMs[1] = static_cast<float>(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    lastLoadRefBases = getLoadedRefLength() - loaded0;
}

// loadRef whose text is already in this device's memory (an ingested contig, or a peer copy of one): nothing crosses
// PCIe.  hostText, when given, feeds the host mirror; without it the mirror must be off (setHostMirror(false)).
void B200SparseEMMatcher::loadRefDevice(const void* dText, const char* hostText, uint64_t refLength, bool loadRCRef,
                                        bool addRegionSeparators, char regionSeparator) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (mirrorEnabled && !hostText && refLength)
        throw MatcherError(MBGCG_E_PARAM, "loadRef from device memory needs the host text too while the host mirror is on");
    const uint8_t* src = static_cast<const uint8_t*>(dText);
    const auto t0 = std::chrono::steady_clock::now();
    const uint64_t loaded0 = getLoadedRefLength();
    check(cudaEventRecord(ev[10], stream), "event");
    loadSegment(hostText, src, refLength, false, addRegionSeparators, regionSeparator);
    if (loadRCRef) loadSegment(hostText, src, refLength, true, addRegionSeparators, regionSeparator);
    check(cudaEventRecord(ev[11], stream), "event");
    check(cudaStreamSynchronize(stream), "loadRef sync");
    check(cudaEventElapsedTime(&lastLoadRefMs[0], ev[10], ev[11]), "event time");
    lastLoadRefMs[1] = static_cast<float>(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    lastLoadRefBases = getLoadedRefLength() - loaded0;
}

void* B200SparseEMMatcher::stageConcat(const StagePiece* pieces, size_t n, uint64_t* total) {
    check(cudaSetDevice(device), "cudaSetDevice");
    uint64_t sum = 0;
    for (size_t i = 0; i < n; i++) sum += pieces[i].len;
    growDevice(reinterpret_cast<void**>(&dStage), &dStageCap, sum + 64, "cudaMalloc(ref staging)");
    uint64_t at = 0;
    for (size_t i = 0; i < n; i++) {
        if (pieces[i].len) {
            launch_ref_segment(static_cast<const uint8_t*>(pieces[i].dText), dStage + at, pieces[i].len, pieces[i].rc, stream);
            totalLaunches++;
        }
        at += pieces[i].len;
    }
    check(cudaGetLastError(), "ref_segment launch");
    check(cudaStreamSynchronize(stream), "stageConcat sync");  // peers are about to read it through their own streams
    if (total) *total = sum;
    return dStage;
}

// staging buffer of at least n bytes on this device (the target of a peer copy before loadRefDevice)
void* B200SparseEMMatcher::refStaging(uint64_t n) {
    check(cudaSetDevice(device), "cudaSetDevice");
    growDevice(reinterpret_cast<void**>(&dStage), &dStageCap, n + 64, "cudaMalloc(ref staging)");
    return dStage;
}

void B200SparseEMMatcher::loadSeparator(char regionSeparator) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (static_cast<uint64_t>(pos1) == maxRefLength && swEnd != maxRefLength) {
        reachedRefLengthCount++;
        pos1 = static_cast<int64_t>(REF_SHIFT);
        samplingPos = REF_SHIFT;
    }
    if (static_cast<uint64_t>(pos1) == maxRefLength) return;
    uint64_t at;
    if (static_cast<uint64_t>(pos1) == swEnd)
        at = static_cast<uint64_t>(pos1) - 1;
    else
        at = static_cast<uint64_t>(pos1++);
    if (mirrorEnabled) hostRef[at] = regionSeparator;
    // (no sync: every later use of the reference is ordered behind this on the handle's stream)
    check(cudaMemsetAsync(dRef + at, static_cast<unsigned char>(regionSeparator), 1, stream), "memset(sep)");
}

// ---- locks: .cpp:361-400 -------------------------------------------------------------------------------

uint64_t B200SparseEMMatcher::acquireWorkerMatchingLockPos() {
    if (swSize == 0 || !circularBuffer) return swEnd;
    uint64_t w = static_cast<uint64_t>(pos1) + swSize;
    if (reachedRefLengthCount || w > maxRefLength) {
        if (w > maxRefLength) w -= maxRefLength - REF_SHIFT;
    } else
        w = maxRefLength;
    if (workersSwEndPositions.empty()) swEnd = w;
    workersSwEndPositions.push_back(w);
    return w;
}

void B200SparseEMMatcher::releaseWorkerMatchingLockPos(uint64_t lockValue) {
    if (swSize == 0 || !circularBuffer) return;
    size_t i = 0;
    while (i < workersSwEndPositions.size() && workersSwEndPositions[i] != lockValue) i++;
    if (i == workersSwEndPositions.size())
        throw MatcherError(MBGCG_E_LOCK, "ERROR: Invalid worker lock value (" + std::to_string(lockValue) + ")");
    if (i == 0) {
        do {
            workersSwEndPositions.pop_front();
        } while (!workersSwEndPositions.empty() && workersSwEndPositions.front() == SW_END_ERASED_FLAG);
        swEnd = workersSwEndPositions.empty() ? swEnd : workersSwEndPositions.front();
    } else
        workersSwEndPositions[i] = SW_END_ERASED_FLAG;
}

// ---- matching ------------------------------------------------------------------------------------------

void B200SparseEMMatcher::matchTexts(std::vector<TextMatch>& resMatches, const char* destText, uint64_t destLen,
                                     bool destIsRef, bool revComplMatching, uint32_t minMatchLength,
                                     uint64_t matchingLockPos) {
    if (minMatchLength < static_cast<uint32_t>(K))  // .cpp:480-483
        throw MatcherError(MBGCG_E_PARAM, "Minimal matching length cannot be smaller than K (" +
                                              std::to_string(minMatchLength) + " < " + std::to_string(K) + ")");
    if (destIsRef || revComplMatching)  // .cpp:185-188
        throw MatcherError(MBGCG_E_PARAM, "Source as destination and reverse-complement matching modes unsupported");
    std::vector<TextView> one{{destText, destLen}};
    matchBatch(one, &matchingLockPos, minMatchLength, false, nullptr, nullptr);
    resMatches.assign(hMatches, hMatches + nMatchesAll);
}

void B200SparseEMMatcher::ensureWork(uint64_t textBytes, uint32_t nTiles, int nContigs, uint32_t seedCap) {
    if (textBytes > dTextCap) {
        cudaFree(dText);
        dText = nullptr;
        dTextCap = 0;
        const uint64_t cap = align_up(textBytes + textBytes / 4, 1 << 20);
        check(cudaMalloc(&dText, cap), "cudaMalloc(text)");
        dTextCap = cap;
    }
    // meta: contigOff u64, contigLock u64, contigLen u32, tileStart u32 (+1)
    const uint64_t metaBytes = align_up(static_cast<uint64_t>(nContigs + 1) * 24 + 64, 256);
    if (metaBytes > dMetaCap) {
        cudaFree(dMeta);
        if (hMeta) cudaFreeHost(hMeta);
        dMeta = nullptr; hMeta = nullptr; dMetaCap = 0;
        const uint64_t cap = metaBytes * 2;
        check(cudaMalloc(&dMeta, cap), "cudaMalloc(meta)");
        check(cudaHostAlloc(&hMeta, cap, cudaHostAllocDefault), "cudaHostAlloc(meta)");
        dMetaCap = hMetaCap = cap;
    }
    const uint64_t countsBytes = align_up(static_cast<uint64_t>(nContigs + 1) * 4 + 64, 256);
    if (countsBytes > hCountsCap) {
        if (hCounts) cudaFreeHost(hCounts);
        hCounts = nullptr;
        check(cudaHostAlloc(&hCounts, countsBytes * 2, cudaHostAllocDefault), "cudaHostAlloc(counts)");
        hCountsCap = countsBytes * 2;
    }
    // work buffers, carved from one allocation
    uint64_t off = 0;
    auto carve = [&](uint64_t bytes) { const uint64_t o = off; off = align_up(off + bytes, 256); return o; };
    const uint64_t scratchN = static_cast<uint64_t>(nTiles) * PROBE_CHUNK;
    const uint64_t oTileCount = carve((static_cast<uint64_t>(nTiles) + 2) * 4);
    const uint64_t oChunkContig = carve((static_cast<uint64_t>(nTiles) + 2) * 4);
    const uint64_t oAnchor = carve((static_cast<uint64_t>(nTiles) + 2) * 8);
    const uint64_t oChunkSegN = carve(static_cast<uint64_t>(nTiles) + 64);
    const uint64_t oWorkList = carve((static_cast<uint64_t>(nTiles) + 2) * 4);
    const uint64_t oChunkSeg = carve((static_cast<uint64_t>(nTiles) + 2) * 8);
    const uint64_t oDirtyMask = carve((static_cast<uint64_t>(nTiles) + 2) * 4);
    const uint64_t oLaneList = carve((static_cast<uint64_t>(nTiles) + 2) * 32 * 4);
    const uint64_t oChunkDesc = carve((static_cast<uint64_t>(nTiles) + 2) * 16);
    const uint64_t oSeedMask = carve((static_cast<uint64_t>(nTiles) + 1) * 32 * 2);
    const uint64_t oScratchEnc = carve(scratchN * 4);
    const uint64_t oSeeds = carve((static_cast<uint64_t>(seedCap) + 2) * sizeof(SeedRec));
    const uint64_t oHeadScan = carve((static_cast<uint64_t>(seedCap) + 2) * 4);
    const uint64_t oRunFirst = carve((static_cast<uint64_t>(seedCap) + 2) * 4);
    const uint64_t oRuns = carve(static_cast<uint64_t>(seedCap) * sizeof(RunRec));
    const uint64_t oRunLink = carve(static_cast<uint64_t>(seedCap) + 16);
    const uint64_t oRunStart = carve((static_cast<uint64_t>(nContigs) + 2) * 4);
    const uint64_t oMatches = carve(static_cast<uint64_t>(seedCap) * sizeof(MatchRec));
    const uint64_t oIslandFwd = carve((static_cast<uint64_t>(seedCap) / 1024 + 4) * 8);
    const uint64_t oIslandBwd = carve((static_cast<uint64_t>(seedCap) / 1024 + 4) * 8);
    const uint64_t oMatchOff = carve((static_cast<uint64_t>(nContigs) + 2) * 4);
    const uint64_t oDense = carve(static_cast<uint64_t>(seedCap) * sizeof(TextMatchRec));
    const uint64_t nScan = std::max<uint64_t>(nTiles, seedCap);
    const uint64_t oBlockSums = carve((nScan / 4096 + 8) * 4);
    const uint64_t oCounters = carve(256);
    const uint64_t scanStride = nScan / 4096 + 8;
    const uint64_t oScanState = carve(3 * scanStride * 8);
    if (off > dWorkCap) {
        cudaFree(dWork);
        dWork = nullptr;
        dWorkCap = 0;
        const uint64_t cap = align_up(off + off / 8, 1 << 20);
        check(cudaMalloc(&dWork, cap), "cudaMalloc(work)");
        dWorkCap = cap;
    }
    wb.tileCount = reinterpret_cast<uint32_t*>(dWork + oTileCount);
    wb.chunkContig = reinterpret_cast<uint32_t*>(dWork + oChunkContig);
    wb.anchor = reinterpret_cast<unsigned long long*>(dWork + oAnchor);
    wb.chunkSegN = dWork + oChunkSegN;
    wb.workList = reinterpret_cast<uint32_t*>(dWork + oWorkList);
    wb.chunkSeg = reinterpret_cast<uint2*>(dWork + oChunkSeg);
    wb.dirtyMask = reinterpret_cast<uint32_t*>(dWork + oDirtyMask);
    wb.laneList = reinterpret_cast<uint32_t*>(dWork + oLaneList);
    wb.chunkDesc = reinterpret_cast<uint4*>(dWork + oChunkDesc);
    wb.seedMask = reinterpret_cast<uint16_t*>(dWork + oSeedMask);
    wb.scratchEnc = reinterpret_cast<uint32_t*>(dWork + oScratchEnc);
    wb.seeds = reinterpret_cast<SeedRec*>(dWork + oSeeds);
    wb.headScan = reinterpret_cast<uint32_t*>(dWork + oHeadScan);
    wb.runFirst = reinterpret_cast<uint32_t*>(dWork + oRunFirst);
    wb.runs = reinterpret_cast<RunRec*>(dWork + oRuns);
    wb.runLink = dWork + oRunLink;
    wb.runStart = reinterpret_cast<uint32_t*>(dWork + oRunStart);
    wb.matches = reinterpret_cast<MatchRec*>(dWork + oMatches);
    wb.islandFwd = reinterpret_cast<unsigned long long*>(dWork + oIslandFwd);
    wb.islandBwd = reinterpret_cast<unsigned long long*>(dWork + oIslandBwd);
    wb.matchOff = reinterpret_cast<uint32_t*>(dWork + oMatchOff);
    wb.matchesDense = reinterpret_cast<TextMatchRec*>(dWork + oDense);
    wb.blockSums = reinterpret_cast<uint32_t*>(dWork + oBlockSums);
    wb.counters = reinterpret_cast<uint32_t*>(dWork + oCounters);
    wb.scanState = reinterpret_cast<unsigned long long*>(dWork + oScanState);
    wb.scanStride = static_cast<uint32_t>(scanStride);
    wb.zeroBytes = static_cast<uint32_t>(oScanState + 3 * scanStride * 8 - oCounters);
    wb.seedCap = seedCap;
}

// stage times of the last pass from its events (lazily: ten cudaEventElapsedTime calls are not free)
void B200SparseEMMatcher::collectTiming() {
    if (!timingPending) return;
    timingPending = false;
    float ms;
    for (int s = 0; s < 6; s++) {
        check(cudaEventElapsedTime(&ms, ev[s], ev[s + 1]), "event time");
        stageMs[s] += ms;
    }
    check(cudaEventElapsedTime(&ms, ev[0], ev[6]), "event time");
    stageMs[MBGCG_T_TOTAL] += ms;
    check(cudaEventElapsedTime(&ms, ev[1], ev[7]), "event time");
    stageMs[MBGCG_T_ANCHOR] += ms;
    check(cudaEventElapsedTime(&ms, ev[7], ev[8]), "event time");
    stageMs[MBGCG_T_CLEAN] += ms;
    check(cudaEventElapsedTime(&ms, ev[8], ev[2]), "event time");
    stageMs[MBGCG_T_FULLPROBE] += ms;
}

void B200SparseEMMatcher::fillIndexState(IndexState& ix) const {
    ix.ref = dRef;
    ix.ht = dHt;
    ix.occ = dOcc;
    ix.filter = dFilter;
    ix.filterBits = PROBE_FILTER_BITS;
    ix.filterWords = PROBE_FILTER_BYTES / 4;
    ix.hashMask = hashSize - 1;
    ix.K = K;
    ix.k1 = k1;
    ix.k1ord = k1ord;
    ix.skipMargin = skipMargin;
    ix.pos1 = static_cast<uint64_t>(pos1);
    ix.refLen = getRefLength();
}

// Fast-path policy (kernels.cu, "Fast path for near-identical texts").  The memo costs one full probe of the
// loaded reference and pays back over the clean chunks of every pass until the index changes again, so it is
// (re)built when  loaded reference <= expected clean fraction x positions of this pass x passes the index has
// already stayed unchanged.  The clean fraction is measured by the fast path itself when it runs and estimated
// from the match lists (share of the text covered by matches >= 1024) when it does not.
bool B200SparseEMMatcher::wantMemo(uint64_t positions) {
    // MBGC_B200_FASTPATH = 0: never, 1: whenever it is applicable (tests), unset: the cost model below
    const char* env = std::getenv("MBGC_B200_FASTPATH");
    const int forced = env ? std::atoi(env) : -1;
    if (forced == 0 || k1 != 16 || k1ord != 4 || K < 16 || reachedRefLengthCount != 0) return false;
    if (pos1 <= static_cast<int64_t>(K) || static_cast<uint64_t>(pos1) >= (1ull << 32) - 65536) return false;
    if (forced == 1) return true;
    if (cleanEstimate < 0.6) return false;  // below that the anchor + compare passes cost more than they save
    if (memoSeq == insertSeq) return true;
    const double benefit = cleanEstimate * static_cast<double>(positions) *
                           static_cast<double>(std::max<uint64_t>(1, passesSinceInsert));
    return static_cast<double>(pos1) <= benefit;
}

// self-probe: the loaded reference [0, pos1) probed as one text against its own index, window test disabled
void B200SparseEMMatcher::buildMemo() {
    const uint64_t len = static_cast<uint64_t>(pos1);
    const uint64_t nValid = len - K + 1;
    const uint32_t tiles = static_cast<uint32_t>((nValid + PROBE_CHUNK - 1) / PROBE_CHUNK);
    const uint64_t words = static_cast<uint64_t>(tiles) + 8;
    if (words > memoCapWords) {
        cudaFree(dMemoX);
        dMemoX = nullptr;
        memoCapWords = 0;
        check(cudaMalloc(&dMemoX, (words + words / 4) * 8), "cudaMalloc(memo)");  // X words, then E words
        memoCapWords = words + words / 4;
    }
    uint32_t* dMemoE = dMemoX + memoCapWords;
    ensureWork(0, tiles, 1, 1024);
    uint64_t* hOff = reinterpret_cast<uint64_t*>(hMeta);
    uint64_t* hLock = hOff + 1;
    uint32_t* hLen = reinterpret_cast<uint32_t*>(hLock + 1);
    uint32_t* hTile = hLen + 1;
    hOff[0] = 0;
    hLock[0] = NO_LOCK;
    hLen[0] = static_cast<uint32_t>(len);
    hTile[0] = 0;
    hTile[1] = tiles;
    check(cudaEventRecord(ev[10], stream), "event");
    check(cudaMemcpyAsync(dMeta, hMeta, 28, cudaMemcpyHostToDevice, stream), "H2D(meta)");
    check(cudaMemsetAsync(dMemoX, 0xFF, memoCapWords * 8, stream), "memset(memo)");
    BatchDesc bd;
    bd.text = dRef;
    bd.contigOff = reinterpret_cast<const uint64_t*>(dMeta);
    bd.contigLock = bd.contigOff + 1;
    bd.contigLen = reinterpret_cast<const uint32_t*>(bd.contigLock + 1);
    bd.tileStart = bd.contigLen + 1;
    bd.nContigs = 1;
    bd.nTiles = tiles;
    IndexState ix;
    fillIndexState(ix);
    ix.pos1 = 1ull << 62;  // every candidate "ends before" the cursor: the window test always passes
    int launches = launch_chunk_desc(ix, bd, wb, stream);
    launches += launch_probe(ix, bd, wb, useSharedFilter(), false, stream);
    launches += launch_memo(bd, wb, static_cast<uint32_t>(nValid), dMemoX, dMemoE, stream);
    check(cudaEventRecord(ev[11], stream), "event");
    check(cudaGetLastError(), "memo launch");
    // the batch pass reuses hMeta: wait until the copy above has read it
    check(cudaStreamSynchronize(stream), "memo sync");
    float ms = 0;
    check(cudaEventElapsedTime(&ms, ev[10], ev[11]), "event time");
    stageMs[MBGCG_T_MEMO] += ms;
    stageMs[MBGCG_T_TOTAL] += ms;
    totalLaunches += static_cast<uint64_t>(launches);
    lastCounters[4] += static_cast<uint64_t>(launches);
    memoSeq = insertSeq;
    memoLen = len;
}

bool B200SparseEMMatcher::useSharedFilter() const {
    static const bool noFilterEnv = std::getenv("MBGC_B200_NO_FILTER") != nullptr;
    // the shared-memory filter only pays while it is sparse (entries <~ bits)
    return samplesInserted < static_cast<uint64_t>(PROBE_FILTER_BITS) * 3 / 2 && !noFilterEnv;
}

void B200SparseEMMatcher::growPinned(void** p, uint64_t* cap, uint64_t need, const char* what) {
    if (need <= *cap) return;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    const uint64_t c = align_up(need + need / 4, 1 << 20);
    check(cudaHostAlloc(p, c, cudaHostAllocDefault), what);
    *cap = c;
}

void B200SparseEMMatcher::growDevice(void** p, uint64_t* cap, uint64_t need, const char* what) {
    if (need <= *cap) return;
    cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const uint64_t c = align_up(need + need / 4, 1 << 20);
    check(cudaMalloc(p, c), what);
    *cap = c;
}

// Host texts -> dText.  Small batches: one DMA per text.  Large ones: blocks of 256 KiB, a share of them DMA'd raw
// from the caller's buffer (only when that buffer is pinned: a pageable source would be staged by the driver at a
// fraction of the speed) while the worker pool packs the others 4:1 into pinned staging memory; packed blocks are
// copied as they finish and expanded on the device.  See hostpack.hpp for why.
//   MBGC_B200_PACK          0 = never pack; unset / 1 = pack batches of MBGC_B200_PACK_MIN_BYTES (default 2 MiB) and more
//   MBGC_B200_RAW_PERMILLE  share of the bytes that travels raw while the CPU packs the rest (default: from the
//                           number of packer threads, 60 with 15 of them)
//   MBGC_B200_PACK_GBPS     effective packing rate per thread the default share is computed from (11)
//   MBGC_B200_PACK_THREADS  worker threads (default: hardware threads / LOCAL_WORLD_SIZE - 1)
int B200SparseEMMatcher::stageTexts(const std::vector<TextView>& texts, size_t first, int nC,
                                    const std::vector<uint64_t>& off) {
    constexpr uint64_t BLK = 256 << 10;
    uint64_t total = 0;
    for (int c = 0; c < nC; c++) total += texts[first + c].len;
    const char* envPack = std::getenv("MBGC_B200_PACK");
    uint64_t packMin = 2ull << 20;
    if (const char* e = std::getenv("MBGC_B200_PACK_MIN_BYTES")) packMin = std::strtoull(e, nullptr, 10);
    const bool pack = !(envPack && std::atoi(envPack) == 0) && total >= packMin;
    if (!pack) {
        for (int c = 0; c < nC; c++)
            if (texts[first + c].len)
                check(cudaMemcpyAsync(dText + off[c], texts[first + c].data, texts[first + c].len, cudaMemcpyHostToDevice,
                                      stream), "H2D(text)");
        lastWire[0] += total;
        return 0;
    }
    if (!pool) {
        int n = static_cast<int>(std::thread::hardware_concurrency());
        if (const char* lw = std::getenv("LOCAL_WORLD_SIZE")) n /= std::max(1, std::atoi(lw));
        n -= 1;  // the calling thread feeds the copy engine
        if (const char* e = std::getenv("MBGC_B200_PACK_THREADS")) n = std::atoi(e);
        pool = new WorkerPool(std::max(1, std::min(n, 64)));
    }
    // Super-blocks of 4 MiB either travel raw (one DMA each: small copies cost ~5 us apiece in launch overhead alone)
    // or are cut into 256 KiB blocks for the packers.
    constexpr uint64_t SUPER = 4ull << 20;
    struct Blk { const uint8_t* src; uint64_t dev; uint32_t n; uint64_t poff; bool raw; };
    std::vector<Blk> blks;      // blocks to pack
    std::vector<Blk> rawBlks;   // super-blocks that travel raw
    bool pinnedSrc = true;
    for (int c = 0; c < nC; c++) {
        const TextView& t = texts[first + c];
        if (!t.len) continue;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, t.data) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            pinnedSrc = false;
        }
    }
    // raw share r that balances the copy engine ((r + (1 - r) / 4) bytes at ~52 GB/s) against the packers ((1 - r)
    // bytes at p GB/s per thread): r = (c - 1/4) / (3/4 + c), c = 52 / (p threads).  With the software prefetch in the
    // packing loop a thread reads ~8.4 GB/s and 15 of them saturate the host's DRAM (~126 GB/s), so every raw byte the
    // copy engine fetches slows them down: the sweep on the B200 box (profiles/README.md r02c) has its optimum at r = 0 ..
    // 0.05, which the formula gives with an effective p = 11; 7 threads (2 ranks on 16 cores) -> 0.30, 3 threads -> 0.57
    double perThread = 11.0;
    if (const char* e = std::getenv("MBGC_B200_PACK_GBPS")) perThread = std::max(0.1, std::atof(e));
    const double cRatio = 52.0 / (perThread * pool->size());
    int rawPermille = static_cast<int>(1000.0 * std::max(0.0, std::min(1.0, (cRatio - 0.25) / (0.75 + cRatio))));
    if (const char* e = std::getenv("MBGC_B200_RAW_PERMILLE")) rawPermille = std::max(0, std::min(1000, std::atoi(e)));
    if (!pinnedSrc) rawPermille = 0;
    uint64_t seen = 0, rawBytes = 0, packedBytes = 0;
    for (int c = 0; c < nC; c++) {
        const TextView& t = texts[first + c];
        for (uint64_t so = 0; so < t.len; so += SUPER) {
            const uint64_t sn = std::min<uint64_t>(SUPER, t.len - so);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(t.data) + so;
            // keep the raw share of the bytes seen so far at rawPermille
            if ((rawBytes + sn) * 1000 <= (seen + sn) * static_cast<uint64_t>(rawPermille)) {
                rawBlks.push_back(Blk{src, off[c] + so, static_cast<uint32_t>(sn), 0, true});
                rawBytes += sn;
            } else {
                for (uint64_t o = 0; o < sn; o += BLK) {
                    const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(BLK, sn - o));
                    blks.push_back(Blk{src + o, off[c] + so + o, n, packedBytes, false});
                    packedBytes += align_up((n + 3) / 4, 16);
                }
            }
            seen += sn;
        }
    }
    const uint32_t nB = static_cast<uint32_t>(blks.size());
    growPinned(&hPacked, &hPackedCap, packedBytes + 64, "cudaHostAlloc(packed)");
    growDevice(&dPacked, &dPackedCap, packedBytes + 64, "cudaMalloc(packed)");
    if (blockRuns.size() < nB) blockRuns.resize(nB);
    std::vector<uint8_t> failed(nB, 0);
    uint8_t* hp = static_cast<uint8_t*>(hPacked);
    Blk* bp = blks.data();
    uint8_t* fp = failed.data();
    std::vector<std::vector<ExcRun>>* runsAll = &blockRuns;
    // the workers hold pointers into blks / failed: whatever throws below must not unwind past them while they run
    struct PoolGuard {
        WorkerPool* p;
        ~PoolGuard() { p->wait(); }
    } poolGuard{pool};
    pool->run(nB, [bp, hp, fp, runsAll](uint32_t b, int) {
        std::vector<ExcRun>& runs = (*runsAll)[b];
        runs.clear();
        if (!pack_block(bp[b].src, bp[b].n, hp + bp[b].poff, bp[b].dev, runs, bp[b].n / 64 + 16)) {
            fp[b] = 1;  // too many non-ACGT bytes: this block travels raw
            runs.clear();
        }
    });
    // feed the copy engine: packed blocks in order as they finish (merged into copies of up to 4 MiB), raw blocks
    // whenever no packed block is ready
    size_t ri = 0;
    uint32_t pi = 0;
    auto copyRaw = [&](const Blk& b) {
        check(cudaMemcpyAsync(dText + b.dev, b.src, b.n, cudaMemcpyHostToDevice, stream), "H2D(raw block)");
        lastWire[0] += b.n;
    };
    while (ri < rawBlks.size() || pi < nB) {
        if (pi < nB && pool->done(pi)) {
            uint64_t lo = UINT64_MAX, hi = 0;
            while (pi < nB && pool->done(pi) && (lo == UINT64_MAX || hi - lo < (4ull << 20))) {
                const uint32_t b = pi++;
                if (failed[b]) {
                    copyRaw(blks[b]);
                    continue;
                }
                lo = std::min(lo, blks[b].poff);
                hi = blks[b].poff + align_up((blks[b].n + 3) / 4, 16);
            }
            if (lo != UINT64_MAX) {
                check(cudaMemcpyAsync(static_cast<uint8_t*>(dPacked) + lo, hp + lo, hi - lo, cudaMemcpyHostToDevice, stream),
                      "H2D(packed blocks)");
                lastWire[0] += hi - lo;
            }
        } else if (ri < rawBlks.size()) {
            copyRaw(rawBlks[ri++]);
        } else {
            std::this_thread::sleep_for(std::chrono::microseconds(20));  // leave the core to a packer
        }
    }
    pool->wait();
    // descriptors of the packed blocks + exception runs -> device, then expand
    uint64_t nRuns = 0, nPacked = 0;
    for (uint32_t b = 0; b < nB; b++)
        if (!failed[b]) {
            nPacked++;
            nRuns += blockRuns[b].size();
        }
    const uint64_t metaBytes = nPacked * sizeof(PackedBlock) + nRuns * sizeof(ExcRunDev) + 64;
    growPinned(&hXferMeta, &hXferMetaCap, metaBytes, "cudaHostAlloc(transfer meta)");
    growDevice(&dXferMeta, &dXferMetaCap, metaBytes, "cudaMalloc(transfer meta)");
    PackedBlock* hb = static_cast<PackedBlock*>(hXferMeta);
    ExcRunDev* hr = reinterpret_cast<ExcRunDev*>(hb + nPacked);
    uint64_t kb = 0, kr = 0;
    for (uint32_t b = 0; b < nB; b++)
        if (!failed[b]) {
            hb[kb++] = PackedBlock{blks[b].dev, blks[b].poff, blks[b].n, 0};
            for (const ExcRun& e : blockRuns[b]) hr[kr++] = ExcRunDev{e.off, e.len, e.byte};
        }
    check(cudaMemcpyAsync(dXferMeta, hXferMeta, metaBytes - 64, cudaMemcpyHostToDevice, stream), "H2D(transfer meta)");
    lastWire[0] += metaBytes - 64;
    const PackedBlock* db = static_cast<const PackedBlock*>(dXferMeta);
    return launch_unpack(db, static_cast<uint32_t>(nPacked), static_cast<const uint8_t*>(dPacked),
                         reinterpret_cast<const ExcRunDev*>(db + nPacked), static_cast<uint32_t>(nRuns), dText, stream);
}

void B200SparseEMMatcher::matchBatch(const std::vector<TextView>& texts, const uint64_t* lockPos, uint32_t minMatchLength,
                                     bool textsOnDevice, const void* dTextBase, const uint64_t* dOffsets,
                                     bool listsStayOnDevice) {
    check(cudaSetDevice(device), "cudaSetDevice");
    // stage boundaries for mbgcg_last_timing; switched off (stageTiming) the stream carries nothing but the work itself
    auto mark = [&](int e) {
        if (stageTiming) check(cudaEventRecord(ev[e], stream), "event");
    };
    if (minMatchLength < static_cast<uint32_t>(K))
        throw MatcherError(MBGCG_E_PARAM, "Minimal matching length cannot be smaller than K (" +
                                              std::to_string(minMatchLength) + " < " + std::to_string(K) + ")");
    nMatchesAll = 0;
    resident.valid = false;
    resCounts.assign(texts.size(), 0);
    for (auto& m : stageMs) m = 0;
    timingPending = false;
    for (auto& c : lastCounters) c = 0;
    lastWire[0] = lastWire[1] = 0;
    if (texts.empty()) return;
    for (const auto& t : texts)
        if (t.len >= (1ull << 32)) throw MatcherError(MBGCG_E_UNSUPPORTED, "target text of 4 GiB or more");

    // split into passes of bounded size (work memory is ~26 bytes per target position)
    uint64_t maxPassBytes = 1024ull << 20;
    if (const char* e = std::getenv("MBGC_B200_MAX_PASS_BYTES")) maxPassBytes = std::max<uint64_t>(1, std::strtoull(e, nullptr, 10));
    if (listsStayOnDevice) {  // only one pass can leave its lists behind: refuse before doing any work
        uint64_t bytes = 0;
        for (size_t i = 0; i < texts.size(); i++) {
            if (i && bytes + texts[i].len > maxPassBytes)
                throw MatcherError(MBGCG_E_PARAM, "match lists can only stay on the device for a batch that fits one pass (" +
                                                      std::to_string(maxPassBytes) + " bytes of text, MBGC_B200_MAX_PASS_BYTES)");
            bytes += texts[i].len;
        }
    }
    size_t first = 0;
    while (first < texts.size()) {
        size_t last = first;
        uint64_t bytes = 0;
        while (last < texts.size() && (last == first || bytes + texts[last].len <= maxPassBytes)) bytes += texts[last++].len;
        const int nC = static_cast<int>(last - first);

        // layout + tiles
        std::vector<uint64_t> off(nC);
        std::vector<uint32_t> tileStart(nC + 1);
        uint64_t cursor = PAD, positions = 0;
        uint32_t tiles = 0;
        for (int c = 0; c < nC; c++) {
            const uint64_t len = texts[first + c].len;
            off[c] = textsOnDevice ? dOffsets[first + c] : cursor;
            cursor = align_up(cursor + len + TEXT_GAP, 16);
            tileStart[c] = tiles;
            const uint64_t nValid = len >= static_cast<uint64_t>(K) ? len - K + 1 : 0;
            tiles += static_cast<uint32_t>((nValid + PROBE_CHUNK - 1) / PROBE_CHUNK);
            positions += nValid;
        }
        tileStart[nC] = tiles;
        const uint64_t textBytes = cursor + TEXT_TAIL;

        bool useFast = wantMemo(positions);
        if (useFast && memoSeq != insertSeq) buildMemo();
        passesSinceInsert++;
        uint32_t seedCap = static_cast<uint32_t>(std::min<uint64_t>(positions / 4 + 65536, 0xFFFFFFF0ull));
        for (int attempt = 0;; attempt++) {
            ensureWork(textsOnDevice ? 0 : textBytes, tiles, nC, seedCap);
            // meta -> device
            uint64_t* hOff = reinterpret_cast<uint64_t*>(hMeta);
            uint64_t* hLock = hOff + nC;
            uint32_t* hLen = reinterpret_cast<uint32_t*>(hLock + nC);
            uint32_t* hTile = hLen + nC;
            for (int c = 0; c < nC; c++) {
                hOff[c] = off[c];
                hLock[c] = lockPos ? lockPos[first + c] : NO_LOCK;
                hLen[c] = static_cast<uint32_t>(texts[first + c].len);
                hTile[c] = tileStart[c];
            }
            hTile[nC] = tileStart[nC];
            const uint64_t metaUsed = static_cast<uint64_t>(nC) * 24 + 4;
            mark(0);
            check(cudaMemcpyAsync(dMeta, hMeta, metaUsed, cudaMemcpyHostToDevice, stream), "H2D(meta)");
            lastWire[0] += metaUsed;
            const uint8_t* textBase;
            if (textsOnDevice) {
                textBase = static_cast<const uint8_t*>(dTextBase);
            } else {
                const int xl = stageTexts(texts, first, nC, off);
                totalLaunches += static_cast<uint64_t>(xl);
                lastCounters[4] += static_cast<uint64_t>(xl);
                lastOffsets = off;
                lastLens.resize(nC);
                for (int c = 0; c < nC; c++) lastLens[c] = texts[first + c].len;
                textBase = dText;
            }
            check(cudaMemsetAsync(wb.counters, 0, wb.zeroBytes, stream), "memset(counters + scan state)");
            mark(1);

            BatchDesc bd;
            bd.text = textBase;
            bd.contigOff = reinterpret_cast<const uint64_t*>(dMeta);
            bd.contigLock = bd.contigOff + nC;
            bd.contigLen = reinterpret_cast<const uint32_t*>(bd.contigLock + nC);
            if (!lockPos) bd.contigLock = nullptr;  // no worker locks in this batch: kernels skip the lookup
            bd.tileStart = bd.contigLen + nC;
            bd.nContigs = nC;
            bd.nTiles = tiles;
            IndexState ix;
            fillIndexState(ix);

            int launches = 0;
            launches += launch_chunk_desc(ix, bd, wb, stream);
            if (useFast) {
                launches += launch_anchor(ix, bd, wb, stream);
                mark(7);
                launches += launch_clean(ix, bd, wb, MemoState{dMemoX, dMemoX + memoCapWords, memoLen}, stream);
                mark(8);
            } else {
                mark(7);
                mark(8);
            }
            launches += launch_probe(ix, bd, wb, useSharedFilter(), useFast, stream);
            mark(2);
            launches += launch_compact_and_runs(ix, bd, wb, useFast, stream);
            mark(3);
            launches += launch_extend(ix, bd, wb, stream);
            mark(4);
            launches += launch_resolve(ix, bd, wb, minMatchLength, stream);
            mark(5);
            check(cudaGetLastError(), "match pass launch");
            totalLaunches += static_cast<uint64_t>(launches);
            lastCounters[4] += static_cast<uint64_t>(launches);

            // counts first, then exactly the matches
            uint32_t* hOffs = hCounts;            // [nC+1]
            uint32_t* hCtr = hCounts + nC + 1;    // [8]
            check(cudaMemcpyAsync(hOffs, wb.matchOff, (static_cast<uint64_t>(nC) + 1) * 4, cudaMemcpyDeviceToHost, stream), "D2H(counts)");
            check(cudaMemcpyAsync(hCtr, wb.counters, 32, cudaMemcpyDeviceToHost, stream), "D2H(counters)");
            // the match list itself: its length is only known after a sync, so a guess (last pass + 1.5 %) travels with
            // the counts and a second copy fetches the rest if the guess was short -- one sync instead of two
            uint64_t guess = std::min<uint64_t>(matchGuess, hMatchesCap > nMatchesAll ? hMatchesCap - nMatchesAll : 0);
            guess = std::min<uint64_t>(guess, seedCap);
            // lists wanted on the device only (a packBatch follows): fetch just a prefix, for the fast-path policy below
            if (listsStayOnDevice) guess = std::min<uint64_t>(guess, 4096);
            if (guess)
                check(cudaMemcpyAsync(hMatches + nMatchesAll, wb.matchesDense, guess * sizeof(TextMatch),
                                      cudaMemcpyDeviceToHost, stream), "D2H(matches, guessed)");
            mark(6);
            check(cudaStreamSynchronize(stream), "match pass sync");
            if (hCtr[3]) {  // seed capacity exceeded: rerun this pass with the worst-case capacity
                if (attempt > 0) throw MatcherError(MBGCG_E_NOMEM, "seed buffer overflow after retry");
                seedCap = static_cast<uint32_t>(std::min<uint64_t>(positions + 65536, 0xFFFFFFF0ull));
                continue;
            }
            if (std::getenv("MBGC_B200_TRACE_ISLANDS"))
                fprintf(stderr, "[mbgc_b200] pass: %u runs, %u islands of more than 256 runs, longest %u runs\n", hCtr[1], hCtr[5], hCtr[4]);
            const uint32_t total = hOffs[nC];
            if (!listsStayOnDevice) matchGuess = static_cast<uint64_t>(total) + total / 64 + 1024;
            const uint32_t fetched = listsStayOnDevice ? static_cast<uint32_t>(std::min<uint64_t>(guess, total)) : total;
            if (nMatchesAll + total > hMatchesCap) {  // grow the pinned result buffer, keeping earlier passes
                const uint64_t cap = (nMatchesAll + total) + (nMatchesAll + total) / 2 + 4096;
                TextMatch* grown = nullptr;
                check(cudaHostAlloc(&grown, cap * sizeof(TextMatch), cudaHostAllocDefault), "cudaHostAlloc(matches)");
                if (nMatchesAll + guess) std::memcpy(grown, hMatches, (nMatchesAll + std::min<uint64_t>(guess, total)) * sizeof(TextMatch));
                if (hMatches) cudaFreeHost(hMatches);
                hMatches = grown;
                hMatchesCap = cap;
            }
            TextMatch* passMatches = hMatches + nMatchesAll;
            if (total > guess && !listsStayOnDevice) {  // the gather kernel already wrote the host layout (TextMatch): no conversion on the host
                check(cudaMemcpyAsync(passMatches + guess, wb.matchesDense + guess,
                                      (static_cast<uint64_t>(total) - guess) * sizeof(TextMatch), cudaMemcpyDeviceToHost,
                                      stream), "D2H(matches)");
                mark(6);
                check(cudaStreamSynchronize(stream), "D2H sync");
            }
            nMatchesAll += total;
            lastWire[1] += (static_cast<uint64_t>(nC) + 1) * 4 + 32 +
                           (listsStayOnDevice ? guess : std::max<uint64_t>(guess, total)) * sizeof(TextMatch);
            for (int c = 0; c < nC; c++) resCounts[first + c] = hOffs[c + 1] - hOffs[c];
            timingPending = stageTiming;  // the events of this pass are read when somebody asks (collectTiming)
            if (last < texts.size()) collectTiming();  // more passes follow and will reuse the events
            // clean-fraction estimate for the fast-path policy (see wantMemo)
            if (useFast && tiles) {  // share of the 16-position lanes that did not need the full probe
                cleanEstimate = 1.0 - static_cast<double>(hCtr[7]) / (32.0 * tiles);
            } else if (positions) {
                uint64_t covered = 0;
                const uint32_t stride = fetched / 4096 + 1;  // a sample is enough for the policy
                for (uint32_t k = 0; k < fetched; k += stride)
                    if (passMatches[k].length >= 1024) covered += passMatches[k].length;
                // (with the lists left on the device only a prefix was fetched: scale it to the whole list)
                const double scale = fetched ? static_cast<double>(total) / fetched : 0.0;
                cleanEstimate = static_cast<double>(covered) * stride * scale / static_cast<double>(positions);
            }
            lastCounters[6] += tiles;
            lastCounters[7] += useFast ? hCtr[7] : static_cast<uint64_t>(tiles) * 32;
            lastCounters[0] += positions;
            lastCounters[1] += hCtr[0];
            lastCounters[2] += hCtr[1];
            lastCounters[3] += total;
            lastCounters[5] += 1;  // passes
            if (first == 0 && last == texts.size()) {  // texts + lists stay on the device for a resident packBatch
                resident.valid = true;
                resident.bd = bd;
                resident.nMatches = total;
                resident.lens.resize(nC);
                for (int c = 0; c < nC; c++) resident.lens[c] = texts[c].len;
            }
            break;
        }
        first = last;
    }
}

void B200SparseEMMatcher::copyTableLow32(uint32_t* out, uint64_t nSlots) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (nSlots > hashSize) nSlots = hashSize;
    uint32_t* d = nullptr;
    check(cudaMalloc(&d, nSlots * 4), "cudaMalloc(table copy)");
    launch_table_low32(dHt, d, nSlots, stream);
    totalLaunches++;
    cudaError_t e = cudaMemcpyAsync(out, d, nSlots * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    check(e, "table copy");
}

void B200SparseEMMatcher::copyStagedText(uint64_t textIndex, char* out) {
    check(cudaSetDevice(device), "cudaSetDevice");
    if (textIndex >= lastOffsets.size()) throw MatcherError(MBGCG_E_PARAM, "no such text in the last host batch");
    if (lastLens[textIndex] == 0) return;
    check(cudaMemcpyAsync(out, dText + lastOffsets[textIndex], lastLens[textIndex], cudaMemcpyDeviceToHost, stream), "D2H(text)");
    check(cudaStreamSynchronize(stream), "text copy sync");
}

void B200SparseEMMatcher::copyDeviceRef(uint64_t from, uint64_t n, char* out) {
    check(cudaSetDevice(device), "cudaSetDevice");
    check(cudaMemcpyAsync(out, dRef + from, n, cudaMemcpyDeviceToHost, stream), "D2H(ref)");
    check(cudaStreamSynchronize(stream), "ref copy sync");
}

}  // namespace mbgc_b200
