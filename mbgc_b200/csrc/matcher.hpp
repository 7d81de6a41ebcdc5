// mbgc_b200/csrc/matcher.hpp -- host side of the B200 match finder.
//
// B200SparseEMMatcher mirrors the public surface of the reference class SlidingWindowSparseEMMatcher
// (/root/reference/matching/SlidingWindowSparseEMMatcher.h:86-126; same method names, argument meaning and
// error behaviour -- errors are thrown as MatcherError instead of exit(EXIT_FAILURE)).  The bookkeeping
// the reference does on the host (circular buffer, sliding window, worker locks, which positions an index
// pass samples) stays on the host and is restated here; all per-base work runs in the CUDA kernels of
// kernels.cu.  There is no CPU implementation of the matching itself in this class.
#pragma once
#include <cstdint>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "hostpack.hpp"
#include "kernels.h"
#include "pack.h"
#include "ingest.h"

constexpr int MBGCG_T_COUNT_ = 12;  // = MBGCG_T_COUNT of include/mbgc_b200.h (checked in capi.cpp)

namespace mbgc_b200 {

struct MatcherError : public std::runtime_error {
    int code;
    MatcherError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

struct TextMatch {  // = PgTools::TextMatch, matching/TextMatchers.h:9-17
    uint64_t posSrcText, length, posDestText, nextSrcRegionLoadingPos;
};

struct FastaRecord {  // one record of an ingested FASTA file (= mbgcg_fasta_rec_t)
    uint64_t hdrOff, hdrLen;   // header line in the file image, without '>' and without the line end
    uint64_t textOff, seqLen;  // the sequence in the device text buffer of the ingest
};

struct TextView {   // one target text (contig / block) of a batch
    const char* data;
    uint64_t len;
};

class B200SparseEMMatcher {
public:
    static constexpr uint64_t REF_SHIFT = 1;                     // ref .h:14
    static constexpr uint64_t SW_END_ERASED_FLAG = UINT64_MAX;   // ref .h:46

    // ref .cpp:325-359 / 494-519; Exp variant iff ctz(k1) != 0 (MultipleGenomeMatchingProcessor.cpp:170-176)
    B200SparseEMMatcher(uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin, int device);
    ~B200SparseEMMatcher();
    B200SparseEMMatcher(const B200SparseEMMatcher&) = delete;
    B200SparseEMMatcher& operator=(const B200SparseEMMatcher&) = delete;

    // ref .h:93-97
    void disableSlidingWindow() { swSize = 0; swEnd = circularBuffer ? 0 : maxRefLength; }
    void disableCircularBuffer() { circularBuffer = false; swEnd = maxRefLength; }
    void setSlidingWindowSize(uint8_t factor) { swSize = maxRefLength / factor; }

    // ref .cpp:453-458
    void loadRef(const char* refText, uint64_t refLength, bool loadRCRef, bool addRegionSeparators, char regionSeparator);
    // ref .cpp:439-451
    void loadSeparator(char regionSeparator);
    // loadRef with the text already in this device's memory (dText; stays valid until the call returns)
    void loadRefDevice(const void* dText, const char* hostText, uint64_t refLength, bool loadRCRef, bool addRegionSeparators,
                       char regionSeparator);
    void* refStaging(uint64_t n);
    // pieces of this device's memory concatenated into the staging buffer (a piece with rc = true goes in as its upper-casing
    // reverse complement): the reference extension of a whole target file, MultipleGenomeMatchingProcessor.cpp:389-398
    struct StagePiece { const void* dText; uint64_t len; bool rc; };
    void* stageConcat(const StagePiece* pieces, size_t n, uint64_t* total);
    // the host mirror behind getRef() is only read by the reference's own CPU packer (MBGC_Encoder.cpp:181,318-330,
    // 376-380); with K6 on the device nobody needs it and loadRef stops filling it
    void setHostMirror(bool on) { mirrorEnabled = on; }
    bool hostMirror() const { return mirrorEnabled; }
    int deviceIndex() const { return device; }
    const uint8_t* deviceRef() const { return dRef; }

    // ref .h:104-113
    const char* getRef() const { return hostRef; }
    uint64_t getMaxRefLength() const { return maxRefLength; }
    uint64_t getRefLength() const { return reachedRefLengthCount ? maxRefLength : static_cast<uint64_t>(pos1); }
    uint64_t getLoadingPosition() const { return static_cast<uint64_t>(pos1); }
    uint64_t getLoadedRefLength() const {
        return static_cast<uint64_t>(reachedRefLengthCount) * (maxRefLength - REF_SHIFT) + (static_cast<uint64_t>(pos1) - REF_SHIFT);
    }
    void setPosition(uint64_t refPos, int reached) {
        pos1 = static_cast<int64_t>(refPos);
        reachedRefLengthCount = reached;
        invalidateMemo();
    }
    // the fast path's self-probe memo describes reference bytes [0, memoLen) under the table as of memoSeq
    void invalidateMemo() { memoSeq = 0; }

    // ref .cpp:361-400
    uint64_t acquireWorkerMatchingLockPos();
    void releaseWorkerMatchingLockPos(uint64_t lockValue);

    // ref .cpp:477-492: one text
    void matchTexts(std::vector<TextMatch>& resMatches, const char* destText, uint64_t destLen, bool destIsRef,
                    bool revComplMatching, uint32_t minMatchLength, uint64_t matchingLockPos = SW_END_ERASED_FLAG);

    // n texts against the same frozen state in one device pass; results concatenated in text order
    // listsStayOnDevice: the match lists are not copied to the host (the counts are) -- for a resident packBatch
    void matchBatch(const std::vector<TextView>& texts, const uint64_t* lockPos, uint32_t minMatchLength,
                    bool textsOnDevice, const void* dTextBase, const uint64_t* dOffsets, bool listsStayOnDevice = false);
    const TextMatch* batchMatches() const { return hMatches; }   // pinned, valid until the next match call
    uint64_t batchMatchCount() const { return nMatchesAll; }
    const std::vector<uint64_t>& batchCounts() const { return resCounts; }

    // K6, MBGC_Encoder::processMatches (mbgccoder/MBGC_Encoder.cpp:143-427) for n texts at once: match lists ->
    // the six streams of the back-end coders.  matches == nullptr: pack the lists (and texts) of the immediately
    // preceding single-pass matchBatch call, which are still on the device.  Results: packStream*/packUnmatched/
    // packKept below, valid until the next pack call.  packer.cpp.
    void packBatch(const PackParams& p, const std::vector<TextView>& texts, const TextMatch* matches, const uint64_t* counts,
                   const uint64_t* lockPos, const uint64_t* loadedPosArr, uint64_t nLoadedPos);
    const uint8_t* packStream(int s) const { return hPackOut[s]; }                    // pinned host memory
    const std::vector<uint64_t>& packStreamOff(int s) const { return packOff[s]; }    // [nTexts+1] per stream
    const std::vector<uint64_t>& packUnmatched() const { return packUnmatchedChars; } // processMatches' return value
    const std::vector<uint64_t>& packKept() const { return packKeptCounts; }          // matches left after pass 1
    // textMatches as pass 1 leaves them (MBGC_Encoder.cpp:197), all texts of the last pack call concatenated
    void packKeptMatches(std::vector<TextMatch>& out);
    // MBGC_Encoder::processLiteral (MBGC_Encoder.cpp:128-135; a developer-build option of the reference, off by default):
    // with minLen != UINT64_MAX every pack call also lists the literal runs that extend the reference, in append order
    // (text = index into the texts of the call).  MGMP_Params.h:68-69,175-187.
    void setLiteralRefExt(uint64_t minLen, uint64_t beforeAfter, uint64_t unmatchedFractionFactor) {
        litExt = LitExtParams{minLen, beforeAfter, unmatchedFractionFactor};
    }
    const std::vector<LitRange>& packLiteralRefExt() const { return packLitRanges; }
    const LitExtParams& literalRefExt() const { return litExt; }
    float packMs[4] = {0, 0, 0, 0};       // last pack call: H2D, kernels, D2H, total (CUDA events)
    uint64_t packLaunches = 0;            // kernel launches of the last pack call
    int packSweeps[2] = {0, 0};           // fixed-point sweeps the last pack call needed (removal, pairing)

    // FASTA ingest, kseq_read_lossless_fasta (utils/kseq.h:233-274) for a whole file image: records + DNA line length +
    // final status, the sequences as a batch of texts in device memory (ingestedTexts(), valid until the next ingest).
    // ingest.cpp / ingest.cu.
    void ingestFasta(const char* bytes, uint64_t n);
    const std::vector<FastaRecord>& ingestedRecords() const { return faRecords; }
    const void* ingestedTexts() const { return dFaText; }
    int ingestStatus() const { return faStatus; }                 // -1 end of file, -3 not FASTA, -4 DNA not well-formed
    uint64_t ingestDnaLineLen() const { return faDnaLineLen; }    // seq->dnaLineLen (UINT64_MAX = DNA_NOT_WELLFORMED)
    uint64_t ingestMaxLastLineLen() const { return faMaxLastLineLen; }
    uint64_t ingestLaunches() const { return faLaunches; }
    float faMs[3] = {0, 0, 0};                                     // last ingest: image H2D, kernels, total (CUDA events)
    void copyIngested(uint64_t record, char* out);                // one sequence back to the host (tests)

    // host-only: the reverse complement loadRef writes for loadRCRef (utils/helper.cpp:405-410 with the table of :312-358)
    static void selftestReverseComplement(const uint8_t* src, uint64_t n, uint8_t* dst);

    // inspection
    void copyTableLow32(uint32_t* out, uint64_t nSlots);
    void copyDeviceRef(uint64_t from, uint64_t n, char* out);
    // text i of the last host-buffer batch as it arrived on the device (single-pass batches; tests)
    void copyStagedText(uint64_t textIndex, char* out);
    void* streamHandle() const { return stream; }

    // state (public like the C struct it is mirrored into)
    int64_t pos1 = 1;
    uint64_t maxRefLength = 0;
    int reachedRefLengthCount = 0;
    bool circularBuffer = true;
    uint64_t swSize = 0, swEnd = 0;
    uint64_t samplingPos = 1;
    uint32_t hashSize = 0;
    int L = 0, K = 0, k1 = 0, k2 = 0, skipMargin = 0, k1ord = -1;

    bool stageTiming = true;               // record the stage events of match calls (mbgcg_set_stage_timing)
    float stageMs[MBGCG_T_COUNT_] = {0};   // valid after collectTiming()
    void collectTiming();
    uint64_t lastCounters[8] = {0};
    uint64_t lastWire[2] = {0, 0};         // last match call: bytes that crossed PCIe host -> device, device -> host
    float lastLoadRefMs[2] = {0, 0};       // last loadRef / loadRefDevice: device time (CUDA events), host wall time
    uint64_t lastLoadRefBases = 0;         // reference bytes it appended (forward + reverse complement)
    uint64_t totalLaunches = 0;

private:
    void loadSegment(const char* text, const uint8_t* dSrc, uint64_t len, bool rc, bool addSep, char sep);
    void indexPass();
    bool wantMemo(uint64_t positions);
    void buildMemo();
    bool useSharedFilter() const;
    void fillIndexState(IndexState& ix) const;
    // host texts -> device text buffer: raw DMA and / or host-packed blocks (hostpack.hpp); returns kernel launches
    int stageTexts(const std::vector<TextView>& texts, size_t first, int nC, const std::vector<uint64_t>& off);
    void growPinned(void** p, uint64_t* cap, uint64_t need, const char* what);
    void growDevice(void** p, uint64_t* cap, uint64_t need, const char* what);
    void ensureWork(uint64_t textBytes, uint32_t nTiles, int nContigs, uint32_t seedCap);
    void check(cudaError_t e, const char* what);

    int device = 0;
    bool mirrorEnabled = true;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[12] = {};
    bool timingPending = false;
    char* hostRef = nullptr;             // host mirror (getRef())
    uint64_t hostRefBytes = 0;           // size of its mapping
    uint8_t* dRefAlloc = nullptr;        // device reference buffer with padding
    uint8_t* dRef = nullptr;             // = dRefAlloc + PAD
    uint8_t* dStage = nullptr; uint64_t dStageCap = 0;   // the text of the current loadRef call on the device
    uint8_t* hStage = nullptr; uint64_t hStageCap = 0;   // pinned staging of the same
    unsigned long long* dHt = nullptr;
    uint32_t* dOcc = nullptr;
    uint32_t* dFilter = nullptr;         // first-level probe filter (copied to shared memory by the probe kernel)
    uint64_t insertSeq = 1;              // program-order sequence number of the next index sample (restarts at an epoch rebase)
    uint64_t seqLimit = 1ull << 32;      // rebase when the numbers would leave 32 bits (MBGC_B200_SEQ_LIMIT lowers it: tests)
    uint64_t samplesInserted = 0;        // index samples over the matcher's life
    uint32_t seqRebases = 0;
    // fast path (kernels.cu, k_anchor / k_clean): self-probe memo of the reference + adaptive policy
    uint32_t* dMemoX = nullptr; uint64_t memoCapWords = 0;
    uint64_t memoSeq = 0;                // insertSeq the memo was built at (0 = none)
    uint64_t memoLen = 0;                // reference positions covered
    uint64_t passesSinceInsert = 0;      // match passes since the index last changed
    double cleanEstimate = 0.75;         // expected share of clean 16-position lanes (updated from every pass)
    std::deque<uint64_t> workersSwEndPositions;

    // text transfer (host-packed blocks)
    WorkerPool* pool = nullptr;
    void* hPacked = nullptr; uint64_t hPackedCap = 0;   // pinned: packed blocks
    void* dPacked = nullptr; uint64_t dPackedCap = 0;
    void* hXferMeta = nullptr; uint64_t hXferMetaCap = 0;  // pinned: block descriptors, then exception runs
    void* dXferMeta = nullptr; uint64_t dXferMetaCap = 0;
    std::vector<std::vector<ExcRun>> blockRuns;
    std::vector<uint64_t> lastOffsets, lastLens;

    // per-pass device buffers (grown on demand)
    uint8_t* dText = nullptr; uint64_t dTextCap = 0;
    uint8_t* dMeta = nullptr; uint64_t dMetaCap = 0;
    uint8_t* hMeta = nullptr; uint64_t hMetaCap = 0;   // pinned
    uint8_t* dWork = nullptr; uint64_t dWorkCap = 0;
    WorkBuffers wb{};
    uint32_t wbTiles = 0; int wbContigs = 0;
    TextMatch* hMatches = nullptr; uint64_t hMatchesCap = 0;  // pinned; all passes of the last call, text order
    uint64_t nMatchesAll = 0;
    uint64_t matchGuess = 0;             // matches expected from the next pass (copied back before their count is known)
    uint32_t* hCounts = nullptr; uint64_t hCountsCap = 0;    // pinned
    std::vector<uint64_t> resCounts;

    // what the last matchBatch call left on the device (single-pass calls only): input of a resident packBatch
    struct Resident {
        bool valid = false;
        BatchDesc bd{};
        uint32_t nMatches = 0;
        std::vector<uint64_t> lens;
    } resident;

    // FASTA ingest (ingest.cpp)
    void releaseIngest();
    uint8_t* dFaIn = nullptr; uint64_t dFaInCap = 0;       // file image (16 bytes of padding in front)
    uint8_t* dFaWork = nullptr; uint64_t dFaWorkCap = 0;   // block counts, record starts, records, gather descriptors
    uint8_t* dFaText = nullptr; uint64_t dFaTextCap = 0;   // the sequences, batch layout
    uint8_t* hFaWork = nullptr; uint64_t hFaWorkCap = 0;   // pinned
    std::vector<FastaRecord> faRecords;
    cudaEvent_t faEv[3] = {};
    int faStatus = -1;
    uint64_t faDnaLineLen = 0, faMaxLastLineLen = 0, faLaunches = 0;

    // K6 (packer.cpp)
    void releasePack();
    cudaEvent_t pkEv[4] = {};
    uint8_t* dPackWork = nullptr; uint64_t dPackWorkCap = 0;
    uint8_t* dPackIn = nullptr; uint64_t dPackInCap = 0;       // matches + meta + loaded positions of a host call
    uint8_t* hPackIn = nullptr; uint64_t hPackInCap = 0;       // pinned staging of the same
    uint8_t* hPackSum = nullptr; uint64_t hPackSumCap = 0;     // pinned: flags, totals, offsets, sums
    uint8_t* hPackOut[PACK_STREAMS] = {}; uint64_t hPackOutCap[PACK_STREAMS] = {};
    std::vector<uint64_t> packOff[PACK_STREAMS];
    std::vector<uint64_t> packUnmatchedChars, packKeptCounts;
    struct KeptSlice { const KeptRec* dev; uint32_t n; };  // kept list of the last pass (device) for packKeptMatches
    std::vector<TextMatch> packKeptHost;                        // earlier passes of a multi-pass call
    KeptSlice packKeptLast{nullptr, 0};
    LitExtParams litExt{UINT64_MAX, 0, 0};
    std::vector<LitRange> packLitRanges;
};

}  // namespace mbgc_b200
