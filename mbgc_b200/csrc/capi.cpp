// mbgc_b200/csrc/capi.cpp -- the extern "C" layer of include/mbgc_b200.h over B200SparseEMMatcher.
#include <chrono>
#include <cerrno>
#include <unistd.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <thread>
#include <mutex>
#include <memory>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbgc_b200.h"
#include "collection.hpp"
#include "decode.hpp"
#include "hostpack.hpp"
#include "matcher.hpp"

using mbgc_b200::B200SparseEMMatcher;
using mbgc_b200::ReplicaSet;
using mbgc_b200::SequentialPass;
using mbgc_b200::MatcherError;
using mbgc_b200::TextMatch;
using mbgc_b200::TextView;

namespace mbgc_b200 {  // coders.cpp
uint64_t symbols_pack(B200SparseEMMatcher* m, const char* symbols, int nSymbols, const char* src, uint64_t n, uint8_t* dst);
void symbols_unpack(B200SparseEMMatcher* m, const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, char* dst);
uint64_t varlen_encode(B200SparseEMMatcher* m, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap);
void varlen_decode(B200SparseEMMatcher* m, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t expLen);
// rcmatch.cpp
uint64_t rc_match_sequence(B200SparseEMMatcher* m, uint8_t* seq, uint64_t N, uint32_t targetMatchLength, uint32_t minMatchLength,
                           uint8_t* mapOff, uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, float deviceMs[2],
                           uint64_t* launches);
}

struct mbgcg {
    ReplicaSet* rs = nullptr;            // the replicas (one per device); m = replica 0
    B200SparseEMMatcher* m = nullptr;
    SequentialPass* pass = nullptr;      // the collection pass in progress, if any
    mbgc_b200::Decoder* dec = nullptr;   // the last decode call (its DNA and reference stay on the device)
    struct RunResult {                    // what mbgcg_seq_run accumulates (valid until mbgcg_seq_end)
        std::string stream[MBGCG_S_COUNT];
        std::vector<uint64_t> seqCounts, dnaLineLens, refExtSizes, lockPos;
        std::vector<int32_t> status;
        std::string headers;
    } run;
    std::string err;
    std::vector<TextMatch> single;
    std::vector<TextMatch> kept;
};

static_assert(sizeof(mbgcg_match_t) == sizeof(TextMatch), "TextMatch layout");
static_assert(MBGCG_T_COUNT == MBGCG_T_COUNT_, "timing slots");
static_assert(MBGCG_S_COUNT == mbgc_b200::PACK_STREAMS, "stream ids");

static thread_local std::string g_createError;

template <typename F>
static int guarded(mbgcg_t* h, F&& f) {
    try {
        f();
        return MBGCG_OK;
    } catch (const MatcherError& e) {
        h->err = e.what();
        return e.code;
    } catch (const std::exception& e) {
        h->err = e.what();
        return MBGCG_E_NOMEM;
    }
}

extern "C" {

int mbgcg_create_multi(mbgcg_t** out, uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin,
                       const int* devices, int ndev) {
    if (!out) return MBGCG_E_PARAM;
    *out = nullptr;
    try {
        if (!devices || ndev < 1) throw MatcherError(MBGCG_E_PARAM, "mbgcg_create_multi: no devices");
        mbgcg_t* h = new mbgcg();
        try {
            h->rs = new ReplicaSet(refLengthLimit, targetMatchLength, k1, k2, skipMargin, devices, ndev);
            h->m = h->rs->primary();
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
        return MBGCG_OK;
    } catch (const MatcherError& e) {
        g_createError = e.what();
        return e.code;
    } catch (const std::exception& e) {
        g_createError = e.what();
        return MBGCG_E_NOMEM;
    }
}

int mbgcg_create(mbgcg_t** out, uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin,
                 int device) {
    return mbgcg_create_multi(out, refLengthLimit, targetMatchLength, k1, k2, skipMargin, &device, 1);
}

int mbgcg_device_count(const mbgcg_t* h) { return h->rs->size(); }

void mbgcg_destroy(mbgcg_t* h) {
    if (!h) return;
    delete h->pass;
    delete h->dec;
    delete h->rs;   // owns the matchers
    delete h;
}

const char* mbgcg_last_error(const mbgcg_t* h) { return h ? h->err.c_str() : ""; }
const char* mbgcg_create_error(void) { return g_createError.c_str(); }

// state-changing calls reach every replica: each does the identical host bookkeeping and the identical ordered inserts
void mbgcg_disable_sliding_window(mbgcg_t* h) {
    for (int r = 0; r < h->rs->size(); r++) h->rs->at(r)->disableSlidingWindow();
}
void mbgcg_disable_circular_buffer(mbgcg_t* h) {
    for (int r = 0; r < h->rs->size(); r++) h->rs->at(r)->disableCircularBuffer();
}
void mbgcg_set_sliding_window_size(mbgcg_t* h, int factor) {
    for (int r = 0; r < h->rs->size(); r++) h->rs->at(r)->setSlidingWindowSize(static_cast<uint8_t>(factor));
}

int mbgcg_load_ref(mbgcg_t* h, const char* text, uint64_t len, int loadRC, int addSeparators, char separator) {
    return guarded(h, [&] {
        if (len && !text) throw MatcherError(MBGCG_E_PARAM, "load_ref: no text");
        h->rs->loadRef(text, len, loadRC != 0, addSeparators != 0, separator);
    });
}

int mbgcg_load_separator(mbgcg_t* h, char separator) {
    return guarded(h, [&] { h->rs->loadSeparator(separator); });
}

void mbgcg_set_host_mirror(mbgcg_t* h, int on) { h->rs->setHostMirror(on != 0); }

int mbgcg_match(mbgcg_t* h, const char* text, uint64_t len, uint32_t minMatchLength, uint64_t lockPos,
                const mbgcg_match_t** out, uint64_t* n) {
    *out = nullptr;
    *n = 0;
    return guarded(h, [&] {
        h->m->matchTexts(h->single, text, len, false, false, minMatchLength, lockPos);
        *out = reinterpret_cast<const mbgcg_match_t*>(h->single.data());
        *n = h->single.size();
    });
}

int mbgcg_match_batch(mbgcg_t* h, int nTexts, const char* const* texts, const uint64_t* lens, uint32_t minMatchLength,
                      const uint64_t* lockPos, const mbgcg_match_t** out, const uint64_t** counts) {
    if (out) *out = nullptr;
    if (!counts) return MBGCG_E_PARAM;
    *counts = nullptr;
    return guarded(h, [&] {
        if (nTexts < 0 || (nTexts > 0 && (!texts || !lens))) throw MatcherError(MBGCG_E_PARAM, "match_batch: texts and lengths are required");
        std::vector<TextView> tv(static_cast<size_t>(nTexts));
        for (int i = 0; i < nTexts; i++) tv[i] = TextView{texts[i], lens[i]};
        h->m->matchBatch(tv, lockPos, minMatchLength, false, nullptr, nullptr, out == nullptr);
        if (out) *out = reinterpret_cast<const mbgcg_match_t*>(h->m->batchMatches());
        *counts = h->m->batchCounts().data();
    });
}

int mbgcg_match_batch_device(mbgcg_t* h, const void* dText, int nTexts, const uint64_t* offsets, const uint64_t* lens,
                             uint32_t minMatchLength, const uint64_t* lockPos, const mbgcg_match_t** out,
                             const uint64_t** counts) {
    if (out) *out = nullptr;
    if (!counts) return MBGCG_E_PARAM;
    *counts = nullptr;
    return guarded(h, [&] {
        if (nTexts < 0 || (nTexts > 0 && (!dText || !offsets || !lens)))
            throw MatcherError(MBGCG_E_PARAM, "match_batch_device: device text, offsets and lengths are required");
        if (reinterpret_cast<uintptr_t>(dText) % 16)
            throw MatcherError(MBGCG_E_PARAM, "match_batch_device: the device text buffer must be 16-byte aligned");
        std::vector<TextView> tv(static_cast<size_t>(nTexts));
        for (int i = 0; i < nTexts; i++) {
            if (offsets[i] % 16) throw MatcherError(MBGCG_E_PARAM, "device text offsets must be multiples of 16");
            tv[i] = TextView{nullptr, lens[i]};
        }
        h->m->matchBatch(tv, lockPos, minMatchLength, true, dText, offsets, out == nullptr);
        if (out) *out = reinterpret_cast<const mbgcg_match_t*>(h->m->batchMatches());
        *counts = h->m->batchCounts().data();
    });
}

int mbgcg_pack_batch(mbgcg_t* h, const mbgcg_pack_params_t* p, int nTexts, const char* const* texts, const uint64_t* lens,
                     const mbgcg_match_t* matches, const uint64_t* counts, const uint64_t* lockPos,
                     const uint64_t* loadedPosArr, uint64_t nLoadedPos, mbgcg_pack_result_t* out) {
    if (!out) return MBGCG_E_PARAM;
    std::memset(out, 0, sizeof(*out));
    return guarded(h, [&] {
        if (!p || (nTexts > 0 && !lens)) throw MatcherError(MBGCG_E_PARAM, "pack_batch: parameters and text lengths are required");
        if (matches && (!texts || !counts)) throw MatcherError(MBGCG_E_PARAM, "pack_batch: match lists without texts / counts");
        mbgc_b200::PackParams pp{};
        pp.enableExt = p->enableExtensionsWithMismatches;
        pp.gapDepthOffset = p->gapDepthOffsetEncoding;
        pp.gapDepthMismatches = p->gapDepthMismatchesEncoding;
        pp.lazy = p->lazyDecompressionSupport;
        pp.frugalLen = p->frugal64bitLenEncoding;
        pp.enable40bit = p->enable40bitReference;
        pp.withExclusion = p->mismatchesWithExclusion;
        pp.initialScore = p->mmsMismatchesInitialScore;
        pp.scoreThreshold = p->mmsMismatchesScoreThreshold;
        pp.mismatchPenalty = p->mmsMismatchPenalty;
        pp.matchBonus = p->mmsMatchBonus;
        pp.gapBreakingMinLen = p->gapBreakingMatchMinLength;
        std::vector<TextView> tv(static_cast<size_t>(nTexts));
        for (int i = 0; i < nTexts; i++) tv[i] = TextView{texts ? texts[i] : nullptr, lens[i]};
        h->m->packBatch(pp, tv, reinterpret_cast<const TextMatch*>(matches), counts, lockPos, loadedPosArr, nLoadedPos);
        for (int s = 0; s < MBGCG_S_COUNT; s++) {
            out->stream[s] = h->m->packStream(s);
            out->streamOff[s] = h->m->packStreamOff(s).data();
        }
        out->unmatched = h->m->packUnmatched().data();
        out->kept = h->m->packKept().data();
    });
}

int mbgcg_pack_kept_matches(mbgcg_t* h, const mbgcg_match_t** out, uint64_t* n) {
    *out = nullptr;
    *n = 0;
    return guarded(h, [&] {
        h->m->packKeptMatches(h->kept);
        *out = reinterpret_cast<const mbgcg_match_t*>(h->kept.data());
        *n = h->kept.size();
    });
}

int mbgcg_set_literal_ref_ext(mbgcg_t* h, uint64_t refLiteralMinimalLengthExt, uint64_t refLiteralBeforeAfterExt,
                              int unmatchedFractionFactor) {
    return guarded(h, [&] {
        if (unmatchedFractionFactor < 0) throw MatcherError(MBGCG_E_PARAM, "negative unmatchedFractionFactor");
        if (refLiteralMinimalLengthExt != UINT64_MAX && refLiteralBeforeAfterExt >= (1ull << 31))
            throw MatcherError(MBGCG_E_PARAM, "refLiteralBeforeAfterExt out of range");
        h->m->setLiteralRefExt(refLiteralMinimalLengthExt, refLiteralBeforeAfterExt, static_cast<uint64_t>(unmatchedFractionFactor));
    });
}

int mbgcg_pack_literal_ref_ext(mbgcg_t* h, const mbgcg_lit_range_t** ranges, uint64_t* n) {
    if (!ranges || !n) return MBGCG_E_PARAM;
    *ranges = nullptr;
    *n = 0;
    return guarded(h, [&] {
        static_assert(sizeof(mbgcg_lit_range_t) == sizeof(mbgc_b200::LitRange), "range layout");
        const auto& r = h->m->packLiteralRefExt();
        *ranges = reinterpret_cast<const mbgcg_lit_range_t*>(r.data());
        *n = r.size();
    });
}

void mbgcg_pack_timing(const mbgcg_t* h, float ms[4], uint64_t* launches, int sweeps[2]) {
    for (int i = 0; i < 4; i++) ms[i] = h->m->packMs[i];
    if (launches) *launches = h->m->packLaunches;
    if (sweeps) {
        sweeps[0] = h->m->packSweeps[0];
        sweeps[1] = h->m->packSweeps[1];
    }
}

int mbgcg_ingest_fasta(mbgcg_t* h, const char* bytes, uint64_t n, const mbgcg_fasta_rec_t** recs, uint64_t* nRecs,
                       const void** dTexts, uint64_t* dnaLineLen, int* status) {
    *recs = nullptr;
    *nRecs = 0;
    return guarded(h, [&] {
        static_assert(sizeof(mbgcg_fasta_rec_t) == sizeof(mbgc_b200::FastaRecord), "FastaRecord layout");
        if (n && !bytes) throw MatcherError(MBGCG_E_PARAM, "ingest_fasta: no file image");
        h->m->ingestFasta(bytes, n);
        *recs = reinterpret_cast<const mbgcg_fasta_rec_t*>(h->m->ingestedRecords().data());
        *nRecs = h->m->ingestedRecords().size();
        if (dTexts) *dTexts = h->m->ingestedTexts();
        if (dnaLineLen) *dnaLineLen = h->m->ingestDnaLineLen();
        if (status) *status = h->m->ingestStatus();
    });
}

int mbgcg_copy_ingested(mbgcg_t* h, uint64_t record, char* out) {
    return guarded(h, [&] { h->m->copyIngested(record, out); });
}

void mbgcg_ingest_timing(const mbgcg_t* h, float ms[3], uint64_t* launches) {
    for (int i = 0; i < 3; i++) ms[i] = h->m->faMs[i];
    if (launches) *launches = h->m->ingestLaunches();
}

uint64_t mbgcg_acquire_lock(mbgcg_t* h) { return h->rs->acquireLock(); }
int mbgcg_release_lock(mbgcg_t* h, uint64_t lockValue) {
    return guarded(h, [&] { h->rs->releaseLock(lockValue); });
}

const char* mbgcg_host_ref(const mbgcg_t* h) { return h->m->getRef(); }
uint64_t mbgcg_max_ref_length(const mbgcg_t* h) { return h->m->getMaxRefLength(); }
uint64_t mbgcg_ref_length(const mbgcg_t* h) { return h->m->getRefLength(); }
uint64_t mbgcg_loading_position(const mbgcg_t* h) { return h->m->getLoadingPosition(); }
uint64_t mbgcg_loaded_ref_length(const mbgcg_t* h) { return h->m->getLoadedRefLength(); }
void mbgcg_set_position(mbgcg_t* h, uint64_t refPos, int reached) {
    for (int r = 0; r < h->rs->size(); r++) h->rs->at(r)->setPosition(refPos, reached);
}

void mbgcg_get_state(const mbgcg_t* h, mbgcg_state_t* out) {
    const B200SparseEMMatcher* m = h->m;
    out->pos1 = m->pos1;
    out->maxRefLength = m->maxRefLength;
    out->reachedRefLengthCount = m->reachedRefLengthCount;
    out->circularBuffer = m->circularBuffer ? 1 : 0;
    out->swSize = m->swSize;
    out->swEnd = m->swEnd;
    out->samplingPos = m->samplingPos;
    out->hashSize = m->hashSize;
    out->K = m->K;
}

void mbgcg_set_state(mbgcg_t* h, const mbgcg_state_t* in) {
    B200SparseEMMatcher* m = h->m;
    if (m->pos1 != in->pos1 || m->reachedRefLengthCount != in->reachedRefLengthCount) m->invalidateMemo();
    m->pos1 = in->pos1;
    m->reachedRefLengthCount = in->reachedRefLengthCount;
    m->circularBuffer = in->circularBuffer != 0;
    m->swSize = in->swSize;
    m->swEnd = in->swEnd;
    m->samplingPos = in->samplingPos;
    if (h->rs->size() > 1) h->rs->syncStateFromPrimary();
}

int mbgcg_copy_table(mbgcg_t* h, uint32_t* out, uint64_t nSlots) {
    return guarded(h, [&] { h->m->copyTableLow32(out, nSlots); });
}
int mbgcg_copy_device_ref(mbgcg_t* h, uint64_t from, uint64_t n, char* out) {
    return guarded(h, [&] { h->m->copyDeviceRef(from, n, out); });
}

int mbgcg_copy_staged_text(mbgcg_t* h, uint64_t textIndex, char* out) {
    return guarded(h, [&] { h->m->copyStagedText(textIndex, out); });
}

void mbgcg_selftest_revcomp(const char* text, uint64_t n, char* out) {
    B200SparseEMMatcher::selftestReverseComplement(reinterpret_cast<const uint8_t*>(text), n, reinterpret_cast<uint8_t*>(out));
}

int mbgcg_selftest_pack(const char* text, uint64_t n, char* out) {
    std::vector<mbgc_b200::ExcRun> runs;
    std::vector<uint8_t> packed((n + 3) / 4 + 16);
    if (!mbgc_b200::pack_block(reinterpret_cast<const uint8_t*>(text), n, packed.data(), 0, runs, n / 64 + 16)) return 1;
    static const char lut[4] = {'A', 'C', 'T', 'G'};  // code = (c >> 1) & 3
    for (uint64_t i = 0; i < n; i++) out[i] = lut[(packed[i >> 2] >> (2 * (i & 3))) & 3];
    for (const auto& r : runs)
        for (uint32_t k = 0; k < r.len; k++) out[r.off + k] = static_cast<char>(r.byte);
    return 0;
}

int mbgcg_selftest_pack_image(const char* image, uint64_t n, char* out, uint64_t* packedBytes, uint64_t* nRuns) {
    const uint64_t cap = mbgc_b200::packed_image_bound(n);
    std::vector<uint8_t> buf(cap + 16);
    uint8_t* dst = buf.data() + ((16 - (reinterpret_cast<uintptr_t>(buf.data()) & 15)) & 15);
    const uint64_t used = mbgc_b200::pack_image_into(reinterpret_cast<const uint8_t*>(image), n, dst, cap);
    if (!used) return 1;
    mbgc_b200::PackedImageHeader h;
    std::memcpy(&h, dst, sizeof(h));
    if (packedBytes) *packedBytes = used;
    if (nRuns) *nRuns = h.nRuns;
    static const char lut[4] = {'A', 'C', 'T', 'G'};
    const uint8_t* packed = dst + h.packedOff;
    for (uint64_t i = 0; i < n; i++) out[i] = lut[(packed[i >> 2] >> (2 * (i & 3))) & 3];
    const auto* runs = reinterpret_cast<const mbgc_b200::ExcRun*>(dst + h.runsOff);
    for (uint32_t r = 0; r < h.nRuns; r++) {
        const uint64_t step = (runs[r].byte >> 8) ? (runs[r].byte >> 8) : 1;
        for (uint32_t k = 0; k < runs[r].len; k++) out[runs[r].off + k * step] = static_cast<char>(runs[r].byte & 0xFF);
    }
    return h.firstByte == static_cast<uint8_t>(image[0]) && h.n == n ? 0 : 2;
}

void mbgcg_set_stage_timing(mbgcg_t* h, int on) { h->m->stageTiming = on != 0; }

void mbgcg_last_timing(const mbgcg_t* h, float ms[MBGCG_T_COUNT]) {
    try {
        h->m->collectTiming();
    } catch (const std::exception&) {
    }
    for (int i = 0; i < MBGCG_T_COUNT; i++) ms[i] = h->m->stageMs[i];
}
void mbgcg_last_counters(const mbgcg_t* h, uint64_t out[8]) {
    for (int i = 0; i < 8; i++) out[i] = h->m->lastCounters[i];
}
uint64_t mbgcg_total_launches(const mbgcg_t* h) { return h->m->totalLaunches; }
void mbgcg_last_wire_bytes(const mbgcg_t* h, uint64_t out[2]) {
    out[0] = h->m->lastWire[0];
    out[1] = h->m->lastWire[1];
}
void mbgcg_load_ref_timing(const mbgcg_t* h, float ms[2], uint64_t* appended) {
    ms[0] = h->m->lastLoadRefMs[0];
    ms[1] = h->m->lastLoadRefMs[1];
    if (appended) *appended = h->m->lastLoadRefBases;
}
void* mbgcg_stream(const mbgcg_t* h) { return h->m->streamHandle(); }

// ---- the reverse-complement redundancy pass of mode 3 (row f3) -----------------------------------------------------------
int mbgcg_rc_match_sequence(mbgcg_t* h, uint8_t* seq, uint64_t n, uint32_t targetMatchLength, uint32_t minMatchLength, uint64_t* newLen,
                            uint8_t* mapOff, uint64_t* mapOffLen, uint8_t* mapLen, uint64_t* mapLenLen, float deviceMs[2]) {
    return guarded(h, [&] {
        if ((n && !seq) || !newLen || !mapOff || !mapOffLen || !mapLen || !mapLenLen)
            throw MatcherError(MBGCG_E_PARAM, "rc_match_sequence: null argument");
        *newLen = mbgc_b200::rc_match_sequence(h->m, seq, n, targetMatchLength, minMatchLength, mapOff, mapOffLen, mapLen, mapLenLen,
                                              deviceMs, nullptr);
    });
}

// ---- the decoder's match-copy path (row f4) ---------------------------------------------------------------------------------
int mbgcg_decode_targets(mbgcg_t* h, const mbgcg_pack_params_t* p, const mbgcg_decode_input_t* in, mbgcg_decode_result_t* out) {
    if (!out) return MBGCG_E_PARAM;
    std::memset(out, 0, sizeof(*out));
    return guarded(h, [&] {
        if (!p || !in) throw MatcherError(MBGCG_E_PARAM, "decode_targets: null argument");
        if (in->nTargets && (!in->seqCounts || !in->unmatchedFractionFactors))
            throw MatcherError(MBGCG_E_PARAM, "decode_targets: sequence counts and extension factors are required");
        mbgc_b200::DecodeStreamsHost hs{};
        for (int i = 0; i < MBGCG_S_COUNT; i++) {
            if (in->streamLen[i] && !in->stream[i]) throw MatcherError(MBGCG_E_PARAM, "decode_targets: a stream without data");
            hs.p[i] = in->stream[i];
            hs.n[i] = in->streamLen[i];
        }
        mbgc_b200::DcParams dp{};
        dp.enableExt = p->enableExtensionsWithMismatches;
        dp.gapDepthOffset = p->gapDepthOffsetEncoding;
        dp.gapDepthMismatches = p->gapDepthMismatchesEncoding;
        dp.lazy = p->lazyDecompressionSupport;
        dp.initialScore = p->mmsMismatchesInitialScore;
        dp.scoreThreshold = p->mmsMismatchesScoreThreshold;
        dp.mismatchPenalty = p->mmsMismatchPenalty;
        dp.matchBonus = p->mmsMatchBonus;
        dp.refTotalLength = in->refTotalLength;
        mbgc_b200::DecodeLayout L{};
        L.refTotalLength = in->refTotalLength;
        L.g0IsTarget = in->g0IsTarget;
        L.g0SeqCount = in->g0SeqCount;
        L.rcInReference = in->rcInReference;
        L.contigsIndividuallyReversed = in->contigsIndividuallyReversed;
        L.frugalLen = p->frugal64bitLenEncoding;
        L.nTargets = in->nTargets;
        L.seqCounts = in->seqCounts;
        L.factors = in->unmatchedFractionFactors;
        L.lockPos = in->lockPos;
        if (!h->dec) h->dec = new mbgc_b200::Decoder(h->m);
        h->dec->run(dp, hs, L);
        const mbgc_b200::Decoder& d = *h->dec;
        out->nContigs = d.contigLen.size() - d.firstOutputContig;
        out->contigLen = d.contigLen.data() + d.firstOutputContig;
        out->contigUnmatched = d.contigUnmatched.data() + d.firstOutputContig;
        out->dnaLen = d.dnaLen - d.contigBase[d.firstOutputContig];
        out->refPos = d.refPos;
        out->reachedRefLengthCount = d.reachedCount;
        out->batches = d.batches;
        out->parseMs = d.ms[0];
        out->copyMs = d.ms[1];
        out->launches = d.launches;
    });
}
int mbgcg_decode_fetch(mbgcg_t* h, uint64_t firstContig, uint64_t n, uint8_t* dst) {
    return guarded(h, [&] {
        if (!h->dec) throw MatcherError(MBGCG_E_PARAM, "decode_fetch: no decode call on this handle");
        if (!dst && n) throw MatcherError(MBGCG_E_PARAM, "decode_fetch: null destination");
        if (firstContig + n >= (1ull << 31)) throw MatcherError(MBGCG_E_PARAM, "decode_fetch: contig range");
        h->dec->fetch(static_cast<uint32_t>(firstContig + h->dec->firstOutputContig), static_cast<uint32_t>(n), dst);
    });
}
const void* mbgcg_decode_device_dna(mbgcg_t* h, uint64_t firstContig) {
    if (!h || !h->dec || firstContig + h->dec->firstOutputContig >= h->dec->contigBase.size()) return nullptr;
    return h->dec->deviceDna() + h->dec->contigBase[firstContig + h->dec->firstOutputContig];
}
int mbgcg_decode_fetch_ref(mbgcg_t* h, uint64_t pos, uint64_t n, uint8_t* dst) {
    return guarded(h, [&] {
        if (!h->dec) throw MatcherError(MBGCG_E_PARAM, "decode_fetch_ref: no decode call on this handle");
        if (!dst && n) throw MatcherError(MBGCG_E_PARAM, "decode_fetch_ref: null destination");
        h->dec->fetchRef(pos, n, dst);
    });
}

// ---- the two legacy symbol packers (row a14) ---------------------------------------------------------------------------
int mbgcg_symbols_pack(mbgcg_t* h, const char* symbols, int nSymbols, const char* src, uint64_t n, uint8_t* dst, uint64_t* nOut) {
    return guarded(h, [&] {
        if (!symbols || (n && (!src || !dst)) || !nOut) throw MatcherError(MBGCG_E_PARAM, "symbols_pack: null argument");
        *nOut = mbgc_b200::symbols_pack(h->m, symbols, nSymbols, src, n, dst);
    });
}
int mbgcg_symbols_unpack(mbgcg_t* h, const char* symbols, int nSymbols, const uint8_t* packed, uint64_t pos, uint64_t length, char* dst) {
    return guarded(h, [&] {
        if (!symbols || (length && (!packed || !dst))) throw MatcherError(MBGCG_E_PARAM, "symbols_unpack: null argument");
        mbgc_b200::symbols_unpack(h->m, symbols, nSymbols, packed, pos, length, dst);
    });
}
int mbgcg_varlen_encode(mbgcg_t* h, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t dstCap, uint64_t* nOut) {
    return guarded(h, [&] {
        if (!codeBook || (n && !src) || !dst || !nOut) throw MatcherError(MBGCG_E_PARAM, "varlen_encode: null argument");
        *nOut = mbgc_b200::varlen_encode(h->m, codeBook, src, n, dst, dstCap);
    });
}
int mbgcg_varlen_decode(mbgcg_t* h, const char* codeBook, const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t expLen) {
    return guarded(h, [&] {
        if (!codeBook || (n && !src) || (expLen && !dst)) throw MatcherError(MBGCG_E_PARAM, "varlen_decode: null argument");
        mbgc_b200::varlen_decode(h->m, codeBook, src, n, dst, expLen);
    });
}

// ---- pinned host memory for file images -----------------------------------------------------------------------------
void* mbgcg_image_alloc(mbgcg_t* h, uint64_t n) {
    try {
        return h->rs->arena().alloc(n);
    } catch (...) {
        return nullptr;
    }
}
uint64_t mbgcg_image_max(mbgcg_t* h) { return h->rs->arena().maxPiece(); }
int mbgcg_image_free(mbgcg_t* h, const void* p) { return h->rs->arena().release(p) ? MBGCG_OK : MBGCG_E_PARAM; }

// ---- the sequential schedule over a collection, batched and sharded over the handle's devices ---------------------------
int mbgcg_seq_begin(mbgcg_t* h, const mbgcg_seq_config_t* c, const uint64_t* loadedPosArr, uint64_t nLoadedPos) {
    return guarded(h, [&] {
        if (!c) throw MatcherError(MBGCG_E_PARAM, "seq_begin: no configuration");
        if (h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_begin: a pass is already in progress on this handle");
        mbgc_b200::SeqConfig sc;
        const mbgcg_pack_params_t* p = &c->pack;
        sc.pack.enableExt = p->enableExtensionsWithMismatches;
        sc.pack.gapDepthOffset = p->gapDepthOffsetEncoding;
        sc.pack.gapDepthMismatches = p->gapDepthMismatchesEncoding;
        sc.pack.lazy = p->lazyDecompressionSupport;
        sc.pack.frugalLen = p->frugal64bitLenEncoding;
        sc.pack.enable40bit = p->enable40bitReference;
        sc.pack.withExclusion = p->mismatchesWithExclusion;
        sc.pack.initialScore = p->mmsMismatchesInitialScore;
        sc.pack.scoreThreshold = p->mmsMismatchesScoreThreshold;
        sc.pack.mismatchPenalty = p->mmsMismatchPenalty;
        sc.pack.matchBonus = p->mmsMatchBonus;
        sc.pack.gapBreakingMinLen = p->gapBreakingMatchMinLength;
        sc.minMatchLength = c->minMatchLength;
        sc.unmatchedFractionFactor = c->unmatchedFractionFactor;
        sc.unmatchedFractionRCFactor = c->unmatchedFractionRCFactor;
        sc.rcDisabled = c->rcInReferenceDisabled != 0;
        sc.refRegionSeparators = c->refRegionSeparators != 0;
        sc.separator = c->regionSeparator;
        sc.separatorPerFile = c->separatorPerFile != 0;
        if (c->chunkBytes) sc.chunkBytes = c->chunkBytes;
        if (c->maxWindowContigs) sc.maxWindowContigs = c->maxWindowContigs;
        sc.parallelSchedule = c->parallelSchedule != 0;
        if (sc.parallelSchedule) {
            sc.contigsIndividuallyReversed = c->contigsIndividuallyReversed != 0;
            if (c->allowedTargetsOutrun > 0) sc.allowedTargetsOutrun = c->allowedTargetsOutrun;
            sc.allowedTargetsOutrunForDissimilarContigs = c->allowedTargetsOutrunForDissimilarContigs;
            if (c->unmatchedFractionFactorTweakForDissimilarContigs > 0)
                sc.unmatchedFractionFactorTweakForDissimilarContigs = c->unmatchedFractionFactorTweakForDissimilarContigs;
            sc.minimalLengthForDissimilarContigs = c->minimalLengthForDissimilarContigs;
        }
        if (nLoadedPos && !loadedPosArr) throw MatcherError(MBGCG_E_PARAM, "seq_begin: loaded positions missing");
        h->pass = new SequentialPass(*h->rs, sc, loadedPosArr, nLoadedPos);
    });
}

int mbgcg_seq_push_file(mbgcg_t* h, const char* bytes, uint64_t n) {
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_push_file: no pass in progress");
        if (n && !bytes) throw MatcherError(MBGCG_E_PARAM, "seq_push_file: no file image");
        h->pass->pushFile(bytes, n);
    });
}

int mbgcg_seq_push_packed(mbgcg_t* h, const void* packedImage) {
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_push_packed: no pass in progress");
        const auto* ph = static_cast<const mbgc_b200::PackedImageHeader*>(packedImage);
        if (!ph || ph->magic != mbgc_b200::PACKED_IMAGE_MAGIC) throw MatcherError(MBGCG_E_PARAM, "seq_push_packed: not a packed image");
        h->pass->pushFile(static_cast<const char*>(packedImage), ph->n, true);
    });
}

uint64_t mbgcg_image_pack_bound(uint64_t n) { return mbgc_b200::packed_image_bound(n); }

int mbgcg_image_pack(const void* src, uint64_t n, void* dst, uint64_t dstCap, uint64_t* used) {
    if (!used) return MBGCG_E_PARAM;
    *used = 0;
    if ((n && !src) || !dst) return MBGCG_E_PARAM;
    *used = mbgc_b200::pack_image_into(static_cast<const uint8_t*>(src), n, static_cast<uint8_t*>(dst), dstCap);
    return *used ? MBGCG_OK : MBGCG_E_UNSUPPORTED;
}

int mbgcg_image_pack_file(const char* path, void* dst, uint64_t dstCap, uint64_t* used, uint64_t* fileSize) {
    if (!used || !fileSize || !path || !dst) return MBGCG_E_PARAM;
    *used = 0;
    try {
        *used = mbgc_b200::pack_image_from_file(path, static_cast<uint8_t*>(dst), dstCap, fileSize);
    } catch (const MatcherError& e) {
        return e.code;
    } catch (const std::exception&) {
        return MBGCG_E_NOMEM;
    }
    return *used ? MBGCG_OK : MBGCG_E_UNSUPPORTED;
}

int mbgcg_seq_finish_input(mbgcg_t* h) {
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_finish_input: no pass in progress");
        h->pass->finishInput();
    });
}

int mbgcg_seq_wants_input(mbgcg_t* h) { return h->pass && h->pass->wantsInput() ? 1 : 0; }

int mbgcg_seq_next(mbgcg_t* h, mbgcg_seq_event_t* ev) {
    if (!ev) return MBGCG_E_PARAM;
    std::memset(ev, 0, sizeof(*ev));
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_next: no pass in progress");
        mbgc_b200::SeqEvent e;
        h->pass->next(e);
        ev->type = e.type;
        ev->status = e.status;
        ev->file = e.file;
        ev->hdr = e.hdr;
        ev->hdrLen = e.hdrLen;
        ev->seqLen = e.seqLen;
        ev->nMatches = e.nMatches;
        ev->kept = e.kept;
        ev->unmatched = e.unmatched;
        ev->extended = e.extended;
        ev->extendedRC = e.extendedRC;
        ev->contigLock = e.contigLock;
        ev->window = e.window;
        for (int s = 0; s < MBGCG_S_COUNT; s++) {
            ev->seg[s] = e.seg[s];
            ev->segLen[s] = e.segLen[s];
        }
        ev->nContigs = e.nContigs;
        ev->dnaLineLen = e.dnaLineLen;
        ev->refExtSize = e.refExtSize;
        ev->lockPos = e.lockPos;
        ev->loadedRefLength = e.loadedRefLength;
    });
}

int mbgcg_seq_run(mbgcg_t* h, mbgcg_next_image_fn next_image, void* user, int keepStreams, mbgcg_seq_result_t* out) {
    if (!out) return MBGCG_E_PARAM;
    std::memset(out, 0, sizeof(*out));
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_run: no pass in progress");
        if (!next_image) throw MatcherError(MBGCG_E_PARAM, "seq_run: no image source");
        h->run = mbgcg::RunResult();
        auto& R = h->run;
        bool finished = false;
        auto feed = [&](bool block) {   // returns false when nothing was pushed
            const char* bytes = nullptr;
            uint64_t n = 0;
            const int rc = next_image(user, block ? 1 : 0, &bytes, &n);
            if (rc == 3) return true;       // (internal) the source pushes its images itself and has made progress
            if (rc == -2) {                 // (internal) ... and has pushed the last one and closed the input
                finished = true;
                return true;
            }
            if (rc > 0) {
                if (n & MBGCG_IMAGE_PACKED) {
                    const auto* ph = reinterpret_cast<const mbgc_b200::PackedImageHeader*>(bytes);
                    if (!ph || ph->magic != mbgc_b200::PACKED_IMAGE_MAGIC) throw MatcherError(MBGCG_E_PARAM, "seq_run: not a packed image");
                    h->pass->pushFile(bytes, ph->n, true);
                } else {
                    h->pass->pushFile(bytes, n);
                }
                return true;
            }
            if (rc < 0) {
                h->pass->finishInput();
                finished = true;
                return true;
            }
            if (block) throw MatcherError(MBGCG_E_PARAM, "seq_run: the image source returned nothing to a blocking request");
            return false;
        };
        const char SEQ_MARK = static_cast<char>(0xA2), FILE_MARK = static_cast<char>(0xBB);  // MBGC_Params.h:46-47
        uint64_t contigs = 0, matches = 0, unmatched = 0, extensions = 0, digest = 0;
        uint64_t sizes[MBGCG_S_COUNT] = {0, 0, 0, 0, 0, 0};
        mbgc_b200::SeqEvent e;
        const bool trace = std::getenv("MBGC_B200_TRACE") != nullptr;
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double feedMs = 0, nextMs = 0;
        const double tStart = now();
        for (;;) {
            double t0 = now();
            while (!finished && h->pass->wantsInput() && feed(false)) {}
            double t1 = now();
            h->pass->next(e);
            feedMs += t1 - t0;
            nextMs += now() - t1;
            if (e.type == mbgc_b200::SeqEvent::DONE) {
                if (trace)
                    fprintf(stderr, "[mbgc_b200] seq_run: %.1f ms, of which feeding images %.1f, waiting for / producing events %.1f\n",
                            now() - tStart, feedMs, nextMs);
                break;
            }
            if (e.type == mbgc_b200::SeqEvent::NEED_INPUT) {
                if (finished) throw MatcherError(MBGCG_E_PARAM, "seq_run: the pass asks for input after the last file");
                feed(true);
                continue;
            }
            if (e.type == mbgc_b200::SeqEvent::CONTIG) {
                uint64_t hsh = 1469598103934665603ull;
                for (int s = 0; s < MBGCG_S_COUNT; s++) {
                    for (uint64_t k = 0; k < e.segLen[s]; k++) hsh = (hsh ^ e.seg[s][k]) * 1099511628211ull;
                    hsh = (hsh ^ 0xFF) * 1099511628211ull;
                    sizes[s] += e.segLen[s];
                    if (keepStreams && e.segLen[s]) R.stream[s].append(reinterpret_cast<const char*>(e.seg[s]), e.segLen[s]);
                }
                digest += hsh;
                sizes[MBGCG_S_LITERALS]++;
                if (keepStreams) {
                    R.stream[MBGCG_S_LITERALS].push_back(SEQ_MARK);
                    R.headers.append(e.hdr, e.hdrLen);
                    R.headers.push_back('\n');
                }
                contigs++;
                matches += e.nMatches;
                unmatched += e.unmatched;
                extensions += e.extended ? 1 : 0;
            } else {
                if (h->pass->config().pack.enableExt) {
                    sizes[MBGCG_S_GAPFLAGS]++;
                    if (keepStreams) R.stream[MBGCG_S_GAPFLAGS].push_back(FILE_MARK);
                }
                R.seqCounts.push_back(e.nContigs);
                R.dnaLineLens.push_back(e.dnaLineLen);
                R.refExtSizes.push_back(e.refExtSize);
                R.lockPos.push_back(e.lockPos);
                R.status.push_back(e.status);
            }
        }
        for (int s = 0; s < MBGCG_S_COUNT; s++) {
            out->stream[s] = reinterpret_cast<const uint8_t*>(R.stream[s].data());
            out->streamLen[s] = sizes[s];   // (the bytes themselves only with keepStreams)
        }
        out->nFiles = R.seqCounts.size();
        out->nContigs = contigs;
        out->seqCounts = R.seqCounts.data();
        out->dnaLineLens = R.dnaLineLens.data();
        out->refExtSizes = R.refExtSizes.data();
        out->lockPos = R.lockPos.data();
        out->status = R.status.data();
        out->headers = R.headers.data();
        out->headersLen = R.headers.size();
        out->matches = matches;
        out->unmatched = unmatched;
        out->extensions = extensions;
        out->digest = digest;
    });
}

// ---- the pass over files on disk: reader threads inside the library ---------------------------------------------------------
namespace {

// Reader threads bring the files into pinned memory (ahead of the pass, in order of need), packed 4 bases per byte on the way
// when asked to; the pass pulls them through next().  Pinned memory is handed out in FILE ORDER, so however small the arena is,
// the file the pass is waiting for always has its memory and the readers behind it wait for space.
class FileFeeder {
public:
    FileFeeder(mbgcg_t* h, const char* const* paths, uint64_t n, int readers, bool pack)
        : h(h), paths(paths), n(n), pack(pack), ready(n), bytes(n, nullptr), sizes(n, 0) {
        arenaMax = h->rs->arena().maxPiece();
        if (readers < 1) readers = 1;
        if (n == 0) {
            h->pass->finishInput();
            inputClosed = true;
        }
        for (int t = 0; t < readers && static_cast<uint64_t>(t) < n; t++) threads.emplace_back([this] { work(); });
    }
    ~FileFeeder() {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : threads) t.join();
        // images that were read but never handed to the pass go back
        for (uint64_t i = handed; i < n; i++)
            if (ready[i] && bytes[i]) h->rs->arena().release(bytes[i]);
    }
    static int next(void* user, int block, const char** bytesOut, uint64_t* nOut) {
        return static_cast<FileFeeder*>(user)->nextImage(block != 0, bytesOut, nOut);
    }
    double busyMs = 0, feedWaitMs = 0;

private:
    static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    // The readers push their images into the pass themselves, in file order, as soon as they are complete (the pass's entry
    // points are thread-safe): the thread that runs the commit loop spends its time in next(), and a file must not wait for it.
    // For the loop this source therefore never has an image; it only reports progress (3) and the end of the input (-2).
    int nextImage(bool block, const char**, uint64_t*) {
        std::unique_lock<std::mutex> lk(mu);
        if (!error.empty()) throw MatcherError(MBGCG_E_PARAM, error);
        if (inputClosed) return -2;
        if (!block) {
            seenPushes = pushes;
            return 0;
        }
        const double t0 = now();
        cv.wait(lk, [&] { return pushes != seenPushes || inputClosed || !error.empty(); });
        feedWaitMs += now() - t0;
        if (!error.empty()) throw MatcherError(MBGCG_E_PARAM, error);
        seenPushes = pushes;
        return inputClosed ? -2 : 3;
    }
    // (mu held on entry and exit) pushes every complete file that is next in line; one thread at a time
    void pushReady(std::unique_lock<std::mutex>& lk) {
        while (!pushing && handed < n && ready[handed] && !stop) {
            pushing = true;
            const uint64_t i = handed;
            lk.unlock();
            try {
                while (!h->pass->wantsInput()) {
                    {
                        std::lock_guard<std::mutex> g(mu);
                        if (stop) break;
                    }
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
                }
                if (sizes[i] & MBGCG_IMAGE_PACKED)
                    h->pass->pushFile(bytes[i], reinterpret_cast<const mbgc_b200::PackedImageHeader*>(bytes[i])->n, true);
                else
                    h->pass->pushFile(bytes[i], sizes[i]);
            } catch (...) {
                lk.lock();
                pushing = false;
                throw;
            }
            lk.lock();
            handed++;
            pushes++;
            pushing = false;
            cv.notify_all();
        }
        if (handed == n && !inputClosed && !pushing && !stop) {
            pushing = true;
            lk.unlock();
            try {
                h->pass->finishInput();
            } catch (...) {
                lk.lock();
                pushing = false;
                throw;
            }
            lk.lock();
            pushing = false;
            inputClosed = true;
            cv.notify_all();
        }
    }
    void fail(const std::string& what) {
        std::lock_guard<std::mutex> g(mu);
        if (error.empty()) error = what;
        stop = true;
        cv.notify_all();
    }
    void work() {
        for (;;) {
            uint64_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (stop || nextFile >= n) return;
                i = nextFile++;
                cv.wait(lk, [&] { return stop || turn == i; });  // pinned memory goes out in file order
                if (stop) return;
            }
            try {
                struct stat sb;
                if (::stat(paths[i], &sb) != 0) throw MatcherError(MBGCG_E_PARAM, std::string("cannot stat ") + paths[i]);
                const uint64_t size = static_cast<uint64_t>(sb.st_size);
                const uint64_t want = pack ? mbgc_b200::packed_image_bound(size) : size;
                uint8_t* p = nullptr;
                if (size && want <= arenaMax) {
                    for (;;) {
                        p = static_cast<uint8_t*>(h->rs->arena().alloc(want));
                        if (p) break;
                        {
                            std::lock_guard<std::mutex> g(mu);
                            if (stop) return;
                        }
                        std::this_thread::sleep_for(std::chrono::microseconds(50));
                    }
                }
                {
                    std::lock_guard<std::mutex> g(mu);
                    turn = i + 1;
                }
                cv.notify_all();
                const double t0 = now();
                uint64_t outSize = size;
                bool done = false;
                if (p && pack) {
                    uint64_t fsize = 0;
                    const uint64_t used = mbgc_b200::pack_image_from_file(paths[i], p, want, &fsize);
                    if (used) {
                        if (fsize != size) throw MatcherError(MBGCG_E_PARAM, std::string(paths[i]) + " changed size while it was read");
                        outSize = used | MBGCG_IMAGE_PACKED;
                        done = true;
                    } else {  // too many exceptional bytes: the file travels as it is (from pageable memory: the piece is too small)
                        h->rs->arena().release(p);
                        p = nullptr;
                    }
                }
                if (!done && size) {
                    if (!p) {
                        std::unique_ptr<uint8_t[]> mem(new uint8_t[size]);
                        p = mem.get();
                        std::lock_guard<std::mutex> g(mu);
                        pageable.push_back(std::move(mem));
                    }
                    const int fd = ::open(paths[i], O_RDONLY);
                    if (fd < 0) throw MatcherError(MBGCG_E_PARAM, std::string("cannot open ") + paths[i]);
                    uint64_t got = 0;
                    while (got < size) {
                        const ssize_t k = ::read(fd, p + got, size - got);
                        if (k < 0 && errno == EINTR) continue;
                        if (k <= 0) {
                            ::close(fd);
                            throw MatcherError(MBGCG_E_PARAM, std::string("short read on ") + paths[i]);
                        }
                        got += static_cast<uint64_t>(k);
                    }
                    ::close(fd);
                }
                const double dt = now() - t0;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    bytes[i] = reinterpret_cast<const char*>(p);
                    sizes[i] = outSize;
                    ready[i] = 1;
                    busyMs += dt;
                    pushReady(lk);
                }
                cv.notify_all();
            } catch (const std::exception& e) {
                fail(e.what());
                return;
            }
        }
    }
    mbgcg_t* h;
    const char* const* paths;
    uint64_t n;
    bool pack;
    uint64_t arenaMax = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<uint8_t> ready;
    std::vector<const char*> bytes;
    std::vector<uint64_t> sizes;
    std::vector<std::unique_ptr<uint8_t[]>> pageable;  // files that did not fit the arena: they stay until the pass is over
    std::vector<std::thread> threads;
    uint64_t nextFile = 0, turn = 0, handed = 0, pushes = 0, seenPushes = 0;
    bool stop = false, pushing = false, inputClosed = false;
    std::string error;
};

}  // namespace

int mbgcg_seq_run_files(mbgcg_t* h, const char* const* paths, uint64_t nFiles, int readers, int packImages, int keepStreams,
                        mbgcg_seq_result_t* out, double hostMs[2]) {
    if (!out) return MBGCG_E_PARAM;
    std::memset(out, 0, sizeof(*out));
    if (hostMs) hostMs[0] = hostMs[1] = 0;
    if (!h->pass) {
        h->err = "seq_run_files: no pass in progress";
        return MBGCG_E_PARAM;
    }
    if (nFiles && !paths) return MBGCG_E_PARAM;
    try {
        FileFeeder feeder(h, paths, nFiles, readers, packImages != 0);
        const int rc = mbgcg_seq_run(h, &FileFeeder::next, &feeder, keepStreams, out);
        if (hostMs) {
            hostMs[0] = feeder.busyMs;
            hostMs[1] = feeder.feedWaitMs;
        }
        return rc;
    } catch (const std::exception& e) {
        h->err = e.what();
        return MBGCG_E_NOMEM;
    }
}

int mbgcg_seq_end(mbgcg_t* h, mbgcg_seq_stats_t* stats) {
    return guarded(h, [&] {
        if (!h->pass) throw MatcherError(MBGCG_E_PARAM, "seq_end: no pass in progress");
        if (stats) {
            const mbgc_b200::SeqStats& s = h->pass->stats();
            stats->files = s.files;
            stats->contigs = s.contigs;
            stats->windows = s.windows;
            stats->runs = s.runs;
            stats->rerunContigs = s.rerunContigs;
            stats->extensions = s.extensions;
            stats->imageBytes = s.imageBytes;
            stats->textBytes = s.textBytes;
            stats->pageableBytes = s.pageableBytes;
            stats->launches = s.launches + s.ingestLaunches;
            stats->chunks = s.chunks;
            stats->ingestMs = s.ingestMs;
            stats->matchMs = s.matchMs;
            stats->packMs = s.packMs;
            stats->loadRefMs = s.loadRefMs;
            stats->waitStageMs = s.waitStageMs;
            stats->wireBytes = s.wireBytes;
        }
        delete h->pass;
        h->pass = nullptr;
    });
}

}  // extern "C"
