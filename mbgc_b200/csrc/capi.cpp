// mbgc_b200/csrc/capi.cpp -- the extern "C" layer of include/mbgc_b200.h over B200SparseEMMatcher.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbgc_b200.h"
#include "matcher.hpp"

using mbgc_b200::B200SparseEMMatcher;
using mbgc_b200::MatcherError;
using mbgc_b200::TextMatch;
using mbgc_b200::TextView;

struct mbgcg {
    B200SparseEMMatcher* m = nullptr;
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

int mbgcg_create(mbgcg_t** out, uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin,
                 int device) {
    *out = nullptr;
    try {
        mbgcg_t* h = new mbgcg();
        try {
            h->m = new B200SparseEMMatcher(refLengthLimit, targetMatchLength, k1, k2, skipMargin, device);
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

void mbgcg_destroy(mbgcg_t* h) {
    if (!h) return;
    delete h->m;
    delete h;
}

const char* mbgcg_last_error(const mbgcg_t* h) { return h ? h->err.c_str() : ""; }
const char* mbgcg_create_error(void) { return g_createError.c_str(); }

void mbgcg_disable_sliding_window(mbgcg_t* h) { h->m->disableSlidingWindow(); }
void mbgcg_disable_circular_buffer(mbgcg_t* h) { h->m->disableCircularBuffer(); }
void mbgcg_set_sliding_window_size(mbgcg_t* h, int factor) { h->m->setSlidingWindowSize(static_cast<uint8_t>(factor)); }

int mbgcg_load_ref(mbgcg_t* h, const char* text, uint64_t len, int loadRC, int addSeparators, char separator) {
    return guarded(h, [&] { h->m->loadRef(text, len, loadRC != 0, addSeparators != 0, separator); });
}

int mbgcg_load_separator(mbgcg_t* h, char separator) {
    return guarded(h, [&] { h->m->loadSeparator(separator); });
}

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
    *counts = nullptr;
    return guarded(h, [&] {
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
    *counts = nullptr;
    return guarded(h, [&] {
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

uint64_t mbgcg_acquire_lock(mbgcg_t* h) { return h->m->acquireWorkerMatchingLockPos(); }
int mbgcg_release_lock(mbgcg_t* h, uint64_t lockValue) {
    return guarded(h, [&] { h->m->releaseWorkerMatchingLockPos(lockValue); });
}

const char* mbgcg_host_ref(const mbgcg_t* h) { return h->m->getRef(); }
uint64_t mbgcg_max_ref_length(const mbgcg_t* h) { return h->m->getMaxRefLength(); }
uint64_t mbgcg_ref_length(const mbgcg_t* h) { return h->m->getRefLength(); }
uint64_t mbgcg_loading_position(const mbgcg_t* h) { return h->m->getLoadingPosition(); }
uint64_t mbgcg_loaded_ref_length(const mbgcg_t* h) { return h->m->getLoadedRefLength(); }
void mbgcg_set_position(mbgcg_t* h, uint64_t refPos, int reached) { h->m->setPosition(refPos, reached); }

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
void* mbgcg_stream(const mbgcg_t* h) { return h->m->streamHandle(); }

}  // extern "C"
