// mbgc_b200/dropin/MBGC_Encoder_processMatches_b200.cpp
//
// Link-time substitute for ONE member function of the reference encoder: MBGC_Encoder::processMatches
// (/root/reference/mbgccoder/MBGC_Encoder.cpp:143-308, with extendMatchLeft/Right :310-427 behind it).
//
// It is compiled against the UNMODIFIED reference header mbgccoder/MBGC_Encoder.h.  The rest of MBGC_Encoder.cpp is
// the reference's own object file, in which the symbol of the stock processMatches has been made weak
// (`objcopy --weaken-symbol`, oracle/Makefile target `dropin_full`), so the vtable slot that
// MultipleGenomeMatchingProcessor.cpp:276,381 calls through resolves to the definition below.  It forwards to
// mbgcg_pack_batch (include/mbgc_b200.h, CUDA kernels of mbgc_b200/csrc/pack.cu) and appends the returned bytes to
// the encoder's per-target streams exactly where the stock function would have.
//
// Kept on the host, because it needs encoder state: the dissimilar-contig bail-out (:202-205) and the statistics
// counters (:289-302; developer output only).  processLiteral (:128-135) is switched off in every preset
// (refLiteralMinimalLengthExt = SIZE_MAX, matching/MGMP_Params.h:45); a developer build (-N / -F) gets it from the device
// packer as a list of text ranges (mbgcg_pack_literal_ref_ext).
#include "MBGC_Encoder.h"

#include <cstdio>
#include <cstdlib>

#include "dropin_shared.h"
#include "mbgc_b200.h"

extern size_t extensionsMatchedCharsAll;        // MBGC_Encoder.cpp:124-126
extern size_t extensionsMismatchesAll;
extern size_t removedGapBreakingMatchesAll;

namespace {
[[noreturn]] void die(const char* msg) {
    fprintf(stderr, "%s\n\n", msg);
    exit(EXIT_FAILURE);
}
}  // namespace

size_t MBGC_Encoder::processMatches(vector<PgTools::TextMatch>& textMatches, char* destStart, size_t destLen, int targetIdx,
                                    size_t matchingLockPos) {
    // ---- the batched target loop (MGMP_processTargets_b200.cpp) on the line ----
    if (mbgc_b200_dropin::EncoderState* es = mbgc_b200_dropin::encoder_state_request()) {
        static_assert(sizeof(size_t) == sizeof(uint64_t), "size_t");
        es->loadedPosArr = reinterpret_cast<const uint64_t*>(refExtLoadedPosArr.data());
        es->nLoadedPos = refExtLoadedPosArr.size();
        return 0;
    }
    if (const mbgcg_seq_event_t* ev = mbgc_b200_dropin::prepacked_event()) {
        // a committed contig of the collection pass: its streams were packed on the device against exactly the state the
        // stock loop would have had; append them where the stock function does (:225-287) and keep the counters (:289-302)
        auto seg = [&](int s) { return reinterpret_cast<const char*>(ev->seg[s]); };
        auto len = [&](int s) { return static_cast<size_t>(ev->segLen[s]); };
        if (len(MBGCG_S_LITERALS)) targetLiterals[targetIdx].append(seg(MBGCG_S_LITERALS), len(MBGCG_S_LITERALS));
        if (len(MBGCG_S_MAPOFF))
            targetMapOffDests[targetIdx].write(seg(MBGCG_S_MAPOFF), static_cast<std::streamsize>(len(MBGCG_S_MAPOFF)));
        if (len(MBGCG_S_MAPOFF5)) targetMapOff5thByte[targetIdx].append(seg(MBGCG_S_MAPOFF5), len(MBGCG_S_MAPOFF5));
        if (len(MBGCG_S_MAPLEN))
            targetMapLenDests[targetIdx].write(seg(MBGCG_S_MAPLEN), static_cast<std::streamsize>(len(MBGCG_S_MAPLEN)));
        if (len(MBGCG_S_GAPDELTA)) targetGapDeltas[targetIdx].append(seg(MBGCG_S_GAPDELTA), len(MBGCG_S_GAPDELTA));
        size_t extMatched = 0;
        if (len(MBGCG_S_GAPFLAGS)) {
            targetGapMismatchesFlags[targetIdx].append(seg(MBGCG_S_GAPFLAGS), len(MBGCG_S_GAPFLAGS));
            const char* f = seg(MBGCG_S_GAPFLAGS);
            for (size_t i = 0, e = len(MBGCG_S_GAPFLAGS); i < e; i++) extMatched += f[i] == 0;
        }
        extensionsMatchedCharsAll += extMatched;
        unmatchedCharsAll += ev->unmatched;
        totalMatchedAll += ev->seqLen - ev->unmatched;
        totalDestLenAll += ev->seqLen;
        removedGapBreakingMatchesAll += ev->nMatches - ev->kept;
        return ev->unmatched;
    }
    // processLiteral (:128-135): the device packer lists the literal runs that extend the reference (row a12); only a
    // developer build of the reference can ask for it (main.cpp:308-313)
    const bool literalExt = params->refLiteralMinimalLengthExt != SIZE_MAX;
    mbgcg_pack_params_t pp{};
    pp.enableExtensionsWithMismatches = params->enableExtensionsWithMismatches;
    pp.gapDepthOffsetEncoding = params->gapDepthOffsetEncoding;
    pp.gapDepthMismatchesEncoding = params->gapDepthMismatchesEncoding;
    pp.lazyDecompressionSupport = params->lazyDecompressionSupport;
    pp.frugal64bitLenEncoding = params->frugal64bitLenEncoding;
    pp.enable40bitReference = params->enable40bitReference;
    pp.mismatchesWithExclusion = params->mismatchesWithExclusion;
    pp.mmsMismatchesInitialScore = params->mmsMismatchesInitialScore;
    pp.mmsMismatchesScoreThreshold = params->mmsMismatchesScoreThreshold;
    pp.mmsMismatchPenalty = params->mmsMismatchPenalty;
    pp.mmsMatchBonus = params->mmsMatchBonus;
    pp.gapBreakingMatchMinLength = params->gapBreakingMatchMinLength;

    mbgc_b200_dropin::Trace tr("processMatches", destLen);
    const uint64_t n0 = textMatches.size();
    uint64_t unmatchedChars = 0, kept = 0;
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        mbgcg_t* h = mbgc_b200_dropin::synced_handle(matcher);
        // the list and the text of the matchTexts call right before this one are still on the device
        const mbgc_b200_dropin::LastMatch lm = mbgc_b200_dropin::last_match();
        const bool residentRoute = lm.matcher == matcher && lm.text == destStart && lm.len == destLen && lm.n == n0;
        mbgc_b200_dropin::forget_last_match();
        const char* texts[1] = {destStart};
        const uint64_t lens[1] = {destLen};
        const uint64_t counts[1] = {n0};
        const uint64_t locks[1] = {matchingLockPos};
        static_assert(sizeof(PgTools::TextMatch) == sizeof(mbgcg_match_t), "TextMatch layout");
        mbgcg_pack_result_t r;
        const int factor0 = unmatchedFractionFactors[2 * (targetsCountShift + targetIdx)];
        if (mbgcg_set_literal_ref_ext(h, literalExt ? params->refLiteralMinimalLengthExt : UINT64_MAX,
                                      literalExt ? params->refLiteralBeforeAfterExt : 0, factor0) != MBGCG_OK)
            die(mbgcg_last_error(h));
        const int rc = mbgcg_pack_batch(h, &pp, 1, texts, lens,
                                        residentRoute ? nullptr : reinterpret_cast<const mbgcg_match_t*>(textMatches.data()),
                                        counts, locks, refExtLoadedPosArr.data(), refExtLoadedPosArr.size(), &r);
        if (rc != MBGCG_OK) die(mbgcg_last_error(h));
        unmatchedChars = r.unmatched[0];
        kept = r.kept[0];
        const int unmatchedFractionFactor = unmatchedFractionFactors[2 * (targetsCountShift + targetIdx)];
        if (processedTargetsCount < targetIdx - params->allowedTargetsOutrunForDissimilarContigs &&
            params->isContigDissimilar(destLen, unmatchedChars, unmatchedFractionFactor)) {  // :202-205
            const mbgcg_match_t* km = nullptr;
            uint64_t nk = 0;
            if (mbgcg_pack_kept_matches(h, &km, &nk) != MBGCG_OK) die(mbgcg_last_error(h));
            textMatches.clear();  // the list as pass 1 left it (:197)
            for (uint64_t i = 0; i < nk; i++)
                textMatches.push_back(PgTools::TextMatch(km[i].posSrcText, km[i].length, km[i].posDestText));
            return PROCESSING_MATCHES_SKIPPED_DUE_TO_CONTIG_DISSIMILARITY;
        }
        if (literalExt) {  // :228 / :290, in append order
            const mbgcg_lit_range_t* lr = nullptr;
            uint64_t nl = 0;
            if (mbgcg_pack_literal_ref_ext(h, &lr, &nl) != MBGCG_OK) die(mbgcg_last_error(h));
            for (uint64_t i = 0; i < nl; i++) targetRefExtensions[targetIdx].append(destStart + lr[i].start, lr[i].len);
        }
        auto seg = [&](int s) { return reinterpret_cast<const char*>(r.stream[s] + r.streamOff[s][0]); };
        auto len = [&](int s) { return static_cast<size_t>(r.streamOff[s][1] - r.streamOff[s][0]); };
        // like the stock function, a per-target stream is only touched when something is appended to it (the
        // encoder does not even size the 5th-byte / gap vectors for modes that never write them)
        if (len(MBGCG_S_LITERALS)) targetLiterals[targetIdx].append(seg(MBGCG_S_LITERALS), len(MBGCG_S_LITERALS));
        if (len(MBGCG_S_MAPOFF))
            targetMapOffDests[targetIdx].write(seg(MBGCG_S_MAPOFF), static_cast<std::streamsize>(len(MBGCG_S_MAPOFF)));
        if (len(MBGCG_S_MAPOFF5)) targetMapOff5thByte[targetIdx].append(seg(MBGCG_S_MAPOFF5), len(MBGCG_S_MAPOFF5));
        if (len(MBGCG_S_MAPLEN))
            targetMapLenDests[targetIdx].write(seg(MBGCG_S_MAPLEN), static_cast<std::streamsize>(len(MBGCG_S_MAPLEN)));
        if (len(MBGCG_S_GAPDELTA)) targetGapDeltas[targetIdx].append(seg(MBGCG_S_GAPDELTA), len(MBGCG_S_GAPDELTA));
        size_t extMatched = 0;  // every matched position of an extension is a 0 flag
        if (len(MBGCG_S_GAPFLAGS)) {
            targetGapMismatchesFlags[targetIdx].append(seg(MBGCG_S_GAPFLAGS), len(MBGCG_S_GAPFLAGS));
            const char* f = seg(MBGCG_S_GAPFLAGS);
            for (size_t i = 0, e = len(MBGCG_S_GAPFLAGS); i < e; i++) extMatched += f[i] == 0;
        }
#pragma omp atomic update
        extensionsMatchedCharsAll += extMatched;
    }
    textMatches.clear();
#pragma omp atomic update
    unmatchedCharsAll += unmatchedChars;
#pragma omp atomic update
    totalMatchedAll += destLen - unmatchedChars;
#pragma omp atomic update
    totalDestLenAll += destLen;
#pragma omp atomic update
    removedGapBreakingMatchesAll += n0 - kept;
    return unmatchedChars;
}
