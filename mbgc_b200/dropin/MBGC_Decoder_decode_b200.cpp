// mbgc_b200/dropin/MBGC_Decoder_decode_b200.cpp
//
// Link-time substitute for the target loop of the reference decoder (row f4 of SURVEY.md 8): explicit specialisations of
// MBGC_Decoder<lazyMode>::decode (/root/reference/mbgccoder/MBGC_Decoder.cpp:1174-1222) and ::extractFilesSequentially
// (:678-735), compiled against the UNMODIFIED header -- member functions of the class, so they reach its private members the
// way the stock ones do.  The template's own instantiations in MBGC_Decoder.o are weak symbols; these are strong.
//
// The stock loop decodes target after target on the host: decodeTarget (:532-649) = per contig header, decodeSequence...
// (:319-432), writeDNA, extension decision, loadRef.  Here the sequences of ALL targets come from one call of
// mbgcg_decode_targets (include/mbgc_b200.h; CUDA kernels of mbgc_b200/csrc/decode.cu); what stays is what belongs to the
// archive's container: header templates, line breaks, file names, the output files -- done with the class's own members
// (decodeHeader, writeDNA, moveToFileIfSelected), in the stock order.
//
// Modes the device path does not cover run the stock code: the stock extractFilesSequentially is kept under another
// name in the reference's own object (objcopy --redefine-sym, oracle/Makefile target `dropin_decode`) and
// extractFilesParallel is untouched.  Covered: archives of version 2.x with more than one file, whole extraction
// (no -f / -F pattern, no append / repack, no single-FASTA mode).
#include "MBGC_Decoder.h"

#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

#include "mbgc_b200.h"

// the stock member functions under their new names (Itanium ABI: `this` is the first argument)
extern "C" void mbgc_stock_extractFilesSequentially_lazy(void* self);
extern "C" void mbgc_stock_extractFilesSequentially_plain(void* self);

namespace {

[[noreturn]] void die(const char* msg) {
    fprintf(stderr, "%s\n\n", msg);
    exit(EXIT_FAILURE);
}

mbgcg_t* decode_handle() {
    static mbgcg_t* h = nullptr;
    if (!h) {
        int device = 0;
        if (const char* e = std::getenv("MBGC_B200_DEVICES")) device = std::atoi(e);
        if (mbgcg_create(&h, 1 << 20, 32, 16, 1, 16, device) != MBGCG_OK) die(mbgcg_create_error());
    }
    return h;
}

}  // namespace

// what decodeTarget + extractNextSequentially do around the sequence decode, for every target, from ready-made sequences
#define MBGC_B200_EXTRACT_ALL(STOCK)                                                                                              \
    const bool covered = !params->appendCommand && !params->repackCommand && params->masterFilterPattern.empty() &&               \
                         !isFilterListActive && filesCount > 1 && params->isVersionAtLeast(2, 0) && !params->enableCustomDNAformatting &&\
                         !std::getenv("MBGC_B200_STOCK_DECODE");                                                                  \
    if (!covered) {                                                                                                                \
        STOCK(this);                                                                                                               \
        return;                                                                                                                    \
    }                                                                                                                              \
    mbgcg_t* h = decode_handle();                                                                                                  \
    mbgcg_pack_params_t pp{};                                                                                                      \
    pp.enableExtensionsWithMismatches = params->enableExtensionsWithMismatches;                                                    \
    pp.gapDepthOffsetEncoding = params->gapDepthOffsetEncoding;                                                                    \
    pp.gapDepthMismatchesEncoding = params->gapDepthMismatchesEncoding;                                                            \
    pp.lazyDecompressionSupport = params->lazyDecompressionSupport;                                                                \
    pp.frugal64bitLenEncoding = params->frugal64bitLenEncoding;                                                                    \
    pp.mmsMismatchesInitialScore = params->mmsMismatchesInitialScore;                                                              \
    pp.mmsMismatchesScoreThreshold = params->mmsMismatchesScoreThreshold;                                                          \
    pp.mmsMismatchPenalty = params->mmsMismatchPenalty;                                                                            \
    pp.mmsMatchBonus = params->mmsMatchBonus;                                                                                      \
    mbgcg_decode_input_t in{};                                                                                                     \
    const std::string* src[MBGCG_S_COUNT] = {&literalStr, &mapOffStream, &mapOff5thByteStream, &mapLenStream, &gapDeltaStr,       \
                                             &gapMismatchesFlagsStr};                                                             \
    for (int s = 0; s < MBGCG_S_COUNT; s++) {                                                                                      \
        in.stream[s] = reinterpret_cast<const uint8_t*>(src[s]->data());                                                           \
        in.streamLen[s] = src[s]->size();                                                                                          \
    }                                                                                                                              \
    in.refTotalLength = refTotalLength;                                                                                            \
    in.g0IsTarget = params->g0IsTarget ? 1 : 0;                                                                                    \
    in.g0SeqCount = params->g0IsTarget ? 0 : seqsCountArr[-1]; /* decodeReference stepped over G0's entry (:272) */                \
    in.rcInReference = params->isRCinReferenceDisabled() ? 0 : 1;                                                                  \
    in.contigsIndividuallyReversed = params->contigsIndivduallyReversed ? 1 : 0;                                                   \
    in.nTargets = static_cast<uint32_t>(targetsCount);                                                                             \
    in.seqCounts = seqsCountArr;                                                                                                   \
    in.unmatchedFractionFactors = unmatchedFractionFactorsArray;                                                                   \
    std::vector<uint64_t> locks;                                                                                                   \
    if (matchingLocksPosArr) {                                                                                                     \
        locks.assign(matchingLocksPosArr, matchingLocksPosArr + targetsCount);                                                     \
        in.lockPos = locks.data();                                                                                                 \
    }                                                                                                                              \
    mbgcg_decode_result_t r;                                                                                                       \
    if (mbgcg_decode_targets(h, &pp, &in, &r) != MBGCG_OK) die(mbgcg_last_error(h));                                               \
    std::string dna(r.dnaLen, '\0');                                                                                               \
    if (mbgcg_decode_fetch(h, 0, r.nContigs, reinterpret_cast<uint8_t*>(&dna[0])) != MBGCG_OK) die(mbgcg_last_error(h));           \
    if (std::getenv("MBGC_B200_TRACE"))                                                                                            \
        fprintf(stderr, "[mbgc_b200] decode: %llu contigs, %llu bytes of DNA, %u copy batches, %llu launches; device: parse %.2f ms, copy %.2f ms\n", \
                (unsigned long long) r.nContigs, (unsigned long long) r.dnaLen, r.batches, (unsigned long long) r.launches, r.parseMs, r.copyMs); \
    uint64_t contig = params->g0IsTarget ? 0 : in.g0SeqCount; /* G0's own contigs were written by decodeReference */              \
    size_t at = 0;                                                                                                                 \
    for (uint64_t c = 0; c < contig; c++) at += r.contigLen[c];                                                                    \
    const int tId = MASTER_THREAD_ID;                                                                                              \
    for (masterThreadTargetIdx = 0; masterThreadTargetIdx < static_cast<int>(targetsCount); masterThreadTargetIdx++) {            \
        const size_t t = static_cast<size_t>(masterThreadTargetIdx);                                                               \
        namePos = namesStr.find(MBGC_Params::FILE_SEPARATOR_MARK, tIdNamesPos[tId]);             /* extractNextSequentially :739 */\
        currentName = namesStr.substr(tIdNamesPos[tId], namePos - tIdNamesPos[tId]);                                               \
        size_t hTemplateEnd = headersTemplates.find(MBGC_Params::FILE_SEPARATOR_MARK, tIdHTemplatesPos[tId]);  /* decodeTarget :555 */ \
        string headerTemplate(headersTemplates, tIdHTemplatesPos[tId], hTemplateEnd - tIdHTemplatesPos[tId]);                      \
        const uint32_t seqCount = seqsCountArr[t];                                                                                 \
        int64_t dnaLineLength = dnaLineLengthsArr == nullptr ? 0 : static_cast<int64_t>(dnaLineLengthsArr[t]);                     \
        dnaLineLength = dnaLineLength ? dnaLineLength : INT64_MAX;                                                                 \
        for (uint32_t i = 0; i < seqCount; i++, contig++) {                                                                        \
            threadOutBuffer[tId].push_back('>');                                                                                   \
            decodeHeader(headerTemplate, tId);                                                                                     \
            threadOutBuffer[tId].push_back('\n');                                                                                  \
            writeDNA(dna.data() + at, static_cast<int64_t>(r.contigLen[contig]), dnaLineLength, tId);                              \
            at += r.contigLen[contig];                                                                                             \
        }                                                                                                                          \
        if (params->storeFileSeparatorMarksInHeadersStream) tIdHeadersPos[tId]++;                                                  \
        tIdHTemplatesPos[tId] = hTemplateEnd + 1;                                                                                  \
        tIdNamesPos[tId] = namesStr.find(MBGC_Params::FILE_SEPARATOR_MARK, tIdNamesPos[tId]) + 1;                                  \
        decodedTargetsCount++;                                                                                                     \
        masterTargetsStats++;                                                                                                      \
        moveToFileIfSelected(currentName, threadOutBuffer[MASTER_THREAD_ID], 0, masterThreadTargetIdx, false);                    \
    }

template <>
void MBGC_Decoder<true>::extractFilesSequentially() {
    MBGC_B200_EXTRACT_ALL(mbgc_stock_extractFilesSequentially_lazy)
}

template <>
void MBGC_Decoder<false>::extractFilesSequentially() {
    MBGC_B200_EXTRACT_ALL(mbgc_stock_extractFilesSequentially_plain)
}

// decode (:1174-1222): the stock dispatcher holds an inlined copy of extractFilesParallel, so it is restated here -- with the
// device path taking the place of both extraction loops whenever it covers the request
#define MBGC_B200_DECODE(LAZY)                                                                                                     \
    if (params->infoCommand) {                                                                                                     \
        list();                                                                                                                    \
        return;                                                                                                                    \
    }                                                                                                                              \
    decodeInit();                                                                                                                  \
    const bool covered = !params->appendCommand && !params->repackCommand && params->masterFilterPattern.empty() &&               \
                         !isFilterListActive && filesCount > 1 && params->isVersionAtLeast(2, 0) && !params->enableCustomDNAformatting &&\
                         !std::getenv("MBGC_B200_STOCK_DECODE");                                                                  \
    if (sequentialDecoding || covered)                                                                                             \
        extractFilesSequentially();                                                                                                \
    else                                                                                                                           \
        extractFilesParallel();                                                                                                    \
    if (filesCount == 1 && params->decompressionToGzCoderLevel)                                                                    \
        moveToFileIfSelected(g0name, threadGzOut[MASTER_THREAD_ID], MASTER_THREAD_ID);                                             \
    refStr.resize(refTotalLength);                                                                                                 \
    uint32_t totalExtractedFilesCount = std::accumulate(extractedFilesCount.begin(), extractedFilesCount.end(), 0);               \
    if (filesCount == 1 && totalExtractedFilesCount) totalExtractedFilesCount = 1;                                                 \
    if (!params->appendCommand)                                                                                                    \
        *PgHelpers::appout << "extracted " << totalExtractedFilesCount << (totalExtractedFilesCount == 1 ? " file" : " files") << endl; \
    *PgHelpers::logout << PgHelpers::time_millis() << "       \t";                                                                 \
    *PgHelpers::logout << "     \t              \t           \t";

template <>
void MBGC_Decoder<true>::decode() {
    MBGC_B200_DECODE(true)
}

template <>
void MBGC_Decoder<false>::decode() {
    MBGC_B200_DECODE(false)
}
