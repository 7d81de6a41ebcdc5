// oracle/ref_decoder_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" face over the reference's own decoder class, MBGC_Decoder<false> (mbgccoder/MBGC_Decoder.{h,cpp}, compiled
// unmodified from /root/reference by oracle/Makefile into oracle/_ref/libref_decoder.so), so that the match-copy path of the
// decoder (row f4 of SURVEY.md 8: initReference / decodeReference :243-317, decodeTarget :532-649 with
// decodeSequenceAndReturnUnmatchedChars :319-432, extendMatchRight/Left :434-527, loadRef :651-675) can be driven on streams
// a test supplies instead of an archive.  Those members are private: the class is included with `private` spelled `public`
// (access control does not change layout or mangled names, the object file is the stock one).  No reference source is copied.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#define private public
#define protected public
#include "mbgccoder/MBGC_Decoder.h"
#undef private
#undef protected

// asmlib is a non-PIC archive; the reference itself maps these to the libc calls on ARM (utils/helper.cpp:598-608)
extern "C" void* A_memcpy(void* dest, const void* src, size_t n) { return memcpy(dest, src, n); }
extern "C" int A_memcmp(const void* s1, const void* s2, size_t n) { return memcmp(s1, s2, n); }

namespace {
struct Handle {
    MBGC_Params params;
    MBGC_Decoder<false>* dec = nullptr;
    std::vector<uint32_t> seqCounts;
    std::vector<size_t> locks;
    std::string factors;
};
}  // namespace

extern "C" {

// pp: enableExtensionsWithMismatches, gapDepthOffsetEncoding, gapDepthMismatchesEncoding, lazyDecompressionSupport,
//     frugal64bitLenEncoding, mmsMismatchesInitialScore, mmsMismatchesScoreThreshold, mmsMismatchPenalty, mmsMatchBonus,
//     rcInReference, contigsIndividuallyReversed
void* refd_create(const int* pp, uint64_t refTotalLength, const uint8_t* lit, uint64_t nLit, const uint8_t* off, uint64_t nOff,
                  const uint8_t* off5, uint64_t nOff5, const uint8_t* len, uint64_t nLen, const uint8_t* gd, uint64_t nGd,
                  const uint8_t* fl, uint64_t nFl, const uint32_t* seqCounts, const uint8_t* factors2, const uint64_t* locks,
                  uint32_t nTargets) {
    Handle* h = new Handle();
    MBGC_Params& p = h->params;
    p.enableExtensionsWithMismatches = pp[0];
    p.gapDepthOffsetEncoding = static_cast<uint8_t>(pp[1]);
    p.gapDepthMismatchesEncoding = static_cast<uint8_t>(pp[2]);
    p.lazyDecompressionSupport = pp[3];
    p.frugal64bitLenEncoding = pp[4];
    p.mmsMismatchesInitialScore = pp[5];
    p.mmsMismatchesScoreThreshold = pp[6];
    p.mmsMismatchPenalty = pp[7];
    p.mmsMatchBonus = pp[8];
    p.refExtensionStrategy = pp[9] ? MBGC_Params::FRACTION_REF_EXTENSION_STRATEGY
                                   : (MBGC_Params::FRACTION_REF_EXTENSION_STRATEGY | MBGC_Params::NO_RC_REF_EXTENSION_STRATEGY_MASK);
    p.contigsIndivduallyReversed = pp[10];
    p.enableCustomDNAformatting = false;
    p.storeFileSeparatorMarksInHeadersStream = false;
    p.masterFilterPattern = std::string("\x01no file is selected");  // moveToFileIfSelected only clears its buffer
    MBGC_Decoder<false>* d = new MBGC_Decoder<false>(&p, nullptr);
    h->dec = d;
    // what decodeInit (:1064-1172) sets up from the archive
    d->refTotalLength = refTotalLength;
    d->filesCount = nTargets + 1;
    d->targetsCount = nTargets;
    d->largestFileLength = 1 << 20;
    d->largestContigLength = 1 << 20;
    d->threadSeqStr.resize(1);
    d->threadExtStr.resize(1);
    d->threadGzOut.resize(1);
    d->refStr.resize(refTotalLength + p.maxConsecutiveMismatches);
    d->threadOutBuffer.resize(1);
    d->literalStr.assign(reinterpret_cast<const char*>(lit), nLit);
    d->mapOffStream.assign(reinterpret_cast<const char*>(off), nOff);
    d->mapOff5thByteStream.assign(reinterpret_cast<const char*>(off5), nOff5);
    d->mapLenStream.assign(reinterpret_cast<const char*>(len), nLen);
    d->gapDeltaStr.assign(reinterpret_cast<const char*>(gd), nGd);
    d->gapMismatchesFlagsStr.assign(reinterpret_cast<const char*>(fl), nFl);
    // one constant header template per file ("h": no MATCH_MARK, so nothing is read from the headers stream); an EMPTY first
    // template would make initReference return before it seeds the reference (:249)
    for (uint32_t i = 0; i <= nTargets; i++) {
        d->headersTemplates.push_back('h');
        d->headersTemplates.push_back(static_cast<char>(MBGC_Params::FILE_SEPARATOR_MARK));
    }
    d->literalsLength = nLit;
    if (p.enableExtensionsWithMismatches) {
        d->threadMatchExtension.emplace_back("");
        d->threadMatchExtension[0].resize(MBGC_Params::MAX_EXTEND_MATCH_LEFT_LENGTH);
    }
    h->seqCounts.assign(seqCounts, seqCounts + nTargets + 1);  // entry 0: G0 (only read by decodeReference)
    h->factors.assign(reinterpret_cast<const char*>(factors2), 2 * static_cast<size_t>(nTargets));
    d->unmatchedFractionFactorsStr = h->factors;
    d->unmatchedFractionFactorsArray = (uint8_t*) d->unmatchedFractionFactorsStr.data();
    d->seqsCountArr = h->seqCounts.data() + 1;
    d->tIdMapOffSrc.emplace_back(d->mapOffStream);
    d->tIdMapOff5thBytePtr.emplace_back(refTotalLength > UINT32_MAX ? (uint8_t*) d->mapOff5thByteStream.data() : nullptr);
    d->tIdGapDeltaPtr.push_back((uint8_t*) d->gapDeltaStr.data());
    d->gapDeltaGuard = (uint8_t*) d->gapDeltaStr.data() + d->gapDeltaStr.size();
    d->tIdGapMismatchesFlagsPtr.push_back(p.enableExtensionsWithMismatches ? (uint8_t*) d->gapMismatchesFlagsStr.data() : nullptr);
    d->gapMismatchesFlagsGuard = (uint8_t*) d->gapMismatchesFlagsStr.data() + d->gapMismatchesFlagsStr.size();
    d->decodeMapLenStream(d->mapLenStream, d->mapLenArr);
    d->dnaLineLengthsArr = nullptr;
    if (locks) {
        h->locks.assign(locks, locks + nTargets);
        d->matchingLocksPosArr = h->locks.data();
    } else
        d->matchingLocksPosArr = nullptr;
    d->tIdNamesPos.push_back(0);
    d->extractedFilesCount.push_back(0);
    d->tIdHeadersPos.push_back(0);
    d->tIdHTemplatesPos.push_back(0);
    d->tIdLiteralPos.push_back(0);
    d->sequentialDecoding = true;
    for (int i = 0; i < MBGC_Params::MAX_GAP_DEPTH; i++) d->pairedGapOffsetDeltaInit[i] = INT64_MAX;
    return h;
}

void refd_destroy(void* hv) {
    Handle* h = static_cast<Handle*>(hv);
    delete h->dec;
    delete h;
}

// initReference (:243-262); note it pushes one more literal cursor, exactly as in decodeInit's call
void refd_init_reference(void* hv) {
    Handle* h = static_cast<Handle*>(hv);
    h->dec->tIdLiteralPos.clear();
    h->dec->initReference("");
}

// decodeReference (:264-317): the G0 contigs themselves are discarded (no file is selected); reference and cursors remain
void refd_decode_reference(void* hv) {
    Handle* h = static_cast<Handle*>(hv);
    h->dec->seqsCountArr = h->seqCounts.data();
    h->dec->decodeReference("");
    h->dec->threadOutBuffer[0].clear();
}

// decodeTarget(targetIdx, MASTER_THREAD_ID): returns the size of the output buffer (">h\n" + sequence + "\n" per contig), copied to out when it fits
uint64_t refd_decode_target(void* hv, uint32_t targetIdx, uint8_t* out, uint64_t cap) {
    Handle* h = static_cast<Handle*>(hv);
    std::string& buf = h->dec->threadOutBuffer[0];
    buf.clear();
    h->dec->decodeTarget(targetIdx, 0, false);
    const uint64_t n = buf.size();
    if (n <= cap) memcpy(out, buf.data(), n);
    return n;
}

// decodeSequenceAndReturnUnmatchedChars alone (no reference growth)
uint64_t refd_decode_sequence(void* hv, uint64_t refLockPos, uint8_t* out, uint64_t cap, uint64_t* unmatched) {
    Handle* h = static_cast<Handle*>(hv);
    std::string dest;
    *unmatched = h->dec->decodeSequenceAndReturnUnmatchedChars(dest, refLockPos, 0);
    if (dest.size() <= cap) memcpy(out, dest.data(), dest.size());
    return dest.size();
}

void refd_load_ref(void* hv, const uint8_t* text, uint64_t n, uint64_t refLockPos, int rc) {
    Handle* h = static_cast<Handle*>(hv);
    h->dec->loadRef(reinterpret_cast<const char*>(text), n, refLockPos, h->dec->refPos, rc != 0);
}

const uint8_t* refd_ref(void* hv) { return reinterpret_cast<const uint8_t*>(static_cast<Handle*>(hv)->dec->refStr.data()); }
uint64_t refd_ref_pos(void* hv) { return static_cast<Handle*>(hv)->dec->refPos; }
int refd_reached_count(void* hv) { return static_cast<Handle*>(hv)->dec->reachedRefLengthCount; }
void refd_set_ref_pos(void* hv, uint64_t pos) { static_cast<Handle*>(hv)->dec->refPos = pos; }
void refd_cursors(void* hv, uint64_t out6[6]) {
    MBGC_Decoder<false>* d = static_cast<Handle*>(hv)->dec;
    out6[0] = static_cast<uint64_t>(d->tIdLiteralPos[0]);
    out6[1] = static_cast<uint64_t>(d->tIdMapOffSrc[0].getPosition()) / 4;
    out6[2] = d->tIdMapOff5thBytePtr[0] ? static_cast<uint64_t>(d->tIdMapOff5thBytePtr[0] - (uint8_t*) d->mapOff5thByteStream.data()) : 0;
    out6[3] = static_cast<uint64_t>(d->tIdMapLenPtr[0] - (d->mapLenArr.empty() ? (uint32_t*) d->mapLenStream.data() : d->mapLenArr.data()));
    out6[4] = static_cast<uint64_t>(d->tIdGapDeltaPtr[0] - (uint8_t*) d->gapDeltaStr.data());
    out6[5] = d->tIdGapMismatchesFlagsPtr[0] ? static_cast<uint64_t>(d->tIdGapMismatchesFlagsPtr[0] - (uint8_t*) d->gapMismatchesFlagsStr.data()) : 0;
}

}  // extern "C"
