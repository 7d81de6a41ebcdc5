/* oracle/mbgc_oracle_dec.c -- TEST INFRASTRUCTURE ONLY (see mbgc_oracle.h).
 *
 * Plain-C restatement of the decoder's match-copy path (row f4 of SURVEY.md 8): what the reference decoder does
 * between "streams are decompressed" and "the DNA of a target is in a buffer" -- the sequence decode with
 * mismatch-tolerant extensions, the reference growth and its G0 bootstrap.  FASTA formatting (headers, line breaks),
 * file output, lazy / filtered decompression and the back-end coders are not part of the path.
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Parity status: PINNED against the reference's own MBGC_Decoder<false> member functions driven on the same streams
 * (oracle/_ref/libref_decoder.so, tests/test_decode_oracle.py) and by round trip: the streams of the (pinned)
 * encoder restatement decode to the bytes that were encoded.
 */
#include <stdlib.h>
#include <string.h>

#include "mbgc_oracle.h"

#define MATCH_MARK ((uint8_t) ('%' + 128))         /* mbgccoder/MBGC_Params.h:45-48 */
#define SEQ_SEPARATOR_MARK ((uint8_t) ('"' + 128))
#define REF_REGION_SEPARATOR 0                     /* matching/MGMP_Params.h:14 */
#define MAX_GAP_DEPTH 128                          /* mbgccoder/MBGC_Params.h:50 */
#define MAX_EXTEND_MATCH_LEFT_LENGTH (1 << 24)     /* mbgccoder/MBGC_Params.h:55 */
#define REF_SHIFT 1                                /* mbgccoder/MBGC_Decoder.h:35 */
#define NPOS UINT64_MAX

struct mbo_decoder {
    mbo_pack_params_t p;
    int rcInReference, contigsIndividuallyReversed;
    uint8_t* ref;            /* refStr: refTotalLength + 256 bytes (the reference adds maxConsecutiveMismatches) */
    uint64_t refTotalLength;
    uint64_t refPos, refG0InitPos;
    int reachedRefLengthCount;
    mbo_dec_streams_t s;
    uint32_t* mapLen;        /* decodeMapLenStream's array */
    uint64_t nMapLenArr;
    /* cursors (tIdLiteralPos, tIdMapOffSrc, tIdMapOff5thBytePtr, tIdMapLenPtr, tIdGapDeltaPtr, tIdGapMismatchesFlagsPtr) */
    uint64_t litPos, offPos, off5Pos, lenIdx, gdPos, flagsPos;
    uint8_t* ext;            /* threadMatchExtension */
};

/* coders/ContextAwareMismatchesCoder.cpp:29-34 (+ the table built at :8-17 from mis2code, .h:13-17).  The reference indexes
 * code2mis[sym2val[actual]][code] unchecked; codes 0..3 are only ever produced against A/C/G/T/N, anything else is returned
 * as it is here. */
uint8_t mbo_code2mismatch(uint8_t actual, uint8_t code) {
    static const uint8_t mis2code[5][5] = {{'A', 2, 0, 1, 3}, {1, 'C', 2, 0, 3}, {0, 2, 'G', 1, 3},
                                           {1, 0, 2, 'T', 3}, {1, 2, 3, 0, 'N'}};
    static const char val2sym[] = "ACGTN";
    int a;
    if (code >= 32) return code; /* val2symLen */
    switch (actual) { case 'A': a = 0; break; case 'C': a = 1; break; case 'G': a = 2; break;
                      case 'T': a = 3; break; case 'N': a = 4; break; default: return code; }
    for (int j = 0; j < 5; j++)
        if (mis2code[a][j] == code) return (uint8_t) val2sym[j];
    return code;
}

static uint64_t find_byte(const uint8_t* s, uint64_t n, uint8_t c, uint64_t from) { /* std::string::find(char, pos) */
    if (from >= n) return NPOS;
    const uint8_t* q = (const uint8_t*) memchr(s + from, c, n - from);
    return q ? (uint64_t) (q - s) : NPOS;
}

/* MBGC_Decoder.cpp:966-986 (decodeMapLenStream; archive version >= 2.0) */
static void decode_map_len(mbo_decoder* d) {
    const uint8_t* s = d->s.mapLen;
    const uint64_t n = d->s.nMapLen;
    d->mapLen = (uint32_t*) malloc((n / 2 + 1) * sizeof(uint32_t));
    d->nMapLenArr = 0;
    if (!d->p.frugal64bitLenEncoding) {
        memcpy(d->mapLen, s, n / 4 * 4);
        d->nMapLenArr = n / 4;
        return;
    }
    uint64_t at = 0;
    while (at + 2 <= n) { /* utils/helper.h:256-272, readUInt64Frugal<uint32_t> */
        uint16_t w;
        memcpy(&w, s + at, 2);
        at += 2;
        uint32_t v = w;
        if (w == UINT16_MAX) {
            uint32_t y;
            memcpy(&y, s + at, 4);
            at += 4;
            v = y;
            if (y == UINT32_MAX) {
                uint64_t z;
                memcpy(&z, s + at, 8);
                at += 8;
                v = (uint32_t) z;
            }
        }
        d->mapLen[d->nMapLenArr++] = v;
    }
}

mbo_decoder* mbo_dec_create(const mbo_pack_params_t* p, uint64_t refTotalLength, int rcInReference,
                            int contigsIndividuallyReversed, const mbo_dec_streams_t* s) {
    mbo_decoder* d = (mbo_decoder*) calloc(1, sizeof(*d));
    d->p = *p;
    d->rcInReference = rcInReference;
    d->contigsIndividuallyReversed = contigsIndividuallyReversed;
    d->refTotalLength = refTotalLength;
    d->ref = (uint8_t*) calloc(refTotalLength + 256, 1); /* decodeInit :1083 */
    d->s = *s;
    decode_map_len(d);
    d->ext = (uint8_t*) malloc(MAX_EXTEND_MATCH_LEFT_LENGTH);
    return d;
}

void mbo_dec_destroy(mbo_decoder* d) {
    if (!d) return;
    free(d->ref);
    free(d->mapLen);
    free(d->ext);
    free(d);
}

const uint8_t* mbo_dec_ref(const mbo_decoder* d) { return d->ref; }
uint64_t mbo_dec_ref_pos(const mbo_decoder* d) { return d->refPos; }
int mbo_dec_reached_count(const mbo_decoder* d) { return d->reachedRefLengthCount; }
void mbo_dec_cursors(const mbo_decoder* d, uint64_t out6[6]) {
    out6[0] = d->litPos; out6[1] = d->offPos / 4; out6[2] = d->off5Pos; out6[3] = d->lenIdx; out6[4] = d->gdPos; out6[5] = d->flagsPos;
}

/* MBGC_Decoder.cpp:651-675 */
void mbo_dec_load_ref(mbo_decoder* d, const uint8_t* seqText, uint64_t seqLength, uint64_t refLockPos, int loadRCRef) {
    while (seqLength != 0) {
        if (d->refPos == d->refTotalLength && refLockPos != d->refTotalLength) {
            d->reachedRefLengthCount++;
            d->refPos = REF_SHIFT;
        }
        uint64_t tmpLength = seqLength;
        const uint64_t tmpMax = refLockPos < d->refPos ? d->refTotalLength : refLockPos;
        if (d->refPos + tmpLength > tmpMax) tmpLength = tmpMax - d->refPos;
        if (loadRCRef)
            mbo_upper_reverse_complement(seqText + seqLength - tmpLength, tmpLength, d->ref + d->refPos);
        else
            memcpy(d->ref + d->refPos, seqText, tmpLength);
        d->refPos += tmpLength;
        seqText += loadRCRef ? 0 : tmpLength;
        seqLength = d->refPos == refLockPos ? 0 : seqLength - tmpLength;
        if (d->p.lazyDecompressionSupport && d->refPos == refLockPos) d->ref[refLockPos - 1] = REF_REGION_SEPARATOR;
    }
}

/* MBGC_Decoder.cpp:243-262 (initReference: -t 1 archives, G0 is a target too; only its first contig seeds the reference) */
void mbo_dec_init_reference(mbo_decoder* d) {
    d->refG0InitPos = 0;
    d->ref[d->refPos] = 0;
    d->refPos = REF_SHIFT;
    if (d->s.nLiterals == 0) return;
    const uint64_t tmp = find_byte(d->s.literals, d->s.nLiterals, SEQ_SEPARATOR_MARK, d->litPos);
    const uint64_t n = tmp - d->litPos;
    memcpy(d->ref + d->refPos, d->s.literals + d->litPos, n);
    d->refPos += n;
    d->refG0InitPos += n;
    d->litPos = tmp + 1;
    if (d->rcInReference) {
        mbo_upper_reverse_complement(d->ref + 1, d->refG0InitPos, d->ref + 1 + d->refG0InitPos);
        d->refPos += d->refG0InitPos;
    }
}

/* MBGC_Decoder.cpp:264-317 (decodeReference: G0 is not a target; all its contigs come verbatim from the literal stream) */
uint64_t mbo_dec_decode_reference(mbo_decoder* d, uint32_t seqCount, uint8_t* out, uint64_t* contigLens) {
    uint64_t total = 0;
    d->refG0InitPos = 0;
    d->ref[d->refPos] = 0;
    d->refPos = REF_SHIFT;
    for (uint32_t i = 0; i < seqCount; i++) {
        const uint64_t tmp = find_byte(d->s.literals, d->s.nLiterals, SEQ_SEPARATOR_MARK, d->litPos);
        const uint64_t n = tmp - d->litPos;
        memcpy(out + total, d->s.literals + d->litPos, n);
        memcpy(d->ref + d->refPos, d->s.literals + d->litPos, n);
        contigLens[i] = n;
        total += n;
        d->refPos += n;
        d->refG0InitPos += n;
        d->litPos = tmp + 1;
    }
    if (d->rcInReference) {
        mbo_upper_reverse_complement(d->ref + 1, d->refG0InitPos, d->ref + 1 + d->refG0InitPos);
        d->refPos += d->refG0InitPos;
    }
    return total;
}

static int score_step(const mbo_pack_params_t* p, int score, int mismatch) {
    if (mismatch) return score + p->mmsMismatchPenalty;
    score -= p->mmsMatchBonus;
    return score < 0 ? 0 : score;
}

/* MBGC_Decoder.cpp:434-462 */
static uint64_t dec_extend_right(mbo_decoder* d, uint8_t* dest, uint64_t* destSize, int64_t offsetDelta, int isGap, int gapStart,
                                 int gapMiddle, int gapEnd, uint64_t guardLitPos) {
    if (d->litPos == guardLitPos && !gapMiddle) return 0;
    const uint64_t destStart = *destSize;
    int64_t src = (int64_t) *destSize + offsetDelta;
    if (gapStart || !isGap)
        dest[(*destSize)++] = mbo_code2mismatch(d->ref[src], d->s.literals[d->litPos++]);
    else
        src--;
    int score = d->p.mmsMismatchesInitialScore;
    while ((!gapEnd || d->litPos != guardLitPos) && (isGap || score < d->p.mmsMismatchesScoreThreshold)) {
        const int mismatch = d->s.gapFlags[d->flagsPos++] != 0;
        if (mismatch && d->litPos == guardLitPos) break;
        score = score_step(&d->p, score, mismatch);
        ++src;
        dest[(*destSize)++] = mismatch ? mbo_code2mismatch(d->ref[src], d->s.literals[d->litPos++]) : d->ref[src];
    }
    return *destSize - destStart;
}

/* MBGC_Decoder.cpp:464-527; the extension is written backwards from extEnd */
static uint64_t dec_extend_left(mbo_decoder* d, uint8_t* extEnd, uint64_t* matchSrcPos, int skipOffset, uint64_t refLockPos,
                                uint64_t markPos) {
    uint8_t* ptr = extEnd;
    int64_t srcMatch = (int64_t) *matchSrcPos;
    int64_t srcGuard = srcMatch - MAX_EXTEND_MATCH_LEFT_LENGTH;
    if (!skipOffset) {
        if (srcGuard < REF_SHIFT) srcGuard = REF_SHIFT;
        const int64_t srcLock = (int64_t) refLockPos;
        if (srcGuard < srcLock && srcLock <= srcMatch) srcGuard = srcLock;
    }
    if (srcGuard == srcMatch) return 0;
    if (skipOffset) {
        int64_t srcPtr = srcMatch - 1;
        uint64_t length = 0, mismatches = 0;
        int isExtLenKnown = 1;
        int score = d->p.mmsMismatchesInitialScore;
        while (--srcPtr > srcGuard && score < d->p.mmsMismatchesScoreThreshold) {
            const int mismatch = d->s.gapFlags[d->flagsPos + length++] != 0;
            if (mismatch && d->litPos + ++mismatches == markPos) {
                isExtLenKnown = 0;
                break;
            }
            score = score_step(&d->p, score, mismatch);
        }
        if (srcPtr == srcGuard && isExtLenKnown) {
            mismatches++;
            length++;
        }
        const uint64_t matchingChars = length - mismatches;
        *matchSrcPos += matchingChars;
        srcGuard += (int64_t) matchingChars;
        srcMatch += (int64_t) matchingChars;
    }
    int64_t srcPtr = srcMatch - 1;
    *(--ptr) = mbo_code2mismatch(d->ref[srcPtr], d->s.literals[d->litPos++]);
    int score = d->p.mmsMismatchesInitialScore;
    while (--srcPtr >= srcGuard && score < d->p.mmsMismatchesScoreThreshold) {
        const int mismatch = d->s.gapFlags[d->flagsPos++] != 0;
        if (mismatch && d->litPos == markPos) break;
        score = score_step(&d->p, score, mismatch);
        *(--ptr) = mismatch ? mbo_code2mismatch(d->ref[srcPtr], d->s.literals[d->litPos++]) : d->ref[srcPtr];
    }
    return (uint64_t) (extEnd - ptr);
}

/* MBGC_Decoder.cpp:319-432.  dest must hold the decoded contig; returns its length, *unmatched = unmatchedChars. */
uint64_t mbo_dec_decode_sequence(mbo_decoder* d, uint64_t refLockPos, uint8_t* dest, uint64_t* unmatchedOut) {
    const mbo_pack_params_t* p = &d->p;
    const uint8_t* lit = d->s.literals;
    const uint64_t nLit = d->s.nLiterals;
    const uint64_t seqEnd = find_byte(lit, nLit, SEQ_SEPARATOR_MARK, d->litPos);
    uint8_t* extEnd = p->enableExtensionsWithMismatches ? d->ext + MAX_EXTEND_MATCH_LEFT_LENGTH : NULL;
    uint64_t destSize = 0;
    uint32_t unmatched = 0;
    int64_t pairedGapOffsetDelta[MAX_GAP_DEPTH];
    for (int i = 0; i < MAX_GAP_DEPTH; i++) pairedGapOffsetDelta[i] = INT64_MAX;
    const int64_t NO_GAP = -1;
    int64_t gapStartIdx = NO_GAP, gapEndIdx = NO_GAP;
    int gapCurIdx = 0;
    uint64_t matchSrcPos = 0, prevMatchDestPos = 0;
    int64_t offsetDelta = -1;
    uint64_t extLeftLen = 0, extRightLen = 0;
    int isGap = 0;
    int64_t j = 0;
    uint64_t markPos = find_byte(lit, nLit, MATCH_MARK, d->litPos);
    while (markPos != NPOS && markPos < seqEnd) {
        const uint64_t literalsLeft = markPos - d->litPos;
        matchSrcPos = 0;
        const int skipOffset = pairedGapOffsetDelta[gapCurIdx] != INT64_MAX;
        if (skipOffset) {
            matchSrcPos = (uint64_t) (pairedGapOffsetDelta[gapCurIdx] + (int64_t) destSize + (int64_t) literalsLeft);
            pairedGapOffsetDelta[gapCurIdx] = INT64_MAX;
        } else {
            uint32_t v;
            memcpy(&v, d->s.mapOff + d->offPos, 4);
            d->offPos += 4;
            matchSrcPos = v;
            if (d->refTotalLength > UINT32_MAX) matchSrcPos += ((uint64_t) d->s.mapOff5[d->off5Pos++]) << 32;
        }
        extLeftLen = 0;
        if (p->enableExtensionsWithMismatches) {
            if (!isGap && literalsLeft) extLeftLen = dec_extend_left(d, extEnd, &matchSrcPos, skipOffset, refLockPos, markPos);
            if (gapEndIdx == j) {
                gapStartIdx = NO_GAP;
                gapEndIdx = NO_GAP;
            }
        }
        const uint64_t literalLen = markPos - d->litPos + extLeftLen + extRightLen;
        memcpy(dest + destSize, lit + d->litPos, markPos - d->litPos);
        destSize += markPos - d->litPos;
        if (extLeftLen) memcpy(dest + destSize, extEnd - extLeftLen, extLeftLen);
        destSize += extLeftLen;
        unmatched += (uint32_t) literalLen;
        d->litPos = markPos + 1;

        const uint32_t matchLength = d->mapLen[d->lenIdx++];
        prevMatchDestPos = destSize;
        memcpy(dest + destSize, d->ref + matchSrcPos, matchLength);
        destSize += matchLength;

        markPos = find_byte(lit, nLit, MATCH_MARK, d->litPos);
        uint8_t gapDelta = 0;
        if (p->gapDepthOffsetEncoding && markPos < seqEnd && (gapDelta = d->s.gapDelta[d->gdPos++])) {
            int gapIdx = gapCurIdx;
            int g = gapDelta;
            if (!p->lazyDecompressionSupport && gapStartIdx == NO_GAP && markPos - d->litPos == 0) {
                gapIdx = (gapIdx + 1) % MAX_GAP_DEPTH;
                g++;
            }
            while (gapDelta) {
                gapIdx = (gapIdx + 1) % MAX_GAP_DEPTH;
                if (pairedGapOffsetDelta[gapIdx] == INT64_MAX)
                    gapDelta--;
                else
                    g++;
            }
            pairedGapOffsetDelta[gapIdx] = (int64_t) matchSrcPos - (int64_t) prevMatchDestPos;
            if (p->enableExtensionsWithMismatches && gapEndIdx <= j + g && g <= p->gapDepthMismatchesEncoding) {
                gapStartIdx = j;
                gapEndIdx = j + g;
            }
        }
        gapCurIdx = (gapCurIdx + 1) % MAX_GAP_DEPTH;

        const int gapStart = gapStartIdx == j;
        const int gapEnd = gapEndIdx == j + 1;
        const int gapMiddle = gapStartIdx < j && j + 1 < gapEndIdx;
        isGap = gapStart || gapMiddle || gapEnd;
        extRightLen = 0;
        if (p->enableExtensionsWithMismatches) {
            if (!isGap || gapStart) offsetDelta = (int64_t) matchSrcPos + matchLength - (int64_t) destSize;
            extRightLen = dec_extend_right(d, dest, &destSize, offsetDelta, isGap, gapStart, gapMiddle, gapEnd,
                                           markPos < seqEnd ? markPos : seqEnd);
        }
        j++;
    }
    const uint64_t literalLen = seqEnd - d->litPos + extRightLen;
    memcpy(dest + destSize, lit + d->litPos, seqEnd - d->litPos);
    destSize += seqEnd - d->litPos;
    unmatched += (uint32_t) literalLen;
    d->litPos = seqEnd + 1;
    *unmatchedOut = unmatched;
    return destSize;
}

/* MBGC_Decoder.cpp:532-649 (decodeTarget), the sequences only: per contig decode, extension decision
 * (matching/MGMP_Params.h:179-190) and reference loads; after the target the whole-target reverse complement (when contigs are
 * not reversed one by one), the region separator, and the flag stream's file separator.  out receives the contigs back to
 * back; contigLens / contigUnmatched one entry per contig.  Returns the bytes written to out. */
uint64_t mbo_dec_decode_target(mbo_decoder* d, uint32_t seqCount, int unmatchedFractionFactor, int unmatchedFractionRCFactor,
                               uint64_t refLockPos, uint8_t* out, uint64_t* contigLens, uint64_t* contigUnmatched) {
    uint64_t total = 0;
    const uint64_t startPos = d->refPos;
    for (uint32_t i = 0; i < seqCount; i++) {
        uint64_t unmatched;
        uint8_t* seq = out + total;
        const uint64_t n = mbo_dec_decode_sequence(d, refLockPos, seq, &unmatched);
        contigLens[i] = n;
        contigUnmatched[i] = unmatched;
        total += n;
        const int load = unmatched * (uint64_t) unmatchedFractionFactor > n;
        mbo_dec_load_ref(d, seq, load ? n : 0, refLockPos, 0);
        if (d->rcInReference && d->contigsIndividuallyReversed && unmatched * (uint64_t) unmatchedFractionRCFactor > n)
            mbo_dec_load_ref(d, seq, load ? n : 0, refLockPos, 1);
    }
    if (d->rcInReference && !d->contigsIndividuallyReversed) {
        if (d->refPos >= startPos) {
            /* the source is the reference itself: loadRef never overlaps source and destination here */
            mbo_dec_load_ref(d, d->ref + startPos, d->refPos - startPos, refLockPos, 1);
        } else {
            const uint64_t tmpStartPos = startPos;
            mbo_dec_load_ref(d, d->ref + 1, d->refPos - 1, refLockPos, 1);
            mbo_dec_load_ref(d, d->ref + tmpStartPos, d->refTotalLength - tmpStartPos, refLockPos, 1);
        }
    }
    if (d->p.lazyDecompressionSupport) {
        const uint8_t sep = REF_REGION_SEPARATOR;
        mbo_dec_load_ref(d, &sep, 1, refLockPos, 0);
    }
    if (d->p.enableExtensionsWithMismatches) d->flagsPos++;
    return total;
}
