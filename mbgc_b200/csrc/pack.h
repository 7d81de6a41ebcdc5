// mbgc_b200/csrc/pack.h -- host-visible launchers of the stream-packing kernels (pack.cu): K6.
//
// Replaces MBGC_Encoder::processMatches with extendMatchLeft / extendMatchRight and
// ContextAwareMismatchesCoder::mismatch2code (/root/reference/mbgccoder/MBGC_Encoder.cpp:143-427,
// coders/ContextAwareMismatchesCoder.cpp:22-27): the match lists of a batch of texts become the six byte
// streams the reference hands to its back-end coders (literals, mapOff, mapOff5thByte, mapLen, gapDeltas,
// gapMismatchesFlags), bit-exact with the single-threaded reference.  See pack.cu for how the reference's
// sequential state machine is turned into parallel kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace mbgc_b200 {

constexpr int PACK_STREAMS = 6;
enum { PK_LITERALS = 0, PK_MAPOFF = 1, PK_MAPOFF5 = 2, PK_MAPLEN = 3, PK_GAPDELTA = 4, PK_GAPFLAGS = 5 };
constexpr uint32_t PK_NONE = 0xFFFFFFFFu;
constexpr uint32_t PK_LONG_RUN = 4096;   // plain literal runs of this many bytes go to the long-run copy kernel
constexpr int PK_THREADS = 256;

struct PackParams {            // the MBGC_Params members processMatches reads (mbgccoder/MBGC_Params.h:66-97)
    int enableExt;             // enableExtensionsWithMismatches
    int gapDepthOffset;        // gapDepthOffsetEncoding (<= 127 here)
    int gapDepthMismatches;    // gapDepthMismatchesEncoding (<= 2 here)
    int lazy;                  // lazyDecompressionSupport
    int frugalLen;             // frugal64bitLenEncoding
    int enable40bit;           // enable40bitReference
    int withExclusion;         // mismatchesWithExclusion
    int initialScore, scoreThreshold, mismatchPenalty, matchBonus;  // mms* (MBGC_Params.h:90-97)
    uint64_t gapBreakingMinLen;  // gapBreakingMatchMinLength
};

struct KeptRec {               // one match that survived pass 1, left shift applied (16 bytes)
    uint64_t src;
    uint32_t dst;
    uint32_t len;
};

struct LitRange;
struct LongRun {               // a long plain-literal run, copied by k_pk_copy_long
    const uint8_t* src;
    uint8_t* dst;
    uint64_t n;
};

struct PackState {             // matcher state the extension guards read (MBGC_Encoder.cpp:318-330,376-380)
    const uint8_t* ref;        // getRef()
    uint64_t loadingPos;       // getLoadingPosition()
    uint64_t maxRefLength;     // getMaxRefLength()
    uint64_t refLength;        // getRefLength()
    const uint64_t* loadedArr; // refExtLoadedPosArr (device copy), may be null when nLoaded == 0
    uint32_t nLoaded;
};

struct PackBatch {
    const uint8_t* text;             // device text buffer
    const uint64_t* contigOff;       // [nContigs] byte offset of every text
    const uint64_t* contigLock;      // [nContigs] matchingLockPos per text, or null (all SIZE_MAX)
    const uint32_t* contigLen;       // [nContigs]
    int nContigs;
    const TextMatchRec* matches;     // [nMatches] all lists, text order (the match stage's dense output layout)
    const uint32_t* matchOff;        // [nContigs+1]
    uint32_t nMatches;
};

struct PackWork {
    uint32_t* mContig;         // [nM] text of every match
    uint8_t* removed;          // [nM] pass 1: gap-breaking match dropped
    uint32_t* keptScan;        // [nM+1] exclusive scan of !removed
    KeptRec* kept;             // [nM]
    uint32_t* kContig;         // [nM]
    uint32_t* kDiag;           // [nM] low 32 bits of posSrcText - posDestText (quick reject in the pairing scan)
    uint32_t* kOff;            // [nC+1] first kept match of every text
    unsigned long long* sumLen;// [nC] sum of kept lengths (unmatched = destLen - sum)
    uint8_t* claim;            // [nM] g of the later match this one pairs with (0 = none)
    uint8_t* delta;            // [nM] the gapDelta byte
    uint32_t* claimedBy;       // [nM] earliest match that claims this one (PK_NONE = its offset is written)
    unsigned long long* nl;    // [nM] nextSrcRegionLoadingPos (lazy mode)
    uint32_t* gapCnt;          // [nGaps*6] bytes every gap appends to every stream
    uint32_t* blockTot;        // [(blocks+1)*6] per-block totals -> exclusive block offsets
    uint32_t* streamOff;       // [6*(nC+1)] start of every text in every stream
    uint32_t* flags;           // [64] 0..15 removal sweeps changed, 16..31 claim sweeps changed, 32 long-run count
    LongRun* longRuns;         // [longCap]
    uint32_t longCap;
    uint8_t* out[PACK_STREAMS];
    // MBGC_Encoder::processLiteral (row a12; all null when the option is off)
    uint2* litSpan;            // [nGaps] (pos, length) of the plain literal run every gap appends (k_pk_emit)
    uint32_t* litFlag;         // [nGaps+1] 1 = the run extends the reference; exclusive scan in place
    LitRange* litRanges;       // [nGaps] ranges in append order
    unsigned long long* litScanState;  // scan scratch, zeroed with the pack state
};

struct LitRange {              // one processLiteral append: text[start, start + len) of text `text`
    uint32_t text, start, len, pad;
};

struct LitExtParams {          // MGMP_Params.h:68-69,175-187
    uint64_t minLen;           // refLiteralMinimalLengthExt (UINT64_MAX = off, the reference's default)
    uint64_t beforeAfter;      // refLiteralBeforeAfterExt
    uint64_t factor;           // unmatchedFractionFactor of the contig-is-proper rule
};

// Each returns the number of kernel launches issued.
// pass 1 (MBGC_Encoder.cpp:153-199): `sweeps` fixed-point sweeps of the removal rule, then compaction.
int launch_pack_pass1(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, int sweeps,
                      cudaStream_t st);
// pairing of pass 2 (MBGC_Encoder.cpp:229-273): `sweeps` fixed-point sweeps over the kept matches (nKeptMax = nM).
int launch_pack_claims(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, int sweeps,
                       cudaStream_t st);
// extensions + stream sizes, scan, emission (MBGC_Encoder.cpp:214-287, 310-427)
int launch_pack_count(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, cudaStream_t st);
int launch_pack_emit(const PackParams& p, const PackState& s, const PackBatch& b, const PackWork& w, cudaStream_t st);
// processLiteral (MBGC_Encoder.cpp:128-135, call sites :228 / :290) over the literal runs k_pk_emit recorded: flags, scan,
// ranges; the number of ranges ends up in w.litFlag[nGaps]
int launch_pack_literal_ext(const LitExtParams& le, const PackBatch& b, const PackWork& w, cudaStream_t st);

}  // namespace mbgc_b200
