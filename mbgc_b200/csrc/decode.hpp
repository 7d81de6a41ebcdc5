// mbgc_b200/csrc/decode.hpp -- the decoder's match-copy path (row f4 of SURVEY.md 8), host interface (decode.cpp)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "decode.h"
#include "matcher.hpp"

namespace mbgc_b200 {

struct DecodeStreamsHost {  // literals, mapOff, mapOff5, mapLen, gapDelta, gapFlags (host memory)
    const uint8_t* p[6];
    uint64_t n[6];
};

struct DecodeLayout {  // what the archive's header and side streams say (MBGC_Decoder.cpp:1291-1322, :1126-1133)
    uint64_t refTotalLength;
    int g0IsTarget;                 // initReference (-t 1 archives) / decodeReference
    uint32_t g0SeqCount;            // contigs of G0 (decodeReference only)
    int rcInReference, contigsIndividuallyReversed, frugalLen;
    uint32_t nTargets;
    const uint32_t* seqCounts;      // contigs per target
    const uint8_t* factors;         // unmatchedFractionFactor, unmatchedFractionRCFactor per target
    const uint64_t* lockPos;        // matchingLocksPosArr, or nullptr (= refTotalLength)
};

struct DevBuf {
    void* p = nullptr;
    uint64_t n = 0, cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf();
    void reset(uint64_t bytes);
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

class Decoder {
   public:
    explicit Decoder(B200SparseEMMatcher* matcher);
    ~Decoder();
    // decodes every target; the DNA and the decoder's reference stay on the device until the next call
    void run(const DcParams& p, const DecodeStreamsHost& hs, const DecodeLayout& L);
    void fetch(uint32_t firstContig, uint32_t n, uint8_t* dst);   // contigs [first, first + n) back to back -> host
    void fetchRef(uint64_t pos, uint64_t n, uint8_t* dst);

    std::vector<uint64_t> contigLen, contigUnmatched, contigBase;  // every contig of the parse (a G0 prologue first)
    uint32_t firstOutputContig = 0;  // contigs in front of it are the -t 1 prologue (G0's first contig, not an output)
    uint64_t dnaLen = 0, refPos = 0, refTotal = 0, launches = 0;
    int reachedCount = 0;
    uint32_t batches = 0;
    float ms[3] = {0, 0, 0};  // markers + parse, copy pass
    const uint8_t* deviceDna() const { return dDna.as<uint8_t>(); }

   private:
    uint8_t* upload(DevBuf& b, const uint8_t* src, uint64_t n, cudaStream_t st);
    int scan(const uint32_t* cnt, uint32_t* out, uint32_t n, cudaStream_t st);
    B200SparseEMMatcher* m;
    DevBuf dLit, dOff, dOff5, dLenSrc, dLen, dGd, dFlags, dScanState, dCnt0, dCnt1, dOff0, dOff1, dSepPos, dSepMarks, dMarkPos, dFlagSep,
        dGdStart, dNzStart, dLenChunk, dLenBlock, dLenEntry, dTargets, dSegs, dContig, dErr, dDna, dRef, dBase, dTileSeg;
};

}  // namespace mbgc_b200
