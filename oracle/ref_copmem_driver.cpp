// oracle/ref_copmem_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the REFERENCE's own reverse-complement redundancy pass, compiled from where it lies:
// SimpleSequenceMatcher::rcMatchSequence (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176) over CopMEMMatcher
// (matching/copmem/CopMEMMatcher.cpp), what mode 3 runs on the literal stream (mbgccoder/MBGC_Encoder.cpp:637-638), and its
// inverse restoreRCMatchedSequence (:178-211).  Row f3 of SURVEY.md 8.  The oracle restatement (oracle/mbgc_oracle.c:
// mbo_rc_match_sequence) and the CUDA path are pinned against it.  PgHelpers::numberOfThreads = 1 selects the
// single-threaded index build (the multi-threaded one races, like everything else outside `-t 1`).
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>

#include "matching/SimpleSequenceMatcher.h"
#include "utils/helper.h"

extern "C" void* A_memcpy(void* dest, const void* src, size_t n) { return memcpy(dest, src, n); }
extern "C" int A_memcmp(const void* s1, const void* s2, size_t n) { return memcmp(s1, s2, n); }

namespace {
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
} g_nullBuf;
std::ostream g_null(&g_nullBuf);
void quiet(int threads) {
    PgHelpers::devout = &g_null;
    PgHelpers::numberOfThreads = threads;
}
}  // namespace

extern "C" {

// seq is rewritten in place; returns its new length.  mapOff / mapLen receive the two streams (caller: >= 4 and >= 10 bytes
// per possible match + 16).
uint64_t refr_rc_match_sequence(uint8_t* seq, uint64_t n, uint64_t targetMatchLength, uint8_t* mapOff, uint64_t* mapOffLen,
                                uint8_t* mapLen, uint64_t* mapLenLen, int threads) {
    quiet(threads);
    std::string s(reinterpret_cast<const char*>(seq), n), off, len;
    PgTools::SimpleSequenceMatcher::rcMatchSequence(s, off, len, targetMatchLength);
    memcpy(seq, s.data(), s.size());
    memcpy(mapOff, off.data(), off.size());
    memcpy(mapLen, len.data(), len.size());
    *mapOffLen = off.size();
    *mapLenLen = len.size();
    return s.size();
}

// restoreRCMatchedSequence: seq (n bytes, with marks) -> the original (orgLen bytes) in out
uint64_t refr_rc_restore(const uint8_t* seq, uint64_t n, const uint8_t* mapOff, uint64_t mapOffLen, const uint8_t* mapLen,
                         uint64_t mapLenLen, uint64_t orgLen, uint8_t* out) {
    quiet(1);
    std::string s(reinterpret_cast<const char*>(seq), n), off(reinterpret_cast<const char*>(mapOff), mapOffLen),
        len(reinterpret_cast<const char*>(mapLen), mapLenLen);
    PgTools::SimpleSequenceMatcher::restoreRCMatchedSequence(s, off, len, orgLen);
    memcpy(out, s.data(), s.size());
    return s.size();
}

}  // extern "C"
