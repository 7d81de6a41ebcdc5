// mbgc_b200/csrc/rcmatch.h -- host-visible launchers of rcmatch.cu (row f3 of SURVEY.md 8: the reverse-complement
// redundancy pass of mode 3, SimpleSequenceMatcher::rcMatchSequence over CopMEMMatcher).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbgc_b200 {

constexpr int RC_MAX_PER_BUCKET = 13;        // HASH_COLLISIONS_PER_POSITION_LIMIT + 1 (CopMEMMatcher.h:11, cpp:155)
constexpr uint32_t RC_EXT_CAP = 4096;        // bytes the device follows a match per side before handing it over as open

struct RcSeed {          // a real seed (K-mer equal) of one query position
    uint32_t pos;        // query position (in the reverse-complemented sequence)
    uint32_t src;        // reference position (in the sequence)
    uint32_t left;       // equal bytes left of the seed the reference's left loop accepts
    uint32_t right;      // equal bytes right of the K-mer
    uint16_t order;      // place of the candidate in its bucket
    uint16_t flags;      // 1 left open, 2 right open (cut at RC_EXT_CAP), 4 seed sits at byte 0 of a text (left loop does not run)
};

void rc_upload_lut(const uint8_t lut[256]);
int launch_rc_revcomp(const uint8_t* src, uint64_t n, uint8_t* dst, cudaStream_t st);
int launch_cm_count(const uint8_t* seq, uint64_t nSamples, int K, int k1, uint32_t mask, uint32_t* bucketOf, uint32_t* count, cudaStream_t st);
int launch_cm_fill(uint64_t nSamples, int k1, const uint32_t* bucketOf, const uint32_t* start, uint32_t* cursor, uint32_t* entries, cudaStream_t st);
int launch_cm_sort(uint32_t hashSize, const uint32_t* start, uint32_t* entries, uint32_t* bigList, uint32_t* bigCount, uint32_t bigCap, cudaStream_t st);
int launch_cm_query(const uint8_t* seq, const uint8_t* q, uint64_t N, int K, int k2, uint32_t mask, const uint32_t* start,
                    const uint32_t* entries, RcSeed* out, uint32_t* outCount, uint32_t outCap, cudaStream_t st);
// exclusive scan of count[0 .. n) into out[0 .. n] (kernels.cu: k_scan_lookback); state = zeroed scratch of (n / 4096 + 8) words
int launch_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* state, cudaStream_t st);

}  // namespace mbgc_b200
