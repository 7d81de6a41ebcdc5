// mbgc_b200/csrc/decode.h -- host-visible launchers of decode.cu (row f4 of SURVEY.md 8: the decoder's match-copy path,
// /root/reference/mbgccoder/MBGC_Decoder.cpp:243-675).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbgc_b200 {

constexpr uint32_t DC_COUNT_BLOCK = 4096;   // bytes of a stream one CTA of the marker passes looks at
constexpr int DC_TILE_SHIFT_MIN = 13, DC_TILE_SHIFT_MAX = 16;   // output bytes one CTA of the copy pass writes: 8 .. 64 KiB
constexpr int DC_LEN_CHUNK = 64;            // 16-bit words of the frugal length stream per chunk map
constexpr int DC_LEN_BLOCK = 256;           // chunk maps per block map

struct DcParams {  // the decoder's view of MBGC_Params (mbgccoder/MBGC_Params.h:66-97)
    int enableExt, gapDepthOffset, gapDepthMismatches, lazy;
    int initialScore, scoreThreshold, mismatchPenalty, matchBonus;
    uint64_t refTotalLength;
};

struct DcStreams {  // the six streams on the device (every buffer padded by 64 readable bytes)
    const uint8_t* lit;
    uint64_t nLit;
    const uint8_t* off;
    uint64_t nOff;     // entries (u32 each)
    const uint8_t* off5;
    uint64_t nOff5;
    const uint32_t* len;  // expanded lengths
    uint64_t nLen;
    const uint8_t* gd;
    uint64_t nGd;
    const uint8_t* flags;
    uint64_t nFlags;
};

struct DcIndex {  // where the contigs / targets start in every stream
    const uint32_t* sepPos;     // literal position of the SEQ_SEPARATOR_MARK that ends contig c
    const uint32_t* sepMarks;   // MATCH_MARKs in front of that separator
    const uint32_t* markPos;    // literal positions of all MATCH_MARKs
    const uint32_t* gdStart;    // gap-delta bytes consumed before contig c
    const uint32_t* nzStart;    // non-zero gap deltas (= matches without an offset) before contig c
    const uint32_t* flagSep;    // flag-stream positions of the FILE_SEPARATOR_MARKs
    uint32_t nContigs, nFlagSep;
};

struct DcTarget {  // one unit of the parse: a target file, or the G0 prologue (flagIdx = UINT32_MAX, nothing read from the flags)
    uint32_t firstContig, nContigs;
    uint32_t flagIdx;  // the target's flags start behind separator flagIdx - 1 (0: at the stream's start)
    uint32_t pad;
    uint64_t lockPos;
};

// what the copy pass does for a stretch of one contig
enum : uint32_t { DC_LITERAL = 0, DC_MATCH = 1, DC_EXT_RIGHT = 2, DC_EXT_LEFT = 3 };
struct DcSeg {
    uint64_t dst;        // contig-relative
    uint64_t src;        // reference position (match, extensions) or literal position
    uint32_t len;
    uint32_t flagsPos;   // extensions: first flag byte
    uint32_t litPos;     // extensions: first mismatch code
    uint32_t meta;       // kind (bits 0-1), first byte is an implicit mismatch (bit 2), contig index (bits 4-31)
};

void dc_upload_tables(cudaStream_t st);
int launch_dc_count(const uint8_t* src, uint64_t n, int b0, int b1, uint32_t* cnt0, uint32_t* cnt1, cudaStream_t st);
int launch_dc_scatter(const uint8_t* src, uint64_t n, int b0, int b1, const uint32_t* off0, const uint32_t* off1, uint32_t* pos0,
                      uint32_t* rank1of0, uint32_t* pos1, cudaStream_t st);
int launch_dc_contig_gd(uint32_t nContigs, const uint32_t* sepMarks, int gapDepthOffset, uint32_t* gdCount, cudaStream_t st);
int launch_dc_contig_nz(uint32_t nContigs, const uint32_t* gdStart, const uint8_t* gd, uint64_t nGd, uint32_t* nzCount, cudaStream_t st);
// frugal u16/u32/u64 words -> u32 lengths (utils/helper.h:256-272); returns launches, *total (device) = entries
uint64_t dc_len_chunks(uint64_t nBytes);
uint64_t dc_len_blocks(uint64_t nBytes);
int launch_dc_len_expand(const uint8_t* src, uint64_t nBytes, uint32_t* chunkMap, uint32_t* blockMap, uint64_t* blockEntry,
                         uint64_t* total, uint32_t* out, uint64_t outCap, cudaStream_t st);
int launch_dc_parse(const DcParams& p, const DcStreams& s, const DcIndex& ix, const DcTarget* targets, uint32_t nTargets, DcSeg* segs,
                    uint64_t* contigLen, uint64_t* contigUnmatched, uint32_t* error, cudaStream_t st);
int launch_dc_tiles(const DcSeg* segs, uint64_t segBegin, uint64_t segEnd, const uint64_t* contigBase, uint64_t dstBegin, int tileShift,
                    uint32_t* tileSeg, cudaStream_t st);
int launch_dc_copy(const DcSeg* segs, uint64_t segEnd, const uint64_t* contigBase, const uint32_t* tileSeg, int tileShift, uint64_t dstBegin,
                   uint64_t dstEnd, const uint8_t* ref, const DcStreams& s, uint8_t* dna, cudaStream_t st);

}  // namespace mbgc_b200
