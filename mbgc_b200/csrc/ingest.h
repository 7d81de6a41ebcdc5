// mbgc_b200/csrc/ingest.h -- host-visible launchers of the FASTA-ingest kernels (ingest.cu).
//
// Replaces the record splitting / line joining / line-length detection of the reference's FASTA reader,
// kseq_read_lossless_fasta (/root/reference/utils/kseq.h:233-274; stream helpers :61-72, :100-149), for a whole file
// image at once: the file travels to the device as it is on disk, the sequences come out as contiguous texts in the
// batch layout the match stage reads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mbgc_b200 {

constexpr int FA_THREADS = 256;
constexpr uint32_t FA_CHUNK = 16;                       // bytes per load / store
constexpr uint32_t FA_PER_THREAD = 4;                   // chunks per thread, FA_THREADS * FA_CHUNK apart (coalesced)
constexpr uint32_t FA_BLOCK_BYTES = FA_THREADS * FA_CHUNK * FA_PER_THREAD;  // 16 KiB of the file per CTA

struct FaRecord {          // per record, filled by k_fa_records / k_fa_check (device) and read by the host
    uint64_t start;        // position of '>'
    uint64_t hdrEnd;       // position of the '\n' that ends the header line (or n)
    uint64_t regionEnd;    // start of the next record (or n)
    uint64_t firstLine;    // length of the first sequence line
    uint32_t endsWithNl;   // the region's last byte is '\n'
    uint32_t bad;          // an empty line, or a line that does not follow the first line's length
};

struct FaGather {          // per accepted record, written by the host
    uint64_t seqStart;     // first sequence byte in the input
    uint64_t outOff;       // byte offset of the text in the output buffer (multiple of 16)
    uint64_t seqLen;
    uint64_t lineLen;      // bytes per full line (0 when seqLen == 0)
    uint64_t chunkStart;   // first 16-byte output chunk of this record
};

// record starts ('>' at a line start): per-block counts, then the positions in file order.  starts must hold the
// total; returns kernel launches.
int launch_fa_count(const uint8_t* in, uint64_t n, uint32_t* blockCnt, cudaStream_t st);   // + scan in place, total at [nBlocks]
int launch_fa_starts(const uint8_t* in, uint64_t n, const uint32_t* blockOff, uint64_t* starts, cudaStream_t st);
// fileEnds != nullptr: `in` holds several file images (batched ingest); fileEnds[f] = end offset of image f, ascending
int launch_fa_records(const uint8_t* in, uint64_t n, const uint64_t* starts, uint32_t nRecs, FaRecord* recs, cudaStream_t st,
                      const uint64_t* fileEnds = nullptr, uint32_t nFiles = 0);
// filler bytes around the images of a batched ingest (spans[2f], spans[2f+1] = start / end offset of image f; the buffer
// has 16 bytes in front of spans[0] and extends to bufEnd)
int launch_fa_filler(uint8_t* in, const uint64_t* spans, uint32_t nFiles, uint64_t bufEnd, cudaStream_t st);
// hdr[3r], hdr[3r+1], hdr[3r+2] = offset of header r in `in`, its length, its offset in `out`
int launch_fa_headers(const uint8_t* in, const uint64_t* hdr, uint32_t nRecs, uint8_t* out, cudaStream_t st);
int launch_fa_gather(const uint8_t* in, const FaGather* g, uint32_t nRecs, uint64_t nChunks, uint8_t* out, cudaStream_t st);

}  // namespace mbgc_b200
