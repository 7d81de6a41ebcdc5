// mbgc_b200/csrc/hostpack.hpp -- host side of the text transfer: 2-bit packing of target bases on the host cores.
//
// Why on the host: once matching costs well under a millisecond per 500 MB, the end-to-end time of a batch that
// starts in host memory is the PCIe copy (~53 GB/s measured on the B200 box, 9.5 ms per 500 MB).  A/C/G/T text
// packs 4:1, so the host cores pack while the copy engine moves bytes: a share of the blocks travels raw (DMA
// straight from the caller's pinned buffer, no CPU involved), the rest is packed by a pool of worker threads into
// a pinned staging buffer and expanded back to the exact ASCII bytes by a small kernel on the device
// (kernels.cu: k_unpack_blocks, k_patch_exceptions).  The matcher itself always sees raw ASCII: the reference
// hashes and compares raw bytes (case, N, IUPAC -- utils/Hashes.h:28-40, SlidingWindowSparseEMMatcher.cpp:298), so
// every byte that is not one of "ACGT" travels as an exception run (offset, length, byte) and is patched back;
// a block with too many of them (soft-masked or protein text) simply travels raw.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace mbgc_b200 {

struct ExcRun {          // a run of identical non-ACGT bytes, device layout (16 bytes)
    uint64_t off;        // byte offset in the device text buffer
    uint32_t len;        // number of bytes
    uint32_t byte;       // bits 0-7 the byte; bits 8-31 the distance between two of them (0: they are adjacent) -- the line
                         // ends of a FASTA file are ONE run (pack_fasta_image)
};

// packs n bytes of src (4 bases per byte, first base in the lowest bits: code = (c >> 1) & 3, i.e. A 0, C 1, T 2,
// G 3) into dst[(n + 3) / 4]; appends exception runs (offsets relative to devOff).  Returns false -- dst contents
// then undefined -- as soon as more than maxRuns runs would be needed.
bool pack_block(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t devOff, std::vector<ExcRun>& runs, size_t maxRuns);

// A whole FASTA file image, packed like pack_block does.  Line ends come at a fixed distance in a well-formed file: they are
// recognised as an arithmetic progression and leave as one run (first offset, count, distance) instead of one run per line;
// header lines and everything else that is not A/C/G/T go the ordinary way.  The image a device rebuilds from (dst, runs) is
// byte-identical to src whatever it contains -- nothing is assumed about the format, only exploited.
bool pack_fasta_image(const uint8_t* src, uint64_t n, uint8_t* dst, std::vector<ExcRun>& runs, size_t maxRuns);
// the same for an image that arrives in pieces (a reader's buffer that stays in the cache): every piece but the last a
// multiple of 64 bytes long
class ImagePacker {
public:
    ImagePacker(uint8_t* dst, std::vector<ExcRun>& runs, size_t maxRuns);
    ~ImagePacker();
    ImagePacker(const ImagePacker&) = delete;
    ImagePacker& operator=(const ImagePacker&) = delete;
    bool add(const uint8_t* src, uint64_t n);   // false: too many runs
    bool finish();
private:
    uint8_t* dst;
    std::vector<ExcRun>& runs;
    size_t maxRuns;
    void* state;
};

// dst[n - 1 - i] = lut[src[i]]: the reverse complement the reference computes with upperReverseComplement
// (utils/helper.cpp:405-410) through its 256-entry table (:312-358).  AVX2 where the CPU has it and the table has the
// shape of that LUT (identity except letters, both cases of a letter mapping to one upper-case letter, 0x7F); plain loop
// otherwise.  src and dst must not overlap.
void reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst, const uint8_t lut[256]);

// Minimal persistent worker pool: run(nItems, fn) calls fn(item, worker) for every item, items handed out in
// index order; returns immediately, wait() blocks until all items are done.  done(i) tells whether item i is.
class WorkerPool {
public:
    explicit WorkerPool(int nThreads);
    ~WorkerPool();
    int size() const { return static_cast<int>(threads.size()); }
    void run(uint32_t nItems, std::function<void(uint32_t, int)> fn);
    bool done(uint32_t item) const { return doneFlags[item].load(std::memory_order_acquire) != 0; }
    void wait();

private:
    void loop(int worker);
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cvStart, cvDone;
    std::function<void(uint32_t, int)> job;
    std::vector<std::atomic<uint8_t>> doneFlags;
    std::atomic<uint32_t> next{0};
    uint32_t nItems = 0;
    uint64_t generation = 0;
    int active = 0;
    bool stop = false;
};

}  // namespace mbgc_b200
