// mbgc_b200/csrc/kernels.h -- host-visible launchers of the sm_100a kernels (kernels.cu).
//
// Kernel set ("New kernel" table) and the reference loop each replaces
// (paths relative to /root/reference):
//   K2 index_insert      matching/SlidingWindowSparseEMMatcher.cpp:146-171 (processIgnoreCollisionsRef)
//   K3 probe             .cpp:200-226 + the K-mer verify of :298, for EVERY target position
//   K4 extend            .cpp:227-246 (right) and the raw left reach of :257-259/:278-280
//   K5 resolve           .cpp:250-315 (overlap resolution, accept, skip) as a sequential-equivalent walk
// plus plumbing kernels (exclusive scan, seed compaction, run detection, match gather).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

namespace mbgc_b200 {

constexpr int PROBE_NPT = 16;                       // positions per lane
constexpr int PROBE_CHUNK = 32 * PROBE_NPT;         // positions per warp-chunk (512)
constexpr int PROBE_THREADS_FILTER = 1024;          // 1 CTA per SM next to the shared-memory filter
constexpr int PROBE_THREADS_PLAIN = 512;
constexpr uint32_t PROBE_FILTER_BYTES = 192 * 1024; // first-level filter, copied to shared memory by every CTA
constexpr uint32_t PROBE_FILTER_BITS = PROBE_FILTER_BYTES * 8;
constexpr uint32_t ANCHOR_STRIDE = 8;               // fast path: one diagonal proposal per 8 chunks (and per contig start)
constexpr uint64_t NO_LOCK = UINT64_MAX;
constexpr uint32_t LINK_WINDOW = 12;                // K4 looks this many runs back / ahead for the same diagonal (< 16)

struct RunRec {          // one seed run = consecutive seeds on one diagonal, spacing k1 (24 bytes)
    uint32_t firstPos;   // target position of the first seed
    uint32_t nSeeds;
    uint32_t enc0;       // table entry (encoded reference position) of the first seed
    uint32_t aFree;      // left end of the maximal equal interval (valid if !leftLinked)
    uint32_t bFree;      // right end (exclusive) of the maximal equal interval (valid if !rightLinked)
    uint32_t flags;      // bit0 leftLinked (aFree = that of the previous run), bit1 rightLinked
};

struct alignas(16) SeedRec {  // one seed segment (16 bytes, read and written as one 128-bit word)
    uint32_t pos;        // first target position of the segment
    uint32_t enc;        // table entry of its first seed
    uint32_t n;          // seeds in the segment (spacing k1, entries ascending by k1 >> k1ord)
    uint32_t contig;     // contig (text) of the segment
};

struct MatchRec {        // 16 bytes, device-internal
    uint64_t src;
    uint32_t dst;
    uint32_t len;
};

struct TextMatchRec {    // = PgTools::TextMatch (matching/TextMatchers.h:9-17), what the host receives
    uint64_t posSrcText, length, posDestText, nextSrcRegionLoadingPos;
};

struct IndexState {      // frozen matcher state a match pass runs against
    const uint8_t* ref;  // device reference buffer (ref[0] is position 0)
    const unsigned long long* ht;  // 64-bit slots: (insert sequence number << 32) | encoded position
    const uint32_t* occ; // occupancy bitmap, 1 bit per slot
    const uint32_t* filter;  // first-level filter: bit mulhi(hash, filterBits) set for every inserted k-mer
    uint32_t filterBits;
    uint32_t filterWords;
    uint32_t hashMask;
    int K;               // hashed prefix length (bytes), multiple of 4
    int k1;              // reference sampling step
    int k1ord;           // >= 0: Exp variant (entry = pos >> k1ord); -1: plain
    int skipMargin;
    uint64_t pos1;       // loader cursor
    uint64_t refLen;     // getRefLength()
};

struct BatchDesc {       // device pointers describing the texts of one pass
    const uint8_t* text;
    const uint64_t* contigOff;   // [nContigs]
    const uint32_t* contigLen;   // [nContigs]
    const uint64_t* contigLock;  // [nContigs] or nullptr
    const uint32_t* tileStart;   // [nContigs+1] first chunk of each contig
    int nContigs;
    uint32_t nTiles;             // number of 512-position chunks
};

struct WorkBuffers {     // device scratch of one pass
    uint32_t* tileCount;     // [nChunks+1] seed segments per chunk -> exclusive scan in place (chunk offsets)
    uint32_t* chunkContig;   // [nChunks]
    unsigned long long* anchor;  // [nChunks] first verified seed among the first 16 positions: (pos in chunk << 32) | entry
    uint8_t* chunkSegN;      // [nChunks] 0x80 | pin = chunk with clean lanes (pin = offset of the grid position in a lane), 0 = none
    uint2* chunkSeg;         // [nChunks] clean chunks: (table entry of group 0's seed, mask of 16-position groups with a seed)
    uint32_t* workList;      // [nChunks] chunks with at least one dirty lane (fast path only)
    uint32_t* dirtyMask;     // [nChunks] lanes (16-position groups) of the chunk that need the full probe
    uint32_t* laneList;      // [nChunks*32] dirty lanes as items chunk * 32 + lane
    uint4* chunkDesc;        // [nChunks] text offset lo/hi, contig-local base position, nValid
    uint16_t* seedMask;      // [nChunks*32] per-lane 16-bit seed masks
    uint32_t* scratchEnc;    // [nChunks*PROBE_CHUNK] position-indexed table entries of the seeds
    SeedRec* seeds;          // [seedCap] seed segments, position-ordered per contig
    uint32_t* headScan;      // [seedCap+2] run-head flags -> scan; later the island-head flag of every run
    uint32_t* runFirst;      // [seedCap+2] first segment of every run; later matches per island -> scan
    RunRec* runs;            // [seedCap]
    uint8_t* runLink;        // [seedCap] distance to the linked run on the same diagonal: low nibble back, high nibble ahead
    uint32_t* runStart;      // [nContigs+1]
    MatchRec* matches;       // [seedCap] per-island regions at the island's first run
    unsigned long long* islandFwd;  // [seedCap/1024+2] per-tile carries of the island scans
    unsigned long long* islandBwd;
    uint32_t* matchOff;      // [nContigs+1]
    TextMatchRec* matchesDense;  // [seedCap] the pass's matches in text order, host layout
    uint32_t* blockSums;     // scan temporaries
    unsigned long long* scanState;  // look-back state of the three scans of a pass (scanStride words each), zeroed with the counters
    uint32_t scanStride;
    uint32_t zeroBytes;      // counters + scanState: one memset at the start of a pass
    uint32_t* counters;      // [16]: 0 nSegments, 1 nRuns, 2 nMatches, 3 overflow flag, 4 longest island, 5 islands, 6 chunks with dirty lanes, 7 dirty lanes (one 64-bit word), 8 long islands
    uint32_t seedCap;
};

// ---- index build ----
struct InsertRange {
    int64_t begin;       // first sampled reference position
    int64_t count;       // number of samples (step k1)
    uint64_t seqBase;    // global insert sequence number of the first sample
};
void launch_index_insert(const uint8_t* ref, unsigned long long* ht, uint32_t* occ, uint32_t* filter, uint32_t filterBits,
                         uint32_t hashMask, int K, int k1, int k1ord, InsertRange r0, InsertRange r1, cudaStream_t st);
void launch_table_rebase(unsigned long long* ht, uint64_t n, cudaStream_t st);
void launch_table_low32(const unsigned long long* ht, uint32_t* out, uint64_t n, cudaStream_t st);
// reference segment from the staged text of a loadRef call: copy or upper-casing reverse complement (table uploaded once
// per process with upload_complement_table)
void upload_complement_table(const uint8_t table[256]);
void launch_ref_segment(const uint8_t* src, uint8_t* dst, uint64_t n, bool reverseComplement, cudaStream_t st);

// ---- match pass ----
// returns the number of kernel launches issued
int launch_chunk_desc(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st);
// full probe of every chunk (useWorkList = false) or of the 16-position lanes the fast path could not prove clean
int launch_probe(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, bool useFilter, bool useWorkList,
                 cudaStream_t st);
// fast path for near-identical texts (needs the self-probe memo of the reference, see kernels.cu):
// anchors (first 16 positions of every chunk) + clean-chunk test; fills the work list for launch_probe
struct MemoState {
    const uint32_t* memoX;   // 1 bit per 16 reference positions: 1 = probing them gives something unusual
    const uint32_t* memoE;   // 1 = probing them gives no seed at all (the sampled k-mer lost its table slot)
    uint64_t memoLen;        // reference positions [0, memoLen) are covered
};
int launch_anchor(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st);
int launch_clean(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, const MemoState& memo, cudaStream_t st);
// memo from a finished self-probe pass (the reference buffer probed as a text, window test disabled)
int launch_memo(const BatchDesc& b, const WorkBuffers& w, uint32_t nValid, uint32_t* memoX, uint32_t* memoE,
                cudaStream_t st);
int launch_compact_and_runs(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, bool segments,
                            cudaStream_t st);
int launch_extend(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, cudaStream_t st);
int launch_resolve(const IndexState& ix, const BatchDesc& b, const WorkBuffers& w, uint32_t minMatchLength,
                   cudaStream_t st);

// ---- text transfer (host-packed blocks) ----
struct PackedBlock {     // one block of a text that travelled 4 bases per byte
    uint64_t dstOff;     // byte offset in the device text buffer (multiple of 16)
    uint64_t srcOff;     // byte offset in the packed staging buffer (multiple of 16)
    uint32_t n;          // bases
    uint32_t pad;
};
struct ExcRunDev {       // = mbgc_b200::ExcRun (hostpack.hpp): a run of one non-ACGT byte
    uint64_t off;
    uint32_t len;
    uint32_t byte;
};
int launch_unpack_image(const uint8_t* packed, uint64_t n, const ExcRunDev* runs, uint32_t nRuns, uint8_t* text, cudaStream_t st);
int launch_unpack(const PackedBlock* blocks, uint32_t nBlocks, const uint8_t* packed, const ExcRunDev* runs, uint32_t nRuns,
                  uint8_t* text, cudaStream_t st);

int launch_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, unsigned long long* state, cudaStream_t st);
bool probe_supports_K(int K);
int device_sm_count();

#ifdef __CUDACC__
// Programmatic dependent launch (sm_90+): the kernels of a match pass form a chain of ~19 dependent launches, most of them a
// few microseconds long.  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, the next kernel's CTAs may take
// their places on the SMs as soon as every CTA of the current one has called griddep_launch_dependents() -- the first
// statement of every pass kernel -- and then sleep in griddep_wait() until the current grid has completed and its writes
// are visible: the launch latency of each link (~2 us) overlaps its predecessor.  Everything a kernel does before
// griddep_wait() must be independent of the kernels in front of it in the pass (only k_probe does something there: it
// starts the bulk copy of the first-level filter, which no pass kernel writes).  Without the launch attribute both
// instructions are no-ops.  MBGC_B200_PDL=0 switches the attribute off.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args&&... args) {
    const char* e = getenv("MBGC_B200_PDL");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    if (!(e && e[0] == '0')) {
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__

}  // namespace mbgc_b200
