// mbgc_b200/csrc/collection.hpp -- the batched, multi-GPU driver of the hot path: replicas of one matcher on several
// devices, FASTA ingest of many files per launch set, and the reference's sequential (-t 1) target schedule run over
// speculative windows.
//
// What it replaces (paths relative to /root/reference):
//   matching/MultipleGenomeMatchingProcessor.cpp:232-313   processTargetsWithParallelIO: per contig matchTexts ->
//                                                          processMatches -> extension decision -> loadRef; per file the
//                                                          meta data and the hook below
//   mbgccoder/MBGC_Encoder.cpp:493-509                     processAfterTargetWithParallelIO: separator, refExtSize, lock
//   matching/MGMP_Params.h:179-196                         isContigProperForRefExtension / ...RCExtension
//
// The schedule is the one SURVEY.md 8e derives ("optimistic batching"): a contig that does not extend the reference
// leaves index and reference untouched, and the one-byte separators between files only move the loading cursor over
// bytes that are zero anyway.  So a window of contigs (across file boundaries, spread over all replicas) is ingested,
// matched and packed against the frozen state; results are committed in target order up to and including the first
// contig that extends the reference; the extension is applied to every replica (device-to-device copy of the resident
// text, then the same ordered index inserts everywhere) and the rest of the window runs again.  Committed output is
// byte-identical to the one-contig-at-a-time schedule.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "matcher.hpp"

namespace mbgc_b200 {

struct HostImage {         // one FASTA file as it is on disk (after gunzip), host memory
    const char* bytes;
    uint64_t n;            // size of the file image
    bool packed = false;   // bytes -> a PackedImageHeader: the image travels 4 bases per byte and is rebuilt on the device
};

// A file image packed by the host for the trip to the device (pack_image_into): header, the 2-bit codes, the exception runs.
constexpr uint64_t PACKED_IMAGE_MAGIC = 0x4942324743474D42ull;  // "MBGCG2BI"
struct PackedImageHeader {  // 64 bytes
    uint64_t magic;
    uint64_t n;             // bytes of the file image
    uint64_t packedOff, packedBytes;   // from the header's start; both multiples of 16
    uint64_t runsOff;
    uint32_t nRuns;
    uint32_t firstByte;     // image[0] (the reader's "is this FASTA" test looks at it)
    uint64_t total;         // bytes of the whole packed image
    uint64_t reserved;
};
uint64_t packed_image_bound(uint64_t n);
// returns the bytes used in dst (16-byte aligned memory of at least packed_image_bound(n) bytes), or 0 when the image has more
// exceptional bytes than fit (it then travels as it is)
uint64_t pack_image_into(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap);
// the same for a file on disk, read in cache-sized pieces; *fileSize = its size (also when 0 is returned)
uint64_t pack_image_from_file(const char* path, uint8_t* dst, uint64_t cap, uint64_t* fileSize);

struct IngestedFile {
    int status = -1;               // what ended the reader's loop: -1 end of file, -3 not FASTA, -4 DNA not well-formed
    uint64_t dnaLineLen = 0;       // seq->dnaLineLen after the last record (UINT64_MAX = DNA_NOT_WELLFORMED)
    uint64_t maxLastLineLen = 0;
    uint32_t firstContig = 0, nContigs = 0;
};

struct IngestedContig {
    uint32_t file;                 // index into the batch
    uint64_t hdrOff, hdrLen;       // header line inside the file image (without '>' and the line end)
    uint64_t textOff, seqLen;      // the sequence inside dTexts()
};

// One ingest slot: the images of a batch of files -> device, record splitting / validation / line joining for all of
// them in one set of launches (ingest.cu), the sequences left on the device in the batch-text layout of the match stage.
// Owns its stream, so that a slot can be filled while the matcher's stream works on another one.
//   begin();  add(image) ...;  finish();      add() returns when the image is on the device (its memory is free again)
class FastaBatchIngest {
public:
    explicit FastaBatchIngest(int device);
    ~FastaBatchIngest();
    FastaBatchIngest(const FastaBatchIngest&) = delete;
    FastaBatchIngest& operator=(const FastaBatchIngest&) = delete;

    void begin(uint64_t expectedBytes);
    // Pinned images are copied by DMA straight from the caller's memory, pageable ones through this slot's pinned bounce
    // buffers.  Does not wait for the copy: the image's memory is free again once reap() has handed its pointer back.
    void add(const HostImage& image);
    // images whose trip to the device is over (all of them after wait = true); call after add() and before finish() returns
    void reap(std::vector<const char*>& released, bool wait);
    // blocking; afterwards files / contigs / headers / dTexts() describe the batch
    void finish();
    void run(const HostImage* images, int nFiles) {
        uint64_t b = 0;
        for (int i = 0; i < nFiles; i++) b += images[i].n;
        begin(b);
        std::vector<const char*> done;
        for (int i = 0; i < nFiles; i++) add(images[i]);
        reap(done, true);
        finish();
    }
    const uint8_t* dTexts() const { return dText; }
    std::vector<IngestedFile> files;
    std::vector<IngestedContig> contigs;
    std::vector<uint64_t> fileBytes;   // image size per file
    const char* header(size_t contig) const { return reinterpret_cast<const char*>(hHdr) + hdrAt[contig]; }
    uint64_t launches = 0;         // kernel launches of the last batch
    uint64_t imageBytes = 0, textBytes = 0, pageableBytes = 0, wireBytes = 0;
    float ms[2] = {0, 0};          // last batch: kernels (incl. the two host round trips) of finish(); begin -> end of finish
    int deviceIndex() const { return device; }

private:
    void check(cudaError_t e, const char* what);
    void grow(uint8_t** p, uint64_t* cap, uint64_t need, const char* what, bool keep = false, uint64_t keepBytes = 0);
    int device;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {};
    cudaEvent_t bounceEv[2] = {};
    struct InFlight { cudaEvent_t ev; const char* bytes; };
    std::deque<InFlight> inflight;       // images on their way, oldest first (ev == nullptr: already there)
    std::vector<cudaEvent_t> evPool;
    uint8_t* hBounce[2] = {nullptr, nullptr};
    int bounceUse = 0;
    uint8_t* dIn = nullptr; uint64_t dInCap = 0;
    uint8_t* dPk = nullptr; uint64_t dPkCap = 0;     // a packed image on its way to dIn
    uint8_t* dWork = nullptr; uint64_t dWorkCap = 0;
    uint8_t* dText = nullptr; uint64_t dTextCap = 0;
    uint8_t* dHdr = nullptr; uint64_t dHdrCap = 0;
    uint8_t* hWork = nullptr; uint64_t hWorkCap = 0;   // pinned
    uint8_t* hHdr = nullptr; uint64_t hHdrCap = 0;     // pinned: the header lines of the batch
    std::vector<uint64_t> hdrAt;
    std::vector<int> live;             // files that travelled
    std::vector<uint64_t> spans;       // start, end per live image
    uint64_t cursor = 0;               // next image offset (behind the 16 leading bytes)
};

// Pinned host memory in large slabs, handed out in pieces (file images).  cudaHostAlloc costs ~0.3 ms per MiB, far too
// much per file; a slab is recycled once every piece cut from it has been released.
class PinnedArena {
public:
    PinnedArena(uint64_t slabBytes, int maxSlabs) : slabBytes(slabBytes), maxSlabs(maxSlabs) {}
    ~PinnedArena();
    void* alloc(uint64_t n);       // nullptr when the arena is full (the caller falls back to pageable memory)
    bool release(const void* p);   // false if p is not from this arena
    bool owns(const void* p);
    uint64_t maxPiece() const { return slabBytes; }
private:
    struct Slab { uint8_t* base; uint64_t used; int live; bool open; };
    std::mutex mu;
    std::vector<Slab> slabs;
    uint64_t slabBytes;
    int maxSlabs;
};

// The replicas of one matcher: the same reference, index and cursors on every device.  State-changing calls go to all
// of them (every replica does the identical ordered inserts, so no table contents ever travel); getters read replica 0.
class ReplicaSet {
public:
    ReplicaSet(uint64_t refLengthLimit, uint32_t targetMatchLength, int k1, int k2, int skipMargin, const int* devices, int ndev);
    ~ReplicaSet();
    int size() const { return static_cast<int>(rep.size()); }
    B200SparseEMMatcher* primary() const { return rep[0]; }
    B200SparseEMMatcher* at(int r) const { return rep[r]; }

    void loadRef(const char* text, uint64_t len, bool rc, bool addSep, char sep);
    // the text is resident on replica `owner` at dText: peers fetch it device-to-device (NVLink), nothing crosses PCIe
    void loadRefFromReplica(int owner, const void* dText, uint64_t len, bool rc, bool addSep, char sep);
    // the same for a text that is the concatenation of several pieces resident on `owner` (a whole target's extension)
    void loadRefConcatFromReplica(int owner, const B200SparseEMMatcher::StagePiece* pieces, size_t n, bool rc, bool addSep, char sep);
    void loadSeparator(char sep);
    uint64_t acquireLock();
    void releaseLock(uint64_t v);
    void setHostMirror(bool on);
    // replica 0's host-visible state -> all others (after mbgcg_set_state / setPosition on the primary)
    void syncStateFromPrimary();

    // run f(replica index) on the worker thread of every replica in `which` (bit mask) and wait; exceptions are rethrown
    void parallel(uint64_t which, const std::function<void(int)>& f);
    PinnedArena& arena();
    // the ingest slots (streams, device and pinned buffers) outlive a pass: the next pass on this handle allocates nothing
    std::vector<std::unique_ptr<FastaBatchIngest>>& ingestSlots() { return slotStore; }

private:
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::function<void()> job;
        bool busy = false, quit = false;
        std::exception_ptr err;
    };
    void workerLoop(int r);
    std::vector<B200SparseEMMatcher*> rep;
    std::vector<std::unique_ptr<Worker>> workers;
    std::unique_ptr<PinnedArena> pinned;
    std::mutex arenaMu;
    std::vector<std::unique_ptr<FastaBatchIngest>> slotStore;
};

struct SeqConfig {                 // what the sequential schedule reads from MGMP_Params / MBGC_Params
    PackParams pack{};
    uint32_t minMatchLength = 32;  // params->k
    int unmatchedFractionFactor = 128;      // currentUnmatchedFractionFactor (MGMP_Params.h:41)
    int unmatchedFractionRCFactor = 8;      // unmatchedFractionRCFactor (:42)
    bool rcDisabled = false;                // isRCinReferenceDisabled()
    bool refRegionSeparators = true;        // params->refRegionSeparators
    char separator = 0;                     // MGMP_Params::REF_REGION_SEPARATOR
    bool separatorPerFile = true;           // lazyDecompressionSupport: loadSeparator + refExtLoadedPosArr per file
    // The batch-frozen schedule of the reference's PARALLEL mode instead (processTargetsParallel / processTarget /
    // finalizeParallelProcessingOfTarget, MultipleGenomeMatchingProcessor.cpp:340-468, 520-555): every target (file) of a
    // window is matched under its own sliding-window lock against the state at the window's start; a target's qualifying
    // contigs (each followed by its reverse complement when contigsIndividuallyReversed) extend the reference when the
    // target is finalized, in target order.  Any such schedule is legal (the decoder replays the reference growth from
    // the lock positions); this one is deterministic.
    bool parallelSchedule = false;
    bool contigsIndividuallyReversed = true;          // MGMP_Params.h:75
    int allowedTargetsOutrun = 64;                    // matcherWorkingThreads * allowedTargetsOutrunFactor: targets per window
    int allowedTargetsOutrunForDissimilarContigs = 1; // MGMP_Params.h:78
    uint64_t minimalLengthForDissimilarContigs = 0;   // MGMP_Params.h:79
    int unmatchedFractionFactorTweakForDissimilarContigs = 16;  // MGMP_Params.h:80
    uint64_t chunkBytes = 64ull << 20;      // file images per ingest slot
    uint32_t maxWindowContigs = 1u << 20;   // speculation window (contigs); follows the commit rate below this
};

struct SeqEvent {
    enum Type { CONTIG = 1, FILE_END = 2, NEED_INPUT = 3, DONE = 4 };
    int type = DONE;
    int status = -1;
    uint64_t file = 0;
    const char* hdr = nullptr; uint64_t hdrLen = 0;
    uint64_t seqLen = 0, nMatches = 0, kept = 0, unmatched = 0;
    int extended = 0, extendedRC = 0;
    uint64_t contigLock = 0;       // parallel schedule: the lock position the contig was matched under
    uint64_t window = 0;
    const uint8_t* seg[PACK_STREAMS] = {}; uint64_t segLen[PACK_STREAMS] = {};
    uint64_t nContigs = 0, dnaLineLen = 0, refExtSize = 0, lockPos = 0, loadedRefLength = 0;
};

struct SeqStats {
    uint64_t files = 0, contigs = 0, windows = 0, runs = 0, rerunContigs = 0, extensions = 0;
    uint64_t imageBytes = 0, textBytes = 0, launches = 0, ingestLaunches = 0, chunks = 0, pageableBytes = 0;
    double ingestMs = 0, matchMs = 0, packMs = 0, loadRefMs = 0, waitStageMs = 0;
    uint64_t wireBytes = 0;
};

// One pass of the sequential schedule over a collection.  Pull interface: push file images in target order, call next()
// until DONE; next() returns NEED_INPUT when it cannot go on without another file (or finishInput()).
class SequentialPass {
public:
    SequentialPass(ReplicaSet& rs, const SeqConfig& cfg, const uint64_t* loadedPosArr, uint64_t nLoaded);
    ~SequentialPass();
    // image memory: from the set's pinned arena -> released by the pass as soon as it is on the device; anything else
    // must stay valid until the file's FILE_END event has been returned
    void pushFile(const char* bytes, uint64_t n, bool packed = false);
    void finishInput();
    bool wantsInput();
    void next(SeqEvent& ev);       // pointers in ev stay valid until the next call
    const SeqStats& stats() const { return st; }
    const SeqConfig& config() const { return cfg; }

private:
    struct Chunk {
        int replica = 0, slot = 0;
        uint64_t firstFile = 0, nFiles = 0, bytes = 0;
        std::vector<uint64_t> locks;   // parallel schedule: the sliding-window lock of every file
        bool closed = false;       // no more files will join (FINISH is queued)
        bool ready = false;        // ingested
        std::exception_ptr err;
    };
    struct StageMsg {
        enum Type { BEGIN, ADD, FINISH } type;
        std::shared_ptr<Chunk> chunk;
        HostImage image;
    };
    struct Item {                  // one step of the committed order
        bool fileEnd;
        int chunk;                 // index into win
        uint32_t contig;           // index into the chunk's ingest contigs (or the file index for fileEnd)
        bool valid = false;
        uint32_t resIndex = 0;     // position in its replica's last run
    };
    void dispatch();
    void stagerLoop(int q);
    bool startWindow();
    void runFrom(size_t from);
    void releaseWindow();

    ReplicaSet& rs;
    SeqConfig cfg;
    SeqStats st;
    std::vector<uint64_t> loaded;              // refExtLoadedPosArr as the schedule maintains it
    uint64_t fileStartLoaded = 0;              // getLoadedRefLength() when the current file began
    bool fileOpen = false;
    uint64_t windowFirstTarget = 0;            // parallel schedule: processedTargetsCount when the current results were matched
    struct ExtPiece { const uint8_t* dText; uint64_t len; bool rc; };
    std::vector<ExtPiece> fileExt;             // parallel schedule: the open target's reference extension (device pieces)
    void loadRefPieces(int owner, const std::vector<ExtPiece>& pieces, bool rc);
    bool litExtOn = false;                     // MBGC_Encoder::processLiteral switched on for this pass (row a12)

    // input
    std::mutex mu;
    std::condition_variable cv;
    std::deque<HostImage> pending;
    std::shared_ptr<Chunk> forming;
    uint64_t pendingBytes = 0, stagedBytes = 0, pushedFiles = 0, dispatchedFiles = 0;
    bool inputDone = false;
    uint64_t chunkSeq = 0;

    // staging: NSLOT ingest slots per replica, one stager thread per slot
    std::vector<std::unique_ptr<FastaBatchIngest>>& slots;  // [replica * NSLOT + slot], owned by the replica set
    std::vector<bool> slotFree;
    std::vector<std::deque<StageMsg>> stageQ;                // per ingest slot
    std::vector<std::thread> stagers;
    std::deque<std::shared_ptr<Chunk>> staged;               // dispatched chunks in file order
    bool quit = false;

    // current window
    std::vector<std::shared_ptr<Chunk>> win;
    std::vector<Item> items;
    size_t cursor = 0;
    bool winActive = false;
    uint32_t curWindow = 0;
    uint32_t committedThisRun = 0;
    std::deque<SeqEvent> evQ;
    // per replica: the contigs of its last run (texts for packBatch) and where each landed
    struct RunState { std::vector<TextView> texts; std::vector<uint64_t> offs, locks; };
    std::vector<RunState> runState;
};

}  // namespace mbgc_b200
