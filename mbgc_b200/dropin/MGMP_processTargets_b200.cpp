// mbgc_b200/dropin/MGMP_processTargets_b200.cpp
//
// Link-time substitutes for the two target loops of the reference driver (both private members of
// MultipleGenomeMatchingProcessor, /root/reference/matching/MultipleGenomeMatchingProcessor.cpp):
//
//   processTargetsWithParallelIO  :232-313   the sequential schedule (`mbgc c -t 1`, `-Q`, `-m 3`)
//   processTargetsParallel        :520-555   the parallel schedule (default), with processTarget :340-429 and
//                                            finalizeParallelProcessingOfTarget :433-468 behind it
//
// Compiled against the UNMODIFIED reference headers; the rest of MultipleGenomeMatchingProcessor.cpp is the reference's
// own object file in which these two symbols have been made weak (`objcopy --weaken-symbol`, oracle/Makefile target
// `dropin_batched`), so performMatching (:568-575) calls the definitions below.
//
// The stock loops handle one contig at a time on a host thread: read a file (kseq), matchTexts, processMatches, decide
// about the reference extension, loadRef.  Here both loops are the collection pass of the C ABI (mbgcg_seq_*,
// include/mbgc_b200.h): reader threads bring whole file images into pinned memory, the pass ingests, matches and packs
// windows of contigs on all devices of the handle and hands back the committed contigs in target order; these functions
// only apply the encoder's hooks to them, in the order the stock loops would have.
//
// Sequential schedule: archives are byte-identical with the stock `mbgc -t 1` (tests/test_gpu_dropin.py).
// Parallel schedule: the stock one is a race by design (no two runs agree); the pass runs a deterministic
// member of the same family -- windows of targets matched against the state at the window's start, finalized in target
// order -- and the archives decode with the stock decoder.
// Environment: MBGC_B200_DEVICES = comma-separated CUDA devices (default "0"), MBGC_B200_READERS = reader threads.
#include "MBGC_Encoder.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "../matching/input_with_libdeflate_wrapper.h"
#include "dropin_shared.h"
#include "mbgc_b200.h"

extern mgmpInFile firstFp;            // MultipleGenomeMatchingProcessor.cpp:51-52
extern bool firstFpIsFirstTarget;

namespace {

[[noreturn]] void die(const char* msg) {
    fprintf(stderr, "%s\n\n", msg);
    exit(EXIT_FAILURE);
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Image {
    const char* bytes = nullptr;
    size_t n = 0;
    bool mallocd = false;     // ours to free() after FILE_END (arena images belong to the pass once pushed)
    bool ready = false;
    bool packed = false;      // bytes = a packed image (mbgcg_image_pack_file): 4 bases per byte cross PCIe, n stays the file's size
};

// Reader threads: file i -> image i, pinned when the arena has room.  Pinned memory goes out in file order: a file whose
// turn it is waits for room (earlier files are on their way to the device and give theirs back), later files wait for
// their turn -- so the arena can be small without ever blocking the file the pass is waiting for.
class ImageReader {
public:
    // names[i] = file of image i; preset[i].bytes != nullptr: the image is already in memory (not ours to free)
    ImageReader(mbgcg_t* h, const std::vector<const char*>& names, const std::vector<Image>& preset, int nReaders, int readAhead)
        : h(h), names(names), images(preset), turnDone(names.size(), 0), arenaMax(mbgcg_image_max(h)), readAhead(readAhead) {
        const int n = static_cast<int>(names.size());
        for (int i = 0; i < n; i++)
            if (images[i].bytes || images[i].ready) {
                images[i].ready = true;
                turnDone[i] = 1;
            }
        while (allocTurn < n && turnDone[allocTurn]) allocTurn++;
        for (int t = 0; t < nReaders; t++) threads.emplace_back([this] { loop(); });
    }
    ~ImageReader() {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
            cv.notify_all();
        }
        for (auto& t : threads) t.join();
    }
    // image i if it is there (block: wait for it); the time spent waiting is added to *waitedMs
    bool get(int i, bool block, Image& out, double* waitedMs) {
        std::unique_lock<std::mutex> lk(mu);
        if (!images[i].ready) {
            if (!block) return false;
            const double w0 = now_ms();
            cv.wait(lk, [&] { return images[i].ready; });
            *waitedMs += now_ms() - w0;
        }
        out = images[i];
        return true;
    }
    size_t size(int i) {
        std::lock_guard<std::mutex> g(mu);
        return images[i].n;
    }
    void consumedUpTo(int i) {  // FILE_END of file i seen: its memory goes back, the read-ahead window moves on
        Image im;
        {
            std::lock_guard<std::mutex> g(mu);
            im = images[i];
            images[i].bytes = nullptr;
            consumed = std::max(consumed, i + 1);
            cv.notify_all();
        }
        if (im.mallocd) free(const_cast<char*>(im.bytes));
    }
    int threadCount() const { return static_cast<int>(threads.size()); }

private:
    void passTurn(int i) {  // (mu held)
        turnDone[i] = 1;
        while (allocTurn < static_cast<int>(names.size()) && turnDone[allocTurn]) allocTurn++;
        cv.notify_all();
    }
    void loop() {
        const int n = static_cast<int>(names.size());
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) return;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (images[i].ready) continue;
                cv.wait(lk, [&] { return stop || i < consumed + readAhead; });
                if (stop) return;
            }
            readOne(i);
        }
    }
    void readOne(int i) {
        Image im;
        bool turnTaken = false;
        const char* name = names[i];
        int fd = iterableDecoder ? -1 : open(name, O_RDONLY);   // repack / append read through the decoder's iterator
        struct stat sb;
        bool done = false;
        if (fd >= 0 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 2) {
            const size_t n = static_cast<size_t>(sb.st_size);
            char* buf = nullptr;
            // plain (not gzip) files are packed 4 bases per byte on their way into pinned memory: the page cache is read once,
            // a quarter is written, a quarter crosses PCIe
            uint8_t magic[2] = {0, 0};
            const bool pack = packImages && pread(fd, magic, 2, 0) == 2 && !(magic[0] == GZIP_ID1 && magic[1] == GZIP_ID2);
            const uint64_t want = pack ? mbgcg_image_pack_bound(n) : n;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || allocTurn == i; });
            }
            if (want <= arenaMax)
                for (int spin = 0; !stop && !(buf = static_cast<char*>(mbgcg_image_alloc(h, want))) && spin < 20000; spin++)
                    std::this_thread::sleep_for(std::chrono::microseconds(50));
            {
                std::lock_guard<std::mutex> g(mu);
                passTurn(i);
            }
            turnTaken = true;
            if (pack && buf) {
                uint64_t used = 0, fsize = 0;
                if (mbgcg_image_pack_file(name, buf, want, &used, &fsize) == MBGCG_OK && fsize == n) {
                    im.bytes = buf;
                    im.n = n;
                    im.packed = true;
                    done = true;
                } else {  // too many exceptional bytes for the piece: the file travels as it is, from pageable memory
                    mbgcg_image_free(h, buf);
                    buf = nullptr;
                }
            }
            const bool arena = buf != nullptr;
            if (done) {
                close(fd);
                std::lock_guard<std::mutex> g(mu);
                im.ready = true;
                images[i] = im;
                cv.notify_all();
                return;
            }
            if (!buf) buf = static_cast<char*>(malloc(n));
            size_t got = 0;
            while (buf && got < n) {
                const ssize_t r = read(fd, buf + got, n - got);
                if (r <= 0) break;
                got += static_cast<size_t>(r);
            }
            if (buf && got == n && !(static_cast<uint8_t>(buf[0]) == GZIP_ID1 && static_cast<uint8_t>(buf[1]) == GZIP_ID2)) {
                im.bytes = buf;
                im.n = n;
                im.mallocd = !arena;
                done = true;
            } else if (buf) {  // gzip (or a short read): the reference's own opener deals with it
                if (arena) mbgcg_image_free(h, buf); else free(buf);
            }
        }
        if (fd >= 0) close(fd);
        if (!done) {  // gz, stdin, "name.gz" fallback, error messages: exactly the reference's behaviour
            mgmpInFile fp = mgmpInOpen(name);
            im.bytes = reinterpret_cast<const char*>(fp.out);
            im.n = fp.size;
            im.mallocd = true;
        }
        std::lock_guard<std::mutex> g(mu);
        if (!turnTaken) passTurn(i);  // a file that never asked for pinned memory still passes the turn on
        im.ready = true;
        images[i] = im;
        cv.notify_all();
    }

    mbgcg_t* h;
    std::vector<const char*> names;
    std::vector<Image> images;
    std::vector<char> turnDone;
    uint64_t arenaMax;
    int readAhead;
    const bool packImages = !(std::getenv("MBGC_B200_PACK_IMAGES") && std::getenv("MBGC_B200_PACK_IMAGES")[0] == '0');
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> next{0};
    int allocTurn = 0, consumed = 0;
    bool stop = false;
    std::vector<std::thread> threads;
};

int reader_threads(int nFiles) {
    // all cores but three: the commit loop, the replicas' workers and the stagers need some too
    int n = std::max(1, static_cast<int>(std::thread::hardware_concurrency()) - 3);
    if (PgHelpers::numberOfThreads > 0 && PgHelpers::numberOfThreads < n) n = PgHelpers::numberOfThreads;
    if (const char* e = std::getenv("MBGC_B200_READERS")) n = std::atoi(e);
    n = std::max(1, std::min(n, std::min(std::max(nFiles, 1), 32)));
    if (iterableDecoder) n = 1;  // the decoder iterator hands out files in order, from one thread
    return n;
}

int read_ahead() {
    if (const char* e = std::getenv("MBGC_B200_READ_AHEAD")) return std::max(1, std::atoi(e));
    return 64;
}

void fill_pack_params(mbgcg_pack_params_t& p, const MBGC_Params* mp) {
    p.enableExtensionsWithMismatches = mp->enableExtensionsWithMismatches;
    p.gapDepthOffsetEncoding = mp->gapDepthOffsetEncoding;
    p.gapDepthMismatchesEncoding = mp->gapDepthMismatchesEncoding;
    p.lazyDecompressionSupport = mp->lazyDecompressionSupport;
    p.frugal64bitLenEncoding = mp->frugal64bitLenEncoding;
    p.enable40bitReference = mp->enable40bitReference;
    p.mismatchesWithExclusion = mp->mismatchesWithExclusion;
    p.mmsMismatchesInitialScore = mp->mmsMismatchesInitialScore;
    p.mmsMismatchesScoreThreshold = mp->mmsMismatchesScoreThreshold;
    p.mmsMismatchPenalty = mp->mmsMismatchPenalty;
    p.mmsMatchBonus = mp->mmsMatchBonus;
    p.gapBreakingMatchMinLength = mp->gapBreakingMatchMinLength;
}

void check_supported(const MGMP_Params* params, bool parallelLoop) {
    if (!params->enableDNALineLengthDetection || params->allowLossyParsing)
        die("mbgc_b200 (batched): the lossy / FASTQ readers are not part of the device ingest; use mbgc_b200_full");
    if (params->uppercaseDNA) die("mbgc_b200 (batched): upper-casing of the input is not part of the device ingest; use mbgc_b200_full");
    if (params->refLiteralMinimalLengthExt != SIZE_MAX && parallelLoop)
        die("mbgc_b200 (batched): literal-based reference extension belongs to the sequential schedule (-t 1)");
#ifdef DEVELOPER_BUILD
    // the experimental extension strategies of the developer build (main.cpp:302-319) other than -N / -F
    if (params->refExtensionStrategy & (MGMP_Params::COMBINED_REF_EXTENSION_STRATEGY_MASK | MGMP_Params::ONLY_LITERAL_REF_EXTENSION_STRATEGY_MASK |
                                        MGMP_Params::BLOCK_REF_EXTENSION_STRATEGY_MASK | MGMP_Params::DYNAMIC_REF_EXT_FACTOR_MASK))
        die("mbgc_b200 (batched): the experimental reference extension strategies (-1 -4 -8 -W) are not part of the collection pass");
#endif
}

// processLiteral (MBGC_Encoder.cpp:128-135; `-N` / `-F` of a developer build): the handle lists the literal runs and the
// sequential pass extends the reference with them
void apply_literal_ref_ext(mbgcg_t* h, const MGMP_Params* params) {
    const bool on = params->refLiteralMinimalLengthExt != SIZE_MAX;
    if (mbgcg_set_literal_ref_ext(h, on ? params->refLiteralMinimalLengthExt : UINT64_MAX, on ? params->refLiteralBeforeAfterExt : 0,
                                  params->currentUnmatchedFractionFactor) != MBGCG_OK)
        die(mbgcg_last_error(h));
}

struct LoopTimes {
    double waitReaders = 0, next = 0, hooks = 0;
};

void trace_pass(const char* what, mbgcg_t* h, const mbgcg_seq_stats_t& st, int nFiles, uint64_t contigs, const LoopTimes& t,
                int readers, double wholeMs) {
    fprintf(stderr,
            "[mbgc_b200] batched pass (%s): %d files, %llu contigs, %d devices, %llu windows, %llu runs, %llu extensions, %llu contigs re-run; "
            "%.1f MB images (%.1f MB pageable), %llu launches; ingest %.1f ms, match %.1f ms, pack %.1f ms, loadRef %.1f ms, "
            "waiting for staging %.1f ms; main thread: waiting for readers %.1f ms, in seq_next %.1f ms, encoder hooks %.1f ms; "
            "%d reader threads; whole loop %.1f ms\n",
            what, nFiles, (unsigned long long) contigs, mbgcg_device_count(h), (unsigned long long) st.windows,
            (unsigned long long) st.runs, (unsigned long long) st.extensions, (unsigned long long) st.rerunContigs,
            st.imageBytes / 1e6, st.pageableBytes / 1e6, (unsigned long long) st.launches, st.ingestMs, st.matchMs, st.packMs,
            st.loadRefMs, st.waitStageMs, t.waitReaders, t.next, t.hooks, readers, wholeMs);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// sequential schedule: processTargetsWithParallelIO, MultipleGenomeMatchingProcessor.cpp:232-313
//     per contig   processHeader, processMatches (the packed streams arrive ready-made), processAfterSequence
//     per file     unmatchedFractionFactors, totalFilesLength / largest*, processTargetMeta,
//                  processAfterTargetWithParallelIO, applyTemplateToHeaders
// ---------------------------------------------------------------------------------------------------------------------
void MultipleGenomeMatchingProcessor::processTargetsWithParallelIO() {
    MBGC_Params* mp = static_cast<MBGC_Params*>(params);  // the driver's MGMP_Params is the encoder's MBGC_Params object
    check_supported(params, false);
    const bool trace = std::getenv("MBGC_B200_TRACE") != nullptr;
    const double tStart = now_ms();

    initProcessTargetsWithParallelIO();
    targetRefExtensions.resize(1);

    mbgcg_t* h;
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        h = mbgc_b200_dropin::synced_handle(matcher);
    }
    // the encoder's refExtLoadedPosArr (private to MBGC_Encoder): asked for through the packer shim
    vector<PgTools::TextMatch> none;
    mbgc_b200_dropin::EncoderState es{};
    mbgc_b200_dropin::set_prepacked(nullptr, &es);
    processMatches(none, nullptr, 0, 0);
    mbgc_b200_dropin::set_prepacked(nullptr, nullptr);

    mbgcg_seq_config_t cfg{};
    fill_pack_params(cfg.pack, mp);
    cfg.minMatchLength = static_cast<uint32_t>(params->k);
    cfg.unmatchedFractionFactor = params->currentUnmatchedFractionFactor;
    cfg.unmatchedFractionRCFactor = params->unmatchedFractionRCFactor;
    cfg.rcInReferenceDisabled = params->isRCinReferenceDisabled();
    cfg.refRegionSeparators = params->refRegionSeparators;
    cfg.separatorPerFile = mp->lazyDecompressionSupport;
    cfg.regionSeparator = MGMP_Params::REF_REGION_SEPARATOR;
    apply_literal_ref_ext(h, params);
    if (mbgcg_seq_begin(h, &cfg, es.loadedPosArr, es.nLoadedPos) != MBGCG_OK) die(mbgcg_last_error(h));

    const int nFiles = static_cast<int>(filesCount);
    std::vector<const char*> names(static_cast<size_t>(nFiles));
    std::vector<Image> preset(static_cast<size_t>(nFiles));
    for (int i = 0; i < nFiles; i++) names[i] = fileNames[i].c_str();
    if (nFiles && firstFpIsFirstTarget) {  // the file loadG0Ref already holds (mgmpInReset, :247)
        mgmpInReset(firstFp);
        preset[0].bytes = reinterpret_cast<const char*>(firstFp.out);
        preset[0].n = firstFp.size;
    }
    LoopTimes lt;
    uint64_t contigsTotal = 0;
    int nReaders = 0;
    {
        ImageReader reader(h, names, preset, reader_threads(nFiles), read_ahead());
        nReaders = reader.threadCount();
        int pushed = 0, openFile = -1, seqCounter = 0;
        size_t fileStartPos = matcher->getLoadedRefLength();
        mbgcg_seq_event_t ev;
        auto pushReady = [&](bool block) {
            for (;;) {
                if (pushed >= nFiles) return;
                if (!block && !mbgcg_seq_wants_input(h)) return;
                Image im;
                if (!reader.get(pushed, block, im, &lt.waitReaders)) return;
                if ((im.packed ? mbgcg_seq_push_packed(h, im.bytes) : mbgcg_seq_push_file(h, im.bytes, im.n)) != MBGCG_OK) die(mbgcg_last_error(h));
                pushed++;
                block = false;
                if (pushed == nFiles && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
            }
        };
        auto beginFile = [&](int i) {  // the head of the ordered section, :249-254
            fileStartPos = matcher->getLoadedRefLength();
            unmatchedFractionFactors.push_back(params->currentUnmatchedFractionFactor < 256 ? params->currentUnmatchedFractionFactor : 0);
            unmatchedFractionFactors.push_back(params->unmatchedFractionRCFactor);
            const size_t n = reader.size(i);
            totalFilesLength += n;
            largestFileLength = n > largestFileLength ? n : largestFileLength;
            seqCounter = 0;
            openFile = i;
        };
        if (nFiles == 0 && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
        for (;;) {
            pushReady(false);
            const double n0 = now_ms();
            if (mbgcg_seq_next(h, &ev) != MBGCG_OK) die(mbgcg_last_error(h));
            const double n1 = now_ms();
            lt.next += n1 - n0;
            if (ev.type == MBGCG_EV_DONE) break;
            if (ev.type == MBGCG_EV_NEED_INPUT) {
                pushReady(true);
                continue;
            }
            const int i = static_cast<int>(ev.file);
            if (openFile != i) beginFile(i);
            if (ev.type == MBGCG_EV_CONTIG) {
                largestContigSize = ev.seqLen > largestContigSize ? ev.seqLen : largestContigSize;
                string header(ev.hdr, ev.hdrLen);                       // readHeader: the lossless reader keeps the whole line in name
                processHeader(i, header);
                seqCounter++;
                contigsTotal++;
                resCount += ev.nMatches;
                mbgc_b200_dropin::set_prepacked(&ev, nullptr);
                processMatches(none, nullptr, ev.seqLen, 0);            // appends the packed segments
                mbgc_b200_dropin::set_prepacked(nullptr, nullptr);
                if (ev.extended) largestRefContigSize = ev.seqLen > largestRefContigSize ? ev.seqLen : largestRefContigSize;
                processAfterSequence(0);
            } else {  // MBGCG_EV_FILE_END
                if (ev.status <= -2) {
                    string last = "";
                    validate_kseq_status(fileNames[i], last.c_str(), ev.status);
                }
                processTargetMeta(seqCounter, ev.dnaLineLen);
                // the pass already moved every replica past the separator and took / released the window lock; the hook
                // only records them: the shim answers its matcher calls from this event
                {
                    std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
                    mbgc_b200_dropin::pull_state(matcher);
                    mbgc_b200_dropin::set_replay(true, ev.lockPos);
                }
                processAfterTargetWithParallelIO(fileStartPos);
                {
                    std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
                    mbgc_b200_dropin::set_replay(false, 0);
                }
                applyTemplateToHeaders(i);
                if (i == 0 && firstFpIsFirstTarget) mgmpInClose(firstFp);      // :309
                reader.consumedUpTo(i);
                openFile = -1;
            }
            lt.hooks += now_ms() - n1;
        }
    }
    mbgcg_seq_stats_t st;
    if (mbgcg_seq_end(h, &st) != MBGCG_OK) die(mbgcg_last_error(h));
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        mbgc_b200_dropin::pull_state(matcher);
    }
    finalizeProcessTargetsWithParallelIO();
    if (trace) trace_pass("sequential schedule", h, st, nFiles, contigsTotal, lt, nReaders, now_ms() - tStart);
}

// ---------------------------------------------------------------------------------------------------------------------
// parallel schedule: processTargetsParallel :520-555, processTarget :340-429, finalizeParallelProcessingOfTarget :433-468
//     per target   initProcessTarget, unmatchedFractionFactors, lock; per contig processHeader, seqsCounts, processMatches
//                  (ready-made streams, under the target's lock), processAfterSequence; processAfterTarget, dnaLineLength,
//                  statistics; then, in target order: processTargetMeta, the encoder's finalize hook, applyTemplateToHeaders
// ---------------------------------------------------------------------------------------------------------------------
void MultipleGenomeMatchingProcessor::processTargetsParallel() {
    MBGC_Params* mp = static_cast<MBGC_Params*>(params);
    check_supported(params, true);
    const bool trace = std::getenv("MBGC_B200_TRACE") != nullptr;
    const double tStart = now_ms();

    initParallelProcessing();
    if (params->matcherWorkingThreads > PgHelpers::numberOfThreads) params->matcherWorkingThreads = PgHelpers::numberOfThreads;  // :527-528
    allowedTargetsOutrun = params->matcherWorkingThreads * (int) params->allowedTargetsOutrunFactor;                             // :532

    mbgcg_t* h;
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        h = mbgc_b200_dropin::synced_handle(matcher);
    }
    vector<PgTools::TextMatch> none;
    mbgc_b200_dropin::EncoderState es{};
    mbgc_b200_dropin::set_prepacked(nullptr, &es);
    processMatches(none, nullptr, 0, 0);
    mbgc_b200_dropin::set_prepacked(nullptr, nullptr);

    mbgcg_seq_config_t cfg{};
    fill_pack_params(cfg.pack, mp);
    cfg.minMatchLength = static_cast<uint32_t>(params->k);
    cfg.unmatchedFractionFactor = params->currentUnmatchedFractionFactor;
    cfg.unmatchedFractionRCFactor = params->unmatchedFractionRCFactor;
    cfg.rcInReferenceDisabled = params->isRCinReferenceDisabled();
    cfg.refRegionSeparators = params->refRegionSeparators;
    cfg.separatorPerFile = mp->lazyDecompressionSupport;
    cfg.regionSeparator = MGMP_Params::REF_REGION_SEPARATOR;
    cfg.parallelSchedule = 1;
    cfg.contigsIndividuallyReversed = params->contigsIndivduallyReversed;
    cfg.allowedTargetsOutrun = std::max(1, allowedTargetsOutrun);
    cfg.allowedTargetsOutrunForDissimilarContigs = params->allowedTargetsOutrunForDissimilarContigs;
    cfg.unmatchedFractionFactorTweakForDissimilarContigs = params->unmatchedFractionFactorTweakForDissimilarContigs;
    cfg.minimalLengthForDissimilarContigs = params->minimalLengthForDissimilarContigs;
    if (mbgcg_seq_begin(h, &cfg, es.loadedPosArr, es.nLoadedPos) != MBGCG_OK) die(mbgcg_last_error(h));

    // targets: files 1.. of the list, or the blocks of the one input file behind the reference block
    // (mgmpInSplit_next, :494-497: at least MIN_BASIC_BLOCK_SIZE bytes, cut at the next '>')
    int nTargets = static_cast<int>(targetsCount);
    std::vector<const char*> names;
    std::vector<Image> preset;
    if (singleFastaFileMode) {
        for (int i = 0; i < nTargets; i++) {
            mgmpInFile fp = mgmpInSplit_next(firstFp, params->MIN_BASIC_BLOCK_SIZE, '>');
            if (fp.size == 0) {
                nTargets = i;
                break;
            }
            Image im;
            im.bytes = reinterpret_cast<const char*>(fp.out);
            im.n = fp.size;
            im.ready = true;
            preset.push_back(im);
        }
        targetsCount = static_cast<uint32_t>(nTargets);  // :499-500
        names.assign(static_cast<size_t>(nTargets), fileNames[0].c_str());
    } else {
        names.resize(static_cast<size_t>(nTargets));
        preset.resize(static_cast<size_t>(nTargets));
        for (int i = 0; i < nTargets; i++) names[i] = fileNames[i + targetsShift].c_str();
        if (nTargets && firstFpIsFirstTarget) {
            mgmpInReset(firstFp);
            preset[0].bytes = reinterpret_cast<const char*>(firstFp.out);
            preset[0].n = firstFp.size;
        }
    }
    LoopTimes lt;
    uint64_t contigsTotal = 0;
    int nReaders = 0;
    {
        ImageReader reader(h, names, preset, reader_threads(nTargets), read_ahead());
        nReaders = reader.threadCount();
        int pushed = 0, openTarget = -1;
        uint32_t tLargestContig = 0, tLargestRefContig = 0, tResCount = 0;
        mbgcg_seq_event_t ev;
        auto pushReady = [&](bool block) {
            for (;;) {
                if (pushed >= nTargets) return;
                if (!block && !mbgcg_seq_wants_input(h)) return;
                Image im;
                if (!reader.get(pushed, block, im, &lt.waitReaders)) return;
                if ((im.packed ? mbgcg_seq_push_packed(h, im.bytes) : mbgcg_seq_push_file(h, im.bytes, im.n)) != MBGCG_OK) die(mbgcg_last_error(h));
                pushed++;
                block = false;
                if (pushed == nTargets && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
            }
        };
        auto beginTarget = [&](int i) {  // the head of processTarget, :341-352
            initProcessTarget(i);
            const int uff = params->currentUnmatchedFractionFactor;
            unmatchedFractionFactors[2 * (targetsCountShift + i)] = uff < 256 ? uff : 0;
            unmatchedFractionFactors[2 * (targetsCountShift + i) + 1] = params->unmatchedFractionRCFactor;
            tLargestContig = tLargestRefContig = tResCount = 0;
            openTarget = i;
        };
        if (nTargets == 0 && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
        for (;;) {
            pushReady(false);
            const double n0 = now_ms();
            if (mbgcg_seq_next(h, &ev) != MBGCG_OK) die(mbgcg_last_error(h));
            const double n1 = now_ms();
            lt.next += n1 - n0;
            if (ev.type == MBGCG_EV_DONE) break;
            if (ev.type == MBGCG_EV_NEED_INPUT) {
                pushReady(true);
                continue;
            }
            const int i = static_cast<int>(ev.file);
            if (openTarget != i) beginTarget(i);
            if (ev.type == MBGCG_EV_CONTIG) {
                matchingLocksPos[i] = ev.contigLock;
                tLargestContig = ev.seqLen > tLargestContig ? ev.seqLen : tLargestContig;
                string header(ev.hdr, ev.hdrLen);
                processHeader(i + targetsShift, header);
                seqsCounts[i]++;
                contigsTotal++;
                tResCount += ev.nMatches;
                mbgc_b200_dropin::set_prepacked(&ev, nullptr);
                processMatches(none, nullptr, ev.seqLen, i, ev.contigLock);
                mbgc_b200_dropin::set_prepacked(nullptr, nullptr);
                if (ev.extended) tLargestRefContig = ev.seqLen > tLargestRefContig ? ev.seqLen : tLargestRefContig;
                processAfterSequence(i);
            } else {  // MBGCG_EV_FILE_END: the rest of processTarget (:412-428), then its finalization (:439-458)
                processAfterTarget(i);
                if (ev.status <= -2) {
                    string last = "";
                    validate_kseq_status(fileNames[i + targetsShift], last.c_str(), ev.status);
                }
                dnaLineLength[i] = ev.dnaLineLen;
                this->resCount += tResCount;
                this->largestRefContigSize = this->largestRefContigSize > tLargestRefContig ? this->largestRefContigSize : tLargestRefContig;
                this->largestContigSize = this->largestContigSize > tLargestContig ? this->largestContigSize : tLargestContig;
                const size_t fsize = reader.size(i);
                largestFileLength = fsize > largestFileLength ? fsize : largestFileLength;
                totalFilesLength += fsize;
                fileNames[i + targetsShift].clear();
                masterTargetsStats++;
                matchingLocksPos[i] = ev.lockPos;
                processTargetMeta(seqsCounts[i], dnaLineLength[i]);
                {
                    std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
                    mbgc_b200_dropin::pull_state(matcher);
                    mbgc_b200_dropin::set_replay(true, ev.lockPos);
                }
                finalizeParallelProcessingOfTarget(static_cast<uint32_t>(i), static_cast<size_t>(ev.loadedRefLength - ev.refExtSize));
                {
                    std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
                    mbgc_b200_dropin::set_replay(false, 0);
                }
                masterRefExtensionsStats++;
                processedTargetsCount = i + 1;
                applyTemplateToHeaders(i + targetsShift);
                if (!singleFastaFileMode) {
                    if (i == 0 && firstFpIsFirstTarget) mgmpInClose(firstFp);
                    reader.consumedUpTo(i);
                }
                openTarget = -1;
            }
            lt.hooks += now_ms() - n1;
        }
    }
    mbgcg_seq_stats_t st;
    if (mbgcg_seq_end(h, &st) != MBGCG_OK) die(mbgcg_last_error(h));
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        mbgc_b200_dropin::pull_state(matcher);
    }
    if (singleFastaFileMode) {  // :550-553
        mgmpInClose(firstFp);
        finalizeParallelProcessingInSingleFastaFileMode();
    }
    if (trace) trace_pass("parallel schedule", h, st, nTargets, contigsTotal, lt, nReaders, now_ms() - tStart);
}

// ---------------------------------------------------------------------------------------------------------------------
// performMatching, MultipleGenomeMatchingProcessor.cpp:568-606: the dispatcher.  The compiler inlines the stock
// processTargetsParallel into the stock performMatching, so the loop above can only be reached by supplying the
// dispatcher too (its symbol is weakened like the other two).  Same decisions, same state left behind
// (refFinalTotalLength, the matcher deleted), same lines on the application and log streams.
// ---------------------------------------------------------------------------------------------------------------------
void MultipleGenomeMatchingProcessor::performMatching() {
    if (params->sequentialMatching) {
        if (params->matcherWorkingThreadsFixed && params->matcherWorkingThreads > 1)
            fprintf(stderr, "WARNING: sequential mode is enabled - only single worker thread is used (although %d workers are set).\n",
                    params->matcherWorkingThreads);
        processTargetsWithParallelIO();
    } else if (params->bruteParallel) {
        processTargetsBruteParallel();  // the reference's own (developer option): one contig per call through the shims
    } else if (targetsCount) {
        processTargetsParallel();
    } else {
        fprintf(stderr, "Error selecting processing mode (no targets for parallel matching?)!\n");
        exit(EXIT_FAILURE);
    }
    std::ostream& dev = *PgHelpers::devout;
    dev << endl << "processed targets count: " << targetsCount << endl;
    refFinalTotalLength = matcher->getRefLength();
    const size_t loadedRefTotalLength = matcher->getLoadedRefLength();
    delete (matcher);
    dev << "sparseEM reference length: " << refFinalTotalLength << endl
        << "loaded reference total length: " << loadedRefTotalLength << endl
        << "largest reference contig length: " << largestRefContigSize << endl
        << "largest contig length: " << largestContigSize << endl
        << "targets dna total length: " << totalDestLenAll << endl
        << "exact matches total: " << resCount << endl
        << "swsMEM unmatched chars: " << unmatchedCharsAll << " (removed: "
        << PgHelpers::getPartShareString(totalMatchedAll, totalDestLenAll) << "; " << totalDestOverlapAll
        << " chars in overlapped target symbol)" << endl;
    printAdditionalMatchingStats();
    *PgHelpers::appout << "matching finished - " << PgHelpers::time_millis() << " [ms]" << endl << endl;
    *PgHelpers::logout << PgHelpers::time_millis() << "       \t";
}
