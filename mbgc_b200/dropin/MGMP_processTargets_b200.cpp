// mbgc_b200/dropin/MGMP_processTargets_b200.cpp
//
// Link-time substitute for ONE member function of the reference driver:
// MultipleGenomeMatchingProcessor::processTargetsWithParallelIO
// (/root/reference/matching/MultipleGenomeMatchingProcessor.cpp:232-313) -- the target loop of the sequential schedule
// (`mbgc c -t 1`, `-Q`, `-m 3`).  Compiled against the UNMODIFIED reference headers; the rest of
// MultipleGenomeMatchingProcessor.cpp is the reference's own object file in which this one symbol has been made weak
// (`objcopy --weaken-symbol`, oracle/Makefile target `dropin_batched`), so performMatching (:568-575) calls the
// definition below.
//
// The stock loop handles one contig at a time: read a file with one thread (kseq), matchTexts, processMatches, decide
// about the reference extension, loadRef.  Here the loop is the collection pass of the C ABI (mbgcg_seq_*,
// include/mbgc_b200.h): reader threads bring whole file images into pinned memory, the pass ingests, matches and packs
// windows of contigs on all devices of the handle and hands back the committed contigs in target order; this function
// only applies the encoder's hooks to them, in the order the stock loop would have:
//
//     per contig   processHeader, processMatches (the packed streams arrive ready-made), processAfterSequence
//     per file     unmatchedFractionFactors, totalFilesLength / largest*, processTargetMeta,
//                  processAfterTargetWithParallelIO, applyTemplateToHeaders
//
// Archives are byte-identical with the stock `mbgc -t 1` (tests/test_gpu_dropin.py).
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
    bool arena = false;       // from mbgcg_image_alloc: the pass owns it once pushed
    bool mallocd = false;     // ours to free() after FILE_END
    bool ready = false;
};

}  // namespace

void MultipleGenomeMatchingProcessor::processTargetsWithParallelIO() {
    MBGC_Params* mp = static_cast<MBGC_Params*>(params);  // the driver's MGMP_Params is the encoder's MBGC_Params object
    if (!params->enableDNALineLengthDetection || params->allowLossyParsing)
        die("mbgc_b200 (batched): the lossy / FASTQ readers are not part of the device ingest; use mbgc_b200_full");
    if (params->uppercaseDNA) die("mbgc_b200 (batched): upper-casing of the input is not part of the device ingest; use mbgc_b200_full");
    if (params->refLiteralMinimalLengthExt != SIZE_MAX)
        die("mbgc_b200 (batched): literal-based reference extension is not part of the device packer");
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
    cfg.pack.enableExtensionsWithMismatches = mp->enableExtensionsWithMismatches;
    cfg.pack.gapDepthOffsetEncoding = mp->gapDepthOffsetEncoding;
    cfg.pack.gapDepthMismatchesEncoding = mp->gapDepthMismatchesEncoding;
    cfg.pack.lazyDecompressionSupport = mp->lazyDecompressionSupport;
    cfg.pack.frugal64bitLenEncoding = mp->frugal64bitLenEncoding;
    cfg.pack.enable40bitReference = mp->enable40bitReference;
    cfg.pack.mismatchesWithExclusion = mp->mismatchesWithExclusion;
    cfg.pack.mmsMismatchesInitialScore = mp->mmsMismatchesInitialScore;
    cfg.pack.mmsMismatchesScoreThreshold = mp->mmsMismatchesScoreThreshold;
    cfg.pack.mmsMismatchPenalty = mp->mmsMismatchPenalty;
    cfg.pack.mmsMatchBonus = mp->mmsMatchBonus;
    cfg.pack.gapBreakingMatchMinLength = mp->gapBreakingMatchMinLength;
    cfg.minMatchLength = static_cast<uint32_t>(params->k);
    cfg.unmatchedFractionFactor = params->currentUnmatchedFractionFactor;
    cfg.unmatchedFractionRCFactor = params->unmatchedFractionRCFactor;
    cfg.rcInReferenceDisabled = params->isRCinReferenceDisabled();
    cfg.refRegionSeparators = params->refRegionSeparators;
    cfg.separatorPerFile = mp->lazyDecompressionSupport;
    cfg.regionSeparator = MGMP_Params::REF_REGION_SEPARATOR;
    if (mbgcg_seq_begin(h, &cfg, es.loadedPosArr, es.nLoadedPos) != MBGCG_OK) die(mbgcg_last_error(h));

    // ---- reader threads: file i -> image i (pinned when the arena has room), in any order ----
    const int nFiles = static_cast<int>(filesCount);
    std::vector<Image> images(static_cast<size_t>(nFiles));
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<int> nextToRead{0};
    int consumed = 0;                       // files whose FILE_END has been seen (under mu)
    int readAhead = 64;
    if (const char* e = std::getenv("MBGC_B200_READ_AHEAD")) readAhead = std::max(1, std::atoi(e));
    int nReaders = static_cast<int>(std::thread::hardware_concurrency());
    if (PgHelpers::numberOfThreads > 0 && PgHelpers::numberOfThreads < nReaders) nReaders = PgHelpers::numberOfThreads;
    if (const char* e = std::getenv("MBGC_B200_READERS")) nReaders = std::atoi(e);
    nReaders = std::max(1, std::min(nReaders, std::min(nFiles, 32)));
    if (iterableDecoder) nReaders = 1;      // the decoder iterator hands out files in order, from one thread
    int allocTurn = 0;                      // next file that may ask for pinned memory (under mu)
    std::vector<char> turnDone(static_cast<size_t>(nFiles), 0);
    auto passTurn = [&](int i) {            // (mu held) file i has had its turn, or does not need one
        turnDone[i] = 1;
        while (allocTurn < nFiles && turnDone[allocTurn]) allocTurn++;
        cv.notify_all();
    };
    const uint64_t arenaMax = mbgcg_image_max(h);
    auto readOne = [&](int i) {
        Image im;
        bool turnTaken = false;
        if (i == 0 && firstFpIsFirstTarget) {  // the file loadG0Ref already holds (mgmpInReset, :247)
            mgmpInReset(firstFp);
            im.bytes = reinterpret_cast<const char*>(firstFp.out);
            im.n = firstFp.size;
        } else {
            const char* name = fileNames[i].c_str();
            int fd = iterableDecoder ? -1 : open(name, O_RDONLY);   // repack / append read through the decoder's iterator
            struct stat sb;
            bool done = false;
            if (fd >= 0 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_size >= 2) {
                const size_t n = static_cast<size_t>(sb.st_size);
                // Pinned memory goes out in file order: a file whose turn it is waits for room (earlier files are on their
                // way to the device and give theirs back), later files wait for their turn -- so the arena can be small
                // without ever blocking the file the pass is waiting for.
                char* buf = nullptr;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return allocTurn == i; });
                }
                if (n <= arenaMax)
                    for (int spin = 0; !(buf = static_cast<char*>(mbgcg_image_alloc(h, n))) && spin < 20000; spin++)
                        std::this_thread::sleep_for(std::chrono::microseconds(50));
                {
                    std::lock_guard<std::mutex> g(mu);
                    passTurn(i);
                }
                turnTaken = true;
                const bool arena = buf != nullptr;
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
                    im.arena = arena;
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
        }
        std::lock_guard<std::mutex> g(mu);
        if (!turnTaken) passTurn(i);  // a file that never asked for pinned memory still passes the turn on
        im.ready = true;
        images[i] = im;
        cv.notify_all();
    };
    std::vector<std::thread> readers;
    for (int t = 0; t < nReaders; t++)
        readers.emplace_back([&] {
            for (;;) {
                const int i = nextToRead.fetch_add(1);
                if (i >= nFiles) return;
                {   // bounded read-ahead: host memory holds at most readAhead images in front of the commit cursor
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return i < consumed + readAhead; });
                }
                readOne(i);
            }
        });

    // ---- the pass: push images in order, apply the hooks to what comes back ----
    int pushed = 0, openFile = -1, seqCounter = 0;
    size_t fileStartPos = matcher->getLoadedRefLength();
    uint64_t contigsTotal = 0;
    double tWaitReaders = 0, tNext = 0, tHooks = 0;
    mbgcg_seq_event_t ev;
    auto pushReady = [&](bool block) {
        for (;;) {
            if (pushed >= nFiles) return;
            if (!block && !mbgcg_seq_wants_input(h)) return;
            Image im;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (!images[pushed].ready) {
                    if (!block) return;
                    const double w0 = now_ms();
                    cv.wait(lk, [&] { return images[pushed].ready; });
                    tWaitReaders += now_ms() - w0;
                }
                im = images[pushed];
            }
            if (mbgcg_seq_push_file(h, im.bytes, im.n) != MBGCG_OK) die(mbgcg_last_error(h));
            pushed++;
            block = false;
            if (pushed == nFiles && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
        }
    };
    auto beginFile = [&](int i) {  // the head of the ordered section, :249-254
        fileStartPos = matcher->getLoadedRefLength();
        unmatchedFractionFactors.push_back(params->currentUnmatchedFractionFactor < 256 ? params->currentUnmatchedFractionFactor : 0);
        unmatchedFractionFactors.push_back(params->unmatchedFractionRCFactor);
        totalFilesLength += images[i].n;
        largestFileLength = images[i].n > largestFileLength ? images[i].n : largestFileLength;
        seqCounter = 0;
        openFile = i;
    };
    if (nFiles == 0 && mbgcg_seq_finish_input(h) != MBGCG_OK) die(mbgcg_last_error(h));
    for (;;) {
        pushReady(false);
        const double n0 = now_ms();
        if (mbgcg_seq_next(h, &ev) != MBGCG_OK) die(mbgcg_last_error(h));
        const double n1 = now_ms();
        tNext += n1 - n0;
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
            const size_t currentUnmatched = processMatches(none, nullptr, ev.seqLen, 0);   // appends the packed segments
            mbgc_b200_dropin::set_prepacked(nullptr, nullptr);
            (void) currentUnmatched;
            if (ev.extended) largestRefContigSize = ev.seqLen > largestRefContigSize ? ev.seqLen : largestRefContigSize;
            processAfterSequence(0);
        } else {  // MBGCG_EV_FILE_END
            if (ev.status <= -2) {
                string last = "";
                validate_kseq_status(fileNames[i], last.c_str(), ev.status);
            }
            processTargetMeta(seqCounter, ev.dnaLineLen);
            // the pass already moved every replica past the separator and took / released the window lock; the hook only
            // records them: the shim answers its matcher calls from this event
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
            Image im;
            {
                std::lock_guard<std::mutex> g(mu);
                im = images[i];
                images[i].bytes = nullptr;
                consumed = i + 1;
                cv.notify_all();
            }
            if (i == 0 && firstFpIsFirstTarget) mgmpInClose(firstFp);      // :309
            else if (im.mallocd) free(const_cast<char*>(im.bytes));
            openFile = -1;
        }
        tHooks += now_ms() - n1;
    }
    for (auto& t : readers) t.join();
    mbgcg_seq_stats_t st;
    if (mbgcg_seq_end(h, &st) != MBGCG_OK) die(mbgcg_last_error(h));
    {
        std::lock_guard<std::mutex> g(mbgc_b200_dropin::handle_lock());
        mbgc_b200_dropin::pull_state(matcher);
    }
    finalizeProcessTargetsWithParallelIO();
    if (trace)
        fprintf(stderr,
                "[mbgc_b200] batched pass: %d files, %llu contigs, %d devices, %llu windows, %llu runs, %llu extensions, %llu contigs re-run; "
                "%.1f MB images (%.1f MB pageable), %llu launches; ingest %.1f ms, match %.1f ms, pack %.1f ms, loadRef %.1f ms, "
                "waiting for staging %.1f ms; main thread: waiting for readers %.1f ms, in seq_next %.1f ms, encoder hooks %.1f ms; "
                "%d reader threads; whole loop %.1f ms\n",
                nFiles, (unsigned long long) contigsTotal, mbgcg_device_count(h), (unsigned long long) st.windows,
                (unsigned long long) st.runs, (unsigned long long) st.extensions, (unsigned long long) st.rerunContigs,
                st.imageBytes / 1e6, st.pageableBytes / 1e6, (unsigned long long) st.launches, st.ingestMs, st.matchMs, st.packMs,
                st.loadRefMs, st.waitStageMs, tWaitReaders, tNext, tHooks, nReaders, now_ms() - tStart);
}
