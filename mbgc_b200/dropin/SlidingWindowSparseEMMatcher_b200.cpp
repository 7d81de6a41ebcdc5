// mbgc_b200/dropin/SlidingWindowSparseEMMatcher_b200.cpp
//
// Link-time substitute for /root/reference/matching/SlidingWindowSparseEMMatcher.cpp.
//
// It is compiled against the UNMODIFIED reference header (matching/SlidingWindowSparseEMMatcher.h) and
// defines every out-of-line member that header declares and the rest of MBGC calls (constructors,
// destructor, loadRef, loadSeparator, matchTexts, acquire/releaseWorkerMatchingLockPos) on top of the C ABI of
// include/mbgc_b200.h, i.e. on the CUDA kernels.  Everything else of `mbgc` is built from the reference's own
// sources (oracle/Makefile target `dropin`), so MultipleGenomeMatchingProcessor and MBGC_Encoder run unchanged:
// same call sites (MultipleGenomeMatchingProcessor.cpp:171-183,274,286,379,441,456; MBGC_Encoder.cpp:181,318-330,
// 376-380,501-508,558-561,883), same archive bytes.
//
// The reference header's inline methods (disableSlidingWindow, setSlidingWindowSize, disableCircularBuffer,
// setPosition, getRef, getRefLength, getLoadingPosition, getLoadedRefLength; .h:93-113) read and write the
// protected members directly, so the members stay authoritative for the host-visible state: it is pushed into
// the handle before, and pulled back after, every call (mbgcg_set_state / mbgcg_get_state).  The handle itself
// lives in the otherwise unused ht64bit pointer; start1 points at the handle's host mirror of the reference.
//
// Errors keep the reference's behaviour: message on stderr, exit(EXIT_FAILURE).
#include "SlidingWindowSparseEMMatcher.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include "mbgc_b200.h"
#include "dropin_shared.h"

namespace {

// the text and list of the most recent matchTexts call (under g_lock): the packer may still find both on the device
mbgc_b200_dropin::LastMatch g_lastMatch{nullptr, nullptr, 0, 0};

// batched target loop: see dropin_shared.h
const mbgcg_seq_event_t* g_prepacked = nullptr;
mbgc_b200_dropin::EncoderState* g_encoderStateRequest = nullptr;
bool g_replay = false;
uint64_t g_replayLock = 0;
int g_firstDevice = 0;

// The reference's parallel mode enters matchTexts from several OpenMP threads while another one runs loadRef
// (SURVEY.md 8b "Threading").  A handle is not thread-safe; serialising is a legal schedule of that mode.
std::mutex g_lock;

inline mbgcg_t* handle_of(uint64_t* slot) { return reinterpret_cast<mbgcg_t*>(slot); }

[[noreturn]] void die(const char* msg) {
    fprintf(stderr, "%s\n\n", msg);
    exit(EXIT_FAILURE);
}

using mbgc_b200_dropin::Trace;

// MBGC_B200_DEVICES = "0,1,2,3": the matcher is replicated on these devices (mbgcg_create_multi); the batched target loop
// shards its windows over them.  MBGC_B200_DEVICE = one device (default 0).
std::vector<int> devices_from_env() {
    std::vector<int> d;
    if (const char* e = std::getenv("MBGC_B200_DEVICES")) {
        for (const char* p = e; *p;) {
            char* end = nullptr;
            const long v = std::strtol(p, &end, 10);
            if (end == p) break;
            d.push_back(static_cast<int>(v));
            p = *end == ',' ? end + 1 : end;
        }
    }
    if (d.empty()) {
        const char* e = std::getenv("MBGC_B200_DEVICE");
        d.push_back(e ? std::atoi(e) : 0);
    }
    // CUDA initialises every VISIBLE device when the first context is created (~0.5-0.8 s apiece on a multi-GPU box):
    // unless the user has set the mask already, show the runtime only the devices this run uses.  This is the first
    // CUDA-related thing the process does, so the mask takes effect.
    if (!std::getenv("CUDA_VISIBLE_DEVICES") && !std::getenv("MBGC_B200_KEEP_VISIBLE")) {
        std::string mask;
        for (size_t i = 0; i < d.size(); i++) mask += (i ? "," : "") + std::to_string(d[i]);
        setenv("CUDA_VISIBLE_DEVICES", mask.c_str(), 1);
        for (size_t i = 0; i < d.size(); i++) d[i] = static_cast<int>(i);
    }
    return d;
}

}  // namespace

// MBGC_B200_TRACE=1: where the process spends its time outside the reference's own timers -- wall-clock stamps (seconds,
// CLOCK_REALTIME) when this translation unit is initialised (the dynamic loader has mapped libmbgc_b200.so and the CUDA
// runtime by then), when main has returned (atexit), and around handle creation / destruction
namespace {
double wall_now() {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return static_cast<double>(ts.tv_sec) + ts.tv_nsec * 1e-9;
}
void stamp(const char* what) {
    if (Trace::on()) fprintf(stderr, "[mbgc_b200] stamp %-22s %.6f\n", what, wall_now());
}
// (Leaving through _exit right after main instead of the CUDA runtime's own teardown was measured: 0.24 s instead of
// 0.25-0.32 s between the end of main and the end of the process -- the rest is the kernel releasing the device
// mappings.  Not worth skipping destructors for.)
void leave(int, void*) { stamp("main returned (on_exit)"); }
struct ProcessStamps {
    ProcessStamps() {
        stamp("shim initialised");
        on_exit(leave, nullptr);
    }
    ~ProcessStamps() { stamp("static destructors"); }
} g_processStamps;
}  // namespace

#define MBGCG_PUSH_STATE(h)                                       \
    do {                                                          \
        mbgcg_state_t st_;                                        \
        mbgcg_get_state((h), &st_);                               \
        st_.pos1 = pos1;                                          \
        st_.reachedRefLengthCount = reachedRefLengthCount;        \
        st_.circularBuffer = circularBuffer ? 1 : 0;              \
        st_.swSize = swSize;                                      \
        st_.swEnd = swEnd;                                        \
        st_.samplingPos = samplingPos;                            \
        mbgcg_set_state((h), &st_);                               \
    } while (0)

#define MBGCG_PULL_STATE(h)                                       \
    do {                                                          \
        mbgcg_state_t st_;                                        \
        mbgcg_get_state((h), &st_);                               \
        pos1 = st_.pos1;                                          \
        reachedRefLengthCount = st_.reachedRefLengthCount;        \
        swEnd = st_.swEnd;                                        \
        samplingPos = st_.samplingPos;                            \
    } while (0)

// ref .cpp:325-359 (skipHtInit is irrelevant here: the table lives on the device either way)
SlidingWindowSparseEMMatcher::SlidingWindowSparseEMMatcher(const size_t refLengthLimit, const uint32_t targetMatchLength,
                                                           int _k1, int _k2, int skipMargin, uint32_t minMatchLength,
                                                           bool skipHtInit)
        : maxRefLength(refLengthLimit), L(targetMatchLength), skipMargin(skipMargin) {
    (void) minMatchLength;
    (void) skipHtInit;
    Trace tr("create", refLengthLimit);
    mbgcg_t* h = nullptr;
    const std::vector<int> devs = devices_from_env();
    g_firstDevice = devs[0];
    if (mbgcg_create_multi(&h, refLengthLimit, targetMatchLength, _k1, _k2, skipMargin, devs.data(), static_cast<int>(devs.size())) != MBGCG_OK)
        die(mbgcg_create_error());
    ht64bit = reinterpret_cast<std::uint64_t*>(h);
    ht32bit = nullptr;
    ht64bitFlag = false;
    mbgcg_state_t st;
    mbgcg_get_state(h, &st);
    start1 = const_cast<char*>(mbgcg_host_ref(h));
    pos1 = st.pos1;
    maxRefLength = st.maxRefLength;
    swSize = st.swSize;
    swEnd = st.swEnd;
    samplingPos = st.samplingPos;
    hash_size = st.hashSize;
    hash_size_minus_one = hash_size - 1;
    K = st.K;
    k1 = _k1;
    k2 = _k2;
    *PgHelpers::devout << "SparseEM (B200): k-mer exact match length = " << L << "; hash pattern length = " << K
                       << "; HASH_SIZE = " << hash_size << "; k1 = " << k1 << "; k2 = " << k2
                       << "; skipMargin = " << skipMargin << std::endl;
}

// ref .cpp:494-519: the C ABI picks the Exp variant from ctz(k1) exactly as MultipleGenomeMatchingProcessor does
SlidingWindowExpSparseEMMatcher::SlidingWindowExpSparseEMMatcher(const size_t refLengthLimit,
                                                                 const uint32_t targetMatchLength, int _k1, int _k2,
                                                                 int skipMargin, uint32_t minMatchLength)
        : SlidingWindowSparseEMMatcher(refLengthLimit, targetMatchLength, _k1, _k2, skipMargin, minMatchLength, true),
          k1ord(__builtin_ctz((uint32_t) _k1)) {
    if (_k1 % 2) die("Error initializing ExpSparseMEM: incorrect k1");
}

SlidingWindowExpRandSparseEMMatcher::SlidingWindowExpRandSparseEMMatcher(const size_t, const uint32_t, int, int, int,
                                                                         uint32_t)
        : SlidingWindowExpSparseEMMatcher(2, 32, 16, 1, 0) {
    die("SlidingWindowExpRandSparseEMMatcher is not part of the accelerated path (never selected by MBGC)");
}

// ref .cpp:460-467
SlidingWindowSparseEMMatcher::~SlidingWindowSparseEMMatcher() {
    std::lock_guard<std::mutex> g(g_lock);
    stamp("destroy begins");
    mbgcg_destroy(handle_of(ht64bit));
    stamp("destroy ends");
    ht64bit = nullptr;
    start1 = nullptr;
}

// ref .cpp:453-458
void SlidingWindowSparseEMMatcher::loadRef(const char* refText, size_t refLength, bool loadRCRef, bool addRegionSeparators,
                                           char regionSeparator) {
    Trace tr("loadRef", refLength);
    std::lock_guard<std::mutex> g(g_lock);
    mbgcg_t* h = handle_of(ht64bit);
    MBGCG_PUSH_STATE(h);
    if (mbgcg_load_ref(h, refText, refLength, loadRCRef, addRegionSeparators, regionSeparator) != MBGCG_OK)
        die(mbgcg_last_error(h));
    MBGCG_PULL_STATE(h);
}

// ref .cpp:439-451
void SlidingWindowSparseEMMatcher::loadSeparator(char regionSeparator) {
    std::lock_guard<std::mutex> g(g_lock);
    if (g_replay) return;  // the collection pass wrote the separator on every replica; the members already show it
    mbgcg_t* h = handle_of(ht64bit);
    MBGCG_PUSH_STATE(h);
    if (mbgcg_load_separator(h, regionSeparator) != MBGCG_OK) die(mbgcg_last_error(h));
    MBGCG_PULL_STATE(h);
}

// ref .cpp:361-378
size_t SlidingWindowSparseEMMatcher::acquireWorkerMatchingLockPos() {
    std::lock_guard<std::mutex> g(g_lock);
    if (g_replay) return g_replayLock;
    mbgcg_t* h = handle_of(ht64bit);
    MBGCG_PUSH_STATE(h);
    const size_t v = mbgcg_acquire_lock(h);
    MBGCG_PULL_STATE(h);
    return v;
}

// ref .cpp:380-400
void SlidingWindowSparseEMMatcher::releaseWorkerMatchingLockPos(size_t lockValue) {
    std::lock_guard<std::mutex> g(g_lock);
    if (g_replay) return;
    mbgcg_t* h = handle_of(ht64bit);
    MBGCG_PUSH_STATE(h);
    if (mbgcg_release_lock(h, lockValue) != MBGCG_OK) die(mbgcg_last_error(h));
    MBGCG_PULL_STATE(h);
}

// ref .cpp:477-492
void SlidingWindowSparseEMMatcher::matchTexts(vector<TextMatch>& resMatches, const char* destText, size_t destLen,
                                              bool destIsRef, bool revComplMatching, uint32_t minMatchLength,
                                              size_t matchingLockPos) {
    if (destIsRef || revComplMatching)  // ref .cpp:185-188
        die("Source as destination and reverse-complement matching modes unsupported");
    Trace tr("matchTexts", destLen);
    std::lock_guard<std::mutex> g(g_lock);
    mbgcg_t* h = handle_of(ht64bit);
    MBGCG_PUSH_STATE(h);
    const mbgcg_match_t* out = nullptr;
    uint64_t n = 0;
    if (mbgcg_match(h, destText, destLen, minMatchLength, matchingLockPos, &out, &n) != MBGCG_OK)
        die(mbgcg_last_error(h));
    resMatches.clear();
    resMatches.reserve(n);
    for (uint64_t i = 0; i < n; i++)
        resMatches.push_back(TextMatch(out[i].posSrcText, out[i].length, out[i].posDestText));
    g_lastMatch = mbgc_b200_dropin::LastMatch{this, destText, destLen, n};
}

// ---- hooks for the other drop-in translation unit (MBGC_Encoder_processMatches_b200.cpp) ----------------------
namespace mbgc_b200_dropin {

// protected members of the reference class, reached through pointers to members named in a derived class
struct Access : SlidingWindowSparseEMMatcher {
    static uint64_t* SlidingWindowSparseEMMatcher::* ht() { return &Access::ht64bit; }
    static void push(SlidingWindowSparseEMMatcher* m, mbgcg_t* h) {
        mbgcg_state_t st;
        mbgcg_get_state(h, &st);
        st.pos1 = m->*(&Access::pos1);
        st.reachedRefLengthCount = m->*(&Access::reachedRefLengthCount);
        st.circularBuffer = m->*(&Access::circularBuffer) ? 1 : 0;
        st.swSize = m->*(&Access::swSize);
        st.swEnd = m->*(&Access::swEnd);
        st.samplingPos = m->*(&Access::samplingPos);
        mbgcg_set_state(h, &st);
    }
    static void pull(SlidingWindowSparseEMMatcher* m, mbgcg_t* h) {
        mbgcg_state_t st;
        mbgcg_get_state(h, &st);
        m->*(&Access::pos1) = st.pos1;
        m->*(&Access::reachedRefLengthCount) = st.reachedRefLengthCount;
        m->*(&Access::swEnd) = st.swEnd;
        m->*(&Access::samplingPos) = st.samplingPos;
    }
};

std::mutex& handle_lock() { return g_lock; }
int first_device() { return g_firstDevice; }

mbgcg_t* synced_handle(SlidingWindowSparseEMMatcher* m) {
    mbgcg_t* h = handle_of(m->*Access::ht());
    Access::push(m, h);
    return h;
}

void pull_state(SlidingWindowSparseEMMatcher* m) { Access::pull(m, handle_of(m->*Access::ht())); }

void set_prepacked(const mbgcg_seq_event_t* ev, EncoderState* es) {
    g_prepacked = ev;
    g_encoderStateRequest = es;
}
const mbgcg_seq_event_t* prepacked_event() { return g_prepacked; }
EncoderState* encoder_state_request() { return g_encoderStateRequest; }
void set_replay(bool on, uint64_t lockPos) {
    g_replay = on;
    g_replayLock = lockPos;
}

LastMatch last_match() { return g_lastMatch; }
void forget_last_match() { g_lastMatch = LastMatch{nullptr, nullptr, 0, 0}; }

}  // namespace mbgc_b200_dropin

// ref .cpp:470-475
void SlidingWindowSparseEMMatcher::matchTexts(vector<TextMatch>& resMatches, const string& destText, bool destIsRef,
                                              bool revComplMatching, uint32_t minMatchLength, size_t matchingLockPos) {
    matchTexts(resMatches, destText.data(), destText.size(), destIsRef, revComplMatching, minMatchLength,
               matchingLockPos);
}
