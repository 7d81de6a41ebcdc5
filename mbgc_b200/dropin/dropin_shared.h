// mbgc_b200/dropin/dropin_shared.h -- what the two drop-in translation units share: the handle behind a reference
// matcher object, the mutex that serialises calls on it, and the identity of the last matchTexts call.
#pragma once
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "mbgc_b200.h"

class SlidingWindowSparseEMMatcher;

namespace mbgc_b200_dropin {

// MBGC_B200_TRACE=1: wall time of every call through the shims on stderr
struct Trace {
    const char* what;
    size_t n;
    std::chrono::steady_clock::time_point t0;
    static bool on() {
        static const bool v = std::getenv("MBGC_B200_TRACE") != nullptr;
        return v;
    }
    Trace(const char* w, size_t bytes) : what(w), n(bytes), t0(std::chrono::steady_clock::now()) {}
    ~Trace() {
        if (on())
            fprintf(stderr, "[mbgc_b200] %-14s %10zu bytes %9.3f ms\n", what, n,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
};

struct LastMatch {
    const SlidingWindowSparseEMMatcher* matcher;
    const char* text;
    size_t len;
    uint64_t n;
};

// ---- hand-over between the batched target loop (MGMP_processTargets_b200.cpp) and the two shims -------------------
// What the loop cannot reach itself because it is private to MBGC_Encoder: asked for through processMatches.
struct EncoderState {
    const uint64_t* loadedPosArr;   // refExtLoadedPosArr
    uint64_t nLoadedPos;
};
// ev != nullptr: the next processMatches call appends the ready-made stream segments of this committed contig instead
// of packing anything; es != nullptr: the next processMatches call only fills *es.
void set_prepacked(const mbgcg_seq_event_t* ev, EncoderState* es);
const mbgcg_seq_event_t* prepacked_event();
EncoderState* encoder_state_request();
// replay: the collection pass has already applied loadSeparator and the lock acquire / release to every replica; while
// this is on, the matcher shim answers those calls of the encoder's per-file hook from the event (take handle_lock()).
void set_replay(bool on, uint64_t lockPos);
// members of the reference matcher object <- the handle (take handle_lock() first)
void pull_state(SlidingWindowSparseEMMatcher* m);

// the CUDA device (as this process numbers them) the matcher shim put replica 0 on; 0 before any matcher was built
int first_device();

std::mutex& handle_lock();
// the C-ABI handle of a matcher built by the shim, after pushing the members the reference header's inline
// setters may have changed (take handle_lock() first)
mbgcg_t* synced_handle(SlidingWindowSparseEMMatcher* m);
LastMatch last_match();
void forget_last_match();

}  // namespace mbgc_b200_dropin
