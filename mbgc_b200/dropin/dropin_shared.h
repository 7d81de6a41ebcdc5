// mbgc_b200/dropin/dropin_shared.h -- what the two drop-in translation units share: the handle behind a reference
// matcher object, the mutex that serialises calls on it, and the identity of the last matchTexts call.
#pragma once
#include <cstddef>
#include <cstdint>
#include <mutex>

#include "mbgc_b200.h"

class SlidingWindowSparseEMMatcher;

namespace mbgc_b200_dropin {

struct LastMatch {
    const SlidingWindowSparseEMMatcher* matcher;
    const char* text;
    size_t len;
    uint64_t n;
};

std::mutex& handle_lock();
// the C-ABI handle of a matcher built by the shim, after pushing the members the reference header's inline
// setters may have changed (take handle_lock() first)
mbgcg_t* synced_handle(SlidingWindowSparseEMMatcher* m);
LastMatch last_match();
void forget_last_match();

}  // namespace mbgc_b200_dropin
