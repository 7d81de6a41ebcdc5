// oracle/ref_matcher_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" driver over the REFERENCE's own matcher class, compiled by oracle/Makefile together
// with /root/reference/matching/SlidingWindowSparseEMMatcher.cpp (unmodified, from where it lies) into
// oracle/_ref/libref_matcher.so.  It is the match-list / hash-table known-answer source that the C
// restatement (oracle/mbgc_oracle.c) and the CUDA path are pinned against.
// It contains no algorithm of its own: every call forwards to the reference class
// (matching/SlidingWindowSparseEMMatcher.h:86-126).  Run with OMP_NUM_THREADS=1: the reference's index
// build is racy with more threads.
#include "matching/SlidingWindowSparseEMMatcher.h"
#include <omp.h>
#include <cstring>

// Agner Fog's asmlib ships only as a non-PIC static archive (libs/libaelf64.a) and cannot be linked into
// a shared object.  Its two entry points used by the matcher are exactly memcmp/memcpy -- the reference
// itself defines them so on ARM (utils/helper.cpp:598-608) -- so the .so provides those.  The stock
// oracle/_ref/mbgc binary links the real asmlib and pins the same results at archive level.
extern "C" void* A_memcpy(void* dest, const void* src, size_t n) { return memcpy(dest, src, n); }
extern "C" int A_memcmp(const void* s1, const void* s2, size_t n) { return memcmp(s1, s2, n); }

namespace {
// Subclasses only to reach the protected table pointer / size for table-level diffs.
struct ExposedExp : public SlidingWindowExpSparseEMMatcher {
    using SlidingWindowExpSparseEMMatcher::SlidingWindowExpSparseEMMatcher;
    const uint32_t* table() const { return ht64bitFlag ? nullptr : ht32bit; }
    uint64_t tableSize() const { return hash_size; }
    uint64_t sampling() const { return samplingPos; }
    int reached() const { return reachedRefLengthCount; }
    int kmer() const { return K; }
};
struct ExposedPlain : public SlidingWindowSparseEMMatcher {
    using SlidingWindowSparseEMMatcher::SlidingWindowSparseEMMatcher;
    const uint32_t* table() const { return ht64bitFlag ? nullptr : ht32bit; }
    uint64_t tableSize() const { return hash_size; }
    uint64_t sampling() const { return samplingPos; }
    int reached() const { return reachedRefLengthCount; }
    int kmer() const { return K; }
};
struct Handle {
    SlidingWindowSparseEMMatcher* m;
    ExposedExp* e;
    ExposedPlain* p;
    std::vector<PgTools::TextMatch> res;
};
}

extern "C" {

void* refm_create(uint64_t refLengthLimit, uint32_t L, int k1, int k2, int skipMargin) {
    PgHelpers::devout = &null_stream;
    PgHelpers::appout = &null_stream;
    omp_set_num_threads(1);
    Handle* h = new Handle();
    // same selection as MultipleGenomeMatchingProcessor.cpp:170-176
    if (__builtin_ctz((uint32_t) k1)) {
        h->e = new ExposedExp(refLengthLimit, L, k1, k2, skipMargin);
        h->p = nullptr;
        h->m = h->e;
    } else {
        h->p = new ExposedPlain(refLengthLimit, L, k1, k2, skipMargin);
        h->e = nullptr;
        h->m = h->p;
    }
    return h;
}

void refm_destroy(void* hv) { Handle* h = (Handle*) hv; delete h->m; delete h; }

void refm_disable_sliding_window(void* hv) { ((Handle*) hv)->m->disableSlidingWindow(); }
void refm_disable_circular_buffer(void* hv) { ((Handle*) hv)->m->disableCircularBuffer(); }
void refm_set_sliding_window_size(void* hv, int factor) { ((Handle*) hv)->m->setSlidingWindowSize((uint8_t) factor); }

void refm_load_ref(void* hv, const char* text, uint64_t len, int loadRC, int addSep, char sep) {
    ((Handle*) hv)->m->loadRef(text, len, loadRC != 0, addSep != 0, sep);
}
void refm_load_separator(void* hv, char sep) { ((Handle*) hv)->m->loadSeparator(sep); }

// returns number of matches; triplets (posSrcText, length, posDestText) via refm_get_matches
uint64_t refm_match(void* hv, const char* text, uint64_t len, uint32_t minMatchLength, uint64_t lockPos) {
    Handle* h = (Handle*) hv;
    h->m->matchTexts(h->res, text, len, false, false, minMatchLength, (size_t) lockPos);
    return h->res.size();
}
void refm_get_matches(void* hv, uint64_t* out3) {
    Handle* h = (Handle*) hv;
    for (size_t i = 0; i < h->res.size(); i++) {
        out3[3 * i] = h->res[i].posSrcText;
        out3[3 * i + 1] = h->res[i].length;
        out3[3 * i + 2] = h->res[i].posDestText;
    }
}

// n independent matchTexts calls against the frozen matcher state on `threads` OpenMP threads -- what the
// reference's parallel workers do between two reference extensions (MultipleGenomeMatchingProcessor.cpp:379).
// Used as the all-cores CPU baseline of the match stage.  Returns the total number of matches.
uint64_t refm_match_many(void* hv, int n, const char* const* texts, const uint64_t* lens, uint32_t minMatchLength,
                         int threads, uint64_t* countsOut) {
    Handle* h = (Handle*) hv;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : total)
    for (int i = 0; i < n; i++) {
        std::vector<PgTools::TextMatch> res;
        h->m->matchTexts(res, texts[i], lens[i], false, false, minMatchLength);
        if (countsOut) countsOut[i] = res.size();
        total += res.size();
    }
    return total;
}

const char* refm_ref(void* hv) { return ((Handle*) hv)->m->getRef(); }
uint64_t refm_ref_length(void* hv) { return ((Handle*) hv)->m->getRefLength(); }
uint64_t refm_loading_position(void* hv) { return ((Handle*) hv)->m->getLoadingPosition(); }
uint64_t refm_loaded_ref_length(void* hv) { return ((Handle*) hv)->m->getLoadedRefLength(); }
uint64_t refm_max_ref_length(void* hv) { return ((Handle*) hv)->m->getMaxRefLength(); }
uint64_t refm_acquire_lock(void* hv) { return ((Handle*) hv)->m->acquireWorkerMatchingLockPos(); }
void refm_release_lock(void* hv, uint64_t v) { ((Handle*) hv)->m->releaseWorkerMatchingLockPos((size_t) v); }
void refm_set_position(void* hv, uint64_t pos, int reached) { ((Handle*) hv)->m->setPosition((size_t) pos, reached); }

const uint32_t* refm_table(void* hv) { Handle* h = (Handle*) hv; return h->e ? h->e->table() : h->p->table(); }
uint64_t refm_table_size(void* hv) { Handle* h = (Handle*) hv; return h->e ? h->e->tableSize() : h->p->tableSize(); }
uint64_t refm_sampling_pos(void* hv) { Handle* h = (Handle*) hv; return h->e ? h->e->sampling() : h->p->sampling(); }
int refm_reached_count(void* hv) { Handle* h = (Handle*) hv; return h->e ? h->e->reached() : h->p->reached(); }
int refm_kmer(void* hv) { Handle* h = (Handle*) hv; return h->e ? h->e->kmer() : h->p->kmer(); }

}  // extern "C"
