// mbgc_b200/csrc/hostpack.cpp -- see hostpack.hpp.  Plain C++ (g++), AVX2 where the CPU has it.
#include "hostpack.hpp"

#include <immintrin.h>

#include <cstring>

namespace mbgc_b200 {

namespace {

inline bool is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// appends byte c at device offset off to the run list; false when the list is full
inline bool add_exception(std::vector<ExcRun>& runs, size_t maxRuns, uint64_t off, uint8_t c) {
    if (!runs.empty()) {
        ExcRun& last = runs.back();
        if (last.off + last.len == off && last.byte == c) {
            last.len++;
            return true;
        }
    }
    if (runs.size() >= maxRuns) return false;
    runs.push_back(ExcRun{off, 1u, c});
    return true;
}

bool pack_scalar(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t devOff, std::vector<ExcRun>& runs, size_t maxRuns,
                 uint64_t from) {
    // from is a multiple of 4
    for (uint64_t i = from; i < n; i += 4) {
        uint8_t b = 0;
        for (uint64_t k = 0; k < 4 && i + k < n; k++) {
            const uint8_t c = src[i + k];
            b |= static_cast<uint8_t>(((c >> 1) & 3) << (2 * k));
            if (!is_acgt(c) && !add_exception(runs, maxRuns, devOff + i + k, c)) return false;
        }
        dst[i >> 2] = b;
    }
    return true;
}

// Software prefetch distance of the packing loops.  On the GPU boxes (16 vCPUs of a Sapphire Rapids host) a core streams
// ~5.6 GB/s out of DRAM on its own and ~8.4 GB/s with a prefetch 4 - 8 KiB ahead: 15 packers read 86 vs 125 GB/s, the
// box's DRAM limit (tools/hostring_probe.cu; sweep of the whole transfer in profiles/README.md r02c: 79 -> 104 GB/s end to
// end).  MBGC_B200_PACK_PREFETCH = bytes, 0 = none.
inline uint64_t prefetch_distance() {
    static const uint64_t d = []() -> uint64_t {
        const char* e = std::getenv("MBGC_B200_PACK_PREFETCH");
        return e ? std::strtoull(e, nullptr, 10) : 8192;
    }();
    return d;
}

__attribute__((target("avx2"))) bool pack_avx2(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t devOff,
                                               std::vector<ExcRun>& runs, size_t maxRuns) {
    const uint64_t pfd = prefetch_distance();
    // expected byte per low nibble: 'A' 0x41, 'C' 0x43, 'T' 0x54, 'G' 0x47; 0xFF where no base has that nibble
    const __m256i lut = _mm256_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1,  //
                                         -1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i lowNib = _mm256_set1_epi8(0x0f);
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);      // bytes (1, 4): c0 + 4 c1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);  // words (1, 16): (c0 + 4 c1) + 16 (c2 + 4 c3)
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,  //
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    // 64 input bytes -> 16 packed bytes per iteration, written with a non-temporal store when dst is 16-byte aligned
    // (it is, for the pinned staging buffer): the whole path is bound by host memory traffic, and a regular store
    // would first read the destination line (read-for-ownership), a quarter more traffic for nothing.
    const bool ntOk = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        uint32_t out[4];
        bool slow = false;
        if (pfd) _mm_prefetch(reinterpret_cast<const char*>(src + i + pfd), _MM_HINT_T0);  // never faults, also past the end
#pragma GCC unroll 2
        for (int h = 0; h < 2; h++) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * h));
            const __m256i expect = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, lowNib));
            const uint32_t valid = static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v)));
            const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
            const __m256i p16 = _mm256_maddubs_epi16(codes, w14);
            const __m256i p32 = _mm256_madd_epi16(p16, w116);
            const __m256i pb = _mm256_shuffle_epi8(p32, gather);
            out[2 * h] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 0));
            out[2 * h + 1] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 4));
            slow |= valid != 0xffffffffu;
        }
        const __m128i o = _mm_loadu_si128(reinterpret_cast<const __m128i*>(out));
        if (ntOk)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
        else
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
        if (slow) {
            for (int h = 0; h < 2; h++) {
                const uint64_t at = i + 32 * h;
                const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + at));
                const __m256i expect = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, lowNib));
                uint32_t bad = ~static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v)));
                if (bad == 0xffffffffu && !runs.empty()) {  // inside a long run of one byte (N): extend it in one step
                    ExcRun& last = runs.back();
                    const __m256i same = _mm256_cmpeq_epi8(v, _mm256_set1_epi8(static_cast<char>(last.byte)));
                    if (last.off + last.len == devOff + at && static_cast<uint32_t>(_mm256_movemask_epi8(same)) == 0xffffffffu) {
                        last.len += 32;
                        continue;
                    }
                }
                while (bad) {
                    const int k = __builtin_ctz(bad);
                    bad &= bad - 1;
                    if (!add_exception(runs, maxRuns, devOff + at + k, src[at + k])) return false;
                }
            }
        }
    }
    if (ntOk) _mm_sfence();
    return pack_scalar(src, n, dst, devOff, runs, maxRuns, i);
}

}  // namespace

bool pack_block(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t devOff, std::vector<ExcRun>& runs, size_t maxRuns) {
    static const bool haveAvx2 = __builtin_cpu_supports("avx2");
    if (haveAvx2) return pack_avx2(src, n, dst, devOff, runs, maxRuns);
    return pack_scalar(src, n, dst, devOff, runs, maxRuns, 0);
}

// ---- FASTA file images -------------------------------------------------------------------------------------

namespace {

struct LineEnds {  // the progression of '\n' being followed
    uint64_t first = 0, last = 0, stride = 0;
    uint32_t count = 0;
};

inline bool flush_line_ends(LineEnds& nl, std::vector<ExcRun>& runs, size_t maxRuns) {
    if (nl.count == 0) return true;
    if (runs.size() >= maxRuns) return false;
    runs.push_back(ExcRun{nl.first, nl.count, static_cast<uint32_t>('\n') | (static_cast<uint32_t>(nl.count > 1 ? nl.stride : 0) << 8)});
    nl.count = 0;
    return true;
}

constexpr uint32_t LINE_RUN_CAP = 512;   // line ends per run: the device patches a run with one warp (16 steps of 32 bytes)

inline bool on_line_end(LineEnds& nl, uint64_t p, std::vector<ExcRun>& runs, size_t maxRuns) {
    if (nl.count >= 1) {
        const uint64_t d = p - nl.last;
        if (nl.stride == 0 && d < (1u << 24)) nl.stride = d;  // the second line end of a run sets the distance
        if (d == nl.stride && nl.count < LINE_RUN_CAP) {
            nl.count++;
            nl.last = p;
            return true;
        }
    }
    const uint64_t keep = nl.count >= 1 && p - nl.last == nl.stride ? nl.stride : 0;  // a full run: the next one goes on alike
    if (!flush_line_ends(nl, runs, maxRuns)) return false;
    nl.first = nl.last = p;
    nl.count = 1;
    nl.stride = keep;
    return true;
}

inline bool image_exception(LineEnds& nl, uint64_t p, uint8_t c, std::vector<ExcRun>& runs, size_t maxRuns) {
    return c == '\n' ? on_line_end(nl, p, runs, maxRuns) : add_exception(runs, maxRuns, p, c);
}

// src[0 .. n) = bytes base .. base + n of the image (base a multiple of 64), dst = where their codes go
bool image_scalar(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t base, LineEnds& nl, std::vector<ExcRun>& runs, size_t maxRuns,
                  uint64_t from) {
    for (uint64_t i = from; i < n; i += 4) {  // from is a multiple of 4
        uint8_t b = 0;
        for (uint64_t k = 0; k < 4 && i + k < n; k++) {
            const uint8_t c = src[i + k];
            b |= static_cast<uint8_t>(((c >> 1) & 3) << (2 * k));
            if (!is_acgt(c) && !image_exception(nl, base + i + k, c, runs, maxRuns)) return false;
        }
        dst[i >> 2] = b;
    }
    return true;
}

__attribute__((target("avx2"))) bool image_avx2(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t base, LineEnds& nl,
                                                std::vector<ExcRun>& runs, size_t maxRuns) {
    const __m256i lut = _mm256_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1,  //
                                         -1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i lowNib = _mm256_set1_epi8(0x0f);
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i newline = _mm256_set1_epi8('\n');
    const __m256i w14 = _mm256_set1_epi16(0x0401);
    const __m256i w116 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,  //
                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    const bool ntOk = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const uint64_t pfd = prefetch_distance();
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        uint32_t out[4];
        uint64_t bad = 0, nls = 0;
        if (pfd) _mm_prefetch(reinterpret_cast<const char*>(src + i + pfd), _MM_HINT_T0);
#pragma GCC unroll 2
        for (int h = 0; h < 2; h++) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * h));
            const __m256i expect = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, lowNib));
            const uint32_t valid = static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v)));
            const uint32_t isNl = static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(newline, v)));
            const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
            const __m256i p16 = _mm256_maddubs_epi16(codes, w14);
            const __m256i p32 = _mm256_madd_epi16(p16, w116);
            const __m256i pb = _mm256_shuffle_epi8(p32, gather);
            out[2 * h] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 0));
            out[2 * h + 1] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 4));
            bad |= static_cast<uint64_t>(~valid) << (32 * h);
            nls |= static_cast<uint64_t>(isNl) << (32 * h);
        }
        const __m128i o = _mm_loadu_si128(reinterpret_cast<const __m128i*>(out));
        if (ntOk)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
        else
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
        if (bad == 0) continue;
        // the one byte that is not a base is the line end the progression predicts: count it and go on
        if (bad == nls && (bad & (bad - 1)) == 0 && nl.count >= 1 && nl.stride && nl.count < LINE_RUN_CAP &&
            base + i + static_cast<uint64_t>(__builtin_ctzll(bad)) == nl.last + nl.stride) {
            nl.count++;
            nl.last += nl.stride;
            continue;
        }
        while (bad) {
            const int k = __builtin_ctzll(bad);
            // a long run of one byte (N): take all of it that lies in this group in one step
            int runLen = 1;
            const uint8_t c = src[i + k];
            if (c != '\n') {
                while (k + runLen < 64 && ((bad >> (k + runLen)) & 1) && src[i + k + runLen] == c) runLen++;
            }
            if (runLen > 1) {
                bool merged = false;
                if (!runs.empty()) {
                    ExcRun& last = runs.back();
                    if (last.off + last.len == base + i + k && last.byte == c && last.len <= 0xFFFFFFFFu - 64) {
                        last.len += static_cast<uint32_t>(runLen);
                        merged = true;
                    }
                }
                if (!merged) {
                    if (runs.size() >= maxRuns) return false;
                    runs.push_back(ExcRun{base + i + static_cast<uint64_t>(k), static_cast<uint32_t>(runLen), c});
                }
                bad &= runLen + k >= 64 ? 0 : ~((1ull << (k + runLen)) - 1);
            } else {
                if (!image_exception(nl, base + i + k, c, runs, maxRuns)) return false;
                bad &= bad - 1;
            }
        }
    }
    if (ntOk) _mm_sfence();
    return image_scalar(src, n, dst, base, nl, runs, maxRuns, i);
}

struct ImagePackState {
    LineEnds nl;
    uint64_t done = 0;
};

}  // namespace

ImagePacker::ImagePacker(uint8_t* dst, std::vector<ExcRun>& runs, size_t maxRuns)
    : dst(dst), runs(runs), maxRuns(maxRuns), state(new ImagePackState()) {}

ImagePacker::~ImagePacker() { delete static_cast<ImagePackState*>(state); }

bool ImagePacker::add(const uint8_t* src, uint64_t n) {
    static const bool haveAvx2 = __builtin_cpu_supports("avx2");
    ImagePackState& st = *static_cast<ImagePackState*>(state);
    if (st.done % 64) return false;  // only the last piece may have a length that is not a multiple of 64
    const bool ok = haveAvx2 ? image_avx2(src, n, dst + st.done / 4, st.done, st.nl, runs, maxRuns)
                             : image_scalar(src, n, dst + st.done / 4, st.done, st.nl, runs, maxRuns, 0);
    st.done += n;
    return ok;
}

bool ImagePacker::finish() { return flush_line_ends(static_cast<ImagePackState*>(state)->nl, runs, maxRuns); }

bool pack_fasta_image(const uint8_t* src, uint64_t n, uint8_t* dst, std::vector<ExcRun>& runs, size_t maxRuns) {
    ImagePacker p(dst, runs, maxRuns);
    return p.add(src, n) && p.finish();
}

// ---- reverse complement ----------------------------------------------------------------------------------

namespace {

// letters[i] = what 0x40 + i and 0x60 + i map to (0 = both stay as they are); false if the table is not of that shape
bool letter_table(const uint8_t lut[256], uint8_t letters[32]) {
    for (int i = 0; i < 256; i++) {
        const bool letterRange = (i & 0xC0) == 0x40;
        if (!letterRange && lut[i] != i) return false;
    }
    for (int i = 0; i < 32; i++) {
        const int up = 0x40 + i, lo = 0x60 + i;
        if (lo == 0x7F) {  // the table's hole: 0x7F -> 0 (handled separately); its upper partner must be identity
            if (lut[lo] != 0 || lut[up] != up) return false;
            letters[i] = 0;
            continue;
        }
        if (lut[up] == up && lut[lo] == lo) letters[i] = 0;
        else if (lut[up] == lut[lo] && lut[up] != 0) letters[i] = lut[up];
        else return false;
    }
    return true;
}

__attribute__((target("avx2"))) void revcomp_avx2(const uint8_t* src, uint64_t n, uint8_t* dst, const uint8_t lut[256],
                                                  const uint8_t letters[32]) {
    const __m256i tblLo = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(letters)));
    const __m256i tblHi = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(letters + 16)));
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,  //
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i c40 = _mm256_set1_epi8(0x40), cC0 = _mm256_set1_epi8(static_cast<char>(0xC0));
    const __m256i c0F = _mm256_set1_epi8(0x0F), c7F = _mm256_set1_epi8(0x7F), zero = _mm256_setzero_si256();
    uint64_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i inRange = _mm256_cmpeq_epi8(_mm256_and_si256(v, cC0), c40);
        const __m256i nib = _mm256_and_si256(v, c0F);
        const __m256i lo = _mm256_shuffle_epi8(tblLo, nib), hi = _mm256_shuffle_epi8(tblHi, nib);
        const __m256i t = _mm256_blendv_epi8(lo, hi, _mm256_slli_epi16(v, 3));  // bit 4 of the byte -> bit 7: upper half
        const __m256i use = _mm256_andnot_si256(_mm256_cmpeq_epi8(t, zero), inRange);
        __m256i c = _mm256_blendv_epi8(v, t, use);
        c = _mm256_andnot_si256(_mm256_cmpeq_epi8(v, c7F), c);  // 0x7F -> 0
        c = _mm256_shuffle_epi8(c, rev);
        c = _mm256_permute2x128_si256(c, c, 1);
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (n - 32 - i)), c);
    }
    for (; i < n; i++) dst[n - 1 - i] = lut[src[i]];
}

}  // namespace

void reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst, const uint8_t lut[256]) {
    static const bool haveAvx2 = __builtin_cpu_supports("avx2");
    uint8_t letters[32];
    if (haveAvx2 && n >= 64 && letter_table(lut, letters)) {
        revcomp_avx2(src, n, dst, lut, letters);
        return;
    }
    for (uint64_t i = 0; i < n; i++) dst[n - 1 - i] = lut[src[i]];
}

// ---- worker pool ---------------------------------------------------------------------------------------

WorkerPool::WorkerPool(int nThreads) {
    if (nThreads < 1) nThreads = 1;
    for (int t = 0; t < nThreads; t++) threads.emplace_back([this, t] { loop(t); });
}

WorkerPool::~WorkerPool() {
    {
        std::lock_guard<std::mutex> lk(mu);
        stop = true;
    }
    cvStart.notify_all();
    for (auto& t : threads) t.join();
}

void WorkerPool::run(uint32_t n, std::function<void(uint32_t, int)> fn) {
    wait();
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(fn);
    nItems = n;
    if (doneFlags.size() < n) doneFlags = std::vector<std::atomic<uint8_t>>(n + n / 2 + 16);
    for (uint32_t i = 0; i < n; i++) doneFlags[i].store(0, std::memory_order_relaxed);
    next.store(0, std::memory_order_relaxed);
    active = static_cast<int>(threads.size());
    generation++;
    cvStart.notify_all();
}

void WorkerPool::wait() {
    std::unique_lock<std::mutex> lk(mu);
    cvDone.wait(lk, [this] { return active == 0; });
}

void WorkerPool::loop(int worker) {
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cvStart.wait(lk, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
        }
        for (;;) {
            const uint32_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= nItems) break;
            job(i, worker);
            doneFlags[i].store(1, std::memory_order_release);
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            if (--active == 0) cvDone.notify_all();
        }
    }
}

}  // namespace mbgc_b200
