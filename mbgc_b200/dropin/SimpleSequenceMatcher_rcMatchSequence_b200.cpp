// mbgc_b200/dropin/SimpleSequenceMatcher_rcMatchSequence_b200.cpp
//
// Link-time substitute for ONE static member of the reference: SimpleSequenceMatcher::rcMatchSequence
// (/root/reference/matching/SimpleSequenceMatcher.cpp:165-176), the reverse-complement redundancy pass mode 3 runs on the
// literal stream right before the back-end coders (mbgccoder/MBGC_Encoder.cpp:637-638).  Compiled against the UNMODIFIED
// header; the symbol of the stock function is made weak in the reference's own object (oracle/Makefile target
// `dropin_batched`).  It forwards to mbgcg_rc_match_sequence (include/mbgc_b200.h; CUDA kernels of
// mbgc_b200/csrc/rcmatch.cu): index build and probe / verify / extend on the device, the match-level bookkeeping on the host.
//
// By the time the encoder gets here the driver has deleted its matcher (MultipleGenomeMatchingProcessor.cpp:594), so a
// small handle of its own is created on the same device (the CUDA context is still there).  Results are those of the
// reference's single-threaded index build, i.e. of `mbgc -t 1`; the stock multi-threaded build races.
#include "SimpleSequenceMatcher.h"

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "dropin_shared.h"
#include "mbgc_b200.h"

namespace PgTools {

void SimpleSequenceMatcher::rcMatchSequence(string& sequence, string& rcMapOff, string& rcMapLen, size_t targetMatchLength,
                                            uint32_t minMatchLength) {
    auto die = [](const char* msg) {
        fprintf(stderr, "%s\n\n", msg);
        exit(EXIT_FAILURE);
    };
    const int device = mbgc_b200_dropin::first_device();  // where the matcher shim put its first replica
    mbgcg_t* h = nullptr;
    if (mbgcg_create(&h, 1 << 20, 32, 16, 1, 16, device) != MBGCG_OK) die(mbgcg_create_error());
    const size_t n = sequence.size();
    const size_t cap = n / 24 + 4;
    std::vector<uint8_t> off(4 * cap), len(10 * cap);
    uint64_t newLen = 0, nOff = 0, nLen = 0;
    float ms[2];
    if (mbgcg_rc_match_sequence(h, reinterpret_cast<uint8_t*>(&sequence[0]), n, static_cast<uint32_t>(targetMatchLength), minMatchLength,
                                &newLen, off.data(), &nOff, len.data(), &nLen, ms) != MBGCG_OK)
        die(mbgcg_last_error(h));
    if (std::getenv("MBGC_B200_TRACE"))
        fprintf(stderr, "[mbgc_b200] rcMatchSequence: %zu -> %llu bytes, %llu matches; device: index %.2f ms, query %.2f ms\n", n,
                (unsigned long long) newLen, (unsigned long long) (nOff / 4), ms[0], ms[1]);
    sequence.resize(newLen);
    rcMapOff.assign(reinterpret_cast<const char*>(off.data()), nOff);
    rcMapLen.assign(reinterpret_cast<const char*>(len.data()), nLen);
    mbgcg_destroy(h);
}

}  // namespace PgTools
