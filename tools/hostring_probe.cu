// tools/hostring_probe.cu -- experiment behind the packed host -> device text transfer.
// Question: the transfer is bound by host memory traffic (read 1 B + write 1/4 B + DMA read 1/4 B per packed base).  If the
// packers write into a SMALL REUSED ring of pinned slots with ordinary stores, do the packed bytes stay in the last-level
// cache until the copy engine has fetched them -- i.e. does the 1/2 B per base of staging traffic disappear?
//   mode A  big pinned staging buffer, non-temporal stores, copies of finished blocks in order (what stageTexts does)
//   mode B  ring of S slots per thread, ordinary stores, one copy + one event per block, slots reused when their copy is done
//   mode R  no packing: one DMA of the raw text
// nvcc -O3 -std=c++17 -Xcompiler -mavx2,-pthread tools/hostring_probe.cu -o tools/hostring_probe
#include <cuda_runtime.h>
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int g_pfd = 0;  // software prefetch distance in bytes (0 = none)

template <bool NT>
static void pack(const uint8_t* src, uint64_t n, uint8_t* dst) {
    const int pfd = g_pfd;
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);
    const __m256i w116 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1,
                                            -1, -1, -1, -1, -1, -1);
    const __m256i lut = _mm256_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0x41, -1, 0x43, 0x54, -1,
                                         -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i lowNib = _mm256_set1_epi8(0x0f);
    uint32_t badAll = 0;
    for (uint64_t i = 0; i + 64 <= n; i += 64) {
        uint32_t out[4];
        if (pfd) _mm_prefetch(reinterpret_cast<const char*>(src + i + pfd), _MM_HINT_T0);
        for (int h = 0; h < 2; h++) {
            const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * h));
            const __m256i expect = _mm256_shuffle_epi8(lut, _mm256_and_si256(v, lowNib));
            badAll |= ~static_cast<uint32_t>(_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v)));
            const __m256i codes = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
            const __m256i pb = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(codes, w14), w116), gather);
            out[2 * h] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 0));
            out[2 * h + 1] = static_cast<uint32_t>(_mm256_extract_epi32(pb, 4));
        }
        const __m128i o = _mm_loadu_si128(reinterpret_cast<const __m128i*>(out));
        if (NT)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
        else
            _mm_store_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), o);
    }
    if (NT) _mm_sfence();
    if (badAll) dst[0] ^= 1;  // keep the validity test alive
}

int main(int argc, char** argv) {
    const uint64_t N = 512ull << 20;
    const int T = argc > 1 ? atoi(argv[1]) : std::max(1, static_cast<int>(std::thread::hardware_concurrency()) - 1);
    const uint64_t BLK = (argc > 2 ? atoi(argv[2]) : 1024) * 1024ull;  // input bytes per block
    const int S = argc > 3 ? atoi(argv[3]) : 2;                          // ring slots per thread
    const int rawPermille = argc > 4 ? atoi(argv[4]) : 0;                // share of the blocks that travels raw in mode B
    g_pfd = argc > 5 ? atoi(argv[5]) : 0;
    cudaFree(0);
    uint8_t *hSrc, *hBig, *hRing, *dText, *dPacked;
    cudaHostAlloc(&hSrc, N, cudaHostAllocDefault);
    cudaHostAlloc(&hBig, N / 4, cudaHostAllocDefault);
    const uint64_t slotBytes = BLK / 4;
    cudaHostAlloc(&hRing, static_cast<uint64_t>(T) * S * slotBytes, cudaHostAllocDefault);
    cudaMalloc(&dText, N);
    cudaMalloc(&dPacked, N / 4);
    uint32_t x = 12345;
    for (uint64_t i = 0; i < N; i++) {
        x = x * 1664525u + 1013904223u;
        hSrc[i] = "ACGT"[x >> 30];
    }
    memset(hBig, 0, N / 4);
    memset(hRing, 0, static_cast<uint64_t>(T) * S * slotBytes);
    cudaStream_t st;
    cudaStreamCreate(&st);
    const uint32_t nB = static_cast<uint32_t>(N / BLK);
    {   // packers alone (no copies): what the cores can read
        for (int pf : {0, 256, 512, 1024, 2048, 4096, 8192}) {
            g_pfd = pf;
            std::atomic<uint32_t> next{0};
            const uint32_t nBk = static_cast<uint32_t>(N / BLK);
            const double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&] {
                    for (;;) {
                        const uint32_t b = next.fetch_add(1);
                        if (b >= nBk) return;
                        pack<true>(hSrc + b * BLK, BLK, hBig + b * (BLK / 4));
                    }
                });
            for (auto& t : th) t.join();
            const double t = now() - t0;
            printf("packers alone, prefetch %5d B: %.2f ms  %.1f GB/s (%.2f GB/s per thread)\n", pf, t, N / t / 1e6, N / t / 1e6 / T);
        }
        g_pfd = argc > 5 ? atoi(argv[5]) : 0;
    }
    printf("threads %d, block %llu KiB in / %llu KiB packed, %d slots per thread (ring %.1f MiB), %u blocks\n", T,
           (unsigned long long) BLK >> 10, (unsigned long long) slotBytes >> 10, S, T * S * slotBytes / 1048576.0, nB);

    for (int rep = 0; rep < 3; rep++) {
        {   // mode R
            cudaStreamSynchronize(st);
            const double t0 = now();
            cudaMemcpyAsync(dText, hSrc, N, cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
            const double t = now() - t0;
            printf("R raw DMA                : %.2f ms  %.1f GB/s\n", t, N / t / 1e6);
        }
        {   // mode A
            std::vector<std::atomic<uint8_t>> done(nB);
            for (auto& d : done) d.store(0);
            std::atomic<uint32_t> next{0};
            const double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&] {
                    for (;;) {
                        const uint32_t b = next.fetch_add(1);
                        if (b >= nB) return;
                        pack<true>(hSrc + b * BLK, BLK, hBig + b * slotBytes);
                        done[b].store(1, std::memory_order_release);
                    }
                });
            uint32_t pi = 0;
            while (pi < nB) {
                if (done[pi].load(std::memory_order_acquire)) {
                    uint32_t lo = pi;
                    while (pi < nB && done[pi].load(std::memory_order_acquire) && (pi - lo) * slotBytes < (4u << 20)) pi++;
                    cudaMemcpyAsync(dPacked + lo * slotBytes, hBig + lo * slotBytes, (pi - lo) * slotBytes, cudaMemcpyHostToDevice, st);
                } else
                    std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
            for (auto& t : th) t.join();
            cudaStreamSynchronize(st);
            const double t = now() - t0;
            printf("A big buffer, NT stores  : %.2f ms  %.1f GB/s of text\n", t, N / t / 1e6);
        }
        {   // mode B
            const int nSlots = T * S;
            std::vector<std::atomic<int>> slotState(nSlots);  // 0 free, 1 ready (packed), 2 in flight
            std::vector<uint32_t> slotBlock(nSlots);
            std::vector<cudaEvent_t> ev(nSlots);
            for (int i = 0; i < nSlots; i++) {
                slotState[i].store(0);
                cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            }
            std::atomic<uint32_t> next{0};
            std::atomic<uint32_t> packedDone{0};
            // raw blocks: the tail of the block list, fed by the main thread whenever nothing packed is ready
            const uint32_t nRaw = static_cast<uint32_t>(static_cast<uint64_t>(nB) * rawPermille / 1000);
            const uint32_t nPk = nB - nRaw;
            const double t0 = now();
            std::vector<std::thread> th;
            for (int t = 0; t < T; t++)
                th.emplace_back([&, t] {
                    int s = 0;
                    for (;;) {
                        const uint32_t b = next.fetch_add(1);
                        if (b >= nPk) return;
                        const int slot = t * S + s;
                        s = (s + 1) % S;
                        while (slotState[slot].load(std::memory_order_acquire) != 0) _mm_pause();
                        pack<false>(hSrc + b * BLK, BLK, hRing + slot * slotBytes);
                        slotBlock[slot] = b;
                        slotState[slot].store(1, std::memory_order_release);
                    }
                });
            uint32_t issued = 0, rawIssued = 0;
            std::vector<int> inflight;
            while (issued < nPk || rawIssued < nRaw || !inflight.empty()) {
                bool any = false;
                for (int slot = 0; slot < nSlots; slot++)
                    if (slotState[slot].load(std::memory_order_acquire) == 1) {
                        cudaMemcpyAsync(dPacked + slotBlock[slot] * slotBytes, hRing + slot * slotBytes, slotBytes, cudaMemcpyHostToDevice, st);
                        cudaEventRecord(ev[slot], st);
                        slotState[slot].store(2, std::memory_order_relaxed);
                        inflight.push_back(slot);
                        issued++;
                        any = true;
                    }
                for (size_t k = 0; k < inflight.size();) {
                    if (cudaEventQuery(ev[inflight[k]]) == cudaSuccess) {
                        slotState[inflight[k]].store(0, std::memory_order_release);
                        inflight[k] = inflight.back();
                        inflight.pop_back();
                    } else
                        k++;
                }
                if (!any && rawIssued < nRaw) {
                    const uint32_t b = nPk + rawIssued++;
                    cudaMemcpyAsync(dText + b * BLK, hSrc + b * BLK, BLK, cudaMemcpyHostToDevice, st);
                    any = true;
                }
                if (!any) _mm_pause();
            }
            for (auto& t : th) t.join();
            cudaStreamSynchronize(st);
            const double t = now() - t0;
            printf("B ring, ordinary stores  : %.2f ms  %.1f GB/s of text (raw share %d permille)\n", t, N / t / 1e6, rawPermille);
            for (auto& e : ev) cudaEventDestroy(e);
        }
    }
    return 0;
}
