// tools/hostxfer_probe.cu -- one-off measurements that size the host side of the batched driver:
// context creation, cudaHostAlloc / cudaHostRegister cost, H2D from pinned vs pageable memory, file read speed.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    double t0 = now();
    cudaFree(0);
    printf("context %.1f ms\n", now() - t0);
    for (size_t mb : {16, 64, 256}) {
        void* p;
        t0 = now();
        cudaHostAlloc(&p, mb << 20, cudaHostAllocDefault);
        double t1 = now();
        memset(p, 1, mb << 20);
        double t2 = now();
        printf("cudaHostAlloc %zu MiB: %.2f ms (%.2f GB/s), first touch memset %.2f ms\n", mb, t1 - t0, (mb / 1024.0) / ((t1 - t0) / 1e3), t2 - t1);
        cudaFreeHost(p);
    }
    for (size_t mb : {64, 256}) {
        char* p = (char*)malloc(mb << 20);
        memset(p, 1, mb << 20);
        t0 = now();
        cudaHostRegister(p, mb << 20, cudaHostRegisterDefault);
        double t1 = now();
        printf("cudaHostRegister %zu MiB (touched): %.2f ms (%.2f GB/s)\n", mb, t1 - t0, (mb / 1024.0) / ((t1 - t0) / 1e3));
        cudaHostUnregister(p);
        free(p);
    }
    const size_t N = 512ull << 20;
    void* d;
    cudaMalloc(&d, N);
    void* hp;
    cudaHostAlloc(&hp, N, cudaHostAllocDefault);
    memset(hp, 2, N);
    char* pg = (char*)malloc(N);
    memset(pg, 3, N);
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int rep = 0; rep < 2; rep++) {
        t0 = now();
        cudaMemcpyAsync(d, hp, N, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        printf("H2D pinned 512 MiB: %.2f ms (%.1f GB/s)\n", now() - t0, 0.5 / ((now() - t0) / 1e3));
        t0 = now();
        cudaMemcpyAsync(d, pg, N, cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        printf("H2D pageable 512 MiB: %.2f ms (%.1f GB/s)\n", now() - t0, 0.5 / ((now() - t0) / 1e3));
        t0 = now();
        for (size_t o = 0; o < N; o += 5u << 20) cudaMemcpyAsync((char*)d + o, pg + o, std::min<size_t>(5u << 20, N - o), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        printf("H2D pageable 5 MiB pieces: %.2f ms (%.1f GB/s)\n", now() - t0, 0.5 / ((now() - t0) / 1e3));
    }
    // memcpy pageable -> pinned with T threads
    for (int T : {1, 2, 4, 8, 16}) {
        t0 = now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back([&, t] { size_t per = N / T; memcpy((char*)hp + t * per, pg + t * per, per); });
        for (auto& x : th) x.join();
        printf("memcpy pageable->pinned, %d threads: %.2f ms (%.1f GB/s)\n", T, now() - t0, 0.5 / ((now() - t0) / 1e3));
    }
    // file read (page cache) into pinned / malloc
    FILE* f = fopen("/tmp/probe.bin", "wb");
    fwrite(pg, 1, 64 << 20, f);
    fclose(f);
    for (int rep = 0; rep < 2; rep++) {
        t0 = now();
        f = fopen("/tmp/probe.bin", "rb");
        size_t r = fread(hp, 1, 64 << 20, f);
        fclose(f);
        printf("fread 64 MiB into pinned: %.2f ms (%.1f GB/s) %zu\n", now() - t0, 0.0625 / ((now() - t0) / 1e3), r);
    }
    int nd = 0;
    cudaGetDeviceCount(&nd);
    printf("devices %d, hw threads %u\n", nd, std::thread::hardware_concurrency());
    for (int dv = 1; dv < nd; dv++) {
        t0 = now();
        cudaSetDevice(dv);
        cudaFree(0);
        printf("context dev %d: %.1f ms\n", dv, now() - t0);
    }
    return 0;
}
