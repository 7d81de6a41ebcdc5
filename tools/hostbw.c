/* tools/hostbw.c -- host memory read bandwidth with all cores (what any host-side pass over the FASTA text,
 * including the reference's own memcmp-style matching of near-identical genomes, is bounded by).
 *   gcc -O3 -mavx2 -fopenmp tools/hostbw.c -o /tmp/hostbw && /tmp/hostbw [MB] */
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char** argv) {
    size_t mb = argc > 1 ? (size_t)atol(argv[1]) : 1024;
    size_t n = mb << 20;
    uint8_t* a = malloc(n);
    uint8_t* b = malloc(n / 4);
#pragma omp parallel for
    for (size_t i = 0; i < n; i += 4096) memset(a + i, 'A' + (i >> 12) % 7, 4096);
    memset(b, 0, n / 4);
    int threads = omp_get_max_threads();
    for (int rep = 0; rep < 3; rep++) {
        uint64_t total = 0;
        double t0 = now();
#pragma omp parallel for reduction(+ : total) schedule(static)
        for (size_t blk = 0; blk < n / 65536; blk++) {
            const uint64_t* p = (const uint64_t*)(a + blk * 65536);
            uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            for (size_t k = 0; k < 8192; k += 4) { s0 += p[k]; s1 += p[k + 1]; s2 += p[k + 2]; s3 += p[k + 3]; }
            total += s0 + s1 + s2 + s3;
        }
        double t1 = now();
        /* read + 1/4 write (the shape of a 2-bit packer) */
#pragma omp parallel for schedule(static)
        for (size_t blk = 0; blk < n / 65536; blk++) {
            const uint64_t* p = (const uint64_t*)(a + blk * 65536);
            uint64_t* q = (uint64_t*)(b + blk * 16384);
            for (size_t k = 0; k < 2048; k++) q[k] = p[4 * k] ^ p[4 * k + 1] ^ p[4 * k + 2] ^ p[4 * k + 3];
        }
        double t2 = now();
        printf("threads %d: read %.1f GB/s (sum %llu), read+quarter-write %.1f GB/s of input\n", threads,
               n / (t1 - t0) / 1e9, (unsigned long long)total, n / (t2 - t1) / 1e9);
    }
    return 0;
}
