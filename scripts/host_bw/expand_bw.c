// host memory experiment: how fast can T threads expand u16 -> i32 (read 2 B, write 4 B per element)?
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
typedef struct { const uint16_t* src; int32_t* dst; size_t n; } job_t;
static void* run(void* p) { job_t* j = (job_t*)p; for (size_t i = 0; i < j->n; ++i) j->dst[i] = j->src[i]; return 0; }
static void* cpy(void* p) { job_t* j = (job_t*)p; memcpy(j->dst, j->src, j->n * 4); return 0; }
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
    int T = argc > 1 ? atoi(argv[1]) : 16;
    size_t n = (size_t)900 << 20;   // 900 Mi elements: 1.9 GB in, 3.8 GB out
    uint16_t* src = malloc(n * 4); int32_t* dst = malloc(n * 4);
    memset(src, 1, n * 4); memset(dst, 0, n * 4);
    pthread_t th[64]; job_t jobs[64];
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            double t0 = now();
            for (int t = 0; t < T; ++t) { size_t lo = n * t / T, hi = n * (t + 1) / T; jobs[t] = (job_t){ src + (mode ? lo * 2 : lo), dst + lo, hi - lo }; pthread_create(&th[t], 0, mode ? cpy : run, &jobs[t]); }
            for (int t = 0; t < T; ++t) pthread_join(th[t], 0);
            double dt = now() - t0;
            printf("%s threads %d: %.1f ms, write %.1f GB/s\n", mode ? "memcpy 3.8GB" : "expand u16->i32", T, dt * 1e3, n * 4 / dt / 1e9);
        }
    return 0;
}
