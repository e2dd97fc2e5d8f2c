// Host micro-benchmark: how fast can T threads expand a uint8 observation transport buffer to float32 (the widening step of the
// packed end-to-end path)?  gcc -O3 -pthread -o expand_probe expand_probe.c ; ./expand_probe
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static const uint8_t *g_in;
static float *g_out;
static size_t g_n;
static int g_T;
static pthread_barrier_t g_bar;

__attribute__((target("avx2"))) static void expand_avx2(const uint8_t *in, float *out, size_t n) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256i b = _mm256_loadu_si256((const __m256i *)(in + i));
        __m128i lo = _mm256_castsi256_si128(b), hi = _mm256_extracti128_si256(b, 1);
        _mm256_stream_ps(out + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(lo)));
        _mm256_stream_ps(out + i + 8, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(lo, 8))));
        _mm256_stream_ps(out + i + 16, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(hi)));
        _mm256_stream_ps(out + i + 24, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(hi, 8))));
    }
    for (; i < n; i++) out[i] = (float)in[i];
    _mm_sfence();
}

static void *worker(void *arg) {
    long t = (long)arg;
    size_t per = ((g_n / g_T) + 31) & ~(size_t)31, a = t * per, b = a + per > g_n ? g_n : a + per;
    for (int rep = 0; rep < 12; rep++) {
        pthread_barrier_wait(&g_bar);
        if (a < g_n) expand_avx2(g_in + a, g_out + a, b - a);
        pthread_barrier_wait(&g_bar);
    }
    return NULL;
}

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }

int main(void) {
    size_t n = 65536ull * 4 * 71;
    uint8_t *in = aligned_alloc(64, n);
    float *out = aligned_alloc(64, n * 4);
    memset(in, 1, n);
    memset(out, 0, n * 4);  // first touch by the main thread, like a caller-allocated pinned buffer
    g_in = in; g_out = out; g_n = n;
    int Ts[] = {1, 4, 8, 16, 32, 64};
    for (int k = 0; k < 6; k++) {
        g_T = Ts[k];
        pthread_barrier_init(&g_bar, NULL, g_T + 1);
        pthread_t th[64];
        for (long t = 0; t < g_T; t++) pthread_create(&th[t], NULL, worker, (void *)t);
        double best = 1e9;
        for (int rep = 0; rep < 12; rep++) {
            pthread_barrier_wait(&g_bar);
            double t0 = now();
            pthread_barrier_wait(&g_bar);
            double dt = now() - t0;
            if (dt < best) best = dt;
        }
        for (int t = 0; t < g_T; t++) pthread_join(th[t], NULL);
        pthread_barrier_destroy(&g_bar);
        printf("threads %2d: %.3f ms  (%.1f GB/s written)\n", g_T, best * 1e3, n * 4 / best / 1e9);
    }
    return 0;
}
