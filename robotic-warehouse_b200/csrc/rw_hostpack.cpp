// rw_hostpack.cpp -- host half of the packed end-to-end transport (rw_step_host).
//
// The step's observations are exact small integers stored as float32 (coordinates < 256, 0/1 flags), and the end-to-end call
// is bound by moving them over PCIe.  rw_step_host therefore ships them as one byte per value and widens them back to float32
// on the host, chunk by chunk while the next chunk is still in flight.  This file holds the widening: a small pool of worker
// threads (created on first use, parked on a condition variable between calls, spinning between the chunks of one call) and
// an AVX2 inner loop with streaming stores.  It is transport code only: no part of the simulation runs on the host.
#include <immintrin.h>
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

__attribute__((target("avx2"))) void widen_avx2(const uint8_t *in, float *out, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(out + i) & 31u)) { out[i] = (float)in[i]; i++; }  // align the stores
    for (; i + 32 <= n; i += 32) {
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(in + i));
        const __m128i lo = _mm256_castsi256_si128(b), hi = _mm256_extracti128_si256(b, 1);
        _mm256_stream_ps(out + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(lo)));
        _mm256_stream_ps(out + i + 8, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(lo, 8))));
        _mm256_stream_ps(out + i + 16, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(hi)));
        _mm256_stream_ps(out + i + 24, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(hi, 8))));
    }
    for (; i < n; i++) out[i] = (float)in[i];
    _mm_sfence();
}

void widen_scalar(const uint8_t *in, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i];
}

void widen(const uint8_t *in, float *out, size_t n) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) widen_avx2(in, out, n);
    else widen_scalar(in, out, n);
}

// NUMA placement: the widening is a pure memory stream, and on a two-socket host half of it crosses the socket interconnect
// unless the workers run next to the memory.  The node of the destination buffer is looked up (move_pages with a NULL target
// only queries) and the workers are confined to that node's CPUs.  Every step is best effort: failures leave the scheduler free.
int node_of(const void *p) {
    void *page = reinterpret_cast<void *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)4095);
    int status = -1;
    const long rc = syscall(SYS_move_pages, 0, 1UL, &page, nullptr, &status, 0);
    return rc == 0 && status >= 0 ? status : -1;
}

bool cpus_of_node(int node, cpu_set_t *set) {
    char path[96], buf[1024];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-31,64-95"
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); count++; }
    }
    return count > 0;
}

struct Pool {
    std::vector<std::thread> workers;
    int node = -2;  // NUMA node the workers are currently confined to (-2: never placed)
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
    bool awake = false, quit = false;      // under mu: workers leave the condition variable while `awake`
    std::atomic<uint64_t> generation{0};   // bumped once per job
    std::atomic<int> done{0};              // workers that finished the current job
    std::atomic<int> ready{0};             // workers that are spinning on `generation` (0 = all parked)
    const uint8_t *in = nullptr;
    float *out = nullptr;
    size_t n = 0;
    int T = 1;  // participants of a job, caller included

    void slice(int t) const {
        const size_t per = ((n + T - 1) / T + 31) & ~(size_t)31;
        const size_t a = (size_t)t * per;
        if (a < n) widen(in + a, out + a, (a + per > n ? n : a + per) - a);
    }

    void worker(int t) {
        for (;;) {
            pthread_mutex_lock(&mu);
            while (!awake && !quit) pthread_cond_wait(&cv, &mu);
            const bool stop = quit;
            pthread_mutex_unlock(&mu);
            if (stop) return;
            // the caller posts its first job only after every worker has announced itself, so `seen` is the value before it
            uint64_t seen = generation.load(std::memory_order_acquire);
            ready.fetch_add(1, std::memory_order_release);
            for (;;) {  // between the chunks of one call: spin
                uint64_t g;
                bool asleep = false;
                while ((g = generation.load(std::memory_order_acquire)) == seen) {
                    if (!__atomic_load_n(&awake, __ATOMIC_ACQUIRE)) { asleep = true; break; }
                    _mm_pause();
                }
                if (asleep) break;
                seen = g;
                slice(t);
                done.fetch_add(1, std::memory_order_release);
            }
            ready.fetch_sub(1, std::memory_order_release);  // back to the condition variable
        }
    }

    void start(int threads) {
        T = threads < 1 ? 1 : threads;
        try {
            for (int t = 1; t < T; t++) workers.emplace_back([this, t] { worker(t); });
        } catch (...) {  // thread limit reached: work with the threads that exist (slices are dealt by T, so fix it first)
        }
        // workers read T only after the first wake(), which happens after this assignment
        T = (int)workers.size() + 1;
    }

    void place(const void *dst) {
        const int want = node_of(dst);
        if (want < 0 || want == node) return;
        cpu_set_t set;
        if (!cpus_of_node(want, &set)) return;
        for (auto &w : workers) pthread_setaffinity_np(w.native_handle(), sizeof(set), &set);
        node = want;
    }

    void wake() {
        if (T == 1) return;
        pthread_mutex_lock(&mu);
        __atomic_store_n(&awake, true, __ATOMIC_RELEASE);
        pthread_cond_broadcast(&cv);
        pthread_mutex_unlock(&mu);
        while (ready.load(std::memory_order_acquire) != T - 1) _mm_pause();
    }

    void sleep() {  // returns once every worker has left the spin loop, so the next wake() starts from a clean state
        if (T == 1) return;
        pthread_mutex_lock(&mu);
        __atomic_store_n(&awake, false, __ATOMIC_RELEASE);
        pthread_mutex_unlock(&mu);
        while (ready.load(std::memory_order_acquire) != 0) _mm_pause();
    }

    void run(const uint8_t *src, float *dst, size_t count) {
        in = src; out = dst; n = count;
        if (T == 1) { slice(0); return; }
        done.store(0, std::memory_order_relaxed);
        generation.fetch_add(1, std::memory_order_release);
        slice(0);
        while (done.load(std::memory_order_acquire) != T - 1) _mm_pause();
    }

    ~Pool() {
        if (T > 1) {
            pthread_mutex_lock(&mu);
            quit = true;
            pthread_cond_broadcast(&cv);
            pthread_mutex_unlock(&mu);
            for (auto &w : workers) w.join();
        }
    }
};

Pool *g_pool = nullptr;
pthread_mutex_t g_pool_mu = PTHREAD_MUTEX_INITIALIZER;  // one widening at a time per process

// CPUs this process may use on average: the cgroup v2 / v1 CFS quota if there is one (a spinning pool larger than the quota
// gets the whole process throttled for the rest of the 100 ms period), else 0
double cgroup_cpu_quota() {
    double quota = 0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // "max 100000" or "<quota> <period>"
        char q[64];
        long period = 0;
        if (fscanf(f, "%63s %ld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) quota = atof(q) / (double)period;
        fclose(f);
        return quota;
    }
    long q = -1, period = 0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(f, "%ld", &q) != 1) q = -1; fclose(f); }
    if (FILE *f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(f, "%ld", &period) != 1) period = 0; fclose(f); }
    return q > 0 && period > 0 ? (double)q / (double)period : 0;
}

int default_threads() {
    const char *ev = getenv("RWARE_B200_HOST_THREADS");
    if (ev && atoi(ev) > 0) return atoi(ev);
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    unsigned ranks = 1;  // one process per GPU under torchrun: share the cores between the ranks of this node
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(lw) > 0 ? (unsigned)atoi(lw) : 1;
    double budget = hw >= 4 ? hw / 2.0 : hw;  // one thread per physical core at most; streaming stores saturate the
    const double quota = cgroup_cpu_quota();  // memory controllers long before all of them are busy
    if (quota > 0 && 0.75 * quota < budget) budget = 0.75 * quota;  // leave a quarter of it to the caller, CUDA and the framework
    int t = (int)(budget / ranks);
    return t < 1 ? 1 : (t > 32 ? 32 : t);
}

}  // namespace

// a forked child inherits the pool object but none of its threads: start over there (the parent's pool is simply abandoned)
static void rw_hostpack_after_fork_in_child() {
    g_pool = nullptr;
    pthread_mutex_init(&g_pool_mu, nullptr);
}

// internal interface used by rw_capi.cu
void rw_hostpack_begin(const void *dst) {
    pthread_mutex_lock(&g_pool_mu);
    if (!g_pool) {
        static const int registered = pthread_atfork(nullptr, nullptr, rw_hostpack_after_fork_in_child);
        (void)registered;
        g_pool = new Pool();
        g_pool->start(default_threads());
    }
    static const bool numa = !(getenv("RWARE_B200_HOST_NUMA") && atoi(getenv("RWARE_B200_HOST_NUMA")) == 0);
    if (numa && dst) g_pool->place(dst);
    g_pool->wake();
}
void rw_hostpack_widen(const uint8_t *src, float *dst, size_t n) { g_pool->run(src, dst, n); }
void rw_hostpack_end() {
    g_pool->sleep();
    pthread_mutex_unlock(&g_pool_mu);
}
int rw_hostpack_threads() { return g_pool ? g_pool->T : default_threads(); }
