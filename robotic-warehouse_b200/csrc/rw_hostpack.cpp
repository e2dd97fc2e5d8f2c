// rw_hostpack.cpp -- host half of the end-to-end transport (rw_step_host).
//
// The caller of rw_step_host wants float32 observation rows in HOST memory (the reference's contract,
// rware/warehouse.py:17-28, 631-674): 75.5 MB per step of 65 536 tiny-4ag worlds.  Every value of such a row is a 0/1 flag
// except the agent's own coordinates, so the device ships one BIT per flag plus one byte per coordinate (11 B instead of
// 284 B per row over PCIe) and this file expands them into the caller's buffer: a pool of worker threads, woken once per call,
// pulls 64 KB output blocks from a shared counter (late or descheduled workers simply take fewer blocks) and writes them with
// non-temporal stores (one 16-bit mask -> one 64-byte line with AVX-512; a 256-entry table with AVX2).  A second job kind
// widens one byte per value (image observations with the AGENT_DIRECTION layer, whose values reach 4).  The expansion is bound
// by the host's DRAM write bandwidth; rw_step_host therefore lets the GPU's copy engine write part of the rows as plain
// float32 at the same time (rw_capi.cu).  Transport code only: no part of the simulation runs on the host.
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "rw_hostpack.h"

#if defined(__x86_64__)
#include <immintrin.h>
#define RW_X86 1
static inline void cpu_relax() { _mm_pause(); }
#else
#define RW_X86 0
static inline void cpu_relax() {
#if defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#endif
}
#endif

namespace {

double now_us() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

// ------------------------------------------------------------------------------------------------ inner loops
// Every job is cut into blocks of RW_HOST_BLOCK floats (64 KB of output); a block is the unit a thread pulls.
constexpr size_t BLOCK = RW_HOST_BLOCK;

// ---- one byte per value -> float32
void widen_scalar(const uint8_t *in, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i];
}
#if RW_X86
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
#endif
void widen(const uint8_t *in, float *out, size_t n) {
#if RW_X86
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return widen_avx2(in, out, n);
#endif
    widen_scalar(in, out, n);
}

// ---- one bit per value (+ one byte for each of the first `nwide` values of a row) -> float32
// Float j of the output is 1.0f iff bit j of `bits` is set, except the "wide" positions (row * L + k, k < nwide), which take
// (float)wide[row * nwide + k] whatever their bit says.  [a, b) is a range of floats; a is a multiple of 16 and out + a is
// 64-byte aligned for the vector forms.
struct WideCursor {  // the next wide position at or after a float index
    size_t w, pos;
    int k;
};
inline WideCursor wide_from(size_t a, int L, int nwide, size_t n_rows) {
    WideCursor c{(size_t)-1, (size_t)-1, 0};
    if (nwide <= 0) return c;
    size_t row = a / (size_t)L;
    int rem = (int)(a - row * (size_t)L);
    if (rem >= nwide) { row++; rem = 0; }
    if (row >= n_rows) return c;
    c.w = row * (size_t)nwide + (size_t)rem;
    c.pos = row * (size_t)L + (size_t)rem;
    c.k = rem;
    return c;
}
inline void wide_next(WideCursor &c, int L, int nwide, size_t n_rows) {
    c.w++;
    if (++c.k == nwide) { c.k = 0; c.pos += (size_t)(L - nwide + 1); }
    else c.pos++;
    if (c.w >= n_rows * (size_t)nwide) c.pos = (size_t)-1;
}

void expand_scalar(const uint8_t *bits, const uint8_t *wide, float *out, size_t a, size_t b, int L, int nwide, size_t n_rows) {
    for (size_t i = a; i < b; i++) out[i] = ((bits[i >> 3] >> (i & 7)) & 1) ? 1.0f : 0.0f;
    for (WideCursor c = wide_from(a, L, nwide, n_rows); c.pos < b; wide_next(c, L, nwide, n_rows)) out[c.pos] = (float)wide[c.w];
}

#if RW_X86
__attribute__((target("avx512f"))) void expand_avx512(const uint8_t *bits, const uint8_t *wide, float *out, size_t a, size_t b,
                                                      int L, int nwide, size_t n_rows) {
    const uint16_t *m16 = reinterpret_cast<const uint16_t *>(bits);
    const __m512 ones = _mm512_set1_ps(1.0f);
    WideCursor c = wide_from(a, L, nwide, n_rows);
    size_t i = a;
    for (; i + 16 <= b; i += 16) {
        __m512 v = _mm512_maskz_mov_ps((__mmask16)m16[i >> 4], ones);
        while (c.pos < i + 16) {
            v = _mm512_mask_mov_ps(v, (__mmask16)(1u << (c.pos - i)), _mm512_set1_ps((float)wide[c.w]));
            wide_next(c, L, nwide, n_rows);
        }
        _mm512_stream_ps(out + i, v);
    }
    _mm_sfence();
    if (i < b) expand_scalar(bits, wide, out, i, b, L, nwide, n_rows);
}

alignas(64) float g_lut8[256][8];  // byte of flags -> eight floats
void build_lut() {
    for (int m = 0; m < 256; m++)
        for (int k = 0; k < 8; k++) g_lut8[m][k] = ((m >> k) & 1) ? 1.0f : 0.0f;
}
__attribute__((target("avx2"))) void expand_avx2(const uint8_t *bits, const uint8_t *wide, float *out, size_t a, size_t b, int L,
                                                 int nwide, size_t n_rows) {
    WideCursor c = wide_from(a, L, nwide, n_rows);
    size_t i = a;
    for (; i + 8 <= b; i += 8) {
        __m256 v = _mm256_load_ps(g_lut8[bits[i >> 3]]);
        if (c.pos < i + 8) {
            alignas(32) float tmp[8];
            _mm256_store_ps(tmp, v);
            while (c.pos < i + 8) { tmp[c.pos - i] = (float)wide[c.w]; wide_next(c, L, nwide, n_rows); }
            v = _mm256_load_ps(tmp);
        }
        _mm256_stream_ps(out + i, v);
    }
    _mm_sfence();
    if (i < b) expand_scalar(bits, wide, out, i, b, L, nwide, n_rows);
}
#endif

int g_isa = -1;  // 0 scalar, 1 avx2, 2 avx512
int detect_isa() {
    int isa = 0;
#if RW_X86
    if (__builtin_cpu_supports("avx2")) isa = 1;
    if (__builtin_cpu_supports("avx512f")) isa = 2;
    build_lut();
#endif
    if (const char *ev = getenv("RWARE_B200_HOST_ISA")) {  // diagnostics: cap the instruction set (0 scalar, 1 avx2, 2 avx512)
        const int cap = atoi(ev);
        if (cap < isa) isa = cap < 0 ? 0 : cap;
    }
    return isa;
}

void expand(const uint8_t *bits, const uint8_t *wide, float *out, size_t a, size_t b, int L, int nwide, size_t n_rows) {
#if RW_X86
    const bool aligned = ((reinterpret_cast<uintptr_t>(out) & 63u) == 0) && (a % 16 == 0);
    if (aligned && g_isa == 2) return expand_avx512(bits, wide, out, a, b, L, nwide, n_rows);
    if (aligned && g_isa >= 1) return expand_avx2(bits, wide, out, a, b, L, nwide, n_rows);
#endif
    expand_scalar(bits, wide, out, a, b, L, nwide, n_rows);
}

// ---- streaming fill (bandwidth probe: what the host's memory system takes from this pool when nothing has to be decoded)
#if RW_X86
__attribute__((target("avx512f"))) void fill_avx512(float *out, size_t a, size_t b, float value) {
    const __m512 v = _mm512_set1_ps(value);
    size_t i = a;
    for (; i + 16 <= b; i += 16) _mm512_stream_ps(out + i, v);
    for (; i < b; i++) out[i] = value;
    _mm_sfence();
}
__attribute__((target("avx2"))) void fill_avx2(float *out, size_t a, size_t b, float value) {
    const __m256 v = _mm256_set1_ps(value);
    size_t i = a;
    for (; i + 8 <= b; i += 8) _mm256_stream_ps(out + i, v);
    for (; i < b; i++) out[i] = value;
    _mm_sfence();
}
#endif
void fill(float *out, size_t a, size_t b, float value) {
#if RW_X86
    const bool aligned = ((reinterpret_cast<uintptr_t>(out) & 63u) == 0) && (a % 16 == 0);
    if (aligned && g_isa == 2) return fill_avx512(out, a, b, value);
    if (aligned && g_isa >= 1) return fill_avx2(out, a, b, value);
#endif
    for (size_t i = a; i < b; i++) out[i] = value;
}

// ------------------------------------------------------------------------------------------------ NUMA helpers
// The expansion is a pure memory stream, bound by the DRAM write bandwidth of the socket(s) that hold the destination pages
// (about 200 GB/s per socket of the 2 x Xeon 8562Y+ hosts this was tuned on, reached by ~10 threads).  Buffers from
// rw_host_alloc_near are therefore striped over the sockets, the workers are spread over the sockets, and every block is
// preferably expanded by a worker of the socket that holds its pages (per-socket queues, stealing when a queue runs dry).
// Every step is best effort: without NUMA information there is one queue and the scheduler is left free.
}  // namespace

bool rw_host_cpus_of_node(int node, cpu_set_t *set) {
    char path[96], buf[1024];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    char *save = nullptr;
    for (char *tok = strtok_r(buf, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {  // "0-31,64-95"
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, set); count++; }
    }
    return count > 0;
}

int rw_host_numa_nodes() {
    static int n = -1;
    if (n < 0) {
        int k = 0;
        cpu_set_t s;
        while (k < RW_HOST_MAX_NODES && rw_host_cpus_of_node(k, &s)) k++;
        if (getenv("RWARE_B200_HOST_NUMA") && atoi(getenv("RWARE_B200_HOST_NUMA")) == 0) k = 1;
        n = k < 1 ? 1 : k;
    }
    return n;
}

// node of every page in `pages` (move_pages with a NULL target only queries); -1 where unknown
void rw_host_nodes_of(void *const *pages, int count, int *status) {
    for (int i = 0; i < count; i++) status[i] = -1;
    for (int a = 0; a < count; a += 512) {
        const int k = count - a < 512 ? count - a : 512;
        if (syscall(SYS_move_pages, 0, (unsigned long)k, pages + a, nullptr, status + a, 0) != 0)
            for (int i = 0; i < k; i++) status[a + i] = -1;
    }
}

namespace {

int node_of_cpu(int cpu) {
    static int table[CPU_SETSIZE];
    static bool built = false;
    if (!built) {
        for (int c = 0; c < CPU_SETSIZE; c++) table[c] = 0;
        for (int nd = 0; nd < rw_host_numa_nodes(); nd++) {
            cpu_set_t s;
            if (!rw_host_cpus_of_node(nd, &s)) continue;
            for (int c = 0; c < CPU_SETSIZE; c++)
                if (CPU_ISSET(c, &s)) table[c] = nd;
        }
        built = true;
    }
    return cpu >= 0 && cpu < CPU_SETSIZE ? table[cpu] : 0;
}

// CPUs this process may use on average: the cgroup v2 / v1 CFS quota if there is one (a pool that keeps more threads busy
// than the quota gets the whole process throttled for the rest of the 100 ms period), else 0
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
    // one thread per physical core at most; ~10 threads per socket saturate its memory controllers
    double budget = hw >= 4 ? hw / 2.0 : hw;
    const double quota = cgroup_cpu_quota();
    if (quota > 0 && quota < budget) budget = quota;  // workers block between calls, so the pool may use the whole quota
    int t = (int)(budget / ranks);
    const int cap = 12 * rw_host_numa_nodes();
    return t < 1 ? 1 : (t > cap ? cap : t);
}

// ------------------------------------------------------------------------------------------------ the pool
constexpr int MAXQ = RW_HOST_MAX_NODES;

struct Slot {  // parameters of one job; slot (generation & 1) is stable from publication until every block is done
    RwHostJob job;
    size_t nblocks = 0;
    int nq = 1;
    std::vector<uint32_t> qblocks[MAXQ];  // blocks whose pages sit on node q, ascending
    std::atomic<int> chunks_ready{0};
    std::atomic<int> cancelled{0};
    std::atomic<size_t> blocks_done{0};
    char pad[64];
};

struct DstInfo {  // node of every block of a destination buffer (looked up once per buffer)
    const float *dst = nullptr;
    size_t n = 0;
    std::vector<uint8_t> node;
};

struct alignas(64) QueueCtr {
    // (generation << 32) | blocks still to hand out.  A thread takes a block by CAS-decrementing the low half; a successful
    // CAS on generation g proves that job g is still incomplete, so reading slot[g & 1] afterwards is safe.
    std::atomic<uint64_t> v{0};
};

struct Pool {
    std::vector<std::thread> workers;
    int T = 1;  // participants of a job, caller included
    int nq = 1;
    pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
    pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
    uint64_t wake_seq = 0;  // under mu
    bool quit = false;      // under mu
    QueueCtr ctr[MAXQ];
    uint64_t gen = 0;       // under the job lock
    Slot slot[2];
    DstInfo known[4];
    int known_next = 0;
    double spin_us = 250;  // a worker without work spins this long (the copies and the kernel of a call are in flight) before it parks

    void run_block(const Slot &s, size_t blk) {
        const RwHostJob &j = s.job;
        const size_t a = blk * BLOCK, b = a + BLOCK > j.n ? j.n : a + BLOCK;
        if (j.nchunks > 1) {  // the bytes this block decodes arrive with the piece that covers it
            int need = 1;
            while (need < j.nchunks && blk >= j.chunk_end[need - 1]) need++;
            while (s.chunks_ready.load(std::memory_order_acquire) < need) cpu_relax();
        }
        if (s.cancelled.load(std::memory_order_relaxed)) return;
        switch (j.kind) {
            case RW_HOSTJOB_WIDEN: widen(j.src + a, j.dst + a, b - a); break;
            case RW_HOSTJOB_BITS: expand(j.src, j.wide, j.dst, a, b, j.L, j.nwide, j.n_rows); break;
            default: fill(j.dst, a, b, j.fill_value); break;
        }
    }

    // pulls blocks -- from the queue of node `home` first, then from the others -- until the current job has none left;
    // returns true if it did any work
    bool drain(int home) {
        bool any = false;
        for (int k = 0; k < nq; k++) {
            const int q = (home + k) % nq;
            uint64_t v = ctr[q].v.load(std::memory_order_acquire);
            for (;;) {
                const uint32_t left = (uint32_t)v;
                if (left == 0) break;
                if (!ctr[q].v.compare_exchange_weak(v, v - 1, std::memory_order_acq_rel, std::memory_order_acquire)) continue;
                Slot &s = slot[(v >> 32) & 1];
                run_block(s, s.qblocks[q][s.qblocks[q].size() - left]);
                s.blocks_done.fetch_add(1, std::memory_order_release);
                any = true;
                k = -1;  // after a block, look at the home queue first again
                break;
            }
        }
        return any;
    }

    void worker(int home) {
        uint64_t seen = 0;
        double idle_since = now_us();
        for (;;) {
            if (drain(home)) { idle_since = now_us(); continue; }
            if (now_us() - idle_since < spin_us) { cpu_relax(); continue; }
            pthread_mutex_lock(&mu);
            while (!quit && wake_seq == seen) pthread_cond_wait(&cv, &mu);
            seen = wake_seq;
            const bool stop = quit;
            pthread_mutex_unlock(&mu);
            if (stop) return;
            idle_since = now_us();
        }
    }

    void start(int threads) {
        T = threads < 1 ? 1 : threads;
        nq = rw_host_numa_nodes();
        if (nq > MAXQ) nq = MAXQ;
        const int home0 = node_of_cpu(sched_getcpu());  // the caller usually runs next to its GPU: the workers start on the other node
        try {
            for (int t = 1; t < T; t++) {
                const int home = (home0 + t) % nq;
                workers.emplace_back([this, home] { worker(home); });
                cpu_set_t set;
                if (nq > 1 && rw_host_cpus_of_node(home, &set)) pthread_setaffinity_np(workers.back().native_handle(), sizeof(set), &set);
            }
        } catch (...) {  // thread limit reached: work with the threads that exist
        }
        T = (int)workers.size() + 1;
    }

    const std::vector<uint8_t> &nodes_of(const float *dst, size_t n) {
        for (auto &k : known)
            if (k.dst == dst && k.n >= n) return k.node;
        DstInfo &k = known[known_next++ & 3];
        k.dst = dst;
        k.n = n;
        const size_t nb = (n + BLOCK - 1) / BLOCK;
        k.node.assign(nb, 0);
        if (nq > 1) {
            std::vector<void *> pages(nb);
            std::vector<int> st(nb);
            for (size_t b = 0; b < nb; b++)
                pages[b] = reinterpret_cast<void *>(reinterpret_cast<uintptr_t>(dst + b * BLOCK + (b + 1 == nb ? 0 : BLOCK / 2)) & ~(uintptr_t)4095);
            rw_host_nodes_of(pages.data(), (int)nb, st.data());
            for (size_t b = 0; b < nb; b++) k.node[b] = (uint8_t)(st[b] >= 0 && st[b] < nq ? st[b] : 0);
        }
        return k.node;
    }

    void wake() {
        if (T == 1) return;
        pthread_mutex_lock(&mu);
        wake_seq++;
        pthread_cond_broadcast(&cv);
        pthread_mutex_unlock(&mu);
    }

    ~Pool() {
        if (T > 1) {
            pthread_mutex_lock(&mu);
            quit = true;
            wake_seq++;
            pthread_cond_broadcast(&cv);
            pthread_mutex_unlock(&mu);
            for (auto &w : workers) w.join();
        }
    }
};

Pool *g_pool = nullptr;
pthread_mutex_t g_create_mu = PTHREAD_MUTEX_INITIALIZER;
pthread_mutex_t g_job_mu = PTHREAD_MUTEX_INITIALIZER;  // one job at a time per process
Slot *g_cur = nullptr;                                 // under g_job_mu: the slot of the job in flight

// a forked child inherits the pool object but none of its threads: start over there (the parent's pool is simply abandoned)
void after_fork_in_child() {
    g_pool = nullptr;
    g_cur = nullptr;
    pthread_mutex_init(&g_create_mu, nullptr);
    pthread_mutex_init(&g_job_mu, nullptr);
}

Pool *pool() {
    Pool *p = __atomic_load_n(&g_pool, __ATOMIC_ACQUIRE);
    if (p) return p;
    pthread_mutex_lock(&g_create_mu);
    if (!g_pool) {
        static const int registered = pthread_atfork(nullptr, nullptr, after_fork_in_child);
        (void)registered;
        if (g_isa < 0) g_isa = detect_isa();
        Pool *np = new Pool();
        if (const char *ev = getenv("RWARE_B200_HOST_SPIN_US")) np->spin_us = atof(ev);
        np->start(default_threads());
        __atomic_store_n(&g_pool, np, __ATOMIC_RELEASE);
    }
    p = g_pool;
    pthread_mutex_unlock(&g_create_mu);
    return p;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ internal interface (rw_hostpack.h)
void rw_hostpool_prewake() { pool()->wake(); }

void rw_hostpool_begin(const RwHostJob &job) {
    Pool *p = pool();
    pthread_mutex_lock(&g_job_mu);
    const uint64_t gen = ++p->gen;
    Slot &s = p->slot[gen & 1];
    s.job = job;
    s.nblocks = (job.n + BLOCK - 1) / BLOCK;
    s.nq = p->nq;
    if (s.job.nchunks < 1) s.job.nchunks = 1;
    if (s.job.nchunks > 8) s.job.nchunks = 8;
    const std::vector<uint8_t> &node = p->nodes_of(job.dst, job.n);
    for (int q = 0; q < p->nq; q++) s.qblocks[q].clear();
    for (size_t b = 0; b < s.nblocks; b++) s.qblocks[node[b]].push_back((uint32_t)b);
    s.chunks_ready.store(0, std::memory_order_relaxed);
    s.cancelled.store(0, std::memory_order_relaxed);
    s.blocks_done.store(0, std::memory_order_relaxed);
    g_cur = &s;
    for (int q = 0; q < p->nq; q++)
        p->ctr[q].v.store((gen << 32) | (uint64_t)(uint32_t)s.qblocks[q].size(), std::memory_order_release);
    p->wake();
    static const bool trace = getenv("RWARE_B200_HOST_TRACE") != nullptr;
    if (trace && (gen % 16) == 1) {
        fprintf(stderr, "[rw_hostpool] job %llu: %zu blocks, per-node queues:", (unsigned long long)gen, s.nblocks);
        for (int q = 0; q < p->nq; q++) fprintf(stderr, " node%d=%zu", q, s.qblocks[q].size());
        fprintf(stderr, " | threads %d, caller on node %d\n", p->T, node_of_cpu(sched_getcpu()));
    }
}

void rw_hostpool_chunk_ready(int chunks) { g_cur->chunks_ready.store(chunks, std::memory_order_release); }

void rw_hostpool_cancel() {
    g_cur->cancelled.store(1, std::memory_order_release);
    g_cur->chunks_ready.store(1 << 30, std::memory_order_release);
}

void rw_hostpool_finish() {
    Pool *p = g_pool;
    Slot &s = *g_cur;
    p->drain(node_of_cpu(sched_getcpu()));  // the caller works too
    while (s.blocks_done.load(std::memory_order_acquire) != s.nblocks) cpu_relax();
    g_cur = nullptr;
    pthread_mutex_unlock(&g_job_mu);
}

int rw_hostpool_threads() { return pool()->T; }
int rw_hostpool_isa() {
    pool();
    return g_isa;
}
