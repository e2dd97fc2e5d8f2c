// rw_capi.cu -- host side of the C ABI declared in include/rware_b200.h.
//
// Owns the packed world state in HBM, the static per-configuration tables, the shared-memory carve-up of the
// fused kernel, and the numpy-compatible seeding (SeedSequence -> PCG64) that makes reset(seed=k) reproduce the
// reference's generator stream (rware/warehouse.py:758-760, 962-964; numpy/random/bit_generator.pyx).
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rware_b200.h"
#include "rw_launch.h"
#include "rw_host_layout.h"
#include "rw_hostpack.h"
#include "rw_views.cuh"

#define RW_DMA_FRAC0 0.08

// What one rw_step_host call does on the device side, decided before anything is enqueued
struct HostPlan {
    const int32_t *actions_host;
    float *obs_host, *reward_host;
    uint8_t *done_host;
    size_t rows_cpu;   // observation rows (world-major) the host pool expands; the rest is delivered by the copy engine
    bool bits, fused;  // bit transport (else bytes); bit form written by the step kernel itself (else packed from float rows)
    int nwide, nch;
    size_t chunk_end[8];
    size_t off_wide;
    int64_t epoch;
    bool same(const HostPlan &o) const {
        return obs_host == o.obs_host && reward_host == o.reward_host && done_host == o.done_host &&  // (actions are copied outside the graph)
               rows_cpu == o.rows_cpu && bits == o.bits && fused == o.fused && nch == o.nch && epoch == o.epoch;
    }
};

struct HostGraph {
    HostPlan plan;
    cudaGraphExec_t exec = nullptr;
    int64_t last_use = 0;
};

#define RW_HOST_LOCAL_TAIL 0.10  // share of a buffer from rw_host_alloc_near that stays on the GPU's NUMA node (the copy engine's rows)

struct rw_handle {
    RwParams p;
    rw_config cfg;
    int device;
    bool ready;  // worlds hold a valid state (after rw_reset / rw_set_state)
    std::string err;
    void *d_highway = nullptr, *d_home = nullptr, *d_home_id = nullptr, *d_cta_tables = nullptr, *d_normx = nullptr, *d_normy = nullptr;
    void *d_state = nullptr, *d_rng = nullptr, *d_err = nullptr;
    // staging for the *_host entry points
    int32_t *d_actions = nullptr;
    float *d_obs = nullptr, *d_reward = nullptr;
    uint8_t *d_done = nullptr;
    cudaStream_t host_stream = nullptr;
    // compressed transport of rw_step_host: bit / byte image of the observations on the device and in pinned host memory
    uint8_t *d_pack = nullptr, *h_pack = nullptr;
    size_t pack_bytes = 0;
    uint32_t *d_flag_ctr = nullptr;            // device counters behind the host-visible flags
    volatile uint32_t *h_flags = nullptr;      // mapped host memory, one flag per 64-byte line: piece c -> h_flags[16 c], copy engine -> [16 * 8]
    uint32_t flag_seen[16] = {};
    struct HostGraph graphs[6];
    int64_t host_calls_graph = 0, host_epoch = 0;  // epoch: bumped by whatever changes what the captured launches would do
    int host_pack = 1;  // 0 plain float32 copy, 1 compressed (bits where the rows allow it, else bytes), 2 compressed, bytes only
    // share of the observation rows that the copy engine delivers as plain float32 while the host pool expands the rest;
    // follows the measured rates of the two unless fixed (rw_set_host_dma_share / RWARE_B200_HOST_DMA_SHARE)
    double dma_frac = RW_DMA_FRAC0, dma_frac_fixed = -1.0;
    int64_t host_calls = 0;
    // fused push all-gather
    void *d_gather = nullptr;
    void *peer_base[RW_MAX_PEERS] = {};
    int gather_world = 0, gather_rank = 0;
    bool gather_open = false;
    size_t gather_off_rew = 0, gather_off_done = 0, gather_slot_bytes = 0;
    int64_t gather_steps = 0;  // rw_step_gather calls so far: call k writes slot k & 1
    rw_info info;
};

static thread_local std::string g_create_error;

static int fail(rw_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(h, RW_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
    } while (0)

static int create_impl(rw_handle *h, const rw_config *cfg, int64_t num_envs, int device, const std::vector<uint16_t> &home);
static cudaError_t free_pinned(void *p);

extern "C" {

const char *rw_last_error(const rw_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rw_create(const rw_config *cfg, int64_t num_envs, int device, rw_handle **out) {
    rw_handle *h = nullptr;
    if (!cfg || !out) return fail(h, RW_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != RW_ABI_VERSION) return fail(h, RW_EINVAL, "rw_config.abi_version mismatch");
    if (num_envs < 1 || num_envs > (1LL << 31)) return fail(h, RW_EINVAL, "num_envs out of range");
    if (cfg->grid_h < 1 || cfg->grid_w < 1 || cfg->grid_h > 255 || cfg->grid_w > 255)
        return fail(h, RW_EINVAL, "grid must be within 255x255");
    if (cfg->n_agents < 1 || cfg->n_agents > RW_MAX_AGENTS) return fail(h, RW_EINVAL, "n_agents must be 1..32");
    if (cfg->msg_bits < 0 || cfg->msg_bits > RW_MAX_MSG_BITS) return fail(h, RW_EINVAL, "msg_bits must be 0..16");
    if (cfg->sensor_range < 0 || cfg->sensor_range > 8) return fail(h, RW_EINVAL, "sensor_range must be 0..8");
    if (cfg->n_goals < 1 || cfg->n_goals > RW_MAX_GOALS) return fail(h, RW_EINVAL, "n_goals must be 1..16");
    if (!cfg->highways || !cfg->goals_xy) return fail(h, RW_EINVAL, "highways/goals_xy missing");
    if (cfg->reward_type < 0 || cfg->reward_type > 2) return fail(h, RW_EINVAL, "bad reward_type");
    if (cfg->observation_type < 0 || cfg->observation_type > 3) return fail(h, RW_EINVAL, "bad observation_type");
    if (cfg->observation_type >= RW_OBS_IMAGE) {
        if (cfg->n_image_layers < 1 || cfg->n_image_layers > 7) return fail(h, RW_EINVAL, "n_image_layers must be 1..7");
        for (int i = 0; i < cfg->n_image_layers; i++) {
            int ly = cfg->image_layers[i];
            if (ly < 0 || ly > 6) return fail(h, RW_EINVAL, "bad image layer");
        }
    }
    const int H = cfg->grid_h, W = cfg->grid_w, cells = H * W, n = cfg->n_agents;
    if (n > cells) return fail(h, RW_EINVAL, "more agents than cells");
    std::vector<uint16_t> home;
    for (int c = 0; c < cells; c++)
        if (!cfg->highways[c]) home.push_back((uint16_t)c);
    const int S = (int)home.size();
    if (S > RW_MAX_SHELVES) return fail(h, RW_EINVAL, "too many shelves");
    if (cfg->request_queue_size < 0 || cfg->request_queue_size > S)
        return fail(h, RW_EINVAL, "request_queue_size must be 0..n_shelves");
    for (int g = 0; g < cfg->n_goals; g++) {
        int gx = cfg->goals_xy[2 * g], gy = cfg->goals_xy[2 * g + 1];
        if (gx < 0 || gy < 0 || gx >= W || gy >= H) return fail(h, RW_EINVAL, "goal outside the grid");
    }

    h = new rw_handle();
    int rc = create_impl(h, cfg, num_envs, device, home);
    if (rc != RW_OK) {
        g_create_error = h->err;
        rw_destroy(h);
        return rc;
    }
    *out = h;
    return RW_OK;
}

}  // extern "C"

static int create_impl(rw_handle *h, const rw_config *cfg, int64_t num_envs, int device, const std::vector<uint16_t> &home) {
    const int H = cfg->grid_h, W = cfg->grid_w, cells = H * W, n = cfg->n_agents;
    const int S = (int)home.size();
    h->cfg = *cfg;
    h->cfg.highways = nullptr;
    h->cfg.goals_xy = nullptr;
    h->device = device;
    h->ready = false;
    RwParams &p = h->p;
    rw_fill_params(p, *cfg, num_envs, S);

    cudaDeviceProp prop;
    CK(cudaSetDevice(device));
    CK(cudaGetDeviceProperties(&prop, device));
    const int smem_cap = (int)prop.sharedMemPerBlockOptin;
    if (p.sm_cta_bytes + p.sm_warp_bytes > smem_cap)
        return fail(h, RW_EINVAL, "configuration needs " + std::to_string(p.sm_warp_bytes) +
                                      " B of shared memory per warp, device allows " + std::to_string(smem_cap));
    {   // warps (= tiles) per CTA: as many as the shared memory of one SM holds, so that a CTA fills an SM and, when the
        // batch allows, the whole launch is a single wave; small batches spread over the SMs instead
        int fit = (smem_cap - p.sm_cta_bytes) / p.sm_warp_bytes;
        if (fit > 14) fit = 14;  // 896 threads per CTA: 72 registers per thread (__launch_bounds__ of rw_main_kernel), named barriers 1..14
        if (fit < 1) fit = 1;
        int spread = (p.tiles + prop.multiProcessorCount - 1) / prop.multiProcessorCount;
        int wpc = spread < fit ? spread : fit;
        const char *ev = getenv("RWARE_B200_WARPS_PER_CTA");
        if (ev) wpc = atoi(ev);
        if (wpc < 1) wpc = 1;
        if (wpc > 14) wpc = 14;
        while (wpc > 1 && p.sm_cta_bytes + wpc * p.sm_warp_bytes > smem_cap) wpc--;
        p.warps_per_cta = wpc;
    }
    // host transport of rw_step_host (see there); RWARE_B200_HOST_PACK = 0 | 1 | 2 and RWARE_B200_HOST_DMA_SHARE override
    if (getenv("RWARE_B200_HOST_PACK")) h->host_pack = atoi(getenv("RWARE_B200_HOST_PACK"));
    if (h->host_pack < 0 || h->host_pack > 2) h->host_pack = 1;
    if (getenv("RWARE_B200_HOST_DMA_SHARE")) h->dma_frac_fixed = atof(getenv("RWARE_B200_HOST_DMA_SHARE"));
    if (getenv("RWARE_B200_FORCE_DYNAMIC")) rw_force_dynamic_kernels(atoi(getenv("RWARE_B200_FORCE_DYNAMIC")) != 0);
    CK(rw_configure_kernels(p, p.sm_cta_bytes + p.sm_warp_bytes * p.warps_per_cta));

    // static tables
    std::vector<float> nx(W), ny(H);
    for (int x = 0; x < W; x++) nx[x] = W > 1 ? (float)((double)x / (double)(W - 1)) : 0.0f;
    for (int y = 0; y < H; y++) ny[y] = H > 1 ? (float)((double)y / (double)(H - 1)) : 0.0f;
    CK(cudaMalloc(&h->d_highway, cells));
    CK(cudaMemcpy(h->d_highway, cfg->highways, cells, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->d_home, sizeof(uint16_t) * (S > 0 ? S : 1)));
    if (S) CK(cudaMemcpy(h->d_home, home.data(), sizeof(uint16_t) * S, cudaMemcpyHostToDevice));
    {
        std::vector<uint16_t> home_id(cells, 0);
        for (int k = 0; k < S; k++) home_id[home[k]] = (uint16_t)(k + 1);
        CK(cudaMalloc(&h->d_home_id, sizeof(uint16_t) * cells));
        CK(cudaMemcpy(h->d_home_id, home_id.data(), sizeof(uint16_t) * cells, cudaMemcpyHostToDevice));
    }
    {   // image of the CTA-shared table region: highway bytes, then (2-byte aligned) home ids
        std::vector<uint8_t> img((size_t)p.sm_cta_bytes, 0);
        for (int c = 0; c < cells; c++) img[c] = cfg->highways[c] ? 1 : 0;
        for (int g = 0; g < cfg->n_goals; g++) img[cfg->goals_xy[2 * g + 1] * W + cfg->goals_xy[2 * g]] |= 2;  // bit 1: goal cell
        uint16_t *hid = reinterpret_cast<uint16_t *>(img.data() + ((cells + 1) & ~1));
        for (int k = 0; k < S; k++) hid[home[k]] = (uint16_t)(k + 1);
        CK(cudaMalloc(&h->d_cta_tables, img.size()));
        CK(cudaMemcpy(h->d_cta_tables, img.data(), img.size(), cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&h->d_normx, sizeof(float) * W));
    CK(cudaMalloc(&h->d_normy, sizeof(float) * H));
    CK(cudaMemcpy(h->d_normx, nx.data(), sizeof(float) * W, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_normy, ny.data(), sizeof(float) * H, cudaMemcpyHostToDevice));
    size_t state_bytes = (size_t)p.tiles * p.nw * RW_TILE * 4;
    CK(cudaMalloc(&h->d_state, state_bytes));
    CK(cudaMemset(h->d_state, 0, state_bytes));
    CK(cudaMalloc(&h->d_rng, (size_t)RW_RNG_WORDS * p.E_pad * 4));
    CK(cudaMemset(h->d_rng, 0, (size_t)RW_RNG_WORDS * p.E_pad * 4));
    CK(cudaMalloc(&h->d_err, 8));
    CK(cudaMemset(h->d_err, 0, 8));
    p.highway = (const uint8_t *)h->d_highway;
    p.home_cell = (const uint16_t *)h->d_home;
    p.home_id = (const uint16_t *)h->d_home_id;
    p.cta_tables = (const uint8_t *)h->d_cta_tables;
    p.normx = (const float *)h->d_normx;
    p.normy = (const float *)h->d_normy;
    p.state = (int32_t *)h->d_state;
    p.rng = (uint32_t *)h->d_rng;
    p.err_count = (unsigned long long *)h->d_err;

    rw_info &inf = h->info;
    memset(&inf, 0, sizeof(inf));
    inf.num_envs = num_envs; inf.n_agents = n; inf.n_shelves = S; inf.obs_len = p.L; inf.act_width = p.A;
    inf.request_queue_size = p.R; inf.grid_h = H; inf.grid_w = W; inf.device = device;
    inf.state_words_per_env = p.nw; inf.envs_per_tile = RW_TILE;
    inf.smem_bytes_per_warp = p.sm_warp_bytes; inf.warps_per_cta = p.warps_per_cta;
    // S_state = 4*(5n + S + R + 2) + 16; B_step = n*L*4 + 2*S_state + (4n + 4n + 1); B_obs = n*L*4 + S_state
    int64_t s_state = 4LL * (5 * n + S + p.R + 2) + 16;
    inf.algorithmic_bytes_step = (int64_t)n * p.L * 4 + 2 * s_state + (8LL * n + 1);
    inf.algorithmic_bytes_obs = (int64_t)n * p.L * 4 + s_state;
    inf.packed_bytes_step = (int64_t)n * p.L * 4 + 2LL * p.nw * 4 + (4LL * n * p.A + 4LL * n + 1);
    inf.packed_bytes_obs = (int64_t)n * p.L * 4 + (int64_t)p.nw * 4;
    return RW_OK;
}

extern "C" {

int rw_destroy(rw_handle *h) {
    if (!h) return RW_OK;
    cudaSetDevice(h->device);
    if (h->d_pack) {
        cudaFree(h->d_pack);
        free_pinned(h->h_pack);
        for (auto &g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
        cudaFree(h->d_flag_ctr);
        if (h->h_flags) cudaFreeHost((void *)h->h_flags);
    }
    cudaFree(h->d_highway); cudaFree(h->d_home); cudaFree(h->d_home_id); cudaFree(h->d_cta_tables); cudaFree(h->d_normx); cudaFree(h->d_normy);
    cudaFree(h->d_state); cudaFree(h->d_rng); cudaFree(h->d_err);
    for (int r = 0; r < h->gather_world; r++)
        if (r != h->gather_rank && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    cudaFree(h->d_gather);
    cudaFree(h->d_actions); cudaFree(h->d_obs); cudaFree(h->d_reward); cudaFree(h->d_done);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    delete h;
    return RW_OK;
}

int rw_get_info(const rw_handle *h, rw_info *out) {
    if (!h || !out) return RW_EINVAL;
    *out = h->info;
    return RW_OK;
}

int rw_debug_set_clock_buffer(rw_handle *h, void *dev_ptr) {  // diagnostics: [CTA][32][8] uint64 phase timestamps
    if (!h) return RW_EINVAL;
    h->p.dbg_clock = (unsigned long long *)dev_ptr;
    h->host_epoch++;
    return RW_OK;
}

int rw_set_auto_reset(rw_handle *h, int enabled) {
    if (!h) return RW_EINVAL;
    h->p.auto_reset = enabled ? 1 : 0;
    h->host_epoch++;
    return RW_OK;
}

int rw_seed(rw_handle *h, const uint64_t *seeds_host, const uint8_t *mask_host, void *stream) {
    if (!h || !seeds_host) return fail(h, RW_EINVAL, "rw_seed: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(h->device));
    const int64_t E = h->p.E;
    std::vector<uint32_t> staging((size_t)RW_RNG_WORDS * E);
    uint32_t w[RW_RNG_WORDS];
    for (int64_t e = 0; e < E; e++) {
        if (mask_host && !mask_host[e]) continue;
        rw_seed_words(seeds_host[e], w);
        for (int k = 0; k < RW_RNG_WORDS; k++) staging[(size_t)k * E + e] = w[k];
    }
    uint32_t *d_staging = nullptr;
    uint8_t *d_mask = nullptr;
    CK(cudaMallocAsync((void **)&d_staging, staging.size() * 4, st));
    CK(cudaMemcpyAsync(d_staging, staging.data(), staging.size() * 4, cudaMemcpyHostToDevice, st));
    if (mask_host) {
        CK(cudaMallocAsync((void **)&d_mask, (size_t)E, st));
        CK(cudaMemcpyAsync(d_mask, mask_host, (size_t)E, cudaMemcpyHostToDevice, st));
    }
    CK(rw_launch_seed_scatter(h->p, d_staging, d_mask, st));
    CK(cudaFreeAsync(d_staging, st));
    if (d_mask) CK(cudaFreeAsync(d_mask, st));
    CK(cudaStreamSynchronize(st));  // `staging` is pageable host memory that must outlive the copy
    return RW_OK;
}

int rw_reset(rw_handle *h, const uint8_t *mask_dev, float *obs_dev, void *stream) {
    if (!h) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    if (mask_dev && !h->ready) return fail(h, RW_ESTATE, "rw_reset: the first reset must cover all worlds (mask == NULL)");
    CK(rw_launch_main(RW_MODE_RESET, h->p, nullptr, obs_dev, nullptr, nullptr, mask_dev, (cudaStream_t)stream));
    h->ready = true;
    return RW_OK;
}

int rw_step(rw_handle *h, const int32_t *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev, void *stream) {
    if (!h) return RW_EINVAL;
    if (!actions_dev || !reward_dev || !done_dev) return fail(h, RW_EINVAL, "rw_step: actions/reward/done are required");
    if (!h->ready) return fail(h, RW_ESTATE, "rw_step before rw_reset / rw_set_state");
    CK(cudaSetDevice(h->device));
    CK(rw_launch_main(RW_MODE_STEP, h->p, actions_dev, obs_dev, reward_dev, done_dev, nullptr, (cudaStream_t)stream));
    return RW_OK;
}

int rw_obs(rw_handle *h, float *obs_dev, void *stream) {
    if (!h || !obs_dev) return fail(h, RW_EINVAL, "rw_obs: null argument");
    if (!h->ready) return fail(h, RW_ESTATE, "rw_obs before rw_reset / rw_set_state");
    CK(cudaSetDevice(h->device));
    CK(rw_launch_main(RW_MODE_OBS, h->p, nullptr, obs_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream));
    return RW_OK;
}

int rw_get_state(rw_handle *h, const rw_state_view *dev, void *stream) {
    if (!h || !dev) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(rw_launch_get_state(h->p, dev, (cudaStream_t)stream));
    return RW_OK;
}

int rw_set_state(rw_handle *h, const rw_state_view *dev, void *stream) {
    if (!h || !dev) return RW_EINVAL;
    if (!h->ready && !(dev->agent_x && dev->agent_y && dev->agent_dir && dev->agent_carry && dev->agent_has_delivered &&
                       dev->shelf_x && dev->shelf_y && (dev->queue || h->p.R == 0) && dev->cur_steps && dev->cur_inactive))
        return fail(h, RW_ESTATE, "rw_set_state on a fresh handle needs every field (rng optional)");
    CK(cudaSetDevice(h->device));
    CK(rw_launch_set_state(h->p, dev, (cudaStream_t)stream));
    h->ready = true;
    return RW_OK;
}

int rw_error_count(rw_handle *h, int64_t *out, void *stream) {
    if (!h || !out) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, h->d_err, 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    *out = (int64_t)v;
    return RW_OK;
}

int rw_flush_l2(rw_handle *h, void *buf_dev, int64_t bytes, int value, void *stream) {
    if (!h || !buf_dev || bytes < 16) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(rw_launch_flush_l2(buf_dev, (size_t)bytes, value, h->p.sm_cta_bytes + h->p.sm_warp_bytes * h->p.warps_per_cta, (cudaStream_t)stream));
    return RW_OK;
}

static void gather_set_peer(rw_handle *h, int r, void *base) { h->peer_base[r] = base; }

// pointers of slot `slot` of rank r's gathered tensors
static void gather_slot_ptrs(const rw_handle *h, int r, int slot, float **obs, float **rew, uint8_t **done) {
    char *base = (char *)h->peer_base[r] + (size_t)slot * h->gather_slot_bytes;
    *obs = (float *)base;
    *rew = (float *)(base + h->gather_off_rew);
    *done = (uint8_t *)(base + h->gather_off_done);
}

int rw_gather_create(rw_handle *h, int world_size, int rank, float **obs_all, float **reward_all, uint8_t **done_all,
                     void *ipc_handle_out) {
    if (!h || world_size < 1 || world_size > RW_MAX_PEERS || rank < 0 || rank >= world_size || !ipc_handle_out)
        return fail(h, RW_EINVAL, "rw_gather_create: bad arguments (world_size 1..8)");
    if (h->d_gather) return fail(h, RW_ESTATE, "rw_gather_create: already created");
    CK(cudaSetDevice(h->device));
    const RwParams &p = h->p;
    const size_t W = (size_t)world_size, E = (size_t)p.E;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    h->gather_off_rew = up(W * E * p.n * p.L * 4);
    h->gather_off_done = h->gather_off_rew + up(W * E * p.n * 4);
    h->gather_slot_bytes = h->gather_off_done + up(W * E);
    // TWO slots in one allocation (one IPC handle): step k lands in slot k & 1, so a peer that is already pushing step k+1
    // cannot touch the tensors this rank's consumer is still reading for step k (see rw_step_gather)
    CK(cudaMalloc(&h->d_gather, 2 * h->gather_slot_bytes));
    CK(cudaMemset(h->d_gather, 0, 2 * h->gather_slot_bytes));
    h->gather_world = world_size;
    h->gather_rank = rank;
    h->gather_steps = 0;
    gather_set_peer(h, rank, h->d_gather);
    cudaIpcMemHandle_t ipc;
    CK(cudaIpcGetMemHandle(&ipc, h->d_gather));
    static_assert(sizeof(ipc) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &ipc, 64);
    float *o; float *r; uint8_t *d;
    gather_slot_ptrs(h, rank, 0, &o, &r, &d);
    if (obs_all) *obs_all = o;
    if (reward_all) *reward_all = r;
    if (done_all) *done_all = d;
    return RW_OK;
}

int rw_gather_open(rw_handle *h, const void *ipc_handles) {
    if (!h || !ipc_handles || !h->d_gather) return fail(h, RW_ESTATE, "rw_gather_open before rw_gather_create");
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < h->gather_world; r++) {
        if (r == h->gather_rank) continue;
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char *)ipc_handles + 64 * (size_t)r, 64);
        void *base = nullptr;
        CK(cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess));
        gather_set_peer(h, r, base);
    }
    h->gather_open = true;
    return RW_OK;
}

int rw_gather_slot(const rw_handle *h, int slot, float **obs_all, float **reward_all, uint8_t **done_all) {
    if (!h || !h->d_gather || slot < 0 || slot > 1) return RW_EINVAL;
    float *o; float *r; uint8_t *d;
    gather_slot_ptrs(h, h->gather_rank, slot, &o, &r, &d);
    if (obs_all) *obs_all = o;
    if (reward_all) *reward_all = r;
    if (done_all) *done_all = d;
    return RW_OK;
}

int rw_step_gather(rw_handle *h, const int32_t *actions_dev, void *stream) {
    if (!h || !actions_dev) return RW_EINVAL;
    if (!h->ready) return fail(h, RW_ESTATE, "rw_step_gather before rw_reset / rw_set_state");
    if (!h->d_gather || (h->gather_world > 1 && !h->gather_open))
        return fail(h, RW_ESTATE, "rw_step_gather before rw_gather_create / rw_gather_open");
    CK(cudaSetDevice(h->device));
    RwParams p = h->p;  // the plain entry points keep n_peers == 0
    const int slot = (int)(h->gather_steps & 1);
    p.n_peers = h->gather_world;
    p.my_rank = h->gather_rank;
    p.gather_env0 = (int64_t)h->gather_rank * p.E;
    for (int r = 0; r < h->gather_world; r++) gather_slot_ptrs(h, r, slot, &p.peer_obs[r], &p.peer_rew[r], &p.peer_done[r]);
    float *obs = p.peer_obs[p.my_rank] + (size_t)p.gather_env0 * p.n * p.L;
    float *rew = p.peer_rew[p.my_rank] + (size_t)p.gather_env0 * p.n;
    uint8_t *done = p.peer_done[p.my_rank] + (size_t)p.gather_env0;
    CK(rw_launch_main(RW_MODE_STEP, p, actions_dev, obs, rew, done, nullptr, (cudaStream_t)stream));
    h->gather_steps++;
    return slot;
}

int rw_random_actions(rw_handle *h, int32_t *actions_dev, uint64_t seed, uint64_t step, void *stream) {
    if (!h || !actions_dev) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(rw_launch_random_actions(h->p, actions_dev, seed, step, (cudaStream_t)stream));
    return RW_OK;
}

int rw_global_image_shape(const rw_handle *h, int32_t n_layers, const int32_t *pad_to_shape, int32_t *shape_out) {
    if (!h || !shape_out || n_layers < 1 || n_layers > 8) return RW_EINVAL;
    const int32_t dims[3] = {n_layers, h->p.H, h->p.W};
    for (int k = 0; k < 3; k++) {
        shape_out[k] = pad_to_shape ? pad_to_shape[k] : dims[k];
        if (shape_out[k] < dims[k]) return RW_EINVAL;  // the reference asserts every padding is >= 0 (warehouse.py:1027)
    }
    return RW_OK;
}

int rw_global_image(rw_handle *h, const int32_t *layers, int32_t n_layers, const int32_t *pad_to_shape, float *out_dev, void *stream) {
    if (!h || !layers || !out_dev) return fail(h, RW_EINVAL, "rw_global_image: null argument");
    if (!h->ready) return fail(h, RW_ESTATE, "rw_global_image before rw_reset / rw_set_state");
    RwViewSpec s;
    memset(&s, 0, sizeof(s));
    int32_t shape[3];
    if (rw_global_image_shape(h, n_layers, pad_to_shape, shape) != RW_OK)
        return fail(h, RW_EINVAL, "rw_global_image: 1..8 layers, pad_to_shape >= (n_layers, H, W) elementwise");
    for (int c = 0; c < n_layers; c++) {
        if (layers[c] < 0 || layers[c] > 6) return fail(h, RW_EINVAL, "rw_global_image: unknown image layer type");  // ValueError 1018-1019
        s.layers[c] = layers[c];
    }
    s.n_layers = n_layers;
    s.C = shape[0]; s.Hp = shape[1]; s.Wp = shape[2];
    s.c0 = (shape[0] - n_layers) / 2; s.y0 = (shape[1] - h->p.H) / 2; s.x0 = (shape[2] - h->p.W) / 2;  // pad_before 1028
    CK(cudaSetDevice(h->device));
    CK(rw_launch_global_image(h->p, s, out_dev, (cudaStream_t)stream));
    return RW_OK;
}

int rw_render_shape(const rw_handle *h, int32_t *height, int32_t *width) {
    if (!h || !height || !width) return RW_EINVAL;
    *height = rw_render_height(h->p);
    *width = rw_render_width(h->p);
    return RW_OK;
}

int rw_render_rgb(rw_handle *h, int64_t first_env, int64_t n_envs, uint8_t *out_dev, void *stream) {
    if (!h || !out_dev) return fail(h, RW_EINVAL, "rw_render_rgb: null argument");
    if (!h->ready) return fail(h, RW_ESTATE, "rw_render_rgb before rw_reset / rw_set_state");
    if (first_env < 0 || n_envs < 1 || first_env + n_envs > h->p.E) return fail(h, RW_EINVAL, "rw_render_rgb: world range out of bounds");
    CK(cudaSetDevice(h->device));
    CK(rw_launch_render(h->p, out_dev, first_env, n_envs, (cudaStream_t)stream));
    return RW_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ host-buffer entry points
// NUMA node the GPU hangs off (sysfs), -1 if unknown
static int gpu_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[96];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    return node;
}

// Page-locked host memory placed for the transport of rw_step_host.  The expansion of the observation rows is bound by the
// DRAM write bandwidth of the socket(s) holding the pages, so a large buffer is STRIPED over the NUMA nodes (2 MB stripes, the
// first one on the GPU's node) and the host pool expands every stripe with threads of its own socket; the last `local_tail`
// of the buffer -- the rows the copy engine delivers -- stays on the node the GPU is attached to, so that the DMA writes do
// not cross the socket interconnect.  Pages are placed by first touch from this thread while it is confined to the node's
// CPUs, then registered with CUDA.  Best effort: without NUMA information this is a plain cudaHostAlloc.
static std::mutex g_reg_mu;
struct RegBlock { void *base; size_t maplen, reglen; };
static std::map<void *, RegBlock> g_registered;  // mmap'ed + cudaHostRegister'ed blocks handed out by alloc_pinned_placed

static cudaError_t alloc_pinned_placed(void **out, size_t bytes, int gpu_node, double local_tail) {
    const int nodes = rw_host_numa_nodes();
    cpu_set_t old;
    if (nodes < 2 || gpu_node < 0 || gpu_node >= nodes || sched_getaffinity(0, sizeof(old), &old) != 0)
        return cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    static const size_t stripe = (size_t)(getenv("RWARE_B200_HOST_STRIPE_MB") ? atoi(getenv("RWARE_B200_HOST_STRIPE_MB")) : 2) << 20;
    if (getenv("RWARE_B200_HOST_LOCAL_TAIL")) local_tail = atof(getenv("RWARE_B200_HOST_LOCAL_TAIL"));
    const size_t huge = (size_t)2 << 20;
    const size_t len = (bytes + 4095) & ~(size_t)4095, maplen = len + huge;
    void *base = mmap(nullptr, maplen, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    char *p = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(base) + huge - 1) & ~(uintptr_t)(huge - 1));
    madvise(p, len, MADV_HUGEPAGE);
    size_t striped = 0;
    if (stripe && len >= ((size_t)8 << 20) && local_tail < 1.0) striped = (size_t)((1.0 - local_tail) * (double)len) / stripe * stripe;
    for (int k = 0; k < nodes; k++) {  // first touch: the default policy places a page on the node of the thread that touches it
        const int node = (gpu_node + k) % nodes;
        cpu_set_t want;
        const bool moved = rw_host_cpus_of_node(node, &want) && sched_setaffinity(0, sizeof(want), &want) == 0;
        for (size_t off = (size_t)k * stripe; stripe && off < striped; off += (size_t)nodes * stripe) memset(p + off, 0, stripe);
        if (k == 0) memset(p + striped, 0, len - striped);
        (void)moved;
    }
    sched_setaffinity(0, sizeof(old), &old);
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        munmap(base, maplen);
        (void)cudaGetLastError();
        return cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    }
    std::lock_guard<std::mutex> g(g_reg_mu);
    g_registered[p] = RegBlock{base, maplen, len};
    *out = p;
    return cudaSuccess;
}

static cudaError_t free_pinned(void *p) {
    if (!p) return cudaSuccess;
    RegBlock blk{nullptr, 0, 0};
    {
        std::lock_guard<std::mutex> g(g_reg_mu);
        auto it = g_registered.find(p);
        if (it != g_registered.end()) { blk = it->second; g_registered.erase(it); }
    }
    if (!blk.base) return cudaFreeHost(p);
    cudaError_t e = cudaHostUnregister(p);
    munmap(blk.base, blk.maplen);
    return e;
}

static int ensure_host_staging(rw_handle *h) {
    if (h->host_stream) return RW_OK;
    const RwParams &p = h->p;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&h->d_actions, (size_t)p.E * p.n * p.A * 4));
    CK(cudaMalloc((void **)&h->d_obs, (size_t)p.E * p.n * p.L * 4));
    CK(cudaMalloc((void **)&h->d_reward, (size_t)p.E * p.n * 4));
    CK(cudaMalloc((void **)&h->d_done, (size_t)p.E));
    return RW_OK;
}

// Everything one call puts on the stream: actions in, the step (one or two launches), the compressed rows out in pieces of
// growing size (each followed by a flag the host can see), rewards / done, and the copy engine's share of the float32 rows.
static int enqueue_step_host(rw_handle *h, const HostPlan &pl, bool pdl_ok) {
    const RwParams &p = h->p;
    cudaStream_t st = h->host_stream;
    const size_t n_rows = (size_t)p.E * p.n, L = (size_t)p.L, n_obs = n_rows * L, n_cpu = pl.rows_cpu * L;
    RwParams float_part;
    int64_t float_env0 = -1;
    if (pl.fused) {
        const int64_t e_cpu = (int64_t)(pl.rows_cpu / p.n);
        auto range = [&](int64_t env0, int64_t count) {  // the same launch for the worlds [env0, env0 + count), env0 % 32 == 0
            RwParams q = p;
            q.state += (size_t)(env0 / RW_TILE) * p.nw * RW_TILE;
            q.rng += env0;  // E_pad (the stride of the generator words) stays
            q.E = count;
            q.tiles = (int)((count + RW_TILE - 1) / RW_TILE);
            return q;
        };
        RwParams a = range(0, e_cpu);
        a.obs_bits_out = (uint32_t *)h->d_pack;
        a.obs_wide_out = h->d_pack + pl.off_wide;
        CK(rw_launch_main(RW_MODE_STEP, a, h->d_actions, nullptr, h->d_reward, h->d_done, nullptr, st, pdl_ok));
        if (e_cpu < p.E) {  // the copy engine's worlds: launched below, behind the first piece of the compressed rows
            float_part = range(e_cpu, p.E - e_cpu);
            float_env0 = e_cpu;
        }
    } else {
        CK(rw_launch_main(RW_MODE_STEP, p, h->d_actions, pl.obs_host ? h->d_obs : nullptr, h->d_reward, h->d_done, nullptr, st, pdl_ok));
    }
    for (int c = 0; c < pl.nch; c++) {
        const size_t a = (c ? pl.chunk_end[c - 1] : 0) * RW_HOST_BLOCK;
        const size_t b = pl.chunk_end[c] * RW_HOST_BLOCK > n_cpu ? n_cpu : pl.chunk_end[c] * RW_HOST_BLOCK;
        if (pl.bits) {
            const size_t ba = a / 8, bb = (b + 7) / 8;
            const size_t ra = a / L, rb = (b + L - 1) / L > pl.rows_cpu ? pl.rows_cpu : (b + L - 1) / L;
            if (!pl.fused)
                CK(rw_launch_pack_bits(h->d_obs, (uint32_t *)h->d_pack, h->d_pack + pl.off_wide, a, b, (int)L, pl.nwide, ra, rb, st));
            CK(cudaMemcpyAsync(h->h_pack + ba, h->d_pack + ba, bb - ba, cudaMemcpyDeviceToHost, st));
            if (pl.nwide)
                CK(cudaMemcpyAsync(h->h_pack + pl.off_wide + ra * pl.nwide, h->d_pack + pl.off_wide + ra * pl.nwide,
                                   (rb - ra) * pl.nwide, cudaMemcpyDeviceToHost, st));
        } else {
            CK(rw_launch_pack_u8(h->d_obs + a, h->d_pack + a, b - a, st));
            CK(cudaMemcpyAsync(h->h_pack + a, h->d_pack + a, b - a, cudaMemcpyDeviceToHost, st));
        }
        CK(rw_launch_host_flag(h->d_flag_ctr + c, h->h_flags + 16 * c, st));
        if (c == 0 && float_env0 >= 0) {
            CK(rw_launch_main(RW_MODE_STEP, float_part, h->d_actions + float_env0 * p.n * p.A, h->d_obs + (size_t)float_env0 * p.n * L,
                              h->d_reward + float_env0 * p.n, h->d_done + float_env0, nullptr, st, pdl_ok));
            float_env0 = -1;
        }
    }
    if (float_env0 >= 0)  // (no pieces: cannot happen with a fused plan, kept for safety)
        CK(rw_launch_main(RW_MODE_STEP, float_part, h->d_actions + float_env0 * p.n * p.A, h->d_obs + (size_t)float_env0 * p.n * L,
                          h->d_reward + float_env0 * p.n, h->d_done + float_env0, nullptr, st, pdl_ok));
    if (pl.reward_host) CK(cudaMemcpyAsync(pl.reward_host, h->d_reward, (size_t)p.E * p.n * 4, cudaMemcpyDeviceToHost, st));
    if (pl.done_host) CK(cudaMemcpyAsync(pl.done_host, h->d_done, (size_t)p.E, cudaMemcpyDeviceToHost, st));
    if (pl.obs_host && n_cpu < n_obs) {  // the copy engine's share (everything, for the plain transport).  Device-to-host copies
        // of different streams do not overlap on this hardware, so it simply follows the compressed pieces on the same stream.
        if (n_cpu) CK(rw_launch_host_flag(h->d_flag_ctr + 9, h->h_flags + 16 * 9, st));  // stamps the start of the copy
        CK(cudaMemcpyAsync(pl.obs_host + n_cpu, h->d_obs + n_cpu, (n_obs - n_cpu) * 4, cudaMemcpyDeviceToHost, st));
        if (n_cpu) CK(rw_launch_host_flag(h->d_flag_ctr + 8, h->h_flags + 16 * 8, st));               // and its end
    }
    return RW_OK;
}

extern "C" {

int rw_step_host(rw_handle *h, const int32_t *actions_host, float *obs_host, float *reward_host, uint8_t *done_host) {
    if (!h || !actions_host) return RW_EINVAL;
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    if (!h->ready) return fail(h, RW_ESTATE, "rw_step_host before rw_reset / rw_set_state");
    const RwParams &p = h->p;
    cudaStream_t st = h->host_stream;
    const size_t n_rows = (size_t)p.E * p.n, L = (size_t)p.L, n_obs = n_rows * L;
    // What the caller gets is float32 rows in host memory.  The plain way is one device-to-host copy of them (PCIe bound).
    // Unless the coordinates are normalised every value is a small integer, and in large batches the rows travel compressed
    // and are expanded by the host pool (rw_hostpack.cpp): one BIT per value plus one byte per coordinate when everything
    // else is a 0/1 flag (FLATTENED rows; image rows without the AGENT_DIRECTION layer), one BYTE per value otherwise.  The
    // expansion is bound by host DRAM writes, so the copy engine delivers a share of the rows as plain float32 at the same
    // time; the share follows the measured finish times of the two (dma_frac).
    bool has_dir_layer = false;
    for (int i = 0; i < p.n_layers; i++) has_dir_layer |= (p.obs_kind != 0 && p.img_layers[i] == RW_LAYER_AGENT_DIRECTION);
    const bool compress = obs_host && h->host_pack != 0 && !p.norm_coords && n_obs >= ((size_t)1 << 20);
    HostPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.actions_host = actions_host; pl.obs_host = obs_host; pl.reward_host = reward_host; pl.done_host = done_host;
    pl.bits = compress && h->host_pack == 1 && !has_dir_layer && (reinterpret_cast<uintptr_t>(obs_host) & 63u) == 0;
    pl.nwide = p.obs_kind == 0 ? 2 : 0;
    pl.epoch = h->host_epoch;
    if (compress) {
        const double f = h->dma_frac_fixed >= 0 ? h->dma_frac_fixed : h->dma_frac;
        // whole tiles of 32 worlds (the fused bit form is produced per tile); in steps of 1/64 of the batch so that the
        // self-adjusting share settles on a handful of distinct plans (each plan has its own CUDA graph)
        size_t gran = 2 * (size_t)RW_TILE * p.n;
        const size_t coarse = n_rows / 64 / gran * gran;
        if (coarse > gran) gran = coarse;
        pl.rows_cpu = ((size_t)((1.0 - f) * (double)n_rows) + gran / 2) / gran * gran;  // nearest step
        if (f <= 0.0 || pl.rows_cpu > n_rows) pl.rows_cpu = n_rows;
        if (pl.rows_cpu * L < ((size_t)1 << 18)) pl.rows_cpu = 0;
    }
    const size_t n_cpu = pl.rows_cpu * L;
    // FLATTENED rows come out of the step kernel in bit form directly (rw_obs_agent_bits): the worlds whose rows the host pool
    // expands are stepped by one launch that writes bits + coordinate bytes, the copy engine's worlds by a second launch that
    // writes float32 rows.  Other row types are packed from the float32 tensor by a separate small kernel.
    static const bool allow_fused = !(getenv("RWARE_B200_HOST_FUSED") && atoi(getenv("RWARE_B200_HOST_FUSED")) == 0);
    pl.fused = pl.bits && pl.rows_cpu && p.obs_kind == 0 && allow_fused &&
               (size_t)p.obs_buf_stride * p.obs_nbuf >= (size_t)p.n * L * 4 + 64u * p.n;
    const size_t rows_pad = (size_t)p.E_pad * p.n;  // the fused form writes whole tiles
    const size_t bits_bytes = ((rows_pad * L + 31) / 32) * 4;
    pl.off_wide = (bits_bytes + 255) & ~(size_t)255;
    if (!h->d_pack) {
        h->pack_bytes = n_obs + 256;  // the byte transport needs n_obs, the bit transport n_obs / 8 + 2 bytes per row
        if (h->pack_bytes < pl.off_wide + rows_pad * 2 + 256) h->pack_bytes = pl.off_wide + rows_pad * 2 + 256;
        CK(cudaMalloc(&h->d_pack, h->pack_bytes));
        CK(alloc_pinned_placed((void **)&h->h_pack, h->pack_bytes, gpu_numa_node(h->device), 1.0));
        CK(cudaMalloc((void **)&h->d_flag_ctr, 16 * 4));
        CK(cudaMemset(h->d_flag_ctr, 0, 16 * 4));
        CK(cudaHostAlloc((void **)&h->h_flags, 16 * 64, cudaHostAllocMapped | cudaHostAllocPortable));
        memset((void *)h->h_flags, 0, 16 * 64);
        memset(h->flag_seen, 0, sizeof(h->flag_seen));
    }
    if (n_cpu) {
        // The compressed rows reach the host in a few pieces of growing size (sizes 1 : 2 : 3 : ...); the pool starts expanding
        // when the first piece has landed.
        static const int want_chunks = getenv("RWARE_B200_HOST_CHUNKS") ? atoi(getenv("RWARE_B200_HOST_CHUNKS")) : 4;
        const size_t nblocks = (n_cpu + RW_HOST_BLOCK - 1) / RW_HOST_BLOCK;
        int nch = want_chunks < 1 ? 1 : (want_chunks > 8 ? 8 : want_chunks);
        if ((size_t)nch > nblocks) nch = (int)nblocks;
        size_t prev = 0;
        int k = 0;
        for (int c = 0; c < nch; c++) {
            size_t end = c + 1 == nch ? nblocks : nblocks * (size_t)(c + 1) * (c + 2) / ((size_t)nch * (nch + 1));
            if (end <= prev) continue;
            pl.chunk_end[k++] = end;
            prev = end;
        }
        pl.nch = k;
    }
    const bool timed_dma = n_cpu && n_cpu < n_obs;
    static const bool trace = getenv("RWARE_B200_HOST_TRACE") != nullptr;  // diagnostics: where the call spends its time
    static const bool use_graph = !(getenv("RWARE_B200_HOST_GRAPH") && atoi(getenv("RWARE_B200_HOST_GRAPH")) == 0);
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double t_call = now_us();
    if (n_cpu) rw_hostpool_prewake();  // parked workers wake up while the copies and the kernel are in flight
    // The ~20 stream operations of a call cost more host time than the GPU needs to execute the first of them, so a plan is
    // captured into a CUDA graph the first time it is seen and replayed with a single launch afterwards.
    CK(cudaMemcpyAsync(h->d_actions, actions_host, (size_t)p.E * p.n * p.A * 4, cudaMemcpyHostToDevice, st));  // not part of the graph: callers rotate action buffers
    bool launched = false;
    if (use_graph) {
        HostGraph *g = nullptr, *victim = nullptr;
        for (auto &cand : h->graphs) {
            if (cand.exec && cand.plan.same(pl)) { g = &cand; break; }
            if (!victim || (victim->exec && (!cand.exec || cand.last_use < victim->last_use))) victim = &cand;  // empty, else least recent
        }
        if (!g) {
            g = victim;
            if (g->exec) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            rc = enqueue_step_host(h, pl, /*pdl_ok=*/false);
            cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == RW_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess) {
                g->plan = pl;
            } else {
                g->exec = nullptr;
                (void)cudaGetLastError();
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (g->exec) {
            g->last_use = ++h->host_calls_graph;
            CK(cudaGraphLaunch(g->exec, st));
            launched = true;
        }
    }
    if (!launched) {
        rc = enqueue_step_host(h, pl, true);
        if (rc) return rc;
    }
    const double t_enq = now_us();
    double t_first = 0, t_last = 0, t_done = 0;
    auto flag_wait = [&](int idx, bool block) -> bool {  // the flag kernel behind piece idx bumps a counter the host can see
        volatile uint32_t *f = h->h_flags + 16 * idx;
        const uint32_t want = h->flag_seen[idx] + 1;
        if (!block) return (int32_t)(*f - want) >= 0;
        unsigned spins = 0;
        while ((int32_t)(*f - want) < 0) {
            if ((++spins & 0x3ff) == 0 && cudaStreamQuery(st) != cudaErrorNotReady) {  // the stream finished or failed
                if ((int32_t)(*f - want) >= 0) break;
                return false;
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        return true;
    };
    bool flags_ok = true;
    if (n_cpu) {
        RwHostJob job;
        job.kind = pl.bits ? RW_HOSTJOB_BITS : RW_HOSTJOB_WIDEN;
        job.src = h->h_pack;
        job.wide = h->h_pack + pl.off_wide;
        job.dst = obs_host;
        job.n = n_cpu;
        job.L = (int)L;
        job.nwide = pl.nwide;
        job.n_rows = pl.rows_cpu;
        job.nchunks = pl.nch;
        for (int c = 0; c < pl.nch; c++) job.chunk_end[c] = pl.chunk_end[c];
        rw_hostpool_begin(job);
        for (int c = 0; c < pl.nch && flags_ok; c++) {
            flags_ok = flag_wait(c, true);
            if (flags_ok) rw_hostpool_chunk_ready(c + 1);
            if (c == 0) t_first = now_us();
        }
        t_last = now_us();
        if (!flags_ok) rw_hostpool_cancel();
        rw_hostpool_finish();
        t_done = now_us();
    }
    bool dma_late = false;
    if (timed_dma && flags_ok) dma_late = !flag_wait(8, false);
    CK(cudaStreamSynchronize(st));
    const double t_end = now_us();
    for (int c = 0; c < 10; c++) h->flag_seen[c] = h->h_flags[16 * c];
    if (!flags_ok) return fail(h, RW_ECUDA, "rw_step_host: the stream went idle before a piece of the compressed rows was flagged");
    double dma_us = 0, dma_end = 0;
    if (timed_dma) {
        auto stamp_ns = [&](int idx) { return (double)*reinterpret_cast<volatile unsigned long long *>(h->h_flags + 16 * idx + 2); };
        dma_us = (stamp_ns(8) - stamp_ns(9)) * 1e-3;  // GPU timer: start / end of the copy engine's share
        dma_end = dma_late ? t_end : t_first + (stamp_ns(8) - stamp_ns(0)) * 1e-3;  // on the host clock (piece 0 was seen at t_first)
        // re-balance so that the expansion and the copy engine FINISH together.  With the rates observed in this call (each
        // includes what the other took from the memory system) and the copy's start d0 relative to the expansion's, the copy
        // engine's bytes x solve (B - x) / r_cpu = d0 + x / r_dma.
        const double B = (double)n_obs * 4, x_now = (double)(n_obs - n_cpu) * 4;
        const double r_cpu = (B - x_now) / ((t_done - t_first) > 1 ? (t_done - t_first) : 1);  // bytes / us
        const double r_dma = x_now / (dma_us > 1 ? dma_us : 1);
        const double d0 = (dma_end - dma_us) - t_first;
        double target = (B / r_cpu - d0) / (1.0 / r_cpu + 1.0 / r_dma) / B;
        if (!(target > 0.02)) target = 0.02;
        if (target > 0.9) target = 0.9;
        h->dma_frac = 0.5 * h->dma_frac + 0.5 * target;
    }
    h->host_calls++;
    if (trace)
        fprintf(stderr, "[rw_step_host] %s%s rows_cpu %zu/%zu pieces %d | enqueued %.0f first-piece %.0f last-piece %.0f expand-done %.0f "
                        "end %.0f us | dma %.0f us, ends at %.0f | dma_frac -> %.3f | threads %d isa %d\n",
                !n_cpu ? "plain" : pl.fused ? "bits(fused)" : pl.bits ? "bits" : "bytes", launched ? " graph" : "", pl.rows_cpu, n_rows,
                pl.nch, t_enq - t_call, n_cpu ? t_first - t_call : 0.0, n_cpu ? t_last - t_call : 0.0, n_cpu ? t_done - t_call : 0.0,
                t_end - t_call, dma_us, timed_dma ? dma_end - t_call : 0.0, h->dma_frac, n_cpu ? rw_hostpool_threads() : 0,
                n_cpu ? rw_hostpool_isa() : 0);
    return RW_OK;
}

int rw_obs_bytes(rw_handle *h, const float *obs_dev, uint8_t *out_dev, void *stream) {
    if (!h || !obs_dev || !out_dev) return fail(h, RW_EINVAL, "rw_obs_bytes: null argument");
    if (h->p.norm_coords) return fail(h, RW_EINVAL, "rw_obs_bytes: normalised coordinates are not integers");
    CK(cudaSetDevice(h->device));
    CK(rw_launch_pack_u8(obs_dev, out_dev, (size_t)h->p.E * h->p.n * h->p.L, (cudaStream_t)stream));
    return RW_OK;
}

int rw_set_host_transport(rw_handle *h, int mode) {
    if (!h || mode < 0 || mode > 2) return RW_EINVAL;
    h->host_pack = mode;
    h->host_calls = 0;
    h->dma_frac = RW_DMA_FRAC0;
    return RW_OK;
}

int rw_set_host_dma_share(rw_handle *h, double share) {
    if (!h || share > 1.0) return RW_EINVAL;
    h->dma_frac_fixed = share < 0 ? -1.0 : share;
    return RW_OK;
}

int rw_get_host_dma_share(const rw_handle *h, double *share) {
    if (!h || !share) return RW_EINVAL;
    *share = h->dma_frac_fixed >= 0 ? h->dma_frac_fixed : h->dma_frac;
    return RW_OK;
}

int rw_host_widen_u8(const uint8_t *src, float *dst, int64_t n) {
    if (!src || !dst || n < 0) return RW_EINVAL;
    RwHostJob job;
    job.kind = RW_HOSTJOB_WIDEN;
    job.src = src;
    job.dst = dst;
    job.n = (size_t)n;
    rw_hostpool_begin(job);
    rw_hostpool_finish();
    return rw_hostpool_threads();
}

int rw_host_expand_bits(const uint8_t *bits, const uint8_t *wide, float *dst, int64_t n_rows, int32_t row_len, int32_t n_wide) {
    if (!bits || !dst || n_rows < 0 || row_len < 1 || n_wide < 0 || n_wide > row_len || (n_wide && !wide)) return RW_EINVAL;
    RwHostJob job;
    job.kind = RW_HOSTJOB_BITS;
    job.src = bits;
    job.wide = wide;
    job.dst = dst;
    job.n = (size_t)n_rows * (size_t)row_len;
    job.L = row_len;
    job.nwide = n_wide;
    job.n_rows = (size_t)n_rows;
    rw_hostpool_begin(job);
    rw_hostpool_finish();
    return rw_hostpool_threads();
}

int rw_host_stream_probe(float *dst, int64_t n, int32_t repeats, double *gb_per_s) {
    if (!dst || n < 1 || repeats < 1 || !gb_per_s) return RW_EINVAL;
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    RwHostJob job;
    job.kind = RW_HOSTJOB_FILL;
    job.dst = dst;
    job.n = (size_t)n;
    double best = 0;
    rw_hostpool_prewake();
    for (int r = 0; r < repeats + 1; r++) {  // the first pass warms the pool up
        job.fill_value = (float)r;
        const double t0 = now_us();
        rw_hostpool_begin(job);
        rw_hostpool_finish();
        const double gbs = (double)n * 4 / ((now_us() - t0) * 1e3);
        if (r > 0 && gbs > best) best = gbs;
    }
    *gb_per_s = best;
    return rw_hostpool_threads();
}

int rw_reset_host(rw_handle *h, const uint64_t *seeds_host, float *obs_host) {
    if (!h) return RW_EINVAL;
    int rc = ensure_host_staging(h);
    if (rc) return rc;
    const RwParams &p = h->p;
    cudaStream_t st = h->host_stream;
    if (seeds_host) {
        rc = rw_seed(h, seeds_host, nullptr, st);
        if (rc) return rc;
    }
    rc = rw_reset(h, nullptr, obs_host ? h->d_obs : nullptr, st);
    if (rc) return rc;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, h->d_obs, (size_t)p.E * p.n * p.L * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return RW_OK;
}

int rw_host_alloc(void **out, int64_t bytes) {
    rw_handle *h = nullptr;
    if (!out || bytes <= 0) return RW_EINVAL;
    CK(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return RW_OK;
}

int rw_host_alloc_near(rw_handle *h, void **out, int64_t bytes) {
    if (!h || !out || bytes <= 0) return RW_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(alloc_pinned_placed(out, (size_t)bytes, gpu_numa_node(h->device), RW_HOST_LOCAL_TAIL));
    return RW_OK;
}

int rw_host_free(void *ptr) {
    rw_handle *h = nullptr;
    CK(free_pinned(ptr));
    return RW_OK;
}

int rw_host_numa_node(const void *ptr) {
    void *page = reinterpret_cast<void *>(reinterpret_cast<uintptr_t>(ptr) & ~(uintptr_t)4095);
    int status = -1;
    const long r = syscall(SYS_move_pages, 0, 1UL, &page, nullptr, &status, 0);
    return r == 0 && status >= 0 ? status : -1;
}

int rw_gpu_numa_node(const rw_handle *h) { return h ? gpu_numa_node(h->device) : -1; }

}  // extern "C"
