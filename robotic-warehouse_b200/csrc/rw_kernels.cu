// rw_kernels.cu -- cold-path kernels (state views, seeding, synthetic actions, L2 flush), kernel selection and the launch
// wrappers of the batched RWARE step path.  The fused kernel itself is rw_main_kernel.cuh / rw_inst.cu.
#include <cstdlib>

#include "rw_launch.h"
#include "rw_logic.cuh"
#include "../../include/rware_b200.h"
#include "rw_state_io.h"

// ---------------------------------------------------------------- cold-path kernels (state views, seeding, synthetic actions)
__global__ void rw_get_state_kernel(const RwParams p, const rw_state_view v) {
    int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    int32_t *t = p.state + (size_t)(env >> 5) * p.nw * RW_TILE + (env & 31);
    RwLane l{t, nullptr, nullptr, 0, nullptr, nullptr};
    const RwOff o = rw_off<RwDynamic>(p);
    for (int i = 0; i < p.n; i++) {
        uint32_t aw = rw_agent(l, i);
        if (v.agent_x) v.agent_x[env * p.n + i] = rw_ax(aw);
        if (v.agent_y) v.agent_y[env * p.n + i] = rw_ay(aw);
        if (v.agent_dir) v.agent_dir[env * p.n + i] = rw_adir(aw);
        if (v.agent_carry) v.agent_carry[env * p.n + i] = rw_acarry(aw);
        if (v.agent_has_delivered) v.agent_has_delivered[env * p.n + i] = rw_adeliv(aw);
        if (v.agent_msg && p.msg_bits) {
            uint32_t bits = (uint32_t)t[(p.w_msg + i) * RW_TILE];
            for (int b = 0; b < p.msg_bits; b++) v.agent_msg[(env * p.n + i) * p.msg_bits + b] = (bits >> b) & 1u;
        }
    }
    for (int k = 0; k < p.S; k++) {
        uint32_t c = (uint32_t)rw_shelf_cell(o, l, k);
        uint32_t y = rw_div(c, p.divW);
        if (v.shelf_x) v.shelf_x[env * p.S + k] = (int32_t)(c - y * p.W);
        if (v.shelf_y) v.shelf_y[env * p.S + k] = (int32_t)y;
    }
    if (v.queue)
        for (int q = 0; q < p.R; q++) v.queue[env * p.R + q] = rw_queue(o, l, q);
    if (v.cur_steps) v.cur_steps[env] = t[p.w_steps * RW_TILE];
    if (v.cur_inactive) v.cur_inactive[env] = (int32_t)((uint32_t)t[p.w_inact * RW_TILE] & ~RW_OVERLAP_BIT);
    if (v.rng)
        for (int w = 0; w < RW_RNG_WORDS; w++) v.rng[env * RW_RNG_WORDS + w] = p.rng[(size_t)w * p.E_pad + env];
}

__global__ void rw_set_state_kernel(const RwParams p, const rw_state_view v) {  // see rw_state_io.h
    int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    if (rw_set_state_world(p, v, env)) atomicAdd(p.err_count, 1ULL);
}

__global__ void rw_seed_scatter_kernel(const RwParams p, const uint32_t *__restrict__ staging, const uint8_t *__restrict__ mask) {
    int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    if (mask && !mask[env]) return;
    for (int w = 0; w < RW_RNG_WORDS; w++) p.rng[(size_t)w * p.E_pad + env] = staging[(size_t)w * p.E + env];
}

__global__ void rw_random_actions_kernel(const RwParams p, int32_t *__restrict__ actions, uint64_t seed, uint64_t step) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = p.E * p.n;
    if (idx >= total) return;
    uint64_t z = seed * 0x9e3779b97f4a7c15ULL + step * 0xbf58476d1ce4e5b9ULL + (uint64_t)idx * 0x94d049bb133111ebULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    actions[idx * p.A] = (int32_t)(((z >> 32) * 5ULL) >> 32);
    for (int b = 0; b < p.msg_bits; b++) actions[idx * p.A + 1 + b] = (int32_t)((z >> b) & 1ULL);
}

// Transport packing for rw_step_host: observations are exact integers < 256 stored as float32; four of them become one 32-bit
// word of bytes (widened back on the host, rw_hostpack.cpp).
__global__ void rw_pack_u8_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, size_t n) {
    const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = reinterpret_cast<const float4 *>(in)[i];
        reinterpret_cast<uint32_t *>(out)[i] = (uint32_t)v.x | ((uint32_t)v.y << 8) | ((uint32_t)v.z << 16) | ((uint32_t)v.w << 24);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[n4 * 4 + threadIdx.x] = (uint8_t)in[n4 * 4 + threadIdx.x];
}

cudaError_t rw_launch_pack_u8(const float *in, uint8_t *out, size_t n, cudaStream_t stream) {
    rw_pack_u8_kernel<<<148 * 8, 256, 0, stream>>>(in, out, n);
    return cudaGetLastError();
}

// Bit transport for rw_step_host: every value of a FLATTENED observation row is a 0/1 flag except the agent's own
// coordinates (rware/warehouse.py:636-674), so the host only needs one bit per value plus one byte per coordinate to rebuild
// the float32 rows (rw_hostpack.cpp).  bits word w, bit i <- in[32 w + i] != 0 (the bits of the "wide" positions are ignored by
// the host); wide[row * nwide + k] <- (uint8_t)in[row * L + k].  A warp turns 1024 consecutive floats into 32 words: 32
// coalesced 128-byte loads in flight, one ballot each, one coalesced 128-byte store.
__global__ void rw_pack_bits_kernel(const float *__restrict__ in, uint32_t *__restrict__ bits, uint8_t *__restrict__ wide, size_t f_lo,
                                    size_t f_hi, int L, int nwide, size_t row_lo, size_t row_hi) {
    // floats [f_lo, f_hi) (f_lo a multiple of 1024) -> bit words; rows [row_lo, row_hi) -> wide bytes
    const int lane = threadIdx.x & 31;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    const size_t w_lo = f_lo / 32, w_hi = (f_hi + 31) / 32;
    for (size_t g = (w_lo >> 5) + (tid >> 5); g * 32 < w_hi; g += nthreads >> 5) {
        const size_t f0 = g * 1024;
        uint32_t mine = 0;
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const size_t idx = f0 + (size_t)j * 32 + lane;
            const float v = idx < f_hi ? __ldcs(in + idx) : 0.0f;
            const uint32_t b = __ballot_sync(0xffffffffu, v != 0.0f);
            if (lane == j) mine = b;
        }
        if (g * 32 + lane < w_hi) bits[g * 32 + lane] = mine;
    }
    const size_t nw = (row_hi - row_lo) * (size_t)nwide;
    for (size_t i = tid; i < nw; i += nthreads) {
        const size_t row = row_lo + i / (size_t)nwide, k = i % (size_t)nwide;
        wide[row * (size_t)nwide + k] = (uint8_t)in[row * (size_t)L + k];
    }
}

cudaError_t rw_launch_pack_bits(const float *in, uint32_t *bits, uint8_t *wide, size_t f_lo, size_t f_hi, int L, int nwide,
                                size_t row_lo, size_t row_hi, cudaStream_t stream) {
    const size_t groups = (f_hi - f_lo + 1023) / 1024;
    int blocks = (int)((groups + 7) / 8);  // 8 warps per block, one 1024-float group per warp and pass
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    rw_pack_bits_kernel<<<blocks, 256, 0, stream>>>(in, bits, wide, f_lo, f_hi, L, nwide, row_lo, row_hi);
    return cudaGetLastError();
}

// Host-visible progress flag of rw_step_host: bumps a device counter and publishes it, with the GPU's nanosecond timer, to a line
// of mapped host memory; stream order puts it behind the device-to-host copy it announces.  (A counter rather than a value passed
// by the caller, and a timestamp rather than CUDA events, so that the launch can be replayed unchanged from a CUDA graph.)
__global__ void rw_host_flag_kernel(uint32_t *counter, volatile uint32_t *host_line) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const uint32_t v = *counter + 1;
    *counter = v;
    *reinterpret_cast<volatile unsigned long long *>(host_line + 2) = t;
    __threadfence_system();
    *host_line = v;
}

cudaError_t rw_launch_host_flag(uint32_t *counter, volatile uint32_t *host_line, cudaStream_t stream) {
    rw_host_flag_kernel<<<1, 1, 0, stream>>>(counter, host_line);
    return cudaGetLastError();
}

// L2 flush for benchmarks: streams `bytes` of writes through the L2 with the SAME shared-memory carve-out as the step
// kernel (a flush kernel with a different carve-out makes the next launch pay an SM reconfiguration).
__global__ void rw_flush_l2_kernel(int4 *buf, size_t n16, int value) {
    extern __shared__ uint8_t flush_smem[];
    if (threadIdx.x == 0 && n16 == 0) flush_smem[0] = 0;
    const int4 v = make_int4(value, value, value, value);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) buf[i] = v;
}

cudaError_t rw_launch_flush_l2(void *buf, size_t bytes, int value, int smem_bytes, cudaStream_t stream) {
    static int configured = -1;
    if (configured != smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(rw_flush_l2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        configured = smem_bytes;
    }
    rw_flush_l2_kernel<<<148, 512, smem_bytes, stream>>>((int4 *)buf, bytes / 16, value);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- launch wrappers
// Kernel selection.  The instantiations live in rw_inst.cu, compiled as several translation units (RW_INST_GROUP);
// group order = priority: fully static BASELINE configurations, then the registered geometries, then RwDynamic.
typedef void (*rw_kernel_fn)(const RwParams, const int32_t *, float *, float *, uint8_t *, const uint8_t *);
#define RW_DECLARE_GROUP(g) rw_kernel_fn rw_select_group_##g(RwMode mode, const RwParams &p, bool force_dynamic);
RW_DECLARE_GROUP(0) RW_DECLARE_GROUP(1) RW_DECLARE_GROUP(2) RW_DECLARE_GROUP(3) RW_DECLARE_GROUP(4) RW_DECLARE_GROUP(5) RW_DECLARE_GROUP(6)

static rw_kernel_fn select_kernel(RwMode mode, const RwParams &p, bool force_dynamic) {
    rw_kernel_fn (*groups[])(RwMode, const RwParams &, bool) = {rw_select_group_0, rw_select_group_1, rw_select_group_2,
                                                               rw_select_group_3, rw_select_group_4, rw_select_group_5,
                                                               rw_select_group_6};
    for (auto g : groups) {
        rw_kernel_fn f = g(mode, p, force_dynamic);
        if (f) return f;
    }
    return nullptr;  // unreachable: group 6 ends with RwDynamic
}

static bool g_force_dynamic = false;
void rw_force_dynamic_kernels(bool on) { g_force_dynamic = on; }

cudaError_t rw_configure_kernels(const RwParams &p, int smem_bytes_per_cta) {
    // The attribute is a property of the FUNCTION, shared by every handle of the process: it is raised to the device's opt-in
    // maximum once and never lowered (a later handle with a smaller CTA used to lower it under the earlier handles' feet).
    int dev = 0, cap = smem_bytes_per_cta;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int optin = 0;
        if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && optin > cap) cap = optin;
    }
    RwParams q = p;
    for (int bits = 0; bits < 2; bits++) {  // the float kernels of every mode, and the bit-form instantiation of the step kernel
        q.obs_bits_out = bits ? reinterpret_cast<uint32_t *>(uintptr_t(16)) : nullptr;
        for (int m = 0; m < (bits ? 1 : 3); m++) {
            for (int dyn = 0; dyn < 2; dyn++) {
                cudaError_t e = cudaFuncSetAttribute((const void *)select_kernel((RwMode)m, q, dyn != 0),
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
                if (e != cudaSuccess) return e;
            }
        }
    }
    return cudaSuccess;
}

cudaError_t rw_launch_main(RwMode mode, const RwParams &p, const int32_t *actions, float *obs, float *reward, uint8_t *done,
                           const uint8_t *mask, cudaStream_t stream, bool pdl_ok) {
    dim3 block(64 * p.warps_per_cta);  // a warp pair per tile
    dim3 grid((p.tiles + p.warps_per_cta - 1) / p.warps_per_cta);
    size_t smem = (size_t)p.sm_cta_bytes + (size_t)p.sm_warp_bytes * p.warps_per_cta;
    static int pdl = -1;
    if (pdl < 0) {
        const char *ev = getenv("RWARE_B200_PDL");
        pdl = ev ? atoi(ev) : 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_ok) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, select_kernel(mode, p, g_force_dynamic), p, actions, obs, reward, done, mask);
}

static inline dim3 per_env_grid(int64_t n) { return dim3((unsigned)((n + 255) / 256)); }

cudaError_t rw_launch_get_state(const RwParams &p, const void *view, cudaStream_t stream) {
    rw_get_state_kernel<<<per_env_grid(p.E), 256, 0, stream>>>(p, *static_cast<const rw_state_view *>(view));
    return cudaGetLastError();
}
cudaError_t rw_launch_set_state(const RwParams &p, const void *view, cudaStream_t stream) {
    rw_set_state_kernel<<<per_env_grid(p.E), 256, 0, stream>>>(p, *static_cast<const rw_state_view *>(view));
    return cudaGetLastError();
}
cudaError_t rw_launch_seed_scatter(const RwParams &p, const uint32_t *staging, const uint8_t *mask, cudaStream_t stream) {
    rw_seed_scatter_kernel<<<per_env_grid(p.E), 256, 0, stream>>>(p, staging, mask);
    return cudaGetLastError();
}
cudaError_t rw_launch_random_actions(const RwParams &p, int32_t *actions, uint64_t seed, uint64_t step, cudaStream_t stream) {
    rw_random_actions_kernel<<<per_env_grid(p.E * p.n), 256, 0, stream>>>(p, actions, seed, step);
    return cudaGetLastError();
}
