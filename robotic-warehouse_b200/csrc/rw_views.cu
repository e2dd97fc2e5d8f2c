// rw_views.cu -- kernels of the whole-warehouse views (rw_views.cuh): batched get_global_image and rgb_array frames.
// One block per tile of 32 worlds: stage the packed tile state, build one occupancy byte per cell and world, then all threads
// stream the output tensor of each world with coalesced stores (every element is a pure function of the staged bytes).
#include "rw_launch.h"
#include "rw_logic.cuh"
#include "rw_views.cuh"

namespace {

constexpr int kViewThreads = 256;

// shared memory: [nw][32] int32 tile words, then 32 occupancy maps of `gs` bytes each
__device__ __forceinline__ int rw_view_stage(const RwParams &p, int tile, uint8_t *sm, int gs, bool dir_layer, bool load_layer) {
    int32_t *tilew = reinterpret_cast<int32_t *>(sm);
    uint8_t *grid = sm + (size_t)p.nw * RW_TILE * 4;
    const int32_t *src = p.state + (size_t)tile * p.nw * RW_TILE;
    for (int i = threadIdx.x; i < p.nw * RW_TILE; i += blockDim.x) tilew[i] = src[i];
    for (int i = threadIdx.x; i < gs * RW_TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(grid)[i] = 0u;
    __syncthreads();
    int bad = 0;
    if (threadIdx.x < RW_TILE && (int64_t)tile * RW_TILE + threadIdx.x < p.E) {
        RwLane l{tilew + threadIdx.x, grid + (size_t)threadIdx.x * gs, nullptr, 0, nullptr, nullptr};
        bad = rw_view_build(p, l, dir_layer, load_layer);
    }
    __syncthreads();
    return bad;
}

__global__ void __launch_bounds__(kViewThreads) rw_global_image_kernel(const RwParams p, const RwViewSpec s, float *__restrict__ out, int gs) {
    extern __shared__ __align__(16) uint8_t sm[];
    bool dir_layer = false, load_layer = false;
    for (int c = 0; c < s.n_layers; c++) { dir_layer |= s.layers[c] == 3; load_layer |= s.layers[c] == 4; }
    const int tile = blockIdx.x;
    if (rw_view_stage(p, tile, sm, gs, dir_layer, load_layer)) atomicAdd(p.err_count, 1ULL);
    int32_t *tilew = reinterpret_cast<int32_t *>(sm);
    uint8_t *grid = sm + (size_t)p.nw * RW_TILE * 4;
    const int per = s.C * s.Hp * s.Wp;
    for (int w = 0; w < RW_TILE; w++) {
        const int64_t env = (int64_t)tile * RW_TILE + w;
        if (env >= p.E) break;
        RwLane l{tilew + w, grid + (size_t)w * gs, nullptr, 0, nullptr, nullptr};
        float *o = out + (size_t)env * per;
        for (int i = threadIdx.x; i < per; i += blockDim.x) o[i] = rw_gimg_element(p, l, p.cta_tables, s, i);
    }
}

// frames of the worlds [env0, env0 + count): out [count][fh][fw][3] uint8
__global__ void __launch_bounds__(kViewThreads) rw_render_kernel(const RwParams p, uint8_t *__restrict__ out, int gs, int64_t env0,
                                                                 int64_t count) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tile = (int)(env0 / RW_TILE) + blockIdx.x;
    rw_view_stage(p, tile, sm, gs, false, false);
    int32_t *tilew = reinterpret_cast<int32_t *>(sm);
    uint8_t *grid = sm + (size_t)p.nw * RW_TILE * 4;
    const int fw = rw_render_width(p), fh = rw_render_height(p), npix = fw * fh;
    for (int w = 0; w < RW_TILE; w++) {
        const int64_t env = (int64_t)tile * RW_TILE + w;
        if (env >= p.E || env >= env0 + count) break;
        if (env < env0) continue;
        RwLane l{tilew + w, grid + (size_t)w * gs, nullptr, 0, nullptr, nullptr};
        uint8_t *o = out + (size_t)(env - env0) * npix * 3;
        for (int k = threadIdx.x; k < npix; k += blockDim.x) {
            const int i = k / fw, j = k - i * fw;
            const uint32_t c = rw_render_pixel(p, l, p.cta_tables, i, j);
            o[3 * k] = (uint8_t)c; o[3 * k + 1] = (uint8_t)(c >> 8); o[3 * k + 2] = (uint8_t)(c >> 16);
        }
    }
}

int view_smem(const RwParams &p, int *gs) {
    *gs = (p.cells + 3) & ~3;
    return p.nw * RW_TILE * 4 + *gs * RW_TILE;
}

template <class K>
cudaError_t ensure_smem(K kernel, int bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

cudaError_t rw_launch_global_image(const RwParams &p, const RwViewSpec &s, float *out, cudaStream_t stream) {
    int gs, bytes = view_smem(p, &gs);
    cudaError_t e = ensure_smem(rw_global_image_kernel, bytes);
    if (e != cudaSuccess) return e;
    rw_global_image_kernel<<<p.tiles, kViewThreads, bytes, stream>>>(p, s, out, gs);
    return cudaGetLastError();
}

cudaError_t rw_launch_render(const RwParams &p, uint8_t *out, int64_t env0, int64_t count, cudaStream_t stream) {
    int gs, bytes = view_smem(p, &gs);
    cudaError_t e = ensure_smem(rw_render_kernel, bytes);
    if (e != cudaSuccess) return e;
    const int blocks = (int)((env0 + count - 1) / RW_TILE - env0 / RW_TILE + 1);
    rw_render_kernel<<<blocks, kViewThreads, bytes, stream>>>(p, out, gs, env0, count);
    return cudaGetLastError();
}
