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
    // occupancy maps (what rw_view_build does for one world, spread over the block): thread = (world, part); the shelf and
    // request bits are OR-ed into the 32-bit word holding the cell's byte, agents ("last id wins") follow one lane per world
    const int w = threadIdx.x & 31, part = threadIdx.x >> 5, parts = blockDim.x >> 5;
    const bool live = (int64_t)tile * RW_TILE + w < p.E;
    RwLane l{tilew + w, grid + (size_t)w * gs, nullptr, 0, nullptr, nullptr};
    const RwOff o = rw_off<RwDynamic>(p);
    auto or_cell = [&](int c, uint32_t bit) { atomicOr(reinterpret_cast<uint32_t *>(l.g + (c & ~3)), bit << (8 * (c & 3))); };
    if (live) {
        for (int k = part; k < p.S; k += parts) or_cell(rw_shelf_cell(o, l, k), RW_G_SHELF);
        for (int q = part; q < p.R; q += parts) {
            int id = rw_queue(o, l, q);
            if (id > 0) or_cell(rw_shelf_cell(o, l, id - 1), RW_G_REQ);
        }
    }
    __syncthreads();
    int bad = 0;
    if (live && part == 0)
        for (int i = 0; i < p.n; i++) {
            uint32_t aw = rw_agent(l, i);
            int x = rw_ax(aw), y = rw_ay(aw), c = y * p.W + x;
            l.g[c] = (uint8_t)((l.g[c] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i + 1));
            if ((x >= p.H || y >= p.W) && (dir_layer || (load_layer && rw_acarry(aw)))) bad = 1;
        }
    __syncthreads();
    return bad;
}

// exact a / d for 0 <= a < 2^24 with a float reciprocal and one correction step (d >= 1)
__device__ __forceinline__ int fdiv(int a, int d, float inv) {
    int q = (int)((float)a * inv);
    int r = a - q * d;
    if (r < 0) { q--; r += d; }
    if (r >= d) q++;
    return q;
}

// Work item = (group of worlds, element position): the position is classified once, then shaded for every world of the group.
// Consecutive threads own consecutive positions of the same world, so each store instruction of a warp is contiguous.
__global__ void __launch_bounds__(kViewThreads) rw_global_image_kernel(const RwParams p, const RwViewSpec s, float *__restrict__ out, int gs) {
    extern __shared__ __align__(16) uint8_t sm[];
    bool dir_layer = false, load_layer = false;
    for (int c = 0; c < s.n_layers; c++) { dir_layer |= s.layers[c] == 3; load_layer |= s.layers[c] == 4; }
    const int tile = blockIdx.x;
    if (rw_view_stage(p, tile, sm, gs, dir_layer, load_layer)) atomicAdd(p.err_count, 1ULL);
    int32_t *tilew = reinterpret_cast<int32_t *>(sm);
    uint8_t *grid = sm + (size_t)p.nw * RW_TILE * 4;
    const int plane = s.Hp * s.Wp, per = s.C * plane;
    const float inv_plane = 1.0f / (float)plane, inv_w = 1.0f / (float)s.Wp;
    const int64_t env_first = (int64_t)tile * RW_TILE;
    const int nworlds = (int)min((int64_t)RW_TILE, p.E - env_first);
    float *o = out + (size_t)env_first * per;
    const bool vec = (per & 3) == 0 && per < (1 << 24);  // out is 16-byte aligned (caller contract), so is every world's block
    const int npos = vec ? per / 4 : per;                // positions: groups of 4 consecutive elements, or single elements
    int groups = (int)blockDim.x / npos;                 // worlds are dealt to `groups` thread groups when positions are few
    groups = groups < 1 ? 1 : (groups > RW_TILE ? RW_TILE : groups);
    for (int item = threadIdx.x; item < npos * groups; item += blockDim.x) {
        const int g = item / npos, pos = item - g * npos;
        const int e = vec ? pos * 4 : pos;
        int cp = per < (1 << 24) ? fdiv(e, plane, inv_plane) : e / plane;
        const int rem = e - cp * plane;
        int yp = per < (1 << 24) ? fdiv(rem, s.Wp, inv_w) : rem / s.Wp, xp = rem - yp * s.Wp;
        if (vec) {
            uint32_t code[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                code[k] = rw_gimg_classify(p, p.cta_tables, s, cp, yp, xp);
                if (++xp == s.Wp) { xp = 0; if (++yp == s.Hp) { yp = 0; cp++; } }
            }
            for (int w = g; w < nworlds; w += groups) {
                RwLane l{tilew + w, grid + (size_t)w * gs, nullptr, 0, nullptr, nullptr};
                reinterpret_cast<float4 *>(o + (size_t)w * per)[pos] =
                    make_float4(rw_gimg_shade(l, code[0]), rw_gimg_shade(l, code[1]), rw_gimg_shade(l, code[2]), rw_gimg_shade(l, code[3]));
            }
        } else {
            const uint32_t code = rw_gimg_classify(p, p.cta_tables, s, cp, yp, xp);
            for (int w = g; w < nworlds; w += groups) {
                RwLane l{tilew + w, grid + (size_t)w * gs, nullptr, 0, nullptr, nullptr};
                o[(size_t)w * per + pos] = rw_gimg_shade(l, code);
            }
        }
    }
}

// frames of the worlds [env0, env0 + count): out [count][fh][fw][3] uint8.  blockIdx.x = tile, blockIdx.y = slice of the frame's
// pixels.  A thread classifies its pixel once and shades it for every requested world of the tile (3 byte stores per world;
// consecutive lanes write consecutive pixels, i.e. 96 contiguous bytes per warp and store).
__global__ void __launch_bounds__(kViewThreads) rw_render_kernel(const RwParams p, uint8_t *__restrict__ out, int gs, int64_t env0,
                                                                 int64_t count) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int tile = (int)(env0 / RW_TILE) + blockIdx.x;
    rw_view_stage(p, tile, sm, gs, false, false);
    int32_t *tilew = reinterpret_cast<int32_t *>(sm);
    uint8_t *grid = sm + (size_t)p.nw * RW_TILE * 4;
    const int fw = rw_render_width(p), fh = rw_render_height(p), npix = fw * fh;
    const float inv_fw = 1.0f / (float)fw;
    const int64_t first = max(env0, (int64_t)tile * RW_TILE);
    const int64_t last = min(min(env0 + count, (int64_t)(tile + 1) * RW_TILE), p.E);  // exclusive
    if (last <= first) return;
    const int w0 = (int)(first - (int64_t)tile * RW_TILE), nworlds = (int)(last - first);
    const size_t frame_bytes = (size_t)npix * 3;
    uint8_t *o = out + (size_t)(first - env0) * frame_bytes;
    const int per_slice = (npix + gridDim.y - 1) / gridDim.y;
    const int k0 = blockIdx.y * per_slice, k1 = min(npix, k0 + per_slice);
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        const int i = npix < (1 << 24) ? fdiv(k, fw, inv_fw) : k / fw, j = k - i * fw;
        const uint32_t code = rw_render_classify(p, p.cta_tables, i, j);
        for (int w = 0; w < nworlds; w++) {
            RwLane l{tilew + w0 + w, grid + (size_t)(w0 + w) * gs, nullptr, 0, nullptr, nullptr};
            const uint32_t c = rw_render_shade(l, code);
            uint8_t *px = o + (size_t)w * frame_bytes + (size_t)k * 3;
            px[0] = (uint8_t)c; px[1] = (uint8_t)(c >> 8); px[2] = (uint8_t)(c >> 16);
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
    const int tiles = (int)((env0 + count - 1) / RW_TILE - env0 / RW_TILE + 1);
    int slices = (4 * 148 + tiles - 1) / tiles;  // enough blocks to fill the GPU even for a single tile of frames
    if (slices > 64) slices = 64;
    rw_render_kernel<<<dim3(tiles, slices), kViewThreads, bytes, stream>>>(p, out, gs, env0, count);
    return cudaGetLastError();
}
