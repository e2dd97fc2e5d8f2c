// rw_main_kernel.cuh -- the fused sm_100a kernel of the batched RWARE step path (template; instantiated per
// configuration class in rw_inst.cu, several translation units compiled in parallel).
//
// rw_main_kernel<MODE, C>: ONE launch does the whole env.step for every world.  A tile = 32 worlds (rw_common.h) is
// served by a warp pair; a CTA holds as many tiles as fit the SM's shared memory.
//   1. prologue.  Before griddepcontrol.wait (so it overlaps the previous launch's tail under PDL): the CTA-shared static
//      tables are read with plain loads, the helper warp zeroes the tile's occupancy grids, the mbarriers are initialised.
//      After it: each tile's packed state (nw*128 B, contiguous in HBM) arrives through one TMA bulk copy
//      (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) and the primary warp issues its action loads.
//   2. primary: lane e advances world e (rw_logic.cuh): grid build, proposals, collision resolution, apply,
//      deliveries/resample, counters, optional auto-reset -- all on shared memory, no HBM traffic.  (Opt-in: the
//      warp-cooperative form of rw_coop.cuh, a group of lanes per world on both warps.)
//   3. the updated tile goes back with one TMA bulk store (cp.async.bulk.global.shared::cta).
//   4. both warps run the observation gather on alternate chunks: one lane per (world, agent, half row) assembles
//      [chunk][n][L] float32 blocks in a per-warp staging buffer that leave through TMA bulk stores (the block of
//      consecutive worlds is contiguous in the caller's [E][n][L] tensor) -- to the caller's tensor, or, with the
//      fused push all-gather, to this rank's slot of every rank's gathered tensor over NVLink.
//   Rewards/done are tiny and written directly (coalesced).  HBM-bound integer work: no tensor cores.
// MODE: 0 = step, 1 = observation gather only (state read + obs write), 2 = reset (+ obs).
#pragma once
#include <cuda_runtime.h>

#include "rw_coop.cuh"
#include "rw_launch.h"
#include "rw_logic.cuh"

// ---------------------------------------------------------------- PTX helpers (TMA bulk copies + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- the fused kernel
__device__ __forceinline__ unsigned long long rw_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define RW_STAMP(k)                                                                                     \
    do {                                                                                                \
        if (p.dbg_clock && lane == 0) p.dbg_clock[((size_t)blockIdx.x * 32 + warp) * 8 + (k)] = rw_gtime(); \
    } while (0)

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- 4'. bit form of a tile's rows for the host-buffer transport (rw_step_host): both warps of the pair OR their rows' bits
//      into a zeroed bit image of the tile (n * L words + 64 n coordinate bytes, in the staging area), which leaves with
//      coalesced stores.  Kept out of line so that the float path of the kernel does not pay for it in registers.
template <class C>
__device__ __noinline__ void rw_tile_bits(const RwParams &p, uint8_t *wb, int32_t *tilew, uint8_t *grid, const uint8_t *s_highway,
                                          const uint16_t *s_home_id, int tile, int tl, int role, int lane) {
    const int n = p.n, L = p.L, P = p.obs_parts;
    const int64_t tile_env0 = (int64_t)tile * RW_TILE;
    uint32_t *sb = reinterpret_cast<uint32_t *>(wb + p.sm_obs);
    uint8_t *sw = reinterpret_cast<uint8_t *>(sb + n * L);
    for (int i = role * 32 + lane; i < n * L; i += 64) sb[i] = 0;
    pair_barrier(1 + tl);
    const int items = RW_TILE * n * P;
    for (int it = role * 32 + lane; it < items; it += 64) {
        const int part = P == 1 ? 0 : (P == 2 ? (it & 1) : (it & 3));
        const int ea = P == 1 ? it : (P == 2 ? (it >> 1) : (it >> 2));
        const int e = (int)rw_div((uint32_t)ea, p.divN);
        RwLane le;
        le.tile = tilew + e;
        le.g = grid + e * p.grid_env_stride;
        le.sc = nullptr;
        le.sc_step = 0;
        le.highway = s_highway;
        le.home_id = s_home_id;
        if (tile_env0 + e < p.E) rw_obs_agent_bits<C>(p, le, ea - e * n, (uint32_t)ea, sb, sw, part, P);
        else if (part == 0) { sw[2 * ea] = 0; sw[2 * ea + 1] = 0; }
    }
    pair_barrier(1 + tl);
    uint32_t *gb = p.obs_bits_out + (size_t)tile * n * L;
    for (int i = role * 32 + lane; i < n * L; i += 64) gb[i] = sb[i];
    uint32_t *gw = reinterpret_cast<uint32_t *>(p.obs_wide_out + (size_t)tile * RW_TILE * n * 2);
    const uint32_t *sw32 = reinterpret_cast<const uint32_t *>(sw);
    for (int i = role * 32 + lane; i < RW_TILE * n / 2; i += 64) gw[i] = sw32[i];
}

// Two warps serve one tile.  The PRIMARY warp (even) runs the per-world logic, one lane per world.  The HELPER warp
// (odd) zeroes the grids, then sleeps on a named barrier until the post-step state is in shared memory; after that
// both warps run the observation gather on alternate chunks, each with its own staging buffer and its own TMA bulk
// stores.  p.warps_per_cta counts TILES per CTA (blockDim = 64 * tiles).  MAXT bounds the CTA size: 896 threads (14 tiles) leave
// 72 registers per thread; configurations whose tiles are large enough that at most 7 fit an SM use the 448-thread build of the
// step kernel, which may keep up to 144 registers and does not spill.
// BITS: the launch writes the bit form of the rows (p.obs_bits_out / p.obs_wide_out) instead of float rows -- a separate
// instantiation of the step kernel, so that the float kernels carry none of it.
template <int MODE, class C, int MAXT = 896, bool BITS = false>
__global__ void __launch_bounds__(MAXT, 1) rw_main_kernel(const __grid_constant__ RwParams p, const int32_t *__restrict__ actions,
                                                       float *__restrict__ obs, float *__restrict__ reward,
                                                       uint8_t *__restrict__ done, const uint8_t *__restrict__ mask) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tl = warp >> 1, role = warp & 1;  // tile slot in this CTA; 0 = primary, 1 = helper
    const int tile = blockIdx.x * p.warps_per_cta + tl;
    const bool active = tile < p.tiles;
    uint8_t *wb = smem + p.sm_cta_bytes + (size_t)tl * p.sm_warp_bytes;
    int32_t *tilew = reinterpret_cast<int32_t *>(wb + p.sm_tile);
    uint8_t *grid = wb + p.sm_grid;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wb + p.sm_bar);
    const uint32_t tile_bytes = (uint32_t)p.nw * RW_TILE * 4u;
    int32_t *gtile = p.state + (size_t)(active ? tile : 0) * p.nw * RW_TILE;
    const int64_t env = (int64_t)tile * RW_TILE + lane;
    const bool valid = active && env < p.E;
    uint8_t *s_highway = smem;
    uint16_t *s_home_id = reinterpret_cast<uint16_t *>(smem + ((p.cells + 1) & ~1));
    RwLane l;
    l.tile = tilew + lane;
    l.g = grid + lane * p.grid_env_stride;
    l.sc = reinterpret_cast<uint32_t *>(wb + p.sm_scratch) + lane;
    l.sc_step = RW_TILE;
    l.highway = s_highway;
    l.home_id = s_home_id;

    // the fully static specialisations carry only the one-lane-per-world step (select_kernel routes p.coop elsewhere)
    const bool coop = !C::full_static && (MODE == RW_MODE_STEP) && p.coop;
    // ---- 1. prologue, part one: everything that does NOT depend on what earlier work on the stream wrote.  Under
    //         programmatic dependent launch this grid's CTAs become resident while the previous grid is still draining, and this
    //         part runs in that shadow: the CTA-shared static tables (highway bytes + home ids, a few hundred bytes: every thread
    //         copies 16 bytes with plain loads -- a TMA bulk copy bought nothing for so few bytes, and compute-sanitizer's
    //         racecheck cannot follow "bulk copy issued by thread 0 -> mbarrier -> reads by other warps"), the zeroed occupancy
    //         grids (helper warp) and the tile's mbarrier (primary lane 0).
    {
        const int4 *g4 = reinterpret_cast<const int4 *>(p.cta_tables);
        int4 *s4 = reinterpret_cast<int4 *>(smem);
        for (int i = threadIdx.x; i < (p.sm_cta_tx >> 4); i += blockDim.x) s4[i] = __ldg(g4 + i);
    }
    if (role == 0) {
        if (active && lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
    } else if (active) {
        int4 *g4 = reinterpret_cast<int4 *>(grid);
        const int n16 = (p.grid_env_stride * RW_TILE) >> 4;
        for (int i = lane; i < n16; i += 32) g4[i] = make_int4(0, 0, 0, 0);
    }
    // Programmatic dependent launch: nothing the previous grid wrote (world state, the caller's actions) may be read before
    // this point.  Dependents of this grid may start launching right away (they wait at the same instruction).  Both are
    // no-ops for an ordinary launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    RW_STAMP(0);
    // ---- 1. prologue, part two: the tile's packed state arrives through a TMA bulk copy on the tile's mbarrier (primary lane 0)
    //         while the primary's lanes issue their action loads.
    RwActRegs<C> areg;
    const bool pre = rw_can_prefetch_actions<C>(p);
    // agent-split step (rw_logic.cuh, rw_pair_*): both warps of the pair step the tile, each owning half of the agents
    constexpr bool PAIRC = (MODE == RW_MODE_STEP) && rw_pair_capable<C>();
    constexpr int NHP = RwPairRegs<C>::NH;
    constexpr bool pair = PAIRC;  // (the fully static classes have no message bits: rw_inst.cu)
    int32_t pact[NHP];
    if (role == 0 && active && lane == 0) {
        mbar_expect_tx(bar, tile_bytes);
        bulk_g2s(tilew, gtile, tile_bytes, bar);
    }
    if constexpr (pair) {
        if (valid) {
#pragma unroll
            for (int j = 0; j < NHP; j++) pact[j] = actions[env * p.n + role * NHP + j];  // in flight, consumed in the step
        }
    } else if (role == 0 && MODE == RW_MODE_STEP && !coop && valid) {
        if (pre) rw_prefetch_actions<C>(p, actions + env * p.n * p.A, areg);  // loads in flight, consumed below
        else rw_load_actions<C>(p, l, actions + env * p.n * p.A);
    }
    __syncthreads();  // static tables, mbarrier inits and the zeroed grids are visible CTA-wide
    if (!active) return;
    if constexpr (!pair) {
        if (role == 0 && MODE == RW_MODE_STEP && !coop && valid && pre) rw_commit_actions<C>(p, l, areg);
    }
    RW_STAMP(1);

    // Observation staging of this warp: chunk of CH worlds starting at world c0 of the tile -> its slice of the caller's
    // [E][n][L] tensor.  One lane per (world, agent, half row); the chunk leaves as one TMA bulk store.
    const int CH = p.obs_chunk_envs;
    const int64_t tile_env0 = (int64_t)tile * RW_TILE;
    const bool fast = (tile_env0 + RW_TILE <= p.E) && p.obs_bulk && (RW_TILE % CH == 0);
    const int nbw = p.obs_nbuf >> 1;  // staging buffers of this warp (1, or 2 = double buffering)
    float *buf0 = reinterpret_cast<float *>(wb + p.sm_obs + role * nbw * p.obs_buf_stride);
    int stores_in_flight = 0, chunk_no = 0;
    auto emit_chunk = [&](int c0) {
        const int n = p.n, L = p.L, P = p.obs_parts;
        const int items = CH * n * P;
        float *buf = buf0 + (nbw == 2 ? (chunk_no & 1) * (p.obs_buf_stride >> 2) : 0);
        chunk_no++;
        if (stores_in_flight >= nbw) {  // the bulk store that last read this buffer must have drained it
            if (lane == 0) {
                if (nbw == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
            __syncwarp();
            stores_in_flight = nbw - 1;
        }
        for (int it = lane; it < items; it += 32) {
            const int part = P == 4 ? (it & 3) : P == 2 ? (it & 1) : 0;
            const int ea = P == 4 ? (it >> 2) : P == 2 ? (it >> 1) : it;  // (world, agent) row inside the chunk
            const int el = (int)rw_div((uint32_t)ea, p.divN);
            const int e = c0 + el;
            if (e < RW_TILE && tile_env0 + e < p.E) {
                RwLane le;
                le.tile = tilew + e;
                le.g = grid + e * p.grid_env_stride;
                le.sc = nullptr;
                le.sc_step = 0;
                le.highway = s_highway;
                le.home_id = s_home_id;
                if (P == 4) rw_obs_row<C, 4>(p, le, ea - el * n, buf + ea * L, part);
                else if (P == 2) rw_obs_row<C, 2>(p, le, ea - el * n, buf + ea * L, part);
                else rw_obs_row<C, 1>(p, le, ea - el * n, buf + ea * L, 0);
            }
        }
        // destinations: the caller's tensor, or (fused push all-gather) this rank's slot of every rank's gathered tensor,
        // starting with the local copy and then spreading over the NVLink peers
        const int ndst = p.n_peers > 0 ? p.n_peers : 1;
        auto dst_ptr = [&](int d) -> float * {
            if (p.n_peers > 0)
                return p.peer_obs[(p.my_rank + d) % p.n_peers] + (size_t)(p.gather_env0 + tile_env0 + c0) * n * L;
            return obs + (tile_env0 + c0) * n * L;
        };
        if (fast && p.obs_tma) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                for (int d = 0; d < ndst; d++) bulk_s2g(dst_ptr(d), buf, (uint32_t)p.obs_chunk_bytes);
                bulk_commit();
            }
            stores_in_flight++;
        } else if (fast) {  // 16-byte coalesced stores: 512 B per warp instruction
            __syncwarp();
            const int4 *src = reinterpret_cast<const int4 *>(buf);
            const int n16 = p.obs_chunk_bytes >> 4;
            for (int d = 0; d < ndst; d++) {
                int4 *dst = reinterpret_cast<int4 *>(dst_ptr(d));
#pragma unroll 4
                for (int i = lane; i < n16; i += 32) __stcs(dst + i, src[i]);
            }
            __syncwarp();
        } else {  // ragged tail / unaligned geometry: plain coalesced stores
            __syncwarp();
            int64_t envs_here = p.E - (tile_env0 + c0);
            if (envs_here > CH) envs_here = CH;
            if (envs_here > RW_TILE - c0) envs_here = RW_TILE - c0;
            const int nf = (int)(envs_here > 0 ? envs_here : 0) * n * L;
            for (int d = 0; d < ndst; d++) {
                float *gdst = dst_ptr(d);
                for (int i = lane; i < nf; i += 32) gdst[i] = buf[i];
            }
            __syncwarp();
        }
    };

    // ---- 2. the step
    const int passes = 1 << p.gshift;
    const bool interleave = coop && p.coop == 2 && passes >= 2 && CH <= RW_TILE / 2;  // measured slower: opt-in (RWARE_B200_COOP=2)
    if (coop) {
        // a group of 2^gshift lanes per world (rw_coop.cuh).  Each warp of the pair owns half of the tile's worlds and
        // emits the observations of a chunk as soon as its worlds are stepped, so that the TMA stores of one pass drain
        // while the next pass computes.
        mbar_wait(bar, 0);
        RW_STAMP(2);
        RwCoopCtx cx;
        cx.tilew = tilew; cx.grid = grid;
        cx.tiebuf = reinterpret_cast<uint32_t *>(wb + p.sm_tie) + 32 * role;
        cx.highway = s_highway; cx.home_id = s_home_id;
        cx.tile_env0 = tile_env0;
        cx.actions = actions + tile_env0 * p.n * p.A;
        cx.reward = reward + tile_env0 * p.n;
        cx.done = done + tile_env0;
        if (interleave) {
            const int half = passes >> 1, epp = RW_TILE >> p.gshift;
            int ready = role * (RW_TILE / 2), emitted = ready;
            for (int pass = role * half; pass < (role + 1) * half; pass++) {
                rw_step_coop_pass<C>(p, cx, pass, p.gshift);
                ready += epp;
                if (obs != nullptr)
                    for (; ready - emitted >= CH; emitted += CH) emit_chunk(emitted);
            }
        } else {
            for (int pass = role; pass < passes; pass += 2) rw_step_coop_pass<C>(p, cx, pass, p.gshift);
        }
        RW_STAMP(3);
        pair_barrier(1 + tl);  // every world of the tile is in its post-step state
        RW_STAMP(4);
    } else if constexpr (pair) {
        {
            // both warps: lane e advances world e, warp `role` owning the agents [role * N/2, (role + 1) * N/2); the pair
            // barriers separate the phases (see rw_logic.cuh).  Worlds flagged RW_OVERLAP_BIT -- and every world when the
            // form is switched off (RWARE_B200_PAIR=0) -- are stepped by the primary alone with the one-lane form, out of line.
            mbar_wait(bar, 0);
            RW_STAMP(2);
            uint32_t *xw = reinterpret_cast<uint32_t *>(wb + p.sm_tie) + lane;  // [0]: helper -> primary, [RW_TILE]: primary -> helper
            const bool go = valid && p.pair && !((uint32_t)l.tile[rw_off<C>(p).w_inact * RW_TILE] & RW_OVERLAP_BIT);
            RwPairRegs<C> r;
            uint32_t commit = 0;
            if (go) rw_pair_propose_a<C>(p, l, role, pact, r);
            pair_barrier(1 + tl);  // shelf bits complete
            if (go && role == 0) rw_grid_requests_agents<C>(p, l);
            pair_barrier(1 + tl);  // grid complete
            if (go) {
                rw_pair_propose_b<C>(p, l, role, r);
                if (role == 1) xw[0] = r.stay;
            }
            pair_barrier(1 + tl);  // every proposal is in the scratch
            if (go && role == 0) {
                commit = rw_resolve_lane<C>(p, l, r.stay | (xw[0] << NHP));
                xw[RW_TILE] = commit;
            }
            pair_barrier(1 + tl);  // the resolver's mask is out
            if (go) {
                if (role == 1) commit = xw[RW_TILE];
                rw_pair_apply_a<C>(p, l, role, commit, r);
            }
            pair_barrier(1 + tl);  // every mover has left its start cell
            if (go) {
                rw_pair_apply_b<C>(p, l, role, r);
                if (role == 1) xw[0] = r.rew;
            }
            pair_barrier(1 + tl);  // moves applied
            if (role == 0) {
                if (valid) {
                    const bool d = go ? rw_pair_finish<C>(p, l, r, xw[0], reward + env * p.n, env)
                                      : rw_step_world_serial<C>(p, l, actions + env * p.n * p.A, reward + env * p.n, env);
                    done[env] = d ? 1 : 0;
                    if (d && p.auto_reset) {
                        rw_reset_lane<C>(p, l, env);
                        uint32_t *gz = reinterpret_cast<uint32_t *>(l.g);
                        for (int i = 0; i < (p.grid_env_stride >> 2); i++) gz[i] = 0;
                        rw_build_grid<C>(p, l);
                    }
                }
                __syncwarp();
            }
            RW_STAMP(3);
        }
    } else if (role == 0) {
        // primary: lane e advances world e
        mbar_wait(bar, 0);
        RW_STAMP(2);
        if (valid) {
            if (MODE == RW_MODE_RESET) {
                if (mask == nullptr || mask[env]) rw_reset_lane<C>(p, l, env);
            }
            rw_build_grid<C>(p, l);
            if (MODE == RW_MODE_STEP) {
                bool d = rw_step_lane<C>(p, l, reward + env * p.n, env);
                done[env] = d ? 1 : 0;
                if (d && p.auto_reset) {
                    rw_reset_lane<C>(p, l, env);
                    uint32_t *gz = reinterpret_cast<uint32_t *>(l.g);
                    for (int i = 0; i < (p.grid_env_stride >> 2); i++) gz[i] = 0;
                    rw_build_grid<C>(p, l);
                }
            }
        }
        __syncwarp();
        RW_STAMP(3);
    }
    // ---- 3. updated tile shared memory -> HBM
    if (role == 0 && MODE != RW_MODE_OBS) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(gtile, tilew, tile_bytes);
            bulk_commit();
        }
    }
    if ((obs != nullptr || BITS) && !interleave) {
        if (!coop) {
            pair_barrier(1 + tl);  // post-step tile + grids are visible to the helper
            RW_STAMP(4);
        }
        if constexpr (BITS) {
            rw_tile_bits<C>(p, wb, tilew, grid, s_highway, s_home_id, tile, tl, role, lane);
        } else {
            // ---- 4. observation gather after the fact: this warp takes every second chunk
            for (int c0 = role * CH; c0 < RW_TILE; c0 += 2 * CH) emit_chunk(c0);
        }
    }
    if (MODE == RW_MODE_STEP && p.n_peers > 1 && role == 1) {
        // rewards / done of this tile -> every peer's gathered tensors (the primary wrote the local copy; the pair
        // barrier above ordered those stores before these loads)
        const int n = p.n;
        const size_t r0 = (size_t)(p.gather_env0 + tile_env0) * n, d0 = (size_t)(p.gather_env0 + tile_env0);
        const int live = (int)((p.E - tile_env0) < RW_TILE ? (p.E - tile_env0) : RW_TILE);
        for (int d = 1; d < p.n_peers; d++) {
            const int dst = (p.my_rank + d) % p.n_peers;
            for (int i = lane; i < live * n; i += 32) p.peer_rew[dst][r0 + i] = p.peer_rew[p.my_rank][r0 + i];
            if (lane < live) p.peer_done[dst][d0 + lane] = p.peer_done[p.my_rank][d0 + lane];
        }
    }
    RW_STAMP(5);
    if (lane == 0) bulk_wait_read<0>();  // shared memory must stay valid until the TMA has read it
    __syncwarp();
    RW_STAMP(6);
}


typedef void (*rw_kernel_fn)(const RwParams, const int32_t *, float *, float *, uint8_t *, const uint8_t *);

template <class C>
static rw_kernel_fn rw_kernel_for_mode(RwMode mode, bool small_cta, bool bits = false) {
    if (mode == RW_MODE_STEP && bits)
        return small_cta ? rw_main_kernel<RW_MODE_STEP, C, 448, true> : rw_main_kernel<RW_MODE_STEP, C, 896, true>;
    switch (mode) {
        case RW_MODE_STEP: return small_cta ? rw_main_kernel<RW_MODE_STEP, C, 448> : rw_main_kernel<RW_MODE_STEP, C>;
        case RW_MODE_OBS: return rw_main_kernel<RW_MODE_OBS, C>;
        default: return rw_main_kernel<RW_MODE_RESET, C>;
    }
}
