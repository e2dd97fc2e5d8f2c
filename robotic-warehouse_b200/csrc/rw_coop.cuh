// rw_coop.cuh -- warp-cooperative form of the hot part of the step: a GROUP of G = 2^k >= n_agents lanes advances one
// world (one lane per agent), so a warp advances 32/G worlds per pass and the per-world work is spread over the
// agents instead of being a serial loop in one lane.  Proposals, the collision fixed point (successor chains, chain
// heights, per-cell winners, commit propagation) and the apply phase run on registers with __shfl_sync /
// __ballot_sync / __match_any_sync inside the group; the occupancy grid and the packed tile stay in shared memory.
// Rare or inherently serial pieces (tie order emulation, request-queue resample, reset, worlds flagged with
// overlapping shelves) are delegated to the leader lane of the group and reuse rw_logic.cuh.
//
// The code is written in an explicit-SPMD style so that the very same source runs on the host (tests/hostsim):
//   RW_EACH_LANE { ... LV(x) ... }   per-lane code; on the device the block runs once for this thread's lane, on the
//                                    host it is a loop over the 32 lanes of the emulated warp
//   RwV<T>                           a per-lane variable: a scalar on the device, T[32] on the host
//   rw_shfl / rw_ballot / rw_match_any / rw_any / rw_syncwarp   the warp collectives, called between lane blocks
// Semantics: rware/warehouse.py:804-946 (cited as ref:LINE), same as rw_step_lane.
#pragma once
#include "rw_logic.cuh"

#if defined(__CUDA_ARCH__)
template <class T>
using RwV = T;
#define RW_EACH_LANE for ([[maybe_unused]] int lane_ = (int)(threadIdx.x & 31u), once_ = 1; once_; once_ = 0)
#define LV(x) (x)
__device__ __forceinline__ void rw_shfl(uint32_t &dst, const uint32_t &src, const int &srclane) {
    dst = __shfl_sync(0xffffffffu, src, srclane);
}
__device__ __forceinline__ uint32_t rw_ballot(const uint32_t &pred) { return __ballot_sync(0xffffffffu, pred != 0); }
__device__ __forceinline__ void rw_match_any(uint32_t &dst, const uint32_t &key) { dst = __match_any_sync(0xffffffffu, key); }
__device__ __forceinline__ void rw_syncwarp() { __syncwarp(); }
#else
template <class T>
struct RwV {
    T v[32];
};
#define RW_EACH_LANE for (int lane_ = 0; lane_ < 32; lane_++)
#define LV(x) ((x).v[lane_])
inline void rw_shfl(RwV<uint32_t> &dst, const RwV<uint32_t> &src, const RwV<int> &srclane) {
    RwV<uint32_t> tmp = src;
    for (int l = 0; l < 32; l++) dst.v[l] = tmp.v[srclane.v[l] & 31];
}
inline uint32_t rw_ballot(const RwV<uint32_t> &pred) {
    uint32_t m = 0;
    for (int l = 0; l < 32; l++) m |= (pred.v[l] ? 1u : 0u) << l;
    return m;
}
inline void rw_match_any(RwV<uint32_t> &dst, const RwV<uint32_t> &key) {
    for (int l = 0; l < 32; l++) {
        uint32_t m = 0;
        for (int k = 0; k < 32; k++) m |= (key.v[k] == key.v[l] ? 1u : 0u) << k;
        dst.v[l] = m;
    }
}
inline void rw_syncwarp() {}
#endif

#define RW_COOP_NONE 0xffu  // "no agent" marker in successor fields (never equals a lane's own index)

struct RwCoopCtx {
    int32_t *tilew;          // tile words in shared memory (word w of world e at tilew[w*32 + e])
    uint8_t *grid;           // occupancy grids, world e at grid + e*grid_env_stride
    uint32_t *tiebuf;        // 32 words private to this warp: a G-word slice per group (tie path, reset sampler)
    const uint8_t *highway;
    const uint16_t *home_id;
    const int32_t *actions;  // this tile's actions [32][n][A] (global memory)
    float *reward;           // this tile's rewards [32][n] (global memory)
    uint8_t *done;           // this tile's done flags [32]
    int64_t tile_env0;
};

RW_HD RwLane rw_coop_lane(const RwParams &p, const RwCoopCtx &c, int e, int base) {
    RwLane l;
    l.tile = c.tilew + e;
    l.g = c.grid + e * p.grid_env_stride;
    l.sc = c.tiebuf + base;
    l.sc_step = 1;
    l.highway = c.highway;
    l.home_id = c.home_id;
    return l;
}

// One pass: worlds [pass*(32/G), (pass+1)*(32/G)) of the tile.  G = 1 << gshift.
template <class C>
RW_HD void rw_step_coop_pass(const RwParams &p, const RwCoopCtx &c, int pass, int gshift) {
    const RwOff o = rw_off<C>(p);
    const int n = o.n, W = C::W(p), H = C::H(p), S = C::S(p), mb = C::mb(p), A = 1 + mb;
    const int G = 1 << gshift, epp = 32 >> gshift;  // group size, worlds per pass
    RwV<int> e, a, base;          // world (index in the tile), agent, first lane of the group
    RwV<uint32_t> live, on;       // world exists / this lane owns an agent of an existing world
    RwV<uint32_t> ovl;            // world flagged with overlapping shelves (leader handles its grid serially)
    RW_EACH_LANE {
        LV(base) = lane_ & ~(G - 1);
        LV(a) = lane_ & (G - 1);
        LV(e) = pass * epp + (lane_ >> gshift);
        LV(live) = (c.tile_env0 + LV(e) < p.E) ? 1u : 0u;
        LV(on) = (LV(live) && LV(a) < n) ? 1u : 0u;
        LV(ovl) = LV(live) ? (((uint32_t)c.tilew[o.w_inact * RW_TILE + LV(e)] & RW_OVERLAP_BIT) ? 1u : 0u) : 0u;
    }
    // ---- 1. occupancy grid (ref:749-755), shelves spread over the lanes of the group
    RW_EACH_LANE {
        if (LV(live) && !LV(ovl)) {
            uint8_t *g = c.grid + LV(e) * p.grid_env_stride;
            for (int k2 = LV(a); k2 < (S >> 1); k2 += G) {
                uint32_t w = (uint32_t)c.tilew[(o.w_shelf + k2) * RW_TILE + LV(e)];
                g[w & 0xffffu] = (uint8_t)RW_G_SHELF;
                g[w >> 16] = (uint8_t)RW_G_SHELF;
            }
            if ((S & 1) && LV(a) == 0)
                g[(uint32_t)c.tilew[(o.w_shelf + (S >> 1)) * RW_TILE + LV(e)] & 0xffffu] = (uint8_t)RW_G_SHELF;
        }
    }
    rw_syncwarp();
    RW_EACH_LANE {
        if (LV(live) && !LV(ovl)) {
            RwLane l = rw_coop_lane(p, c, LV(e), LV(base));
            for (int q = LV(a); q < o.R; q += G) {
                int id = rw_queue(o, l, q);
                if (id) l.g[rw_shelf_cell(o, l, id - 1)] |= (uint8_t)RW_G_REQ;
            }
        }
    }
    rw_syncwarp();
    RwV<uint32_t> aw, act, s, t, nxt;
    RW_EACH_LANE {
        LV(aw) = 0; LV(act) = RW_ACT_NOOP; LV(s) = 0; LV(t) = 0; LV(nxt) = RW_NXT_SELF;
        if (LV(live) && LV(ovl) && LV(a) == 0) rw_build_grid<C>(p, rw_coop_lane(p, c, LV(e), LV(base)));
        if (LV(on)) {
            LV(aw) = (uint32_t)c.tilew[LV(a) * RW_TILE + LV(e)];
            if (!LV(ovl)) {
                uint8_t *g = c.grid + LV(e) * p.grid_env_stride;
                uint32_t cell = (uint32_t)(rw_ay(LV(aw)) * W + rw_ax(LV(aw)));
                g[cell] = (uint8_t)((g[cell] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(LV(a) + 1));
            }
            // action decode (ref:809-814); invalid values count as NOOP
            const int32_t *ap = c.actions + ((size_t)LV(e) * n + LV(a)) * A;
            int av = ap[0];
            if ((unsigned)av > 4u) {
                av = RW_ACT_NOOP;
#if defined(__CUDA_ARCH__)
                atomicAdd(p.err_count, 1ULL);
#else
                *p.err_count += 1;
#endif
            }
            LV(act) = (uint32_t)av;
            if (mb) {
                uint32_t bits = 0;
                for (int b = 0; b < mb; b++) bits |= (uint32_t)(ap[1 + b] & 1) << b;
                c.tilew[(o.w_msg + LV(a)) * RW_TILE + LV(e)] = (int32_t)bits;
            }
        }
    }
    rw_syncwarp();
    // ---- 2. proposals (ref:825-846)
    RwV<uint32_t> occ_aw;
    RwV<int> src;
    RW_EACH_LANE {
        LV(src) = lane_;
        if (LV(on)) {
            int x = rw_ax(LV(aw)), y = rw_ay(LV(aw)), tx = x, ty = y;
            if (LV(act) == RW_ACT_FORWARD) {  // Agent.req_location ref:102-116
                int d = rw_adir(LV(aw));
                if (d == 0) ty = y > 0 ? y - 1 : 0;
                else if (d == 1) ty = y < H - 1 ? y + 1 : H - 1;
                else if (d == 2) tx = x > 0 ? x - 1 : 0;
                else tx = x < W - 1 ? x + 1 : W - 1;
            }
            LV(s) = (uint32_t)(y * W + x);
            LV(t) = (uint32_t)(ty * W + tx);
            if (LV(t) != LV(s)) {
                uint32_t gt = c.grid[LV(e) * p.grid_env_stride + LV(t)];
                uint32_t occ = gt & RW_G_AGENT;
                LV(nxt) = occ ? occ - 1 : RW_NXT_SINK;
                // stash the shelf bit of the target for the cancel rule below
                LV(nxt) |= (gt & RW_G_SHELF) << 8;
                if (occ) LV(src) = LV(base) + (int)occ - 1;
            }
        }
    }
    rw_shfl(occ_aw, aw, src);
    RwV<uint32_t> commit, mover;
    RW_EACH_LANE {
        LV(commit) = 0; LV(mover) = 0;
        if (LV(on)) {
            uint32_t shelf_at_t = (LV(nxt) >> 8) & RW_G_SHELF;
            LV(nxt) &= 0xffu;
            // ref:829-844 a carrier may not enter a cell holding a shelf unless an agent there carries it
            if (LV(t) != LV(s) && rw_acarry(LV(aw)) && shelf_at_t &&
                !(LV(nxt) < 32u && rw_acarry(LV(occ_aw)))) {
                LV(act) = RW_ACT_NOOP; LV(t) = LV(s); LV(nxt) = RW_NXT_SELF;
            }
            if (LV(nxt) == RW_NXT_SELF) LV(commit) = 1; else LV(mover) = 1;
        }
    }
    // ---- 3. resolve (ref:848-876, SURVEY.md App. B agent-centric form)
    RwV<uint32_t> dag;
    RW_EACH_LANE { LV(dag) = 0; }
    if (rw_ballot(mover)) {
        RwV<uint32_t> cur, pending, nj;
        RW_EACH_LANE { LV(cur) = LV(nxt); LV(pending) = LV(mover); }
        for (int hop = 1; hop <= n; hop++) {
            RW_EACH_LANE { LV(src) = LV(base) + (LV(cur) < 32u ? (int)LV(cur) : LV(a)); }
            rw_shfl(nj, nxt, src);
            RW_EACH_LANE {
                if (LV(pending)) {
                    if (LV(cur) == RW_NXT_SINK) { LV(dag) = 1; LV(pending) = 0; }               // chain ends in a free cell
                    else if (LV(cur) == (uint32_t)LV(a)) { if (hop != 2) LV(commit) = 1; LV(pending) = 0; }  // cycle; swaps fail
                    else if (LV(nj) == RW_NXT_SELF || hop >= n) LV(pending) = 0;                // blocked / feeds a cycle
                    else LV(cur) = LV(nj);
                }
            }
            if (!rw_ballot(pending)) break;
        }
        const uint32_t dagmask = rw_ballot(dag);
        if (dagmask) {
            // predecessor sets: lanes of the group whose (dag) successor is this agent
            RwV<uint32_t> pm, nd, hk, h, ch;
            RW_EACH_LANE { LV(pm) = 0; LV(h) = 0; LV(nd) = LV(dag) ? LV(nxt) : RW_COOP_NONE; }
            for (int k = 0; k < G; k++) {
                RW_EACH_LANE { LV(src) = LV(base) + k; }
                rw_shfl(hk, nd, src);
                RW_EACH_LANE { if (LV(hk) == (uint32_t)LV(a) && LV(on)) LV(pm) |= 1u << k; }
            }
            // heights h[j] = 1 + max h[i] over predecessors (dag.py:1073-1101 `dist`), to the fixed point
            for (int round = 0; round < n; round++) {
                RwV<uint32_t> nh;
                RW_EACH_LANE { LV(nh) = LV(h); }
                for (int k = 0; k < G; k++) {
                    RW_EACH_LANE { LV(src) = LV(base) + k; }
                    rw_shfl(hk, h, src);
                    RW_EACH_LANE { if ((LV(pm) >> k) & 1u) { if (LV(hk) + 1 > LV(nh)) LV(nh) = LV(hk) + 1; } }
                }
                RW_EACH_LANE { LV(ch) = LV(nh) != LV(h); LV(h) = LV(nh); }
                if (!rw_ballot(ch)) break;
            }
            // per contested cell: the contender with the greatest height wins; ties by networkx's node order
            RwV<uint32_t> key, peers, mx, eq, win;
            RW_EACH_LANE { LV(key) = LV(dag) ? ((uint32_t)(lane_ >> gshift) << 16 | LV(t)) : (0xffff0000u | (uint32_t)lane_); }
            rw_match_any(peers, key);
            RW_EACH_LANE { LV(peers) &= dagmask; LV(mx) = LV(h); }
            for (int k = 0; k < G; k++) {
                RW_EACH_LANE { LV(src) = LV(base) + k; }
                rw_shfl(hk, h, src);
                RW_EACH_LANE { if (((LV(peers) >> (LV(base) + k)) & 1u) && LV(hk) > LV(mx)) LV(mx) = LV(hk); }
            }
            RW_EACH_LANE { LV(eq) = (LV(dag) && LV(h) == LV(mx)) ? 1u : 0u; }
            const uint32_t eqmask = rw_ballot(eq);
            RwV<uint32_t> tie;
            RW_EACH_LANE {
                uint32_t m = eqmask & LV(peers);
                LV(tie) = (LV(eq) && (m & (m - 1))) ? 1u : 0u;
                LV(win) = LV(eq);
            }
            if (rw_ballot(tie)) {  // rare: emulate networkx's component-copy node order (rw_tie_first)
                RW_EACH_LANE {
                    if (LV(on)) c.tiebuf[lane_] = LV(s) | (LV(t) << 16);  // rw_st(l, a) of this group's slice
                }
                rw_syncwarp();
                RW_EACH_LANE {
                    if (LV(tie)) {
                        uint32_t ties = ((eqmask & LV(peers)) >> LV(base)) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
                        LV(win) = (rw_tie_first<C>(p, rw_coop_lane(p, c, LV(e), LV(base)), ties) == LV(a)) ? 1u : 0u;
                    }
                }
                rw_syncwarp();
            }
            // commit along the selected path: commit_i = win_i and (target free or its occupant commits)
            RwV<uint32_t> cd, cn;
            RW_EACH_LANE { LV(cd) = (LV(dag) && LV(win) && LV(nxt) == RW_NXT_SINK) ? 1u : 0u; }
            for (int round = 0; round < n; round++) {
                RW_EACH_LANE { LV(src) = LV(base) + (LV(nxt) < 32u ? (int)LV(nxt) : LV(a)); }
                rw_shfl(cn, cd, src);
                RW_EACH_LANE {
                    uint32_t nc = (LV(dag) && LV(win) && (LV(nxt) == RW_NXT_SINK || (LV(nxt) < 32u && LV(cn)))) ? 1u : 0u;
                    LV(ch) = nc != LV(cd);
                    LV(cd) = nc;
                }
                if (!rw_ballot(ch)) break;
            }
            RW_EACH_LANE { LV(commit) |= LV(cd); }
        }
    }
    // ---- 4. apply (ref:878-899); failed agents become NOOP (ref:871-876)
    RwV<uint32_t> moved, lifted, drop_half;
    RW_EACH_LANE {
        LV(moved) = 0; LV(lifted) = 0; LV(drop_half) = 0;
        if (LV(on)) {
            uint8_t *g = c.grid + LV(e) * p.grid_env_stride;
            uint32_t av = LV(act);
            if (av == RW_ACT_FORWARD) {
                if (LV(commit) && LV(t) != LV(s)) {
                    LV(moved) = 1;
                    uint32_t gsb = g[LV(s)];
                    if (rw_acarry(LV(aw))) { LV(lifted) = gsb & (RW_G_SHELF | RW_G_REQ); g[LV(s)] = 0; }
                    else g[LV(s)] = (uint8_t)(gsb & (RW_G_SHELF | RW_G_REQ));
                }
            } else if (av == RW_ACT_LEFT || av == RW_ACT_RIGHT) {  // Agent.req_direction ref:118-125
                uint32_t d = (uint32_t)rw_adir(LV(aw));
                uint32_t nd = ((av == RW_ACT_RIGHT ? 0x4Bu : 0x1Eu) >> (2 * d)) & 3u;
                LV(aw) = (LV(aw) & ~(3u << 16)) | (nd << 16);
            } else if (av == RW_ACT_TOGGLE) {  // ref:889-899
                if (!rw_acarry(LV(aw))) {
                    if (g[LV(s)] & RW_G_SHELF) LV(aw) |= (uint32_t)rw_shelf_at<C>(p, rw_coop_lane(p, c, LV(e), LV(base)), LV(s)) << 19;
                } else if (!c.highway[LV(s)]) {
                    if (rw_adeliv(LV(aw)) && p.reward_type == 2) LV(drop_half) = 1;
                    LV(aw) = LV(aw) & 0x3ffffu & ~(1u << 18);
                }
            }
        }
    }
    rw_syncwarp();
    RW_EACH_LANE {
        if (LV(on)) {
            if (LV(moved)) {
                uint8_t *g = c.grid + LV(e) * p.grid_env_stride;
                uint32_t ty = C::div_w(p, LV(t)), tx = LV(t) - ty * (uint32_t)W;
                LV(aw) = (LV(aw) & ~0xffffu) | tx | (ty << 8);
                int carry = rw_acarry(LV(aw));
                if (carry) {
                    g[LV(t)] = (uint8_t)((uint32_t)(LV(a) + 1) | LV(lifted));
                    RwLane l = rw_coop_lane(p, c, LV(e), LV(base));
                    rw_set_shelf_cell(o, l, carry - 1, (int)LV(t));
                } else {
                    g[LV(t)] = (uint8_t)((g[LV(t)] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(LV(a) + 1));
                }
            }
            c.tilew[LV(a) * RW_TILE + LV(e)] = (int32_t)LV(aw);
        }
    }
    rw_syncwarp();
    // ---- 5. leader of each group: deliveries / resample / counters (ref:903-941), auto-reset
    const uint32_t movedmask = rw_ballot(moved);
    RwV<uint32_t> r0, r1, r2, r3, r4;  // reward accumulators of the world, broadcast from the leader
    RW_EACH_LANE {
        LV(r0) = LV(r1) = LV(r2) = LV(r3) = LV(r4) = 0;
        if (LV(live) && LV(a) == 0) {
            RwLane l = rw_coop_lane(p, c, LV(e), LV(base));
            uint32_t gm = (movedmask >> LV(base)) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            if (LV(ovl) && gm) rw_rebuild_grid_slow<C>(p, l);  // degenerate injected state: redo ref:901 literally
            uint64_t lo = 0, hi = 0;
            int glob = 0;
            const int64_t env = c.tile_env0 + LV(e);
            bool d = rw_finish_lane<C>(p, l, env, lo, hi, glob);
            c.done[LV(e)] = d ? 1 : 0;
            if (d && p.auto_reset) {
                rw_reset_lane<C>(p, l, env);
                uint32_t *gz = reinterpret_cast<uint32_t *>(l.g);
                for (int i = 0; i < (p.grid_env_stride >> 2); i++) gz[i] = 0;
                rw_build_grid<C>(p, l);
            }
            LV(r0) = (uint32_t)lo; LV(r1) = (uint32_t)(lo >> 32); LV(r2) = (uint32_t)hi; LV(r3) = (uint32_t)(hi >> 32);
            LV(r4) = (uint32_t)glob;
        }
    }
    RW_EACH_LANE { LV(src) = LV(base); }
    if (G > 1) {
        rw_shfl(r0, r0, src); rw_shfl(r1, r1, src); rw_shfl(r2, r2, src); rw_shfl(r3, r3, src); rw_shfl(r4, r4, src);
    }
    RW_EACH_LANE {
        if (LV(on)) {
            int i = LV(a);
            uint32_t word = (i < 8) ? LV(r0) : (i < 16) ? LV(r1) : (i < 24) ? LV(r2) : LV(r3);
            uint32_t half = ((word >> (4 * (i & 7))) & 15u) + 2u * LV(r4) + LV(drop_half);
            c.reward[(size_t)LV(e) * n + i] = 0.5f * (float)half;
        }
    }
    rw_syncwarp();
}
