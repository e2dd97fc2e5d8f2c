// rw_common.h -- shared (host + device) definitions of the packed world-state layout and kernel parameters.
//
// HBM layout ("tile" = 32 consecutive worlds = one warp's worth):
//   state[tile][word][lane]  int32, lane fastest  -> one tile is ONE contiguous block of nw*128 bytes, moved
//   with a single TMA bulk copy in each direction, and lane e reads word w at bank e (conflict free).
// Words of one world (reference fields in brackets, rware/warehouse.py):
//   w_agent + i      : agent i+1  = x | y<<8 | dir<<16 | has_delivered<<18 | carried_shelf_id<<19   [82-93]
//   w_msg   + i      : agent i+1 message bits (only when msg_bits > 0)                               [89]
//   w_shelf + k/2    : two uint16 cell indices (y*W+x) of shelves 2*(k/2)+1, +2                      [128-133]
//   w_queue + q/2    : two uint16 shelf ids of request_queue[q]                                      [262-263]
//   w_steps, w_inact : _cur_steps, _cur_inactive_steps                                               [249-250]
// The numpy-compatible PCG64 generator state (10 words per world) is cold data and lives in a separate
// struct-of-arrays rng[word][world] that is only touched by worlds that deliver or reset.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RW_HD __host__ __device__ __forceinline__
#define RW_HD_NOINLINE __host__ __device__ __noinline__
#else
#define RW_HD inline
#define RW_HD_NOINLINE
#endif

#define RW_TILE 32
#define RW_MAX_AGENTS 32
#define RW_MAX_GOALS 16
#define RW_MAX_MSG_BITS 16
#define RW_RNG_WORDS 10
#define RW_MAX_PEERS 8
#define RW_MAX_SHELVES 8191
// occupancy grid byte (shared memory, rebuilt every step): bits 0-5 id of the agent standing on the cell (0 = none),
// bit 6 a shelf is on the cell (standing or carried), bit 7 that shelf is in the request queue
#define RW_G_AGENT 0x3fu
#define RW_G_SHELF 0x40u
#define RW_G_REQ 0x80u
// bit 31 of the w_inact word: two shelves share a cell somewhere in this world (only reachable through state
// injection); such worlds take the literal "_recalc_grid" slow path
#define RW_OVERLAP_BIT 0x80000000u

#define RW_ACT_NOOP 0
#define RW_ACT_FORWARD 1
#define RW_ACT_LEFT 2
#define RW_ACT_RIGHT 3
#define RW_ACT_TOGGLE 4

#define RW_NXT_SINK 62u
#define RW_NXT_SELF 63u

struct RwDiv {  // exact x / d for 0 <= x < 2^16, 1 <= d < 2^16 via one mulhi
    uint32_t d, magic;
};

struct RwParams {
    // static configuration
    int32_t H, W, cells, n, msg_bits, A, r, win, K, cellF, L, R, S;
    int32_t reward_type, max_steps, max_inact, norm_coords, n_goals, auto_reset;
    int32_t goal_cell[RW_MAX_GOALS];
    int32_t obs_kind;             // 0 flattened vector, 2 image, 3 image + features (ObservationType values)
    int32_t img_directional, n_layers, img_layers[8];
    uint32_t img_mask[8];         // per image layer: test mask over (grid byte | table byte << 8 | outside-the-map << 16), bit 31 = invert;
                                  // 0 for the transposed agent layers (AGENT_DIRECTION / AGENT_LOAD), which a second pass fills in
    int32_t img_transposed;       // bit 0: AGENT_DIRECTION configured, bit 1: AGENT_LOAD configured
    // tile layout
    int32_t nw, w_agent, w_msg, w_shelf, w_queue, w_steps, w_inact;
    // static device tables
    const uint8_t *highway;     // [cells]
    const uint16_t *home_cell;  // [S]
    const uint16_t *home_id;    // [cells]: inverse of home_cell (0 = highway)
    const uint8_t *cta_tables;  // device only: image of the CTA-shared table region (highway bytes, then home_id), sm_cta_bytes long
    const float *normx, *normy; // [W], [H]: float32(x / (W-1)), float32(y / (H-1))
    // world state
    int32_t *state;  // [tiles][nw][32]
    uint32_t *rng;   // [RW_RNG_WORDS][E_pad]
    unsigned long long *err_count;
    unsigned long long *dbg_clock;  // optional [CTA][warp][8] phase timestamps (globaltimer ns), NULL = off
    int64_t E, E_pad;
    int32_t tiles;
    // per-warp shared memory carve-up (byte offsets from the warp's base)
    int32_t sm_tile, sm_grid, sm_scratch, sm_obs, sm_bar, sm_tie, sm_warp_bytes;
    int32_t sm_cta_tx;     // bytes of the table image moved by the TMA (multiple of 16, < sm_cta_bytes - 8: the CTA mbarrier sits at the end)
    int32_t sm_cta_bytes;  // CTA-shared prologue region: highway[cells] u8 then home_id[cells] u16 (device only)
    // fused push all-gather (multi-GPU, optional): every observation chunk is TMA-stored into the gathered buffer of
    // every rank (peer-mapped over NVLink), rewards/done likewise with plain stores.  n_peers == 0: off.
    int32_t n_peers, my_rank;
    int64_t gather_env0;           // first world of this rank in the gathered tensors
    float *peer_obs[RW_MAX_PEERS];
    float *peer_rew[RW_MAX_PEERS];
    uint8_t *peer_done[RW_MAX_PEERS];
    int32_t coop, gshift;  // cooperative step: a group of 1 << gshift lanes per world (rw_coop.cuh); 0 = one lane per world
    int32_t pair;          // agent-split step (rw_logic.cuh, rw_pair_*): both warps of a pair step the tile, half of the agents each
    int32_t obs_parts;     // lanes per agent row in the observation gather (1 or 2)
    int32_t obs_tma;       // 1: observation chunks leave through cp.async.bulk, 0: 16-byte st.global from all lanes
    int32_t grid_env_stride;  // bytes per world of the occupancy grid: one byte per cell (see RW_G_*)
    int32_t obs_chunk_envs, obs_chunk_bytes, obs_nbuf, obs_env_bytes;
    int32_t obs_buf_stride;  // bytes between the primary's and the helper's staging buffer
    int32_t warps_per_cta;  // TILES per CTA (each tile is served by a warp pair)
    RwDiv divW, divN;
    int32_t obs_bulk;  // 1: observation chunks leave through cp.async.bulk (16-byte aligned), 0: generic stores
    // bit form of the FLATTENED observations for the host-buffer transport (rw_step_host): when obs_bits_out != NULL the
    // launch writes, instead of float rows, bit (row * L + k) = "value k of the row is non-zero" and the two coordinate bytes
    // of every row (rw_obs_agent_bits); tile t owns the words [t * n * L, (t + 1) * n * L) and the bytes [t * 64 n, ...).
    uint32_t *obs_bits_out;
    uint8_t *obs_wide_out;
};

RW_HD uint32_t rw_div(uint32_t x, const RwDiv &dv) {
    if (dv.magic == 0) return x;  // d == 1 (ceil(2^32 / 1) does not fit 32 bits)
#if defined(__CUDA_ARCH__)
    return __umulhi(x, dv.magic);
#else
    return (uint32_t)(((uint64_t)x * dv.magic) >> 32);
#endif
}

// ---- agent word helpers
RW_HD int rw_ax(uint32_t aw) { return (int)(aw & 0xffu); }
RW_HD int rw_ay(uint32_t aw) { return (int)((aw >> 8) & 0xffu); }
RW_HD int rw_adir(uint32_t aw) { return (int)((aw >> 16) & 3u); }
RW_HD int rw_adeliv(uint32_t aw) { return (int)((aw >> 18) & 1u); }
RW_HD int rw_acarry(uint32_t aw) { return (int)(aw >> 19); }
RW_HD uint32_t rw_apack(int x, int y, int dir, int deliv, int carry) {
    return (uint32_t)x | ((uint32_t)y << 8) | ((uint32_t)dir << 16) | ((uint32_t)deliv << 18) | ((uint32_t)carry << 19);
}

// per-world view used by the lane logic; `tile` already points at this world's lane
struct RwLane {
    int32_t *tile;   // word w at tile[w * RW_TILE]
    uint8_t *g;      // occupancy grid of this world, one byte per cell (RW_G_*)
    uint32_t *sc;    // resolver scratch.  One-lane-per-world step: [2n][32] words interleaved over the tile's lanes
                     // (sc_step = 32: start|target of agent i at sc[i*64], meta at sc[i*64+32]).  Cooperative step: a
                     // private n-word buffer per group (sc_step = 1: start|target of agent i at sc[i], no meta words).
    int sc_step;
    const uint8_t *highway;   // [cells] 1 = highway (CTA-shared copy in shared memory on the device)
    const uint16_t *home_id;  // [cells] id of the shelf whose home is this cell, 0 for highways (CTA-shared copy)
};

// Compile-time specialisation of the layout-dependent constants.  RwDynamic reads them from RwParams; RwStatic<...>
// bakes in the registered warehouse sizes so that the cell loops unroll and the index arithmetic folds.
struct RwDynamic {
    static constexpr bool is_static = false;
    static constexpr bool full_static = false;
    static constexpr int static_n = 0;
    static constexpr int static_mb = 0;
    RW_HD static int n(const RwParams &p) { return p.n; }
    RW_HD static int R(const RwParams &p) { return p.R; }
    RW_HD static int H(const RwParams &p) { return p.H; }
    RW_HD static int W(const RwParams &p) { return p.W; }
    RW_HD static int S(const RwParams &p) { return p.S; }
    RW_HD static int r(const RwParams &p) { return p.r; }
    RW_HD static int mb(const RwParams &p) { return p.msg_bits; }
    RW_HD static uint32_t div_w(const RwParams &p, uint32_t c) { return rw_div(c, p.divW); }
};
// N_ / RQ_ = 0 keep the number of agents / the request-queue size dynamic
template <int H_, int W_, int S_, int R_, int MB_, int N_ = 0, int RQ_ = -1>
struct RwStatic {
    static constexpr bool is_static = true;
    static constexpr bool full_static = (N_ > 0 && RQ_ >= 0);
    static constexpr int static_n = N_;
    static constexpr int static_mb = MB_;
    RW_HD static int n(const RwParams &p) { return N_ > 0 ? N_ : p.n; }
    RW_HD static int R(const RwParams &p) { return RQ_ >= 0 ? RQ_ : p.R; }
    RW_HD static int H(const RwParams &) { return H_; }
    RW_HD static int W(const RwParams &) { return W_; }
    RW_HD static int S(const RwParams &) { return S_; }
    RW_HD static int r(const RwParams &) { return R_; }
    RW_HD static int mb(const RwParams &) { return MB_; }
    RW_HD static uint32_t div_w(const RwParams &, uint32_t c) { return c / (uint32_t)W_; }
};

// tile word offsets of one world (compile-time constants for fully static configurations)
struct RwOff {
    int n, R, w_msg, w_shelf, w_queue, w_steps, w_inact;
};
template <class C>
RW_HD RwOff rw_off(const RwParams &p) {
    RwOff o;
    o.n = C::n(p);
    o.R = C::R(p);
    if (C::full_static) {
        const int mb = C::mb(p), S = C::S(p);
        o.w_msg = o.n;
        o.w_shelf = o.n + (mb ? o.n : 0);
        o.w_queue = o.w_shelf + (S + 1) / 2;
        o.w_steps = o.w_queue + (o.R + 1) / 2;
        o.w_inact = o.w_steps + 1;
    } else {
        o.w_msg = p.w_msg; o.w_shelf = p.w_shelf; o.w_queue = p.w_queue; o.w_steps = p.w_steps; o.w_inact = p.w_inact;
    }
    return o;
}

RW_HD uint32_t rw_agent(const RwLane &l, int i) { return (uint32_t)l.tile[i * RW_TILE]; }  // w_agent == 0
RW_HD void rw_set_agent(const RwLane &l, int i, uint32_t aw) { l.tile[i * RW_TILE] = (int32_t)aw; }
RW_HD uint16_t *rw_half(int32_t *tile, int word0, int k) {
    return reinterpret_cast<uint16_t *>(&tile[(word0 + (k >> 1)) * RW_TILE]) + (k & 1);
}
RW_HD int rw_shelf_cell(const RwOff &o, const RwLane &l, int k) { return *rw_half(l.tile, o.w_shelf, k); }
RW_HD void rw_set_shelf_cell(const RwOff &o, const RwLane &l, int k, int cell) { *rw_half(l.tile, o.w_shelf, k) = (uint16_t)cell; }
RW_HD int rw_queue(const RwOff &o, const RwLane &l, int q) { return *rw_half(l.tile, o.w_queue, q); }
RW_HD void rw_set_queue(const RwOff &o, const RwLane &l, int q, int id) { *rw_half(l.tile, o.w_queue, q) = (uint16_t)id; }
