// rw_host_layout.h -- host-only helpers shared by the C-ABI implementation (rw_capi.cu) and the test-only host
// twin (tests/hostsim): the packed tile layout / shared-memory carve-up for a configuration, and the
// numpy-compatible seeding SeedSequence(seed) -> PCG64 (numpy/random/bit_generator.pyx, pcg64.h).
#pragma once
#include <cstdlib>
#include <cstring>

#include "../../include/rware_b200.h"
#include "rw_common.h"

static inline RwDiv rw_make_div(uint32_t d) {
    RwDiv r;
    r.d = d;
    r.magic = (uint32_t)((0x100000000ULL + d - 1) / d);  // ceil(2^32 / d): exact for x*d < 2^32 with x < 2^16
    return r;
}

static inline int rw_round_up(int x, int a) { return (x + a - 1) / a * a; }

// ---- numpy SeedSequence(seed).generate_state(8 words) -> PCG64 initial state, 10-word export
static inline void rw_seed_words(uint64_t seed, uint32_t out[RW_RNG_WORDS]) {
    const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
    const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
    uint32_t ent[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int n_ent = ent[1] ? 2 : 1;
    uint32_t pool[4], hc = INIT_A;
    auto hashmix = [&](uint32_t v) {
        v ^= hc;
        hc *= MULT_A;
        v *= hc;
        v ^= v >> 16;
        return v;
    };
    auto mix = [&](uint32_t x, uint32_t y) {
        uint32_t r = MIX_L * x - MIX_R * y;
        return r ^ (r >> 16);
    };
    for (int i = 0; i < 4; i++) pool[i] = hashmix(i < n_ent ? ent[i] : 0u);
    for (int s = 0; s < 4; s++)
        for (int d = 0; d < 4; d++)
            if (s != d) pool[d] = mix(pool[d], hashmix(pool[s]));
    uint32_t w[8], hb = INIT_B;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pool[i & 3];
        v ^= hb;
        hb *= MULT_B;
        v *= hb;
        v ^= v >> 16;
        w[i] = v;
    }
    typedef unsigned __int128 u128;
    uint64_t v64[4];
    for (int i = 0; i < 4; i++) v64[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    const u128 MULT = (((u128)2549297995355413924ULL) << 64) | (u128)4865540595714422341ULL;
    u128 initstate = (((u128)v64[0]) << 64) | v64[1], initseq = (((u128)v64[2]) << 64) | v64[3];
    u128 inc = (initseq << 1) | 1u;
    u128 st = 0;
    st = st * MULT + inc;
    st += initstate;
    st = st * MULT + inc;
    for (int i = 0; i < 4; i++) out[i] = (uint32_t)(st >> (32 * i));
    for (int i = 0; i < 4; i++) out[4 + i] = (uint32_t)(inc >> (32 * i));
    out[8] = 0;
    out[9] = 0;
}


// Fills everything in RwParams that does not depend on device pointers.  S = number of non-highway cells.
static inline void rw_fill_params(RwParams &p, const rw_config &c, int64_t num_envs, int S) {
    const rw_config *cfg = &c;
    const int H = cfg->grid_h, W = cfg->grid_w, cells = H * W, n = cfg->n_agents;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.cells = cells; p.n = n; p.msg_bits = cfg->msg_bits; p.A = 1 + cfg->msg_bits;
    p.r = cfg->sensor_range; p.win = 2 * p.r + 1; p.K = p.win * p.win; p.cellF = 7 + p.msg_bits;
    p.L = 8 + p.K * p.cellF;  // warehouse.py:432-443
    p.obs_kind = 0;
    if (cfg->observation_type == RW_OBS_IMAGE || cfg->observation_type == RW_OBS_IMAGE_DICT) {
        p.obs_kind = cfg->observation_type;
        p.img_directional = cfg->image_directional ? 1 : 0;
        p.n_layers = cfg->n_image_layers;
        for (int i = 0; i < 8; i++) {
            p.img_layers[i] = cfg->image_layers[i];
            switch (cfg->image_layers[i]) {  // see rw_obs_image_agent
                case 0: p.img_mask[i] = RW_G_SHELF; break;
                case 1: p.img_mask[i] = RW_G_REQ; break;
                case 2: p.img_mask[i] = RW_G_AGENT; break;
                case 5: p.img_mask[i] = 2u << 8; break;                        // goal flag of the static table byte
                case 3: p.img_mask[i] = 0; if (i < cfg->n_image_layers) p.img_transposed |= 1; break;  // second pass,
                case 4: p.img_mask[i] = 0; if (i < cfg->n_image_layers) p.img_transposed |= 2; break;  // rw_obs_image_transposed
                default: p.img_mask[i] = RW_G_AGENT | (1u << 16) | (1u << 31);  // ACCESSIBLE: inside the map and no agent
            }
        }
        p.L = p.n_layers * p.K + (p.obs_kind == RW_OBS_IMAGE_DICT ? 6 : 0);  // warehouse.py:365-388, 400-427
    }
    p.R = cfg->request_queue_size; p.S = S;
    p.reward_type = cfg->reward_type; p.max_steps = cfg->max_steps; p.max_inact = cfg->max_inactivity_steps;
    p.norm_coords = cfg->normalised_coordinates ? 1 : 0; p.n_goals = cfg->n_goals; p.auto_reset = 0;
    for (int g = 0; g < cfg->n_goals; g++) p.goal_cell[g] = cfg->goals_xy[2 * g + 1] * W + cfg->goals_xy[2 * g];
    // tile layout
    int w = 0;
    p.w_agent = w; w += n;
    p.w_msg = w; if (p.msg_bits) w += n;
    p.w_shelf = w; w += (S + 1) / 2;
    p.w_queue = w; w += (p.R + 1) / 2;
    p.w_steps = w++;
    p.w_inact = w++;
    p.nw = rw_round_up(w, 4);  // tile bytes = nw*128: keep it a multiple of 512 B
    p.E = num_envs;
    p.tiles = (int)((num_envs + RW_TILE - 1) / RW_TILE);
    p.E_pad = (int64_t)p.tiles * RW_TILE;
    p.divW = rw_make_div((uint32_t)W);
    p.divN = rw_make_div((uint32_t)n);

    // shared memory carve-up per warp
    int stride = rw_round_up(cells, 4);       // one byte per cell
    if (((stride / 4) & 1) == 0) stride += 4;  // odd word stride: same-cell accesses of the 32 lanes hit 32 banks
    p.grid_env_stride = stride;
    p.obs_env_bytes = n * p.L * 4;
    // Observation staging: a chunk of C consecutive worlds is assembled in shared memory and leaves as one TMA bulk
    // store (C*n*L*4 bytes, must be a 16-byte multiple).  Two lanes share an agent row (obs_parts = 2) so that a
    // full warp of work needs only C*n*2 >= 32 rows' worth of staging, which makes room for double buffering.
    {
        const int budget = 24 * 1024;  // per buffer
        p.obs_parts = p.K >= 2 ? 2 : 1;
        {
            const char *ev = getenv("RWARE_B200_OBS_PARTS");
            if (ev && (atoi(ev) == 1 || atoi(ev) == 2 || atoi(ev) == 4)) p.obs_parts = atoi(ev);
            if (p.obs_parts == 4 && p.K < 4) p.obs_parts = 2;
        }
        int best = 0;
        for (int C = 1; C <= RW_TILE; C *= 2) {
            if (C * p.obs_env_bytes > budget && best) break;
            if ((C * p.obs_env_bytes) % 16 == 0) best = C;
            if (C * n * p.obs_parts >= 32 && best == C) break;  // a full warp of (world, agent, slice) items is enough
        }
        p.obs_bulk = best ? 1 : 0;
        if (!best) {
            best = 1;
            while (best * 2 <= RW_TILE && best * 2 * p.obs_env_bytes <= budget) best *= 2;
        }
        p.obs_chunk_envs = best;
        p.obs_chunk_bytes = best * p.obs_env_bytes;
        // staging buffers per tile: one per warp of the pair; with quarter-row slices (obs_parts == 4) the chunks are half
        // as large, so each warp double-buffers (its next chunk is assembled while the previous TMA store drains)
        p.obs_nbuf = (p.obs_parts == 4 && RW_TILE / best >= 4) ? 4 : 2;
        p.obs_buf_stride = rw_round_up(p.obs_chunk_bytes, 128);
    }
    int off = 0;
    p.sm_tile = off; off += p.nw * RW_TILE * 4; off = rw_round_up(off, 128);
    p.sm_grid = off; off += stride * RW_TILE; off = rw_round_up(off, 16);
    // the resolver scratch (2n words per world) is dead before the observation gather starts: it aliases the staging buffer
    p.sm_scratch = off;
    p.sm_obs = off;
    {
        int obs_bytes = p.obs_buf_stride * p.obs_nbuf, scratch_bytes = 2 * n * RW_TILE * 4;
        off += obs_bytes > scratch_bytes ? obs_bytes : scratch_bytes;
    }
    p.sm_bar = off; off += 8;
    off = rw_round_up(off, 128);
    p.sm_tie = off; off += 2 * 32 * 4;  // 32 words per warp of the pair (rw_coop.cuh)
    p.sm_warp_bytes = rw_round_up(off, 128);
    p.sm_cta_tx = rw_round_up(((cells + 1) & ~1) + 2 * cells, 16);
    p.sm_cta_bytes = rw_round_up(p.sm_cta_tx + 8, 128);
    p.gshift = 0;
    while ((1 << p.gshift) < n) p.gshift++;
    {
        const char *ev = getenv("RWARE_B200_COOP");
        p.coop = ev ? atoi(ev) : 0;  // measured (profiles/phase_timeline.py): one lane per world is ~2.5 us faster per launch
    }
    {
        const char *ev = getenv("RWARE_B200_PAIR");
        p.pair = ev ? atoi(ev) : 1;  // only the fully static kernels with an even agent count <= 8 carry it
    }
    {
        const char *ev = getenv("RWARE_B200_OBS_TMA");
        p.obs_tma = ev ? atoi(ev) : 1;  // measured: TMA bulk stores beat st.global.v4 copy-out (30 vs 38 us under ncu)
    }

}
