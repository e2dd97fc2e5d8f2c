// hostsim.cpp -- TEST-ONLY host twin of the fused CUDA kernel.
//
// Compiles the product's per-world logic (robotic-warehouse_b200/csrc/rw_logic.cuh, the same source the GPU runs)
// with g++ and drives it over the same packed tile layout, emulating what one warp does for its 32 worlds with a
// plain loop over lanes.  It exists so that the agent-centric resolver, the reset sampler and the observation
// writer can be debugged against the oracle on machines without a GPU (`pytest -m "not gpu"`).  It is NOT a CPU
// fallback: nothing in the product package loads it, and the product raises when the CUDA library is missing.
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../robotic-warehouse_b200/csrc/rw_host_layout.h"
#include "../../robotic-warehouse_b200/csrc/rw_logic.cuh"
#include "../../robotic-warehouse_b200/csrc/rw_coop.cuh"
#include "../../robotic-warehouse_b200/csrc/rw_views.cuh"
#include "../../robotic-warehouse_b200/csrc/rw_state_io.h"

struct Sim {
    RwParams p;
    std::vector<uint8_t> highway;
    std::vector<uint16_t> home, home_id;
    std::vector<float> nx, ny;
    std::vector<int32_t> state;
    std::vector<uint32_t> rng;
    unsigned long long err = 0;
    long long pair_steps = 0;
    // one warp's shared memory
    std::vector<uint8_t> smem;
};

static RwLane lane_view(Sim *s, int lane) {
    const RwParams &p = s->p;
    uint8_t *wb = s->smem.data();
    RwLane l;
    l.tile = reinterpret_cast<int32_t *>(wb + p.sm_tile) + lane;
    l.g = wb + p.sm_grid + (size_t)lane * p.grid_env_stride;
    l.sc = reinterpret_cast<uint32_t *>(wb + p.sm_scratch) + lane;
    l.sc_step = RW_TILE;
    l.highway = s->highway.data();
    l.home_id = s->home_id.data();
    return l;
}

extern "C" {

Sim *sim_create(const rw_config *cfg, int64_t num_envs) {
    Sim *s = new Sim();
    int cells = cfg->grid_h * cfg->grid_w;
    s->highway.resize(cells);
    for (int c = 0; c < cells; c++) s->highway[c] = cfg->highways[c] ? 1 : 0;
    for (int g = 0; g < cfg->n_goals; g++) s->highway[cfg->goals_xy[2 * g + 1] * cfg->grid_w + cfg->goals_xy[2 * g]] |= 2;  // goal flag
    for (int c = 0; c < cells; c++)
        if (!cfg->highways[c]) s->home.push_back((uint16_t)c);
    s->home_id.assign(cells, 0);
    for (size_t k = 0; k < s->home.size(); k++) s->home_id[s->home[k]] = (uint16_t)(k + 1);
    rw_fill_params(s->p, *cfg, num_envs, (int)s->home.size());
    RwParams &p = s->p;
    s->nx.resize(p.W);
    s->ny.resize(p.H);
    for (int x = 0; x < p.W; x++) s->nx[x] = p.W > 1 ? (float)((double)x / (double)(p.W - 1)) : 0.0f;
    for (int y = 0; y < p.H; y++) s->ny[y] = p.H > 1 ? (float)((double)y / (double)(p.H - 1)) : 0.0f;
    s->state.assign((size_t)p.tiles * p.nw * RW_TILE, 0);
    s->rng.assign((size_t)RW_RNG_WORDS * p.E_pad, 0);
    s->smem.assign((size_t)p.sm_warp_bytes, 0);
    p.highway = s->highway.data();
    p.cta_tables = s->highway.data();  // bit 1 marks goal cells, as the head of the device's table image
    p.home_cell = s->home.data();
    p.home_id = s->home_id.data();
    p.normx = s->nx.data();
    p.normy = s->ny.data();
    p.state = s->state.data();
    p.rng = s->rng.data();
    p.err_count = &s->err;
    return s;
}

void sim_destroy(Sim *s) { delete s; }
int sim_obs_len(Sim *s) { return s->p.L; }
int sim_n_shelves(Sim *s) { return s->p.S; }
void sim_set_auto_reset(Sim *s, int on) { s->p.auto_reset = on; }
void sim_set_coop(Sim *s, int on) { s->p.coop = on; }
long long sim_error_count(Sim *s) { return (long long)s->err; }
long long sim_pair_steps(Sim *s) { return s->pair_steps; }  // world-steps taken through the agent-split form
void sim_set_pair(Sim *s, int on) { s->p.pair = on; }

void sim_seed(Sim *s, const uint64_t *seeds, const uint8_t *mask) {
    uint32_t w[RW_RNG_WORDS];
    for (int64_t e = 0; e < s->p.E; e++) {
        if (mask && !mask[e]) continue;
        rw_seed_words(seeds[e], w);
        for (int k = 0; k < RW_RNG_WORDS; k++) s->rng[(size_t)k * s->p.E_pad + e] = w[k];
    }
}

}  // extern "C"

// mode 0 step, 1 obs, 2 reset -- mirrors rw_main_kernel<MODE, C>
template <class C>
static void sim_run_t(Sim *s, int mode, const int32_t *actions, float *obs, float *reward, uint8_t *done, const uint8_t *mask) {
    const RwParams &p = s->p;
    for (int tile = 0; tile < p.tiles; tile++) {
        uint8_t *wb = s->smem.data();
        int32_t *tilew = reinterpret_cast<int32_t *>(wb + p.sm_tile);
        int32_t *gtile = p.state + (size_t)tile * p.nw * RW_TILE;
        memcpy(tilew, gtile, (size_t)p.nw * RW_TILE * 4);
        memset(wb + p.sm_grid, 0, (size_t)p.grid_env_stride * RW_TILE);
        if (mode == 0 && p.coop) {  // warp-cooperative step: the host runs the passes one after the other
            RwCoopCtx cx;
            cx.tilew = tilew; cx.grid = wb + p.sm_grid; cx.tiebuf = reinterpret_cast<uint32_t *>(wb + p.sm_tie);
            cx.highway = s->highway.data(); cx.home_id = s->home_id.data();
            cx.tile_env0 = (int64_t)tile * RW_TILE;
            cx.actions = actions + cx.tile_env0 * p.n * p.A;
            cx.reward = reward + cx.tile_env0 * p.n;
            cx.done = done + cx.tile_env0;
            for (int pass = 0; pass < (1 << p.gshift); pass++) rw_step_coop_pass<C>(p, cx, pass, p.gshift);
        }
        for (int lane = 0; lane < RW_TILE && !(mode == 0 && p.coop); lane++) {
            int64_t env = (int64_t)tile * RW_TILE + lane;
            if (env >= p.E) continue;
            RwLane l = lane_view(s, lane);
            if (mode == 2 && (!mask || mask[env])) rw_reset_lane<C>(p, l, env);
            bool pair_done = false, d_pair = false;
            if constexpr (rw_pair_capable<C>()) {
                // the kernel's agent-split step: the phases in the kernel's order, part 0 then part 1 inside each phase
                if (mode == 0) {
                    uint32_t *xw = reinterpret_cast<uint32_t *>(wb + p.sm_tie) + lane;
                    constexpr int NH = RwPairRegs<C>::NH;
                    const int32_t *act = actions + env * p.n;
                    if (p.pair && !((uint32_t)l.tile[rw_off<C>(p).w_inact * RW_TILE] & RW_OVERLAP_BIT)) {
                        RwPairRegs<C> r[2];
                        // inside a phase the two parts are independent: p.pair == 2 runs them in the opposite order (tests)
                        const int rev = p.pair == 2 ? 1 : 0;
                        for (int k = 0; k < 2; k++) { const int part = k ^ rev; rw_pair_propose_a<C>(p, l, part, act + part * NH, r[part]); }
                        rw_grid_requests_agents<C>(p, l);
                        for (int k = 0; k < 2; k++) { const int part = k ^ rev; rw_pair_propose_b<C>(p, l, part, r[part]); }
                        xw[0] = r[1].stay;
                        uint32_t commit = rw_resolve_lane<C>(p, l, r[0].stay | (xw[0] << NH));
                        xw[RW_TILE] = commit;
                        for (int k = 0; k < 2; k++) { const int part = k ^ rev; rw_pair_apply_a<C>(p, l, part, part ? xw[RW_TILE] : commit, r[part]); }
                        for (int k = 0; k < 2; k++) { const int part = k ^ rev; rw_pair_apply_b<C>(p, l, part, r[part]); }
                        xw[0] = r[1].rew;
                        d_pair = rw_pair_finish<C>(p, l, r[0], xw[0], reward + env * p.n, env);
                        s->pair_steps++;
                    } else {
                        d_pair = rw_step_world_serial<C>(p, l, act, reward + env * p.n, env);
                    }
                    pair_done = true;
                }
            }
            if (!pair_done) rw_build_grid<C>(p, l);
            if (mode == 0) {
                bool d;
                if (pair_done) d = d_pair;
                else {
                    rw_load_actions<C>(p, l, actions + env * p.n * p.A);
                    d = rw_step_lane<C>(p, l, reward + env * p.n, env);
                }
                done[env] = d;
                if (d && p.auto_reset) {
                    rw_reset_lane<C>(p, l, env);
                    memset(l.g, 0, (size_t)p.grid_env_stride);
                    rw_build_grid<C>(p, l);
                }
            }
        }
        if (mode != 1) memcpy(gtile, tilew, (size_t)p.nw * RW_TILE * 4);
        if (obs)
            for (int lane = 0; lane < RW_TILE; lane++) {
                int64_t env = (int64_t)tile * RW_TILE + lane;
                if (env >= p.E) continue;
                RwLane l = lane_view(s, lane);
                for (int a = 0; a < p.n; a++)
                    for (int part = 0; part < p.obs_parts; part++) {
                        if (p.obs_parts == 4) rw_obs_row<C, 4>(p, l, a, obs + (env * p.n + a) * p.L, part);
                        else if (p.obs_parts == 2) rw_obs_row<C, 2>(p, l, a, obs + (env * p.n + a) * p.L, part);
                        else rw_obs_row<C, 1>(p, l, a, obs + (env * p.n + a) * p.L, 0);
                    }
            }
    }
}

#include "../../robotic-warehouse_b200/csrc/rw_static_configs.h"

extern "C" {

static int g_force_dynamic = 0;
void sim_force_dynamic(int on) { g_force_dynamic = on; }

// same dispatch rule as rw_kernels.cu:select_kernel
void sim_run(Sim *s, int mode, const int32_t *actions, float *obs, float *reward, uint8_t *done, const uint8_t *mask) {
    const RwParams &p = s->p;
    if (!g_force_dynamic && p.msg_bits == 0) {
#define X(H_, W_, S_, R_, N_, RQ_) \
    if (p.H == H_ && p.W == W_ && p.S == S_ && p.r == R_ && p.n == N_ && p.R == RQ_) \
        return sim_run_t<RwStatic<H_, W_, S_, R_, 0, N_, RQ_>>(s, mode, actions, obs, reward, done, mask);
        RW_FULL_STATIC_CONFIGS(X)
#undef X
#define X(H_, W_, S_, R_) \
    if (p.H == H_ && p.W == W_ && p.S == S_ && p.r == R_) return sim_run_t<RwStatic<H_, W_, S_, R_, 0>>(s, mode, actions, obs, reward, done, mask);
        RW_STATIC_CONFIGS(X)
#undef X
    }
    sim_run_t<RwDynamic>(s, mode, actions, obs, reward, done, mask);
}

// plain-array state access, same field conventions as rw_state_view (host pointers here)
void sim_get_state(Sim *s, const rw_state_view *v) {
    const RwParams &p = s->p;
    for (int64_t env = 0; env < p.E; env++) {
        int32_t *t = p.state + (size_t)(env >> 5) * p.nw * RW_TILE + (env & 31);
        RwLane l{t, nullptr, nullptr, 0, nullptr, nullptr};
        const RwOff o = rw_off<RwDynamic>(p);
        for (int i = 0; i < p.n; i++) {
            uint32_t aw = rw_agent(l, i);
            if (v->agent_x) v->agent_x[env * p.n + i] = rw_ax(aw);
            if (v->agent_y) v->agent_y[env * p.n + i] = rw_ay(aw);
            if (v->agent_dir) v->agent_dir[env * p.n + i] = rw_adir(aw);
            if (v->agent_carry) v->agent_carry[env * p.n + i] = rw_acarry(aw);
            if (v->agent_has_delivered) v->agent_has_delivered[env * p.n + i] = rw_adeliv(aw);
            if (v->agent_msg && p.msg_bits)
                for (int b = 0; b < p.msg_bits; b++)
                    v->agent_msg[(env * p.n + i) * p.msg_bits + b] = ((uint32_t)t[(p.w_msg + i) * RW_TILE] >> b) & 1u;
        }
        for (int k = 0; k < p.S; k++) {
            int c = rw_shelf_cell(o, l, k);
            if (v->shelf_x) v->shelf_x[env * p.S + k] = c % p.W;
            if (v->shelf_y) v->shelf_y[env * p.S + k] = c / p.W;
        }
        if (v->queue)
            for (int q = 0; q < p.R; q++) v->queue[env * p.R + q] = rw_queue(o, l, q);
        if (v->cur_steps) v->cur_steps[env] = t[p.w_steps * RW_TILE];
        if (v->cur_inactive) v->cur_inactive[env] = (int32_t)((uint32_t)t[p.w_inact * RW_TILE] & ~RW_OVERLAP_BIT);
        if (v->rng)
            for (int w = 0; w < RW_RNG_WORDS; w++) v->rng[env * RW_RNG_WORDS + w] = p.rng[(size_t)w * p.E_pad + env];
    }
}

void sim_set_state(Sim *s, const rw_state_view *v) {  // the same per-world function as the CUDA kernel (rw_state_io.h)
    const RwParams &p = s->p;
    for (int64_t env = 0; env < p.E; env++)
        if (rw_set_state_world(p, *v, env)) *p.err_count += 1;
}

// whole-warehouse views (rw_views.cuh), same per-element functions as rw_views.cu
int sim_global_image(Sim *s, const int32_t *layers, int n_layers, const int32_t *pad, float *out) {
    const RwParams &p = s->p;
    RwViewSpec sp;
    memset(&sp, 0, sizeof(sp));
    const int32_t dims[3] = {n_layers, p.H, p.W};
    int32_t shape[3];
    for (int k = 0; k < 3; k++) { shape[k] = pad ? pad[k] : dims[k]; if (shape[k] < dims[k]) return -1; }
    bool dirl = false, loadl = false;
    for (int c = 0; c < n_layers; c++) {
        if (layers[c] < 0 || layers[c] > 6) return -1;
        sp.layers[c] = layers[c]; dirl |= layers[c] == 3; loadl |= layers[c] == 4;
    }
    sp.n_layers = n_layers; sp.C = shape[0]; sp.Hp = shape[1]; sp.Wp = shape[2];
    sp.c0 = (shape[0] - n_layers) / 2; sp.y0 = (shape[1] - p.H) / 2; sp.x0 = (shape[2] - p.W) / 2;
    std::vector<uint8_t> g(p.cells);
    const int per = sp.C * sp.Hp * sp.Wp;
    for (int64_t env = 0; env < p.E; env++) {
        std::fill(g.begin(), g.end(), 0);
        RwLane l{p.state + (size_t)(env >> 5) * p.nw * RW_TILE + (env & 31), g.data(), nullptr, 0, nullptr, nullptr};
        s->err += rw_view_build(p, l, dirl, loadl);
        for (int i = 0; i < per; i++) out[(size_t)env * per + i] = rw_gimg_element(p, l, p.cta_tables, sp, i);
    }
    return 0;
}

void sim_render_shape(Sim *s, int32_t *h, int32_t *w) { *h = rw_render_height(s->p); *w = rw_render_width(s->p); }

int sim_render_rgb(Sim *s, int64_t env0, int64_t count, uint8_t *out) {
    const RwParams &p = s->p;
    if (env0 < 0 || count < 1 || env0 + count > p.E) return -1;
    std::vector<uint8_t> g(p.cells);
    const int fw = rw_render_width(p), fh = rw_render_height(p);
    for (int64_t env = env0; env < env0 + count; env++) {
        std::fill(g.begin(), g.end(), 0);
        RwLane l{p.state + (size_t)(env >> 5) * p.nw * RW_TILE + (env & 31), g.data(), nullptr, 0, nullptr, nullptr};
        rw_view_build(p, l, false, false);
        uint8_t *o = out + (size_t)(env - env0) * fw * fh * 3;
        for (int k = 0; k < fw * fh; k++) {
            uint32_t c = rw_render_pixel(p, l, p.cta_tables, k / fw, k % fw);
            o[3 * k] = (uint8_t)c; o[3 * k + 1] = (uint8_t)(c >> 8); o[3 * k + 2] = (uint8_t)(c >> 16);
        }
    }
    return 0;
}

}  // extern "C"
