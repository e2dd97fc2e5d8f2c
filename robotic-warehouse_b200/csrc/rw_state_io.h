// rw_state_io.h -- world-state injection shared by the CUDA kernel (rw_kernels.cu) and the host test twin (tests/hostsim):
// what the reference's tests do through env.agents[i].x = ..., env._recalc_grid() (tests/test_movement.py:14-47).
#pragma once
#include "../../include/rware_b200.h"
#include "rw_common.h"

// Writes world `env` of the state view into the packed tile.  Out-of-range values (a coordinate outside the grid, a shelf /
// queue id above S) would make the step, observation and view kernels index shared memory out of bounds; the reference raises
// IndexError in Python for the same input.  They are clamped into range here and the function returns true so that the caller
// can count the world (rw_error_count): a bad injection is visible and harmless.
RW_HD int rw_clamp_flag(int v, int lo, int hi, bool &bad) {
    if (v < lo) { bad = true; return lo; }
    if (v > hi) { bad = true; return hi; }
    return v;
}

RW_HD bool rw_set_state_world(const RwParams &p, const rw_state_view &v, int64_t env) {
    int32_t *t = p.state + (size_t)(env >> 5) * p.nw * RW_TILE + (env & 31);
    RwLane l{t, nullptr, nullptr, 0, nullptr, nullptr};
    const RwOff o = rw_off<RwDynamic>(p);
    bool bad = false;
    for (int i = 0; i < p.n; i++) {
        uint32_t aw = rw_agent(l, i);
        int x = v.agent_x ? rw_clamp_flag(v.agent_x[env * p.n + i], 0, p.W - 1, bad) : rw_ax(aw);
        int y = v.agent_y ? rw_clamp_flag(v.agent_y[env * p.n + i], 0, p.H - 1, bad) : rw_ay(aw);
        int d = v.agent_dir ? rw_clamp_flag(v.agent_dir[env * p.n + i], 0, 3, bad) : rw_adir(aw);
        int c = v.agent_carry ? rw_clamp_flag(v.agent_carry[env * p.n + i], 0, p.S, bad) : rw_acarry(aw);
        int hd = v.agent_has_delivered ? (v.agent_has_delivered[env * p.n + i] != 0) : rw_adeliv(aw);
        rw_set_agent(l, i, rw_apack(x, y, d, hd, c));
        if (v.agent_msg && p.msg_bits) {
            uint32_t bits = 0;
            for (int b = 0; b < p.msg_bits; b++) bits |= (uint32_t)(v.agent_msg[(env * p.n + i) * p.msg_bits + b] & 1) << b;
            t[(p.w_msg + i) * RW_TILE] = (int32_t)bits;
        }
    }
    if (v.shelf_x && v.shelf_y)
        for (int k = 0; k < p.S; k++)
            rw_set_shelf_cell(o, l, k, rw_clamp_flag(v.shelf_y[env * p.S + k], 0, p.H - 1, bad) * p.W +
                                           rw_clamp_flag(v.shelf_x[env * p.S + k], 0, p.W - 1, bad));
    if (v.queue)
        for (int q = 0; q < p.R; q++) rw_set_queue(o, l, q, rw_clamp_flag(v.queue[env * p.R + q], 0, p.S, bad));
    if (v.cur_steps) t[p.w_steps * RW_TILE] = v.cur_steps[env];
    if (v.cur_inactive) t[p.w_inact * RW_TILE] = v.cur_inactive[env] & 0x7fffffff;
    // flag worlds where two shelves (or two agents) share a cell (cold path, O(S^2)); see RW_OVERLAP_BIT
    bool overlap = false;
    for (int k = 1; k < p.S && !overlap; k++) {
        int c = rw_shelf_cell(o, l, k);
        for (int j = 0; j < k; j++) overlap |= (rw_shelf_cell(o, l, j) == c);
    }
    for (int i = 1; i < p.n && !overlap; i++)  // two agents on one cell take the same literal path
        for (int j = 0; j < i; j++) overlap |= ((rw_agent(l, i) & 0xffffu) == (rw_agent(l, j) & 0xffffu));
    uint32_t wi = (uint32_t)t[p.w_inact * RW_TILE] & ~RW_OVERLAP_BIT;
    t[p.w_inact * RW_TILE] = (int32_t)(wi | (overlap ? RW_OVERLAP_BIT : 0u));
    if (v.rng)
        for (int w = 0; w < RW_RNG_WORDS; w++) p.rng[(size_t)w * p.E_pad + env] = v.rng[env * RW_RNG_WORDS + w];
    return bad;
}
