// rw_logic.cuh -- per-world ("lane") logic of the fused RWARE step: action decode, move proposal, collision
// resolution, apply, deliveries/resample, counters, reset, and the per-agent observation writer.
//
// One CUDA lane advances one world; a warp advances a tile of 32 worlds in lock step (rw_kernels.cu).  The
// code is written against RwLane (rw_common.h) and contains no warp intrinsics, so the very same source is
// also compiled for the host by tests/hostsim (a test-only twin used to debug the logic without a GPU).
//
// Semantics follow the reference (semitable/robotic-warehouse @ 96fbc64e, rware/warehouse.py; cited inline
// as "ref:LINE") but not its formulation: the reference hands a cell graph to networkx; here the resolver
// is agent-centric (every agent has at most one successor agent), see DESIGN.md section 4.
#pragma once
#include "rw_common.h"

// ------------------------------------------------------------------------------------------------
// numpy-compatible PCG64 (XSL-RR 128/64) + Lemire bounded draws, so that reset()/resample reproduce the
// reference's np_random.choice(...) calls bit for bit (ref:781, 788, 797, 916).
// numpy/random/src/pcg64/pcg64.h, numpy/random/src/distributions/distributions.c.
struct RwPcg {
    uint64_t s_lo, s_hi, i_lo, i_hi;
    uint32_t has32, u32;
};

RW_HD uint64_t rw_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

RW_HD void rw_pcg_load(RwPcg &g, const uint32_t *rng, int64_t stride, int64_t env) {
    const uint32_t *b = rng + env;
    g.s_lo = (uint64_t)b[0 * stride] | ((uint64_t)b[1 * stride] << 32);
    g.s_hi = (uint64_t)b[2 * stride] | ((uint64_t)b[3 * stride] << 32);
    g.i_lo = (uint64_t)b[4 * stride] | ((uint64_t)b[5 * stride] << 32);
    g.i_hi = (uint64_t)b[6 * stride] | ((uint64_t)b[7 * stride] << 32);
    g.has32 = b[8 * stride];
    g.u32 = b[9 * stride];
}

RW_HD void rw_pcg_store(const RwPcg &g, uint32_t *rng, int64_t stride, int64_t env) {
    uint32_t *b = rng + env;
    b[0 * stride] = (uint32_t)g.s_lo; b[1 * stride] = (uint32_t)(g.s_lo >> 32);
    b[2 * stride] = (uint32_t)g.s_hi; b[3 * stride] = (uint32_t)(g.s_hi >> 32);
    b[8 * stride] = g.has32;
    b[9 * stride] = g.u32;
}

RW_HD uint64_t rw_pcg_next64(RwPcg &g) {
    const uint64_t M_HI = 2549297995355413924ULL, M_LO = 4865540595714422341ULL;
    uint64_t lo = g.s_lo * M_LO;
    uint64_t hi = rw_mulhi64(g.s_lo, M_LO) + g.s_lo * M_HI + g.s_hi * M_LO;
    uint64_t nlo = lo + g.i_lo;
    hi += g.i_hi + (uint64_t)(nlo < lo);
    g.s_lo = nlo;
    g.s_hi = hi;
    uint64_t x = hi ^ nlo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}

RW_HD uint32_t rw_pcg_next32(RwPcg &g) {
    if (g.has32) {
        g.has32 = 0;
        return g.u32;
    }
    uint64_t v = rw_pcg_next64(g);
    g.has32 = 1;
    g.u32 = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

// uniform on [0, rng] inclusive (buffered_bounded_lemire_uint32); rng == 0 consumes nothing
RW_HD uint32_t rw_bounded(RwPcg &g, uint32_t rng) {
    if (rng == 0) return 0;
    const uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)rw_pcg_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)rw_pcg_next32(g) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

// ------------------------------------------------------------------------------------------------
// occupancy grid (what ref:749-755 _recalc_grid keeps in self.grid), rebuilt in shared memory every step as one
// byte per cell.  Precondition: this world's grid region is zero.
RW_HD int rw_ffs32(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)b) - 1;
#else
    return __builtin_ctz(b);
#endif
}

// id of the shelf the reference's grid shows on `cell`: the highest id standing there (ref:751-752), 0 if none
template <class C>
RW_HD int rw_shelf_at(const RwParams &p, const RwLane &l, uint32_t cell) {
    const RwOff o = rw_off<C>(p);
    const int S = C::S(p);
    // fast path: without overlaps a cell holds at most one shelf, and almost always it is the one whose home it is
    int id = l.home_id[cell];
    if (id && (uint32_t)rw_shelf_cell(o, l, id - 1) == cell && !((uint32_t)l.tile[o.w_inact * RW_TILE] & RW_OVERLAP_BIT))
        return id;
    id = 0;
#pragma unroll 4
    for (int k2 = 0; k2 < (S >> 1); k2++) {
        uint32_t w = (uint32_t)l.tile[(o.w_shelf + k2) * RW_TILE];
        if ((w & 0xffffu) == cell) id = 2 * k2 + 1;
        if ((w >> 16) == cell) id = 2 * k2 + 2;
    }
    if (S & 1) {
        uint32_t w = (uint32_t)l.tile[(o.w_shelf + (S >> 1)) * RW_TILE];
        if ((w & 0xffffu) == cell) id = S;
    }
    return id;
}

RW_HD bool rw_in_queue(const RwOff &o, const RwLane &l, int id) {
    for (int q = 0; q < o.R; q++)
        if (rw_queue(o, l, q) == id) return true;
    return false;
}

// shelf bits of the state words [k_lo, k_hi) (two shelves per word; `tail`: also the odd last shelf)
template <class C>
RW_HD void rw_grid_shelves(const RwParams &p, const RwLane &l, int k_lo, int k_hi, bool tail) {
    const RwOff o = rw_off<C>(p);
    const int S = C::S(p);
    {   // batches of 8 words: all loads first, then the scattered byte stores
        int k2 = k_lo;
        for (; k2 + 8 <= k_hi; k2 += 8) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = (uint32_t)l.tile[(o.w_shelf + k2 + j) * RW_TILE];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                l.g[w[j] & 0xffffu] = (uint8_t)RW_G_SHELF;
                l.g[w[j] >> 16] = (uint8_t)RW_G_SHELF;
            }
        }
        const int rest = k_hi - k2;  // < 8
        for (int j = 0; j < rest; j++) {
            uint32_t w = (uint32_t)l.tile[(o.w_shelf + k2 + j) * RW_TILE];
            l.g[w & 0xffffu] = (uint8_t)RW_G_SHELF;
            l.g[w >> 16] = (uint8_t)RW_G_SHELF;
        }
    }
    if (tail && (S & 1)) l.g[(uint32_t)l.tile[(o.w_shelf + (S >> 1)) * RW_TILE] & 0xffffu] = (uint8_t)RW_G_SHELF;
}

// request bits and agent ids on top of the shelf bits
template <class C>
RW_HD void rw_grid_requests_agents(const RwParams &p, const RwLane &l) {
    const RwOff o = rw_off<C>(p);
    const bool overlap = ((uint32_t)l.tile[o.w_inact * RW_TILE] & RW_OVERLAP_BIT) != 0;
    for (int q = 0; q < o.R; q++) {
        int id = rw_queue(o, l, q);
        if (!id) continue;
        uint32_t c = (uint32_t)rw_shelf_cell(o, l, id - 1);
        // with overlapping shelves the reference's grid shows the highest id on the cell; only that one counts
        if (!overlap || rw_shelf_at<C>(p, l, c) == id) l.g[c] |= (uint8_t)RW_G_REQ;
    }
    for (int i = 0; i < o.n; i++) {
        uint32_t aw = rw_agent(l, i);
        uint32_t c = (uint32_t)(rw_ay(aw) * C::W(p) + rw_ax(aw));
        l.g[c] = (uint8_t)((l.g[c] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i + 1));
    }
}

template <class C>
RW_HD void rw_build_grid(const RwParams &p, const RwLane &l) {
    rw_grid_shelves<C>(p, l, 0, C::S(p) >> 1, true);
    rw_grid_requests_agents<C>(p, l);
}

// literal _recalc_grid for worlds flagged with overlapping shelves; also refreshes the flag (O(S^2), cold)
template <class C>
RW_HD_NOINLINE void rw_rebuild_grid_slow(const RwParams &p, const RwLane l) {  // by value: `l` must not escape the caller
    const RwOff o = rw_off<C>(p);
    uint32_t *gz = reinterpret_cast<uint32_t *>(l.g);
    for (int i = 0; i < (p.grid_env_stride >> 2); i++) gz[i] = 0;
    bool overlap = false;
    for (int k = 0; k < C::S(p) && !overlap; k++) {
        int c = rw_shelf_cell(o, l, k);
        for (int j = 0; j < k; j++) overlap |= (rw_shelf_cell(o, l, j) == c);
    }
    for (int i = 1; i < o.n && !overlap; i++)  // two agents on one cell (injected): same flag, same literal path
        for (int j = 0; j < i; j++) overlap |= ((rw_agent(l, i) & 0xffffu) == (rw_agent(l, j) & 0xffffu));
    uint32_t w = (uint32_t)l.tile[o.w_inact * RW_TILE];
    l.tile[o.w_inact * RW_TILE] = (int32_t)((w & ~RW_OVERLAP_BIT) | (overlap ? RW_OVERLAP_BIT : 0u));
    rw_build_grid<C>(p, l);
}

// scratch accessors: entry 2i = start | target<<16, entry 2i+1 = meta
RW_HD uint32_t &rw_st(const RwLane &l, int i) { return l.sc[i * (l.sc_step == 1 ? 1 : 2 * RW_TILE)]; }
RW_HD uint32_t &rw_meta(const RwLane &l, int i) { return l.sc[(2 * i + 1) * RW_TILE]; }  // one-lane-per-world step only
// meta: act[0:3) nxt[3:9) h[9:15) sh[16:32)
RW_HD uint32_t rw_m_act(uint32_t m) { return m & 7u; }
RW_HD uint32_t rw_m_nxt(uint32_t m) { return (m >> 3) & 63u; }
RW_HD uint32_t rw_m_h(uint32_t m) { return (m >> 9) & 63u; }

// ------------------------------------------------------------------------------------------------
// CPython 3.12 tuple hash of (x, y) and set insertion/iteration order (Objects/tupleobject.c:tuplehash,
// Objects/setobject.c:set_add_entry / set_insert_clean / set_table_resize).  Needed only to order tied
// contenders the way networkx's component copy does.
RW_HD uint64_t rw_py_hash_xy(uint32_t x, uint32_t y) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    acc += (uint64_t)x * P2;
    acc = (acc << 31) | (acc >> 33);
    acc *= P1;
    acc += (uint64_t)y * P2;
    acc = (acc << 31) | (acc >> 33);
    acc *= P1;
    acc += 2ULL ^ (P5 ^ 3527539ULL);
    return acc == ~0ULL ? 1546275796ULL : acc;
}

struct RwPySet {  // keys are node ids < 64; tables never exceed 128 slots for <= 64 keys
    uint8_t tab[128];
    int mask, fill;
};

RW_HD void rw_pyset_init(RwPySet &s) {
    s.mask = 7;
    s.fill = 0;
    for (int i = 0; i < 8; i++) s.tab[i] = 0xff;
}

RW_HD int rw_pyset_probe(const uint8_t *tab, int mask, uint64_t hash) {
    uint64_t perturb = hash;
    uint32_t i = (uint32_t)hash & (uint32_t)mask;
    for (;;) {
        if (tab[i] == 0xff) return (int)i;
        if (i + 9 <= (uint32_t)mask) {
            for (uint32_t j = 1; j <= 9; j++)
                if (tab[i + j] == 0xff) return (int)(i + j);
        }
        perturb >>= 5;
        i = (uint32_t)((uint64_t)i * 5 + 1 + perturb) & (uint32_t)mask;
    }
}

template <class HashOf>
RW_HD void rw_pyset_add(RwPySet &s, int key, HashOf hash_of) {
    s.tab[rw_pyset_probe(s.tab, s.mask, hash_of(key))] = (uint8_t)key;
    s.fill++;
    if (s.fill * 5 < s.mask * 3) return;
    int newsize = 8;
    while (newsize <= s.fill * 4) newsize <<= 1;
    uint8_t nt[128];
    for (int i = 0; i < newsize; i++) nt[i] = 0xff;
    for (int i = 0; i <= s.mask; i++)
        if (s.tab[i] != 0xff) nt[rw_pyset_probe(nt, newsize - 1, hash_of(s.tab[i]))] = s.tab[i];
    for (int i = 0; i < newsize; i++) s.tab[i] = nt[i];
    s.mask = newsize - 1;
}

// Among the agents in `ties` (bit i = agent i; their start cells are tied predecessors of one cell), return the
// one whose start cell comes first in the node order of networkx's G.subgraph(component).copy():
// G insertion order when 2*|component| >= |G|, CPython set iteration order otherwise
// (NX/classes/coreviews.py:293-300, 342-349; NX/algorithms/components/weakly_connected.py:61-66, 169-196).
template <class C>
RW_HD_NOINLINE int rw_tie_first(const RwParams &p, const RwLane l, uint32_t ties) {
    const RwOff o = rw_off<C>(p);
    const int n = o.n, m = 2 * n;
    // node id of a cell := index of its first appearance in the sequence s_0, t_0, s_1, t_1, ... (DiGraph.add_edge
    // inserts the start node, then the target node, ref:825-846)
    uint8_t first[2 * RW_MAX_AGENTS];
    uint64_t is_node = 0;
    for (int k = 0; k < m; k++) {
        uint32_t ck = (k & 1) ? (rw_st(l, k >> 1) >> 16) : (rw_st(l, k >> 1) & 0xffffu);
        int f = k;
        for (int j = 0; j < k; j++) {
            uint32_t cj = (j & 1) ? (rw_st(l, j >> 1) >> 16) : (rw_st(l, j >> 1) & 0xffffu);
            if (cj == ck) { f = j; break; }
        }
        first[k] = (uint8_t)f;
        if (f == k) is_node |= 1ULL << k;
    }
    int nG = 0;
    for (uint64_t b = is_node; b; b &= b - 1) nG++;
    // weakly connected component containing the tied cells: flood fill over succ/pred
    int cand0 = rw_ffs32(ties);
    uint64_t member = 1ULL << first[2 * cand0], frontier = member;
    while (frontier) {
        uint64_t next = 0;
        for (int a = 0; a < n; a++) {
            uint64_t bs = 1ULL << first[2 * a], bt = 1ULL << first[2 * a + 1];
            if (frontier & bs) next |= bt;
            if (frontier & bt) next |= bs;
        }
        frontier = next & ~member;
        member |= frontier;
    }
    int csize = 0;
    for (uint64_t b = member; b; b &= b - 1) csize++;
    if (2 * csize >= nG) {  // G insertion order
        int best = -1, best_node = 1 << 30;
        for (int a = 0; a < n; a++)
            if (((ties >> a) & 1u) && first[2 * a] < best_node) { best_node = first[2 * a]; best = a; }
        return best;
    }
    // CPython set order.  c = _plain_bfs(G, source = first node of the component in G order)
    auto cell_of = [&](int node) -> uint32_t {
        return (node & 1) ? (rw_st(l, node >> 1) >> 16) : (rw_st(l, node >> 1) & 0xffffu);
    };
    auto hash_of = [&](int node) -> uint64_t {
        uint32_t c = cell_of(node);
        uint32_t y = C::div_w(p, c);
        return rw_py_hash_xy(c - y * (uint32_t)C::W(p), y);
    };
    int src = 0;
    while (!((member >> src) & 1ULL)) src++;
    RwPySet set1;
    rw_pyset_init(set1);
    uint64_t seen = 1ULL << src;
    rw_pyset_add(set1, src, hash_of);
    uint8_t level[2 * RW_MAX_AGENTS], nextl[2 * RW_MAX_AGENTS];
    int nlevel = 1, nnext = 0;
    level[0] = (uint8_t)src;
    while (nlevel) {
        nnext = 0;
        for (int li = 0; li < nlevel; li++) {
            int v = level[li];
            uint32_t cv = cell_of(v);
            int occ = l.g[cv] & RW_G_AGENT;  // agent standing on this cell (pre-move grid) owns its single out-edge
            if (occ) {
                int w = first[2 * (occ - 1) + 1];
                if (!((seen >> w) & 1ULL)) { seen |= 1ULL << w; rw_pyset_add(set1, w, hash_of); nextl[nnext++] = (uint8_t)w; }
            }
            for (int a = 0; a < n; a++)  // G._pred[v] in edge (= agent) order
                if ((rw_st(l, a) >> 16) == cv) {
                    int w = first[2 * a];
                    if (!((seen >> w) & 1ULL)) { seen |= 1ULL << w; rw_pyset_add(set1, w, hash_of); nextl[nnext++] = (uint8_t)w; }
                }
        }
        for (int i = 0; i < nnext; i++) level[i] = nextl[i];
        nlevel = nnext;
    }
    // show_nodes.nodes = set(nbunch_iter(c)): second set filled in the iteration order of the first
    RwPySet set2;
    rw_pyset_init(set2);
    for (int i = 0; i <= set1.mask; i++)
        if (set1.tab[i] != 0xff) rw_pyset_add(set2, set1.tab[i], hash_of);
    for (int i = 0; i <= set2.mask; i++) {
        int node = set2.tab[i];
        if (node == 0xff) continue;
        for (int a = 0; a < n; a++)
            if (((ties >> a) & 1u) && first[2 * a] == node) return a;
    }
    return cand0;  // unreachable
}

// pick the contender with the greatest height; ties by rw_tie_first
template <class C>
RW_HD int rw_pick(const RwParams &p, const RwLane &l, uint32_t cand) {
    uint32_t best = 0, ties = 0;
    int besth = -1;
    for (uint32_t b = cand; b; b &= b - 1) {
        int k = rw_ffs32(b);
        int h = (int)rw_m_h(rw_meta(l, k));
        if (h > besth) { besth = h; best = (uint32_t)k; ties = 1u << k; }
        else if (h == besth) ties |= 1u << k;
    }
    if (ties & (ties - 1)) return rw_tie_first<C>(p, l, ties);
    return (int)best;
}

// ------------------------------------------------------------------------------------------------
// request-queue resample on delivery (ref:915-917): candidates = shelves not in the queue in ascending id,
// new = np_random.choice(candidates), replaces the delivered shelf in place.
template <class C>
RW_HD_NOINLINE void rw_resample(const RwParams &p, const RwLane l, int delivered_id, int64_t env) {
    const RwOff o = rw_off<C>(p);
    const int S = C::S(p);
    int distinct = 0;
    for (int q = 0; q < o.R; q++) {
        int id = rw_queue(o, l, q);
        bool dup = false;
        for (int j = 0; j < q; j++) dup |= (rw_queue(o, l, j) == id);
        distinct += !dup;
    }
    int n_cand = S - distinct;
    int new_id = delivered_id;  // S == R: the reference raises ValueError (empty population); we keep the shelf
    if (n_cand > 0) {
        RwPcg g;
        rw_pcg_load(g, p.rng, p.E_pad, env);
        int pick = (int)rw_bounded(g, (uint32_t)(n_cand - 1));
        rw_pcg_store(g, p.rng, p.E_pad, env);
        for (int s = 1; s <= S; s++) {
            if (rw_in_queue(o, l, s)) continue;
            if (pick-- == 0) { new_id = s; break; }
        }
    }
    for (int q = 0; q < o.R; q++)
        if (rw_queue(o, l, q) == delivered_id) { rw_set_queue(o, l, q, new_id); break; }
    if ((uint32_t)l.tile[o.w_inact * RW_TILE] & RW_OVERLAP_BIT) {
        rw_rebuild_grid_slow<C>(p, l);
        return;
    }
    // request bits of the two shelves involved
    uint32_t cd = (uint32_t)rw_shelf_cell(o, l, delivered_id - 1);
    if (!rw_in_queue(o, l, delivered_id)) l.g[cd] &= (uint8_t)~RW_G_REQ;
    l.g[rw_shelf_cell(o, l, new_id - 1)] |= (uint8_t)RW_G_REQ;
}

// ------------------------------------------------------------------------------------------------
// Warehouse.reset() for one world (ref:757-802), continuing the world's generator stream.  Does NOT touch
// the occupancy grid (the caller rebuilds it).
template <class C>
RW_HD_NOINLINE void rw_reset_lane(const RwParams &p, const RwLane l, int64_t env) {
    const RwOff o = rw_off<C>(p);
    RwPcg g;
    rw_pcg_load(g, p.rng, p.E_pad, env);
    const int S = C::S(p), W = C::W(p);
    for (int k = 0; k < S; k++) rw_set_shelf_cell(o, l, k, p.home_cell[k]);  // ref:771-778
    const int n = o.n;
    // ref:781-785 np_random.choice(arange(H*W), size=n, replace=False): Floyd's algorithm then a shuffle
    const int cells = C::H(p) * W;
    for (int k = 0; k < n; k++) {
        uint32_t j = (uint32_t)(cells - n + k);
        uint32_t val = rw_bounded(g, j);
        bool found = false;
        for (int q = 0; q < k; q++) found |= (l.sc[q * l.sc_step] == val);
        l.sc[k * l.sc_step] = found ? j : val;
    }
    for (int i = n - 1; i >= 1; i--) {
        uint32_t j = rw_bounded(g, (uint32_t)i);
        uint32_t t = l.sc[i * l.sc_step];
        l.sc[i * l.sc_step] = l.sc[j * l.sc_step];
        l.sc[j * l.sc_step] = t;
    }
    for (int i = 0; i < n; i++) {  // ref:786-792: unravel_index -> (y, x); dirs from choice(Direction, size=n)
        uint32_t loc = l.sc[i * l.sc_step];
        uint32_t y = C::div_w(p, loc);
        uint32_t dir = rw_bounded(g, 3);
        rw_set_agent(l, i, rw_apack((int)(loc - y * (uint32_t)W), (int)y, (int)dir, 0, 0));
        if (C::mb(p)) l.tile[(o.w_msg + i) * RW_TILE] = 0;
    }
    // ref:796-800 np_random.choice(self.shelfs, size=R, replace=False)
    for (int k = 0; k < o.R; k++) {
        uint32_t j = (uint32_t)(S - o.R + k);
        uint32_t val = rw_bounded(g, j);
        bool found = false;
        for (int q = 0; q < k; q++) found |= ((uint32_t)rw_queue(o, l, q) == val + 1);
        rw_set_queue(o, l, k, (int)((found ? j : val) + 1));
    }
    for (int i = o.R - 1; i >= 1; i--) {
        int j = (int)rw_bounded(g, (uint32_t)i);
        int t = rw_queue(o, l, i);
        rw_set_queue(o, l, i, rw_queue(o, l, j));
        rw_set_queue(o, l, j, t);
    }
    l.tile[o.w_steps * RW_TILE] = 0;  // ref:764-765
    l.tile[o.w_inact * RW_TILE] = 0;  // (also clears the overlap flag: every shelf is back on its own home cell)
    rw_pcg_store(g, p.rng, p.E_pad, env);
}

// ------------------------------------------------------------------------------------------------
// Warehouse.step() for one world (ref:804-942), everything except the observation gather.
// act: this world's actions [n][A]; reward: this world's [n] floats.  Returns done.
// Precondition: grid built from the current state.  Postcondition: tile words and grid hold the
// post-step state (post-move agents/shelves, post-resample request bits).
// Deliveries, rewards, request-queue resample, counters and termination of one world (ref:903-941), run after the
// moves have been applied to the tile words and the grid.  Rewards accumulate in half units: 4 bits per agent in
// rew_lo (agents 0-15) / rew_hi (16-31), plus rew_global whole units for every agent.  Returns done.
template <class C>
RW_HD bool rw_finish_lane(const RwParams &p, const RwLane &l, int64_t env, uint64_t &rew_lo, uint64_t &rew_hi, int &rew_global) {
    const RwOff o = rw_off<C>(p);
    bool delivered = false;
    for (int gi = 0; gi < p.n_goals; gi++) {
        uint32_t c = (uint32_t)p.goal_cell[gi];
        uint32_t gb = l.g[c];
        if ((gb & (RW_G_SHELF | RW_G_REQ)) != (RW_G_SHELF | RW_G_REQ)) continue;
        delivered = true;
        rw_resample<C>(p, l, rw_shelf_at<C>(p, l, c), env);
        int occ = (int)(gb & RW_G_AGENT);
        int ai = occ ? occ - 1 : o.n - 1;  // python's rewards[-1] when nobody stands on the goal
        if (p.reward_type == 0) rew_global++;
        else {
            uint64_t add = (p.reward_type == 1) ? 2ULL : 1ULL;
            if (ai < 16) rew_lo += add << (4 * ai); else rew_hi += add << (4 * (ai - 16));
            if (p.reward_type == 2) rw_set_agent(l, ai, rw_agent(l, ai) | (1u << 18));
        }
    }
    // counters / termination (ref:929-941)
    uint32_t wi = (uint32_t)l.tile[o.w_inact * RW_TILE];
    int inact = delivered ? 0 : (int)(wi & ~RW_OVERLAP_BIT) + 1;
    int steps = l.tile[o.w_steps * RW_TILE] + 1;
    l.tile[o.w_inact * RW_TILE] = (int32_t)((uint32_t)inact | (wi & RW_OVERLAP_BIT));
    l.tile[o.w_steps * RW_TILE] = steps;
    return (p.max_inact && inact >= p.max_inact) || (p.max_steps && steps >= p.max_steps);
}

// Action decode (ref:809-814), hoisted out of the step so that the global loads are issued back to back while
// the state tile is still in flight: scratch entry 2i+1 <- action | message bits << 8.  Invalid values become NOOP.
template <class C>
RW_HD void rw_load_actions(const RwParams &p, const RwLane &l, const int32_t *act) {
    const RwOff o = rw_off<C>(p);
    const int n = o.n, mb = C::mb(p), A = 1 + mb;
    uint32_t bad = 0;
    for (int i = 0; i < n; i++) {
        int a = act[i * A];
        if ((unsigned)a > 4u) { a = RW_ACT_NOOP; bad++; }
        uint32_t bits = 0;
        for (int b = 0; b < mb; b++) bits |= (uint32_t)(act[i * A + 1 + b] & 1) << b;
        rw_meta(l, i) = (uint32_t)a | (bits << 8);
    }
    if (bad) {
#if defined(__CUDA_ARCH__)
        atomicAdd(p.err_count, (unsigned long long)bad);
#else
        *p.err_count += bad;
#endif
    }
}

// Register prefetch of the actions for fully static configurations without message bits: the loads are issued at the
// very top of the kernel and only consumed (rw_commit_actions) after the shared-memory barriers, so their DRAM latency
// overlaps the TMA transfers instead of serialising in front of them.
template <class C>
struct RwActRegs {
    static constexpr int N = C::full_static ? 8 : 1;
    int32_t v[N];
};
template <class C>
RW_HD bool rw_can_prefetch_actions(const RwParams &p) { return C::full_static && C::mb(p) == 0 && C::n(p) <= 8; }
template <class C>
RW_HD void rw_prefetch_actions(const RwParams &p, const int32_t *act, RwActRegs<C> &r) {
#pragma unroll
    for (int i = 0; i < RwActRegs<C>::N; i++)
        if (i < C::n(p)) r.v[i] = act[i];
}
template <class C>
RW_HD void rw_commit_actions(const RwParams &p, const RwLane &l, const RwActRegs<C> &r) {
    uint32_t bad = 0;
#pragma unroll
    for (int i = 0; i < RwActRegs<C>::N; i++)
        if (i < C::n(p)) {
            int a = r.v[i];
            if ((unsigned)a > 4u) { a = RW_ACT_NOOP; bad++; }
            rw_meta(l, i) = (uint32_t)a;
        }
    if (bad) {
#if defined(__CUDA_ARCH__)
        atomicAdd(p.err_count, (unsigned long long)bad);
#else
        *p.err_count += bad;
#endif
    }
}

// Collision resolution for one world (ref:848-876 restated agent-centrically).  Input: the scratch
// holds start|target and action|successor of every agent, `commit` the agents that stay put.  Returns the final
// commit mask (bit i: agent i's proposed move, if any, is carried out).
template <class C>
RW_HD uint32_t rw_resolve_lane(const RwParams &p, const RwLane &l, uint32_t commit) {
    const RwOff o = rw_off<C>(p);
    const int n = o.n;
    uint32_t dag = 0;
    // ---- resolve (ref:848-876 restated agent-centrically)
    uint32_t movers = ~commit & (n == 32 ? 0xffffffffu : ((1u << n) - 1u));
    if (movers) {
        for (uint32_t b = movers; b; b &= b - 1) {
            int i = rw_ffs32(b);
            uint32_t j = rw_m_nxt(rw_meta(l, i));
            int hops = 1;
            for (;;) {
                if (j == RW_NXT_SINK) { dag |= 1u << i; break; }          // chain ends in a free cell
                if (j == (uint32_t)i) { if (hops != 2) commit |= 1u << i; break; }  // on a cycle; 2-cycles (swaps) fail
                uint32_t nj = rw_m_nxt(rw_meta(l, (int)j));
                if (nj == RW_NXT_SELF || hops >= n) break;               // blocked by a stationary agent / feeds a cycle
                j = nj;
                hops++;
            }
        }
        if (dag) {
            // heights: h[j] = 1 + max h[i] over dag agents i stepping onto j's cell (dag.py:1073-1101 `dist`)
            bool changed = true;
            for (int round = 0; changed && round < n; round++) {
                changed = false;
                for (uint32_t b = dag; b; b &= b - 1) {
                    int i = rw_ffs32(b);
                    uint32_t mi = rw_meta(l, i);
                    uint32_t j = rw_m_nxt(mi);
                    if (j >= RW_NXT_SINK) continue;
                    uint32_t mj = rw_meta(l, (int)j);
                    uint32_t hi1 = rw_m_h(mi) + 1;
                    if (rw_m_h(mj) < hi1) { rw_meta(l, (int)j) = (mj & ~(63u << 9)) | (hi1 << 9); changed = true; }
                }
            }
            // walk back from every sink cell along the longest path
            uint32_t sinkers = 0;
            for (uint32_t b = dag; b; b &= b - 1) {
                int i = rw_ffs32(b);
                if (rw_m_nxt(rw_meta(l, i)) == RW_NXT_SINK) sinkers |= 1u << i;
            }
            uint32_t handled = 0;
            while (sinkers & ~handled) {
                uint32_t rem = sinkers & ~handled;
                int i0 = rw_ffs32(rem);
                uint32_t c = rw_st(l, i0) >> 16, cand = 0;
                for (uint32_t b = rem; b; b &= b - 1) {
                    int k = rw_ffs32(b);
                    if ((rw_st(l, k) >> 16) == c) cand |= 1u << k;
                }
                handled |= cand;
                int cur = (cand & (cand - 1)) ? rw_pick<C>(p, l, cand) : i0;
                for (;;) {
                    commit |= 1u << cur;
                    if (rw_m_h(rw_meta(l, cur)) == 0) break;  // nobody steps onto cur's cell
                    cand = 0;
                    for (uint32_t b = dag; b; b &= b - 1) {
                        int k = rw_ffs32(b);
                        if (rw_m_nxt(rw_meta(l, k)) == (uint32_t)cur) cand |= 1u << k;
                    }
                    cur = rw_pick<C>(p, l, cand);
                }
            }
        }
    }
    return commit;
}

template <class C>
RW_HD bool rw_step_lane(const RwParams &p, const RwLane &l, float *reward, int64_t env) {
    const RwOff o = rw_off<C>(p);
    const int n = o.n, W = C::W(p), H = C::H(p), mb = C::mb(p);
    uint32_t commit = 0;
    const bool overlap = ((uint32_t)l.tile[o.w_inact * RW_TILE] & RW_OVERLAP_BIT) != 0;
    uint64_t rew_lo = 0, rew_hi = 0;  // per-agent reward in half units, 4 bits each
    int rew_global = 0;
    uint32_t moved = 0;
    if constexpr (C::full_static && C::static_n >= 1 && C::static_n <= 8) {
        // ---- fully static agent count: everything per agent lives in registers and every phase is written as explicit
        // batches (all loads, then arithmetic, then the dependent loads, stores last) so that the shared-memory
        // latencies of the agents overlap instead of adding up.  Same semantics as the generic branch below.
        // proposals (ref:825-846); actions were decoded into the scratch beforehand
        constexpr int N = C::static_n;
        uint32_t awv[N], av[N], sv[N], tv[N], gtv[N];
#pragma unroll
        for (int i = 0; i < N; i++) {
            awv[i] = rw_agent(l, i);
            av[i] = rw_meta(l, i);
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (mb) l.tile[(o.w_msg + i) * RW_TILE] = (int32_t)(av[i] >> 8);
            av[i] &= 7u;
            int x = rw_ax(awv[i]), y = rw_ay(awv[i]), tx = x, ty = y;
            if (av[i] == RW_ACT_FORWARD) {  // Agent.req_location ref:102-116, clamped at the walls
                int d = rw_adir(awv[i]);
                if (d == 0) ty = y > 0 ? y - 1 : 0;
                else if (d == 1) ty = y < H - 1 ? y + 1 : H - 1;
                else if (d == 2) tx = x > 0 ? x - 1 : 0;
                else tx = x < W - 1 ? x + 1 : W - 1;
            }
            sv[i] = (uint32_t)(y * W + x);
            tv[i] = (uint32_t)(ty * W + tx);
        }
#pragma unroll
        for (int i = 0; i < N; i++) gtv[i] = l.g[tv[i]];
#pragma unroll
        for (int i = 0; i < N; i++) {
            uint32_t nxt;
            if (tv[i] == sv[i]) { nxt = RW_NXT_SELF; commit |= 1u << i; }
            else {
                uint32_t occ = gtv[i] & RW_G_AGENT;
                // ref:829-844 a carrier may not enter a cell holding a shelf unless an agent there carries it
                if (rw_acarry(awv[i]) && (gtv[i] & RW_G_SHELF) && !(occ && rw_acarry(rw_agent(l, (int)occ - 1)))) {
                    av[i] = RW_ACT_NOOP; tv[i] = sv[i]; nxt = RW_NXT_SELF; commit |= 1u << i;
                } else {
                    nxt = occ ? occ - 1 : RW_NXT_SINK;
                }
            }
            rw_st(l, i) = sv[i] | (tv[i] << 16);
            rw_meta(l, i) = av[i] | (nxt << 3);
        }
        commit = rw_resolve_lane<C>(p, l, commit);
        // apply (ref:878-899); failed agents become NOOP (ref:871-876)
        uint32_t gsv[N], lift[N];
#pragma unroll
        for (int i = 0; i < N; i++) gsv[i] = l.g[sv[i]];
#pragma unroll
        for (int i = 0; i < N; i++) {
            lift[i] = 0;
            const uint32_t a = av[i];
            if (a == RW_ACT_FORWARD) {
                if (((commit >> i) & 1u) && tv[i] != sv[i]) {
                    moved |= 1u << i;
                    if (rw_acarry(awv[i])) { lift[i] = gsv[i] & (RW_G_SHELF | RW_G_REQ); l.g[sv[i]] = 0; }
                    else l.g[sv[i]] = (uint8_t)(gsv[i] & (RW_G_SHELF | RW_G_REQ));
                }
            } else if (a == RW_ACT_LEFT || a == RW_ACT_RIGHT) {  // Agent.req_direction ref:118-125
                uint32_t d = (uint32_t)rw_adir(awv[i]);
                uint32_t nd = ((a == RW_ACT_RIGHT ? 0x4Bu : 0x1Eu) >> (2 * d)) & 3u;
                awv[i] = (awv[i] & ~(3u << 16)) | (nd << 16);
                rw_set_agent(l, i, awv[i]);
            } else if (a == RW_ACT_TOGGLE) {  // ref:889-899
                if (!rw_acarry(awv[i])) {
                    if (gsv[i] & RW_G_SHELF) rw_set_agent(l, i, awv[i] | ((uint32_t)rw_shelf_at<C>(p, l, sv[i]) << 19));
                } else if (!(l.highway[sv[i]] & 1)) {
                    if (rw_adeliv(awv[i]) && p.reward_type == 2) rew_lo += 1ULL << (4 * i);
                    rw_set_agent(l, i, awv[i] & 0x3ffffu & ~(1u << 18));
                }
            }
        }
        if (moved) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                if (!((moved >> i) & 1u)) continue;
                const uint32_t t = tv[i];
                uint32_t ty = C::div_w(p, t), tx = t - ty * (uint32_t)W;
                rw_set_agent(l, i, (awv[i] & ~0xffffu) | tx | (ty << 8));
                int carry = rw_acarry(awv[i]);
                if (carry) {
                    l.g[t] = (uint8_t)((uint32_t)(i + 1) | lift[i]);
                    rw_set_shelf_cell(o, l, carry - 1, (int)t);
                } else {
                    l.g[t] = (uint8_t)((l.g[t] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i + 1));
                }
            }
        }
    } else {
    // ---- proposals (ref:825-846); actions were decoded into the scratch by rw_load_actions
    for (int i = 0; i < n; i++) {
        uint32_t aw = rw_agent(l, i);
        uint32_t pre = rw_meta(l, i);
        int a = (int)(pre & 7u);
        if (mb) l.tile[(o.w_msg + i) * RW_TILE] = (int32_t)(pre >> 8);
        int x = rw_ax(aw), y = rw_ay(aw), tx = x, ty = y;
        if (a == RW_ACT_FORWARD) {  // Agent.req_location ref:102-116, clamped at the walls
            int d = rw_adir(aw);
            if (d == 0) ty = y > 0 ? y - 1 : 0;
            else if (d == 1) ty = y < H - 1 ? y + 1 : H - 1;
            else if (d == 2) tx = x > 0 ? x - 1 : 0;
            else tx = x < W - 1 ? x + 1 : W - 1;
        }
        uint32_t s = (uint32_t)(y * W + x), t = (uint32_t)(ty * W + tx);
        uint32_t nxt;
        if (t == s) { nxt = RW_NXT_SELF; commit |= 1u << i; }
        else {
            uint32_t gt = l.g[t];
            uint32_t o = gt & RW_G_AGENT;
            // ref:829-844 a carrier may not enter a cell holding a shelf unless an agent there carries it
            if (rw_acarry(aw) && (gt & RW_G_SHELF) && !(o && rw_acarry(rw_agent(l, (int)o - 1)))) {
                a = RW_ACT_NOOP; t = s; nxt = RW_NXT_SELF; commit |= 1u << i;
            } else {
                nxt = o ? o - 1 : RW_NXT_SINK;
            }
        }
        rw_st(l, i) = s | (t << 16);
        rw_meta(l, i) = (uint32_t)a | (nxt << 3);
    }
    commit = rw_resolve_lane<C>(p, l, commit);
    // ---- apply (ref:878-899).  Failed agents become NOOP (ref:871-876).
    for (int i = 0; i < n; i++) {
        uint32_t m = rw_meta(l, i);
        uint32_t a = rw_m_act(m);
        if (a == RW_ACT_NOOP) continue;
        uint32_t aw = rw_agent(l, i);
        uint32_t st = rw_st(l, i);
        uint32_t s = st & 0xffffu, t = st >> 16;
        if (a == RW_ACT_FORWARD) {
            if (!((commit >> i) & 1u) || t == s) continue;
            moved |= 1u << i;
            uint32_t gsb = l.g[s];
            if (rw_acarry(aw)) {  // lift the carried shelf's bits; they are put down at t in the second pass
                rw_meta(l, i) = (m & 0xffffu) | ((gsb & (RW_G_SHELF | RW_G_REQ)) << 16);
                l.g[s] = 0;
            } else {
                l.g[s] = (uint8_t)(gsb & (RW_G_SHELF | RW_G_REQ));
            }
        } else if (a == RW_ACT_LEFT || a == RW_ACT_RIGHT) {  // Agent.req_direction ref:118-125
            uint32_t d = (uint32_t)rw_adir(aw);
            uint32_t nd = ((a == RW_ACT_RIGHT ? 0x4Bu : 0x1Eu) >> (2 * d)) & 3u;
            rw_set_agent(l, i, (aw & ~(3u << 16)) | (nd << 16));
        } else {  // TOGGLE_LOAD ref:889-899
            if (!rw_acarry(aw)) {
                if (l.g[s] & RW_G_SHELF) rw_set_agent(l, i, aw | ((uint32_t)rw_shelf_at<C>(p, l, s) << 19));
            } else if (!(l.highway[s] & 1)) {
                if (rw_adeliv(aw) && p.reward_type == 2) {
                    if (i < 16) rew_lo += 1ULL << (4 * i); else rew_hi += 1ULL << (4 * (i - 16));
                }
                rw_set_agent(l, i, aw & 0x3ffffu & ~(1u << 18));
            }
        }
    }
    for (uint32_t b = moved; b; b &= b - 1) {
        int i = rw_ffs32(b);
        uint32_t aw = rw_agent(l, i);
        uint32_t t = rw_st(l, i) >> 16;
        uint32_t ty = C::div_w(p, t), tx = t - ty * (uint32_t)W;
        rw_set_agent(l, i, (aw & ~0xffffu) | tx | (ty << 8));
        int carry = rw_acarry(aw);
        if (carry) {
            l.g[t] = (uint8_t)((uint32_t)(i + 1) | (rw_meta(l, i) >> 16));
            rw_set_shelf_cell(o, l, carry - 1, (int)t);
        } else {
            l.g[t] = (uint8_t)((l.g[t] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i + 1));
        }
    }
    }
    if (overlap && moved) rw_rebuild_grid_slow<C>(p, l);  // degenerate injected state: redo ref:901 literally
    const bool done = rw_finish_lane<C>(p, l, env, rew_lo, rew_hi, rew_global);
    for (int i = 0; i < n; i++) {
        uint32_t half = (uint32_t)(((i < 16 ? rew_lo >> (4 * i) : rew_hi >> (4 * (i - 16))) & 15ULL)) + 2u * (uint32_t)rew_global;
        reward[i] = 0.5f * (float)half;
    }
    return done;
}

// ------------------------------------------------------------------------------------------------
// Agent-split form of the step above for the fully static configurations with 2, 4, 6 or 8 agents and no message bits:
// BOTH warps of a tile's pair run one lane per world and warp `part` owns the agents [part * N/2, (part + 1) * N/2), so the
// per-agent work (proposals, apply) and the shelf marking of the grid halve on the critical path and the helper warp's issue
// slots, idle during the one-warp step, are used.  The phases are separated by pair barriers in the kernel (the host twin runs
// part 0 then part 1 of every phase, one valid interleaving); inside a phase the two parts touch disjoint words, bytes and
// half-words: own agent words, own scratch entries, the start / target cells of own agents (agents stand on distinct cells and
// committed targets are distinct).  The collision resolver and the deliveries stay with the primary.  Two words per world
// (`xw[0]`: helper -> primary, `xw[RW_TILE]`: primary -> helper) carry the masks across.  Worlds flagged RW_OVERLAP_BIT
// (injected degenerate states) never come here: the kernel steps them with the one-lane form.
template <class C>
struct RwPairRegs {
    static constexpr int NH = C::static_n >= 2 ? C::static_n / 2 : 1;
    uint32_t awv[NH];  // agent word
    uint32_t av[NH];   // action (bits 0-2) | lifted shelf / request bits of a moving carrier << 8
    uint32_t st[NH];   // start cell | target cell << 16
    uint32_t stay, moved, rew;  // bit j / nibble j: own agent j
};
template <class C>
RW_HD constexpr bool rw_pair_capable() {
    // (measured with 16 agents, 8 per part: 104.3 vs 99.8 us for rware-large-16ag -- the register arrays spill; not taken)
    return C::full_static && C::static_mb == 0 && C::static_n >= 2 && C::static_n <= 8 && (C::static_n & 1) == 0;
}

// phase 1 (both parts; needs no grid): own half of the shelf bits; own agents' words, actions and start / target cells
template <class C>
RW_HD void rw_pair_propose_a(const RwParams &p, const RwLane &l, int part, const int32_t *act, RwPairRegs<C> &r) {
    constexpr int NH = RwPairRegs<C>::NH;
    const int W = C::W(p), H = C::H(p), words = C::S(p) >> 1, half = (words + 1) >> 1, i0 = part * NH;
    if (part == 0) rw_grid_shelves<C>(p, l, 0, half, false);
    else rw_grid_shelves<C>(p, l, half, words, true);
    uint32_t bad = 0;
    r.stay = 0; r.moved = 0; r.rew = 0;
#pragma unroll
    for (int j = 0; j < NH; j++) {
        r.awv[j] = rw_agent(l, i0 + j);
        int a = act[j];
        if ((unsigned)a > 4u) { a = RW_ACT_NOOP; bad++; }
        r.av[j] = (uint32_t)a;
    }
#pragma unroll
    for (int j = 0; j < NH; j++) {
        int x = rw_ax(r.awv[j]), y = rw_ay(r.awv[j]), tx = x, ty = y;
        if (r.av[j] == RW_ACT_FORWARD) {  // Agent.req_location ref:102-116, clamped at the walls
            int d = rw_adir(r.awv[j]);
            if (d == 0) ty = y > 0 ? y - 1 : 0;
            else if (d == 1) ty = y < H - 1 ? y + 1 : H - 1;
            else if (d == 2) tx = x > 0 ? x - 1 : 0;
            else tx = x < W - 1 ? x + 1 : W - 1;
        }
        r.st[j] = (uint32_t)(y * W + x) | ((uint32_t)(ty * W + tx) << 16);
    }
    if (bad) {
#if defined(__CUDA_ARCH__)
        atomicAdd(p.err_count, (unsigned long long)bad);
#else
        *p.err_count += bad;
#endif
    }
}

// phase 3 (both parts; the grid is complete): classify own proposals (ref:825-846) into the scratch
template <class C>
RW_HD void rw_pair_propose_b(const RwParams &p, const RwLane &l, int part, RwPairRegs<C> &r) {
    constexpr int NH = RwPairRegs<C>::NH;
    const int i0 = part * NH;
    uint32_t gtv[NH];
#pragma unroll
    for (int j = 0; j < NH; j++) gtv[j] = l.g[r.st[j] >> 16];
#pragma unroll
    for (int j = 0; j < NH; j++) {
        const uint32_t s = r.st[j] & 0xffffu;
        uint32_t nxt;
        if ((r.st[j] >> 16) == s) { nxt = RW_NXT_SELF; r.stay |= 1u << j; }
        else {
            uint32_t occ = gtv[j] & RW_G_AGENT;
            // ref:829-844 a carrier may not enter a cell holding a shelf unless an agent there carries it
            if (rw_acarry(r.awv[j]) && (gtv[j] & RW_G_SHELF) && !(occ && rw_acarry(rw_agent(l, (int)occ - 1)))) {
                r.av[j] = RW_ACT_NOOP; r.st[j] = s | (s << 16); nxt = RW_NXT_SELF; r.stay |= 1u << j;
            } else {
                nxt = occ ? occ - 1 : RW_NXT_SINK;
            }
        }
        rw_st(l, i0 + j) = r.st[j];
        rw_meta(l, i0 + j) = r.av[j] | (nxt << 3);
    }
}

// phase 5 (both parts): first half of apply (ref:878-899) for own agents -- turns, load / unload, and the movers leave
// their start cells.  `commit`: the resolver's mask over all agents.
template <class C>
RW_HD void rw_pair_apply_a(const RwParams &p, const RwLane &l, int part, uint32_t commit, RwPairRegs<C> &r) {
    constexpr int NH = RwPairRegs<C>::NH;
    const int i0 = part * NH;
    uint32_t gsv[NH];
#pragma unroll
    for (int j = 0; j < NH; j++) gsv[j] = l.g[r.st[j] & 0xffffu];
#pragma unroll
    for (int j = 0; j < NH; j++) {
        const uint32_t a = r.av[j], s = r.st[j] & 0xffffu;
        if (a == RW_ACT_FORWARD) {
            if (((commit >> (i0 + j)) & 1u) && (r.st[j] >> 16) != s) {
                r.moved |= 1u << j;
                if (rw_acarry(r.awv[j])) { r.av[j] = a | ((gsv[j] & (RW_G_SHELF | RW_G_REQ)) << 8); l.g[s] = 0; }
                else l.g[s] = (uint8_t)(gsv[j] & (RW_G_SHELF | RW_G_REQ));
            }
        } else if (a == RW_ACT_LEFT || a == RW_ACT_RIGHT) {  // Agent.req_direction ref:118-125
            uint32_t d = (uint32_t)rw_adir(r.awv[j]);
            uint32_t nd = ((a == RW_ACT_RIGHT ? 0x4Bu : 0x1Eu) >> (2 * d)) & 3u;
            r.awv[j] = (r.awv[j] & ~(3u << 16)) | (nd << 16);
            rw_set_agent(l, i0 + j, r.awv[j]);
        } else if (a == RW_ACT_TOGGLE) {  // ref:889-899
            if (!rw_acarry(r.awv[j])) {
                if (gsv[j] & RW_G_SHELF) rw_set_agent(l, i0 + j, r.awv[j] | ((uint32_t)rw_shelf_at<C>(p, l, s) << 19));
            } else if (!(l.highway[s] & 1)) {
                if (rw_adeliv(r.awv[j]) && p.reward_type == 2) r.rew += 1u << (4 * j);
                rw_set_agent(l, i0 + j, r.awv[j] & 0x3ffffu & ~(1u << 18));
            }
        }
    }
}

// phase 6 (both parts; every mover has left its start cell): own movers arrive
template <class C>
RW_HD void rw_pair_apply_b(const RwParams &p, const RwLane &l, int part, const RwPairRegs<C> &r) {
    constexpr int NH = RwPairRegs<C>::NH;
    const RwOff o = rw_off<C>(p);
    const int W = C::W(p), i0 = part * NH;
    if (!r.moved) return;
#pragma unroll
    for (int j = 0; j < NH; j++) {
        if (!((r.moved >> j) & 1u)) continue;
        const uint32_t t = r.st[j] >> 16;
        uint32_t ty = C::div_w(p, t), tx = t - ty * (uint32_t)W;
        rw_set_agent(l, i0 + j, (r.awv[j] & ~0xffffu) | tx | (ty << 8));
        int carry = rw_acarry(r.awv[j]);
        if (carry) {
            l.g[t] = (uint8_t)((uint32_t)(i0 + j + 1) | (r.av[j] >> 8));
            rw_set_shelf_cell(o, l, carry - 1, (int)t);
        } else {
            l.g[t] = (uint8_t)((l.g[t] & (RW_G_SHELF | RW_G_REQ)) | (uint32_t)(i0 + j + 1));
        }
    }
}

// phase 7 (primary): deliveries, counters, rewards.  `rew_h`: the helper's reward nibbles.  Returns done.
template <class C>
RW_HD bool rw_pair_finish(const RwParams &p, const RwLane &l, const RwPairRegs<C> &r, uint32_t rew_h, float *reward, int64_t env) {
    constexpr int NH = RwPairRegs<C>::NH, N = 2 * NH;
    uint64_t rew_lo = (uint64_t)r.rew | ((uint64_t)rew_h << (4 * NH)), rew_hi = 0;
    int rew_global = 0;
    const bool done = rw_finish_lane<C>(p, l, env, rew_lo, rew_hi, rew_global);
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t half = (uint32_t)((rew_lo >> (4 * i)) & 15ULL) + 2u * (uint32_t)rew_global;
        reward[i] = 0.5f * (float)half;
    }
    return done;
}

// the one-lane step of a whole world out of line: worlds the agent-split form must not take (RW_OVERLAP_BIT), cold
template <class C>
RW_HD_NOINLINE bool rw_step_world_serial(const RwParams &p, const RwLane l, const int32_t *act, float *reward, int64_t env) {
    rw_build_grid<C>(p, l);
    rw_load_actions<C>(p, l, act);
    return rw_step_lane<C>(p, l, reward, env);
}

// ------------------------------------------------------------------------------------------------
// Flattened observation of agent i of the world in `l` (ref:598-674; layout SURVEY.md App. C) -> out[L].
// The row is produced in `nparts` slices so that two lanes can share one agent (finer staging chunks):
// slice 0 writes the 8 self features and the first ceil(K / nparts) window cells, slice 1 the rest.
template <class C, int nparts>
RW_HD void rw_obs_agent(const RwParams &p, const RwLane &l, int i, float *out, int part) {
    const RwOff o = rw_off<C>(p);
    uint32_t aw = rw_agent(l, i);
    const int x = rw_ax(aw), y = rw_ay(aw);
    const int W = C::W(p), H = C::H(p);
    const int r = C::r(p), mb = C::mb(p), win = 2 * r + 1, K = win * win, cellF = 7 + mb;
    if (part == 0) {
        const int d = rw_adir(aw);
        out[0] = p.norm_coords ? p.normx[x] : (float)x;  // ref:636-643
        out[1] = p.norm_coords ? p.normy[y] : (float)y;
        out[2] = rw_acarry(aw) ? 1.0f : 0.0f;
        out[3] = d == 0 ? 1.0f : 0.0f;                    // ref:644-646
        out[4] = d == 1 ? 1.0f : 0.0f;
        out[5] = d == 2 ? 1.0f : 0.0f;
        out[6] = d == 3 ? 1.0f : 0.0f;
        out[7] = (l.highway[y * W + x] & 1) ? 1.0f : 0.0f;  // ref:647
    }
    // Slice s covers window cells [s*Kp, min(K, (s+1)*Kp)).  `part` varies across lanes: with two slices everything
    // that depends on it is a select between the compile-time values of slice 0 and slice 1 (no divergence, no
    // integer division); with more slices the cell index is computed (division by the compile-time window width).
    const int Kp = (K + nparts - 1) / nparts;
    const bool hi = part != 0;
    const uint8_t *gc = l.g + (y * W + x);
    float *ob = out + 8 + (nparts <= 2 ? (hi ? Kp * cellF : 0) : part * Kp * cellF);
#pragma unroll
    for (int k = 0; k < Kp; k++) {                        // ref:628-629 window cells, rows outer, columns inner
        int dy, dx;
        bool live;
        if (nparts <= 2) {
            const int ca = k, cb = Kp + k;                // cell index in slice 0 / slice 1 (compile time when unrolled)
            const int dya = ca / win - r, dxa = ca % win - r;
            const int dyb = (cb < K ? cb : K - 1) / win - r, dxb = (cb < K ? cb : K - 1) % win - r;
            dy = hi ? dyb : dya;
            dx = hi ? dxb : dxa;
            live = hi ? (cb < K) : (ca < K);
        } else {
            const int c = part * Kp + k;
            const int cc = c < K ? c : K - 1;
            dy = cc / win - r;
            dx = cc - (cc / win) * win - r;
            live = c < K;
        }
        if (live) {
            uint32_t gb = 0;                              // zero padding outside the map (ref:606-626)
            if ((unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W) gb = gc[dy * W + dx];
            const uint32_t ida = gb & RW_G_AGENT;
            uint32_t od = 0, bits = 0;
            if (ida) {
                od = (uint32_t)rw_adir(rw_agent(l, (int)ida - 1));
                if (mb) bits = (uint32_t)l.tile[(o.w_msg + (int)ida - 1) * RW_TILE];
            }
            float *oc = ob + k * cellF;
            oc[0] = ida ? 1.0f : 0.0f;                    // ref:656-667: an empty cell reads as direction one-hot [1,0,0,0]
            oc[1] = od == 0 ? 1.0f : 0.0f;
            oc[2] = od == 1 ? 1.0f : 0.0f;
            oc[3] = od == 2 ? 1.0f : 0.0f;
            oc[4] = od == 3 ? 1.0f : 0.0f;
            for (int b = 0; b < mb; b++) oc[5 + b] = (float)((bits >> b) & 1u);
            oc[5 + mb] = (gb & RW_G_SHELF) ? 1.0f : 0.0f; // ref:668-673
            oc[6 + mb] = (gb & RW_G_REQ) ? 1.0f : 0.0f;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same row as rw_obs_agent, as BITS: every value of a flattened row is a 0/1 flag except the two coordinates, so the
// host-buffer transport (rw_step_host) can carry a row as L bits + 2 bytes instead of L floats.  Bit k of the row is
// "value k is non-zero" (the bits of the coordinates are left 0; the host ignores them); bits are OR-ed into `words`, a
// zeroed bit image of the whole tile in shared memory, at bit offset row * L.  x / y go to wide[2 * row], wide[2 * row + 1].
struct RwBitWriter {
    uint32_t *words;
    uint32_t w;
    uint64_t acc;
    int nacc;
};
RW_HD void rw_bw_or(uint32_t *addr, uint32_t v) {
#if defined(__CUDA_ARCH__)
    if (v) atomicOr(addr, v);  // rows of different lanes share the words at their boundaries
#else
    *addr |= v;
#endif
}
RW_HD void rw_bw_init(RwBitWriter &b, uint32_t *words, uint32_t bitpos) {
    b.words = words; b.w = bitpos >> 5; b.nacc = (int)(bitpos & 31u); b.acc = 0;
}
RW_HD void rw_bw_put(RwBitWriter &b, uint32_t value, int nbits) {  // nbits <= 32
    b.acc |= (uint64_t)value << b.nacc;
    b.nacc += nbits;
    if (b.nacc >= 32) {
        rw_bw_or(b.words + b.w, (uint32_t)b.acc);
        b.w++;
        b.acc >>= 32;
        b.nacc -= 32;
    }
}
RW_HD void rw_bw_flush(RwBitWriter &b) {
    if (b.nacc > 0) rw_bw_or(b.words + b.w, (uint32_t)b.acc);
}

template <class C>
RW_HD void rw_obs_agent_bits(const RwParams &p, const RwLane &l, int i, uint32_t row, uint32_t *words, uint8_t *wide, int part,
                             int nparts) {
    const RwOff o = rw_off<C>(p);
    const uint32_t aw = rw_agent(l, i);
    const int x = rw_ax(aw), y = rw_ay(aw);
    const int W = C::W(p), H = C::H(p);
    const int r = C::r(p), mb = C::mb(p), win = 2 * r + 1, K = win * win, cellF = 7 + mb;
    const int L = 8 + K * cellF;
    const int Kp = (K + nparts - 1) / nparts;
    RwBitWriter bw;
    if (part == 0) {
        wide[2 * row] = (uint8_t)x;
        wide[2 * row + 1] = (uint8_t)y;
        rw_bw_init(bw, words, row * (uint32_t)L);
        const uint32_t d = (uint32_t)rw_adir(aw);
        rw_bw_put(bw, (rw_acarry(aw) ? 4u : 0u) | (8u << d) | ((l.highway[y * W + x] & 1) ? 128u : 0u), 8);  // ref:636-647
    } else {
        rw_bw_init(bw, words, row * (uint32_t)L + 8u + (uint32_t)(part * Kp * cellF));
    }
    const uint8_t *gc = l.g + (y * W + x);
    int c = part * Kp;
    int dy = c / win - r, dx = c - (c / win) * win - r;
    for (int k = 0; k < Kp && c < K; k++, c++) {  // ref:628-629 window cells, rows outer, columns inner
        uint32_t gb = 0;                           // zero padding outside the map (ref:606-626)
        if ((unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W) gb = gc[dy * W + dx];
        const uint32_t ida = gb & RW_G_AGENT;
        uint32_t od = 0, bits = 0;
        if (ida) {
            od = (uint32_t)rw_adir(rw_agent(l, (int)ida - 1));
            if (mb) bits = (uint32_t)l.tile[(o.w_msg + (int)ida - 1) * RW_TILE];
        }
        // ref:656-673: has_agent, direction one-hot (an empty cell reads [1,0,0,0]), message bits, has_shelf, requested
        rw_bw_put(bw, (ida ? 1u : 0u) | (2u << od) | (bits << 5) | ((gb & RW_G_SHELF) ? (1u << (5 + mb)) : 0u) |
                          ((gb & RW_G_REQ) ? (1u << (6 + mb)) : 0u), cellF);
        if (++dx > r) { dx = -r; dy++; }
    }
    rw_bw_flush(bw);
}

// ------------------------------------------------------------------------------------------------
template <class C, int nparts>
RW_HD_NOINLINE void rw_obs_image_transposed(const RwParams &p, const RwLane l, int i, float *out, int part);

// Image observation of agent i (ref:527-596): out = [n_layers][K][K] float32, K = 2r+1, the window centred on the agent,
// rotated towards its heading when `img_directional` (np.rot90, ref:584-595); IMAGE_DICT appends 6 feature floats
// (one-hot direction, on_highway, carrying; ref:726-742).  Layers (ImageLayer values, ref:59-70): 0 SHELVES,
// 1 REQUESTS, 2 AGENTS, 5 GOALS, 6 ACCESSIBLE (0 outside the map: the reference pads with zeros, ref:573).
// Sliced over the window positions like rw_obs_agent.
template <class C, int nparts>
RW_HD void rw_obs_image_agent(const RwParams &p, const RwLane &l, int i, float *out, int part) {
    const uint32_t aw = rw_agent(l, i);
    const int x = rw_ax(aw), y = rw_ay(aw), d = rw_adir(aw);
    const int W = C::W(p), H = C::H(p), r = C::r(p), K = 2 * r + 1, KK = K * K;
    // k quarter turns counter-clockwise: DOWN 2, LEFT 3, RIGHT 1 (ref:586-594)
    const int rot = p.img_directional ? (d == 1 ? 2 : d == 2 ? 3 : d == 3 ? 1 : 0) : 0;
    const int Kp = (KK + nparts - 1) / nparts;
    const int o0 = part * Kp;
#pragma unroll
    for (int k = 0; k < Kp; k++) {
        const int o = o0 + k;
        if (o < KK) {
            const int oi = o / K, oj = o - oi * K;
            int si, sj;  // R = rot90(A, k): k=1 R[i][j] = A[j][K-1-i]; k=2 A[K-1-i][K-1-j]; k=3 A[K-1-j][i]
            if (rot == 0) { si = oi; sj = oj; }
            else if (rot == 1) { si = oj; sj = K - 1 - oi; }
            else if (rot == 2) { si = K - 1 - oi; sj = K - 1 - oj; }
            else { si = K - 1 - oj; sj = oi; }
            const int yy = y - r + si, xx = x - r + sj;
            const bool in = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
            // one word per window cell: occupancy byte | static table byte << 8, or only bit 16 outside the map (the
            // reference pads with zeros, ref:573); every layer is a mask test over it (p.img_mask, bit 31 inverts)
            uint32_t word = 1u << 16;
            if (in) word = (uint32_t)l.g[yy * W + xx] | ((uint32_t)l.highway[yy * W + xx] << 8);
#pragma unroll
            for (int li = 0; li < 8; li++)
                if (li < p.n_layers) {
                    const uint32_t m = p.img_mask[li];  // 0 for the transposed agent layers: zero here, filled in below
                    out[li * KK + o] = (((word & m & 0x1ffffu) != 0) != ((m >> 31) != 0)) ? 1.0f : 0.0f;
                }
        }
    }
    if (part == 0 && p.obs_kind == 3) {
        float *f = out + p.n_layers * KK;
        f[0] = d == 0 ? 1.0f : 0.0f;
        f[1] = d == 1 ? 1.0f : 0.0f;
        f[2] = d == 2 ? 1.0f : 0.0f;
        f[3] = d == 3 ? 1.0f : 0.0f;
        f[4] = (l.highway[y * W + x] & 1) ? 1.0f : 0.0f;
        f[5] = rw_acarry(aw) ? 1.0f : 0.0f;
    }
    if (p.img_transposed) rw_obs_image_transposed<C, nparts>(p, l, i, out, part);
}

// AGENT_DIRECTION / AGENT_LOAD layers of the image observation, a second pass taken only by configurations that ask for them:
// the reference writes layer[ag.x, ag.y] (sic, ref:552, 558), so window cell [yy][xx] shows the agent that stands at column yy,
// row xx of the map.
template <class C, int nparts>
RW_HD_NOINLINE void rw_obs_image_transposed(const RwParams &p, const RwLane l, int i, float *out, int part) {
    const uint32_t aw = rw_agent(l, i);
    const int x = rw_ax(aw), y = rw_ay(aw), d = rw_adir(aw);
    const int W = C::W(p), H = C::H(p), r = C::r(p), K = 2 * r + 1, KK = K * K;
    const int rot = p.img_directional ? (d == 1 ? 2 : d == 2 ? 3 : d == 3 ? 1 : 0) : 0;
    const int Kp = (KK + nparts - 1) / nparts;
    for (int o = part * Kp; o < (part + 1) * Kp && o < KK; o++) {
        const int oi = o / K, oj = o - oi * K;
        int si, sj;
        if (rot == 0) { si = oi; sj = oj; }
        else if (rot == 1) { si = oj; sj = K - 1 - oi; }
        else if (rot == 2) { si = K - 1 - oi; sj = K - 1 - oj; }
        else { si = K - 1 - oj; sj = oi; }
        const int yy = y - r + si, xx = x - r + sj;
        // inside the (H, W) layer, and the transposed position (row xx, column yy) inside the map
        const bool ok = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W && xx < H && yy < W;
        uint32_t ow = 0;
        bool have = false;
        if (ok) {
            const uint32_t id = (uint32_t)l.g[xx * W + yy] & RW_G_AGENT;
            if (id) { ow = rw_agent(l, (int)id - 1); have = true; }
        }
        for (int li = 0; li < p.n_layers; li++) {
            const int ly = p.img_layers[li];
            if (ly == 3) out[li * KK + o] = have ? (float)(rw_adir(ow) + 1) : 0.0f;
            else if (ly == 4) out[li * KK + o] = (have && rw_acarry(ow)) ? 1.0f : 0.0f;
        }
    }
    // where numpy raises IndexError for the transposed writes (an agent with x >= H or y >= W) the world is counted once
    // per observation pass (by the lane of its first agent) and the offending entries are simply absent
    if (i == 0 && part == 0) {
        int bad = 0;
        for (int a = 0; a < C::n(p); a++) {
            const uint32_t ow = rw_agent(l, a);
            if ((rw_ax(ow) >= H || rw_ay(ow) >= W) && ((p.img_transposed & 1) || rw_acarry(ow))) bad = 1;
        }
        if (bad) {
#if defined(__CUDA_ARCH__)
            atomicAdd(p.err_count, 1ULL);
#else
            *p.err_count += 1;
#endif
        }
    }
}

// dispatch on the configured observation type
template <class C, int nparts>
RW_HD void rw_obs_row(const RwParams &p, const RwLane &l, int i, float *out, int part) {
    if (p.obs_kind == 0) rw_obs_agent<C, nparts>(p, l, i, out, part);
    else rw_obs_image_agent<C, nparts>(p, l, i, out, part);
}
