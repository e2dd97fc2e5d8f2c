/* rware_oracle.c -- TEST INFRASTRUCTURE: CPU oracle for the RWARE step path (see rware_oracle.h).
 *
 * Formulation note.  The oracle deliberately follows the reference's *cell graph* formulation
 * (a DiGraph whose nodes are grid cells, exactly what rware/warehouse.py:823-876 hands to networkx) and
 * re-enacts what networkx 3.6.1 computes on it, including the container orders that decide ties.  The
 * CUDA product path uses a different, agent-centric formulation; agreement between
 * the two is therefore a meaningful check and not an identity.
 */
#include "rware_oracle.h"
#include "np_random.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_AGENTS 64
#define MAX_NODES (2 * MAX_AGENTS)
#define PYSET_CAP 1024

struct orc_env {
    orc_config cfg;
    uint8_t *highways;
    int32_t *goals_xy;
    int S;        /* number of shelves = non-highway cells, warehouse.py:771-778 */
    int L;        /* flattened obs length, warehouse.py:432-443 */
    /* agents (warehouse.py:82-93), index = id-1 */
    int32_t ax[MAX_AGENTS], ay[MAX_AGENTS], adir[MAX_AGENTS], acarry[MAX_AGENTS], adeliv[MAX_AGENTS];
    int32_t areq[MAX_AGENTS];
    int32_t *amsg; /* [n][msg_bits] */
    /* shelves (warehouse.py:128-133), index = id-1 */
    int32_t *sx, *sy;
    int32_t *queue; /* request_queue as shelf ids */
    int32_t *grid;  /* [2][H][W], layer 0 agents, layer 1 shelves (warehouse.py:11-14, 302) */
    int32_t cur_steps, cur_inactive;
    np_pcg64 rng;
    int rng_seeded;
    int32_t stat_ties, stat_setorder, stat_deliveries;
};

/* ---------------------------------------------------------------- layout (warehouse.py:294-326) */
int orc_layout_from_params(int shelf_columns, int shelf_rows, int column_height, int32_t *H, int32_t *W,
                           uint8_t *highways, int cap, int32_t *goals_xy) {
    if (shelf_columns % 2 != 1) return -1; /* warehouse.py:295 */
    int h = (column_height + 1) * shelf_rows + 2, w = (2 + 1) * shelf_columns + 1; /* 297-300 */
    if (h * w > cap) return -2;
    *H = h;
    *W = w;
    goals_xy[0] = w / 2 - 1; goals_xy[1] = h - 1; /* 303-306 */
    goals_xy[2] = w / 2;     goals_xy[3] = h - 1;
    for (int x = 0; x < w; x++)
        for (int y = 0; y < h; y++) { /* highway_func 310-322 */
            int vert = x % 3 == 0;
            int horiz = y % (column_height + 1) == 0;
            int delivery = y == h - 1;
            int queue = (y > h - (column_height + 3)) && (x == w / 2 - 1 || x == w / 2);
            highways[y * w + x] = (uint8_t)(vert || horiz || delivery || queue);
        }
    return 0;
}

/* ---------------------------------------------------------------- env lifecycle */
orc_env *orc_create(const orc_config *cfg) {
    if (cfg->n_agents < 1 || cfg->n_agents > MAX_AGENTS || cfg->H < 1 || cfg->W < 1) return NULL;
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->cfg = *cfg;
    int cells = cfg->H * cfg->W;
    e->highways = (uint8_t *)malloc((size_t)cells);
    memcpy(e->highways, cfg->highways, (size_t)cells);
    e->goals_xy = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cfg->n_goals);
    memcpy(e->goals_xy, cfg->goals_xy, sizeof(int32_t) * 2 * (size_t)cfg->n_goals);
    e->cfg.highways = e->highways;
    e->cfg.goals_xy = e->goals_xy;
    e->S = 0;
    for (int c = 0; c < cells; c++) e->S += !e->highways[c];
    /* warehouse.py:432-443: 4+4 self bits, per cell 1+4+msg agent bits and 2 shelf bits */
    int loc = (1 + 2 * cfg->sensor_range) * (1 + 2 * cfg->sensor_range);
    e->L = 8 + loc * (1 + 4 + cfg->msg_bits) + loc * 2;
    e->amsg = (int32_t *)calloc((size_t)cfg->n_agents * (size_t)(cfg->msg_bits > 0 ? cfg->msg_bits : 1), sizeof(int32_t));
    e->sx = (int32_t *)calloc((size_t)(e->S > 0 ? e->S : 1), sizeof(int32_t));
    e->sy = (int32_t *)calloc((size_t)(e->S > 0 ? e->S : 1), sizeof(int32_t));
    e->queue = (int32_t *)calloc((size_t)(cfg->request_queue_size > 0 ? cfg->request_queue_size : 1), sizeof(int32_t));
    e->grid = (int32_t *)calloc((size_t)2 * (size_t)cells, sizeof(int32_t));
    return e;
}

void orc_destroy(orc_env *e) {
    if (!e) return;
    free(e->highways); free(e->goals_xy); free(e->amsg); free(e->sx); free(e->sy); free(e->queue); free(e->grid);
    free(e);
}

int orc_n_shelves(const orc_env *e) { return e->S; }
int orc_obs_len(const orc_env *e) { return e->L; }

/* warehouse.py:749-755 */
static void recalc_grid(orc_env *e) {
    int H = e->cfg.H, W = e->cfg.W;
    memset(e->grid, 0, sizeof(int32_t) * 2 * (size_t)H * (size_t)W);
    for (int s = 0; s < e->S; s++) e->grid[H * W + e->sy[s] * W + e->sx[s]] = s + 1;
    for (int a = 0; a < e->cfg.n_agents; a++) e->grid[e->ay[a] * W + e->ax[a]] = a + 1;
}

static int in_queue(const orc_env *e, int shelf_id) { /* `shelf in self.request_queue` */
    for (int q = 0; q < e->cfg.request_queue_size; q++)
        if (e->queue[q] == shelf_id) return 1;
    return 0;
}

/* ---------------------------------------------------------------- observations (warehouse.py:598-674) */
static void make_obs_agent(const orc_env *e, int a, float *o) {
    const orc_config *c = &e->cfg;
    int H = c->H, W = c->W, r = c->sensor_range, mb = c->msg_bits;
    int idx = 0;
    /* 636-643: coordinates (optionally normalised in float64, stored as float32), carrying flag */
    if (c->normalised_coordinates) {
        o[idx++] = (float)((double)e->ax[a] / (double)(W - 1));
        o[idx++] = (float)((double)e->ay[a] / (double)(H - 1));
    } else {
        o[idx++] = (float)e->ax[a];
        o[idx++] = (float)e->ay[a];
    }
    o[idx++] = e->acarry[a] ? 1.0f : 0.0f;
    for (int d = 0; d < 4; d++) o[idx++] = (e->adir[a] == d) ? 1.0f : 0.0f; /* 644-646 */
    o[idx++] = e->highways[e->ay[a] * W + e->ax[a]] ? 1.0f : 0.0f;          /* 647 */
    /* 599-629: window rows min_y..max_y outer, columns inner; zero padding outside the map */
    for (int yy = e->ay[a] - r; yy <= e->ay[a] + r; yy++)
        for (int xx = e->ax[a] - r; xx <= e->ax[a] + r; xx++) {
            int id_agent = 0, id_shelf = 0;
            if (xx >= 0 && yy >= 0 && xx < W && yy < H) {
                id_agent = e->grid[yy * W + xx];
                id_shelf = e->grid[H * W + yy * W + xx];
            }
            if (id_agent == 0) { /* 656-660 */
                o[idx++] = 0.0f;
                o[idx++] = 1.0f; o[idx++] = 0.0f; o[idx++] = 0.0f; o[idx++] = 0.0f;
                for (int b = 0; b < mb; b++) o[idx++] = 0.0f; /* skip() over the zero-initialised vector */
            } else { /* 661-667 */
                o[idx++] = 1.0f;
                for (int d = 0; d < 4; d++) o[idx++] = (e->adir[id_agent - 1] == d) ? 1.0f : 0.0f;
                for (int b = 0; b < mb; b++) o[idx++] = (float)e->amsg[(id_agent - 1) * mb + b];
            }
            if (id_shelf == 0) { /* 668-669 */
                o[idx++] = 0.0f; o[idx++] = 0.0f;
            } else { /* 670-673 */
                o[idx++] = 1.0f;
                o[idx++] = in_queue(e, id_shelf) ? 1.0f : 0.0f;
            }
        }
}

void orc_obs(const orc_env *e, float *obs) {
    for (int a = 0; a < e->cfg.n_agents; a++) make_obs_agent(e, a, obs + (size_t)a * (size_t)e->L);
}

/* ---------------------------------------------------------------- image observations (warehouse.py:527-596, 726-742) */
int orc_obs_image(const orc_env *e, const int32_t *layers, int n_layers, int directional, float *out, float *features) {
    const orc_config *c = &e->cfg;
    const int H = c->H, W = c->W, r = c->sensor_range, K = 2 * r + 1, n = c->n_agents;
    const int PH = H + 2 * r, PW = W + 2 * r; /* np.pad(layer, sensor_range) 573 */
    float *glob = (float *)calloc((size_t)n_layers * PH * PW, sizeof(float));
    for (int li = 0; li < n_layers; li++) {
        float *L = glob + (size_t)li * PH * PW;
#define AT(row, col) L[((row) + r) * PW + (col) + r]
        switch (layers[li]) {
            case 0: /* SHELVES 533-536 */
                for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) AT(y, x) = e->grid[H * W + y * W + x] > 0 ? 1.0f : 0.0f;
                break;
            case 1: /* REQUESTS 538-541 */
                for (int q = 0; q < c->request_queue_size; q++) { int s = e->queue[q] - 1; AT(e->sy[s], e->sx[s]) = 1.0f; }
                break;
            case 2: /* AGENTS 543-546 */
                for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) AT(y, x) = e->grid[y * W + x] > 0 ? 1.0f : 0.0f;
                break;
            case 3: /* AGENT_DIRECTION 548-552: layer[ag.x, ag.y] (sic) */
                for (int a = 0; a < n; a++) {
                    if (e->ax[a] >= H || e->ay[a] >= W) { free(glob); return -1; }
                    AT(e->ax[a], e->ay[a]) = (float)(e->adir[a] + 1);
                }
                break;
            case 4: /* AGENT_LOAD 554-558: layer[ag.x, ag.y] (sic) */
                for (int a = 0; a < n; a++)
                    if (e->acarry[a]) {
                        if (e->ax[a] >= H || e->ay[a] >= W) { free(glob); return -1; }
                        AT(e->ax[a], e->ay[a]) = 1.0f;
                    }
                break;
            case 5: /* GOALS 560-563: goals are (x, y); layer[y, x] */
                for (int g = 0; g < c->n_goals; g++) AT(e->goals_xy[2 * g + 1], e->goals_xy[2 * g]) = 1.0f;
                break;
            case 6: /* ACCESSIBLE 565-568 */
                for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) AT(y, x) = 1.0f;
                for (int a = 0; a < n; a++) AT(e->ay[a], e->ax[a]) = 0.0f;
                break;
            default: free(glob); return -2;
        }
#undef AT
    }
    for (int a = 0; a < n; a++) {
        /* 578-582: rows agent.y .. agent.y+2r, columns agent.x .. agent.x+2r of the padded layers */
        for (int li = 0; li < n_layers; li++) {
            const float *L = glob + (size_t)li * PH * PW;
            float *o = out + ((size_t)a * n_layers + li) * K * K;
            int k = 0; /* np.rot90(obs, k, axes=(1, 2)), 584-595 */
            if (directional) k = e->adir[a] == ORC_DOWN ? 2 : e->adir[a] == ORC_DIR_LEFT ? 3 : e->adir[a] == ORC_DIR_RIGHT ? 1 : 0;
            for (int i = 0; i < K; i++)
                for (int j = 0; j < K; j++) {
                    int si, sj; /* R = rot90(A, k): k=1 R[i][j]=A[j][K-1-i]; k=2 A[K-1-i][K-1-j]; k=3 A[K-1-j][i] */
                    if (k == 0) { si = i; sj = j; }
                    else if (k == 1) { si = j; sj = K - 1 - i; }
                    else if (k == 2) { si = K - 1 - i; sj = K - 1 - j; }
                    else { si = K - 1 - j; sj = i; }
                    o[i * K + j] = L[(e->ay[a] + si) * PW + e->ax[a] + sj];
                }
        }
        if (features) { /* 727-738 */
            float *f = features + (size_t)a * 6;
            for (int d = 0; d < 4; d++) f[d] = e->adir[a] == d ? 1.0f : 0.0f;
            f[4] = e->highways[e->ay[a] * W + e->ax[a]] ? 1.0f : 0.0f;
            f[5] = e->acarry[a] ? 1.0f : 0.0f;
        }
    }
    free(glob);
    return 0;
}

/* ---------------------------------------------------------------- global image (warehouse.py:966-1040), before padding */
int orc_global_image(const orc_env *e, const int32_t *layers, int n_layers, float *out) {
    const orc_config *c = &e->cfg;
    const int H = c->H, W = c->W, n = c->n_agents;
    memset(out, 0, sizeof(float) * (size_t)n_layers * H * W);
    for (int li = 0; li < n_layers; li++) {
        float *L = out + (size_t)li * H * W;
        switch (layers[li]) {
            case 0: /* SHELVES 988-991: grid[_LAYER_SHELFS] > 0 */
                for (int k = 0; k < H * W; k++) L[k] = e->grid[H * W + k] > 0 ? 1.0f : 0.0f;
                break;
            case 1: /* REQUESTS 992-995 */
                for (int q = 0; q < c->request_queue_size; q++) { int s = e->queue[q] - 1; L[e->sy[s] * W + e->sx[s]] = 1.0f; }
                break;
            case 2: /* AGENTS 996-999 */
                for (int k = 0; k < H * W; k++) L[k] = e->grid[k] > 0 ? 1.0f : 0.0f;
                break;
            case 3: /* AGENT_DIRECTION 1000-1004: layer[ag.x, ag.y] (sic) */
                for (int a = 0; a < n; a++) {
                    if (e->ax[a] >= H || e->ay[a] >= W) return -1;
                    L[e->ax[a] * W + e->ay[a]] = (float)(e->adir[a] + 1);
                }
                break;
            case 4: /* AGENT_LOAD 1005-1009: layer[ag.x, ag.y] (sic) */
                for (int a = 0; a < n; a++)
                    if (e->acarry[a]) {
                        if (e->ax[a] >= H || e->ay[a] >= W) return -1;
                        L[e->ax[a] * W + e->ay[a]] = 1.0f;
                    }
                break;
            case 5: /* GOALS 1010-1013: goals hold (x, y); the loop names them (goal_y, goal_x) and writes layer[goal_x, goal_y] = [y][x] */
                for (int g = 0; g < c->n_goals; g++) L[c->goals_xy[2 * g + 1] * W + c->goals_xy[2 * g]] = 1.0f;
                break;
            case 6: /* ACCESSIBLE 1014-1017 */
                for (int k = 0; k < H * W; k++) L[k] = 1.0f;
                for (int a = 0; a < n; a++) L[e->ay[a] * W + e->ax[a]] = 0.0f;
                break;
            default: return -2; /* ValueError 1018-1019 */
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- seeding / reset (warehouse.py:757-802, 962-964) */
void orc_seed(orc_env *e, uint64_t seed) {
    np_seed(&e->rng, seed);
    e->rng_seeded = 1;
}

void orc_reset(orc_env *e, int64_t seed, float *obs) {
    const orc_config *c = &e->cfg;
    int H = c->H, W = c->W, n = c->n_agents;
    if (seed >= 0) orc_seed(e, (uint64_t)seed); /* 758-760 */
    if (!e->rng_seeded) orc_seed(e, 0);          /* (the reference would draw OS entropy here) */
    e->cur_inactive = 0; /* 764-765 */
    e->cur_steps = 0;
    int s = 0; /* 771-778: shelves on every non-highway cell, row-major, ids 1..S */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (!e->highways[y * W + x]) { e->sx[s] = x; e->sy[s] = y; s++; }
    int64_t locs[MAX_AGENTS], dirs[MAX_AGENTS];
    np_choice_noreplace(&e->rng, (int64_t)H * W, n, locs); /* 781-785 */
    np_integers(&e->rng, 4, n, dirs);                      /* 788 */
    for (int a = 0; a < n; a++) { /* 786, 789-792: unravel_index -> (y, x) */
        e->ay[a] = (int32_t)(locs[a] / W);
        e->ax[a] = (int32_t)(locs[a] % W);
        e->adir[a] = (int32_t)dirs[a];
        e->acarry[a] = 0;
        e->adeliv[a] = 0;
        e->areq[a] = ORC_NOOP;
        for (int b = 0; b < c->msg_bits; b++) e->amsg[a * c->msg_bits + b] = 0;
    }
    recalc_grid(e); /* 794 */
    if (c->request_queue_size > 0) { /* 796-800 */
        int64_t *pick = (int64_t *)malloc(sizeof(int64_t) * (size_t)c->request_queue_size);
        np_choice_noreplace(&e->rng, e->S, c->request_queue_size, pick);
        for (int q = 0; q < c->request_queue_size; q++) e->queue[q] = (int32_t)pick[q] + 1;
        free(pick);
    }
    if (obs) orc_obs(e, obs); /* 802 */
}

/* ---------------------------------------------------------------- CPython 3.12 set / tuple-hash emulation
 * Objects/tupleobject.c:tuplehash (xxHash-style) and Objects/setobject.c:set_add_entry, set_insert_clean,
 * set_table_resize.  Only insertion of distinct keys and iteration are needed. */
uint64_t orc_py_tuple_hash(int64_t x, int64_t y) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P5 = 2870177450012600261ULL;
    uint64_t acc = P5;
    uint64_t lanes[2] = {(uint64_t)x, (uint64_t)y}; /* hash(int v) == v for 0 <= v < 2^61-1 */
    for (int i = 0; i < 2; i++) {
        acc += lanes[i] * P2;
        acc = (acc << 31) | (acc >> 33);
        acc *= P1;
    }
    acc += 2ULL ^ (P5 ^ 3527539ULL);
    if (acc == (uint64_t)-1) return 1546275796ULL;
    return acc;
}

typedef struct {
    int mask, fill;
    int32_t key[PYSET_CAP]; /* -1 = unused */
    uint64_t hash[PYSET_CAP];
} pyset;

static void pyset_init(pyset *s) {
    s->mask = 7; /* PySet_MINSIZE 8 */
    s->fill = 0;
    for (int i = 0; i <= s->mask; i++) s->key[i] = -1;
}

static void pyset_insert_clean(int32_t *keys, uint64_t *hashes, int mask, int32_t key, uint64_t hash) {
    uint64_t perturb = hash;
    uint64_t i = hash & (uint64_t)mask;
    uint64_t slot;
    for (;;) {
        slot = i;
        if (keys[slot] < 0) goto found;
        if (i + 9 <= (uint64_t)mask) { /* LINEAR_PROBES = 9; the probe cursor advances, `i` does not */
            for (int j = 0; j < 9; j++) {
                slot++;
                if (keys[slot] < 0) goto found;
            }
        }
        perturb >>= 5; /* PERTURB_SHIFT */
        i = (i * 5 + 1 + perturb) & (uint64_t)mask;
    }
found:
    keys[slot] = key;
    hashes[slot] = hash;
}

static void pyset_add(pyset *s, int32_t key, uint64_t hash) {
    uint64_t mask = (uint64_t)s->mask;
    uint64_t perturb = hash;
    uint64_t i = hash & mask;
    uint64_t slot;
    for (;;) {
        uint64_t e = i;
        int probes = (i + 9 <= mask) ? 9 : 0;
        int found = 0;
        do {
            if (s->key[e] < 0) { slot = e; found = 1; break; }
            if (s->key[e] == key) return; /* already present */
            e++;
        } while (probes--);
        if (found) break;
        perturb >>= 5;
        i = (i * 5 + 1 + perturb) & mask;
    }
    s->key[slot] = key;
    s->hash[slot] = hash;
    s->fill++;
    if ((uint64_t)s->fill * 5 < mask * 3) return;
    /* set_table_resize(so, used*4): smallest power of two > minused */
    int minused = s->fill * 4;
    int newsize = 8;
    while (newsize <= minused) newsize <<= 1;
    static __thread int32_t nk[PYSET_CAP];
    static __thread uint64_t nh[PYSET_CAP];
    if (newsize > PYSET_CAP) abort();
    for (int k = 0; k < newsize; k++) nk[k] = -1;
    for (int k = 0; k <= s->mask; k++)
        if (s->key[k] >= 0) pyset_insert_clean(nk, nh, newsize - 1, s->key[k], s->hash[k]);
    memcpy(s->key, nk, sizeof(int32_t) * (size_t)newsize);
    memcpy(s->hash, nh, sizeof(uint64_t) * (size_t)newsize);
    s->mask = newsize - 1;
}

static int pyset_iter(const pyset *s, int32_t *out) {
    int n = 0;
    for (int k = 0; k <= s->mask; k++)
        if (s->key[k] >= 0) out[n++] = s->key[k];
    return n;
}

void orc_py_set_order(const int32_t *xy, int n, int32_t *out_index_order) {
    pyset s;
    pyset_init(&s);
    for (int i = 0; i < n; i++) pyset_add(&s, i, orc_py_tuple_hash(xy[2 * i], xy[2 * i + 1]));
    pyset_iter(&s, out_index_order);
}

/* ---------------------------------------------------------------- collision resolver (warehouse.py:821-876 + networkx) */
typedef struct {
    int n_nodes;
    int cell[MAX_NODES];           /* node -> cell, in G insertion order */
    int succ[MAX_NODES];           /* node -> node or -1 (out-degree <= 1) */
    int npred[MAX_NODES];
    int pred[MAX_NODES][MAX_AGENTS + 1]; /* G._pred[v] in edge insertion order */
} cellgraph;

static int cg_node(cellgraph *g, int *node_of_cell, int cell) {
    if (node_of_cell[cell] >= 0) return node_of_cell[cell];
    int v = g->n_nodes++;
    g->cell[v] = cell;
    g->succ[v] = -1;
    g->npred[v] = 0;
    node_of_cell[cell] = v;
    return v;
}

static void cg_add_edge(cellgraph *g, int *node_of_cell, int start_cell, int target_cell) {
    /* DiGraph.add_edge(u, v): u is inserted before v */
    int u = cg_node(g, node_of_cell, start_cell);
    int v = cg_node(g, node_of_cell, target_cell);
    if (g->succ[u] == v) return;
    g->succ[u] = v;
    g->pred[v][g->npred[v]++] = u;
}

/* returns committed[agent] */
static void resolve(orc_env *e, const int *start_cell, const int *target_cell, uint8_t *committed) {
    const orc_config *c = &e->cfg;
    int n = c->n_agents, W = c->W;
    static __thread int node_of_cell_buf[1 << 16];
    int *node_of_cell = node_of_cell_buf;
    cellgraph g;
    g.n_nodes = 0;
    for (int a = 0; a < n; a++) { node_of_cell[start_cell[a]] = -1; node_of_cell[target_cell[a]] = -1; }
    for (int a = 0; a < n; a++) cg_add_edge(&g, node_of_cell, start_cell[a], target_cell[a]); /* 825-846 */
    memset(committed, 0, (size_t)n);

    uint8_t seen_global[MAX_NODES];
    memset(seen_global, 0, sizeof(seen_global));
    int n_seen_global = 0;
    /* nx.weakly_connected_components (NX/algorithms/components/weakly_connected.py:61-66) */
    for (int src = 0; src < g.n_nodes; src++) {
        if (seen_global[src]) continue;
        /* _plain_bfs (weakly_connected.py:169-196): `seen` is a Python set; its insertion order matters later */
        int remaining = g.n_nodes - n_seen_global;
        pyset seen;
        pyset_init(&seen);
        uint8_t in_seen[MAX_NODES];
        memset(in_seen, 0, sizeof(in_seen));
        int level[MAX_NODES], nlevel = 0, next[MAX_NODES], nnext = 0, n_in = 0;
#define SEEN_ADD(v) do { in_seen[v] = 1; n_in++; \
        pyset_add(&seen, (v), orc_py_tuple_hash(g.cell[v] % W, g.cell[v] / W)); } while (0)
        SEEN_ADD(src);
        next[nnext++] = src;
        int done_bfs = 0;
        while (nnext && !done_bfs) {
            memcpy(level, next, sizeof(int) * (size_t)nnext);
            nlevel = nnext;
            nnext = 0;
            for (int li = 0; li < nlevel && !done_bfs; li++) {
                int v = level[li];
                int w = g.succ[v];
                if (w >= 0 && !in_seen[w]) { SEEN_ADD(w); next[nnext++] = w; }
                for (int k = 0; k < g.npred[v]; k++) {
                    w = g.pred[v][k];
                    if (!in_seen[w]) { SEEN_ADD(w); next[nnext++] = w; }
                }
                if (n_in == remaining) done_bfs = 1;
            }
        }
#undef SEEN_ADD
        int32_t order1[MAX_NODES];
        int csize = pyset_iter(&seen, order1); /* iteration order of the component set */
        for (int k = 0; k < csize; k++) { seen_global[order1[k]] = 1; }
        n_seen_global += csize;

        /* G.subgraph(c).copy()  (warehouse.py:848; NX/classes/graph.py:1869, 1671-1679;
           NX/classes/filters.py:67-68; NX/classes/coreviews.py:293-300, 342-349) */
        int32_t comp[MAX_NODES];
        if (2 * csize < g.n_nodes) {
            /* iterate show_nodes.nodes = set(nbunch_iter(c)): a second set filled in order1 */
            pyset s2;
            pyset_init(&s2);
            for (int k = 0; k < csize; k++)
                pyset_add(&s2, order1[k], orc_py_tuple_hash(g.cell[order1[k]] % W, g.cell[order1[k]] / W));
            pyset_iter(&s2, comp);
            e->stat_setorder++;
        } else {
            int k = 0;
            for (int v = 0; v < g.n_nodes; v++)
                if (in_seen[v]) comp[k++] = v;
        }
        /* edges of the copy are added for u in node order => comp._pred[v] order */
        int cnpred[MAX_NODES];
        int cpred[MAX_NODES][MAX_AGENTS + 1];
        for (int k = 0; k < csize; k++) cnpred[comp[k]] = 0;
        for (int k = 0; k < csize; k++) {
            int u = comp[k], v = g.succ[u];
            if (v >= 0) cpred[v][cnpred[v]++] = u;
        }
        /* nx.find_cycle(comp) (warehouse.py:854): out-degree <= 1 => at most one cycle per component and
           every walk along successors reaches it (or the sink) */
        int stamp[MAX_NODES];
        for (int k = 0; k < csize; k++) stamp[comp[k]] = -1;
        int v = comp[0], steps = 0;
        while (v >= 0 && stamp[v] < 0) { stamp[v] = steps++; v = g.succ[v]; }
        if (v >= 0) {
            int cyc_len = steps - stamp[v];
            if (cyc_len == 2) continue; /* 855-858: head-on swap, nobody in the component commits */
            int u = v;                  /* 859-863 */
            do {
                int agent_id = e->grid[g.cell[u]];
                if (agent_id > 0) committed[agent_id - 1] = 1;
                u = g.succ[u];
            } while (u != v);
            continue;
        }
        /* nx.dag_longest_path(comp) (warehouse.py:865; NX/algorithms/dag.py:1067-1101) with
           topological_sort = concatenated topological_generations (dag.py:218-235, 308-309) */
        int indeg[MAX_NODES], topo[MAX_NODES], ntopo = 0;
        int gen[MAX_NODES], ngen = 0, ngen_next;
        int gen_next[MAX_NODES];
        for (int k = 0; k < csize; k++) {
            indeg[comp[k]] = cnpred[comp[k]];
            if (indeg[comp[k]] == 0) gen[ngen++] = comp[k];
        }
        while (ngen) {
            ngen_next = 0;
            for (int k = 0; k < ngen; k++) {
                int node = gen[k];
                topo[ntopo++] = node;
                int child = g.succ[node];
                if (child >= 0 && --indeg[child] == 0) gen_next[ngen_next++] = child;
            }
            memcpy(gen, gen_next, sizeof(int) * (size_t)ngen_next);
            ngen = ngen_next;
        }
        int dist_len[MAX_NODES], dist_from[MAX_NODES];
        uint8_t dist_tie[MAX_NODES];
        for (int k = 0; k < ntopo; k++) {
            int node = topo[k];
            int best = -1, best_u = node, tie = 0;
            for (int p = 0; p < cnpred[node]; p++) { /* max(us, key=x[0]) keeps the FIRST maximum (dag.py:1089) */
                int u = cpred[node][p];
                int d = dist_len[u] + 1;
                if (d > best) { best = d; best_u = u; tie = 0; }
                else if (d == best) tie = 1;
            }
            if (cnpred[node] == 0) { dist_len[node] = 0; dist_from[node] = node; dist_tie[node] = 0; }
            else { dist_len[node] = best; dist_from[node] = best_u; dist_tie[node] = (uint8_t)tie; }
        }
        int end = topo[0]; /* v = max(dist, key=dist[x][0]): first maximum in insertion (= topological) order */
        for (int k = 1; k < ntopo; k++)
            if (dist_len[topo[k]] > dist_len[end]) end = topo[k];
        int u = -1;
        v = end;
        while (u != v) { /* 866-869 */
            int agent_id = e->grid[g.cell[v]];
            if (agent_id > 0) committed[agent_id - 1] = 1;
            if (dist_tie[v]) e->stat_ties++;
            u = v;
            v = dist_from[v];
        }
    }
}

/* ---------------------------------------------------------------- step (warehouse.py:804-946) */
int orc_step(orc_env *e, const int32_t *actions, float *obs, double *reward, uint8_t *done) {
    const orc_config *c = &e->cfg;
    int n = c->n_agents, H = c->H, W = c->W, mb = c->msg_bits;
    e->stat_ties = e->stat_setorder = e->stat_deliveries = 0;
    /* 809-814: decode */
    for (int a = 0; a < n; a++) {
        int act = actions[a * (1 + mb)];
        if (act < 0 || act > 4) return -1; /* Action(action) raises ValueError */
        e->areq[a] = act;
        for (int b = 0; b < mb; b++) e->amsg[a * mb + b] = actions[a * (1 + mb) + 1 + b];
    }
    /* 825-846: proposals, with the carrier-into-standing-shelf cancel rule */
    int start_cell[MAX_AGENTS], target_cell[MAX_AGENTS];
    for (int a = 0; a < n; a++) {
        int x = e->ax[a], y = e->ay[a], tx = x, ty = y;
        if (e->areq[a] == ORC_FORWARD) { /* Agent.req_location, warehouse.py:102-116 */
            switch (e->adir[a]) {
                case ORC_UP: ty = y - 1 > 0 ? y - 1 : 0; break;
                case ORC_DOWN: ty = y + 1 < H - 1 ? y + 1 : H - 1; break;
                case ORC_DIR_LEFT: tx = x - 1 > 0 ? x - 1 : 0; break;
                case ORC_DIR_RIGHT: tx = x + 1 < W - 1 ? x + 1 : W - 1; break;
                default: return -3;
            }
        }
        start_cell[a] = y * W + x;
        int tcell = ty * W + tx;
        int other = e->grid[tcell];
        if (e->acarry[a] && tcell != start_cell[a] && e->grid[H * W + tcell] &&
            !(other && e->acarry[other - 1])) { /* 829-844 */
            e->areq[a] = ORC_NOOP;
            tcell = start_cell[a];
        }
        target_cell[a] = tcell;
    }
    uint8_t committed[MAX_AGENTS];
    resolve(e, start_cell, target_cell, committed); /* 848-869 */
    for (int a = 0; a < n; a++) /* 871-876 */
        if (!committed[a]) {
            if (e->areq[a] != ORC_FORWARD) return -2; /* the reference asserts */
            e->areq[a] = ORC_NOOP;
        }
    for (int a = 0; a < n; a++) reward[a] = 0.0; /* 878 */
    static const int wrap[4] = {ORC_UP, ORC_DIR_RIGHT, ORC_DOWN, ORC_DIR_LEFT}; /* 119 */
    for (int a = 0; a < n; a++) { /* 880-899 */
        int act = e->areq[a];
        if (act == ORC_FORWARD) {
            e->ax[a] = target_cell[a] % W; /* req_location again: same clamp as above */
            e->ay[a] = target_cell[a] / W;
            if (e->acarry[a]) { e->sx[e->acarry[a] - 1] = e->ax[a]; e->sy[e->acarry[a] - 1] = e->ay[a]; }
        } else if (act == ORC_LEFT || act == ORC_RIGHT) { /* req_direction 118-125 */
            int k = 0;
            while (wrap[k] != e->adir[a]) k++;
            e->adir[a] = wrap[(k + (act == ORC_RIGHT ? 1 : 3)) % 4];
        } else if (act == ORC_TOGGLE_LOAD && !e->acarry[a]) {
            int shelf_id = e->grid[H * W + e->ay[a] * W + e->ax[a]]; /* pre-move grid, 890 */
            if (shelf_id) e->acarry[a] = shelf_id;
        } else if (act == ORC_TOGGLE_LOAD && e->acarry[a]) {
            if (!e->highways[e->ay[a] * W + e->ax[a]]) { /* 894-899 */
                e->acarry[a] = 0;
                if (e->adeliv[a] && c->reward_type == ORC_TWO_STAGE) reward[a] += 0.5;
                e->adeliv[a] = 0;
            }
        }
    }
    recalc_grid(e); /* 901 */
    int shelf_delivered = 0;
    for (int gi = 0; gi < c->n_goals; gi++) { /* 904-927 */
        int gx = e->goals_xy[2 * gi], gy = e->goals_xy[2 * gi + 1];
        int shelf_id = e->grid[H * W + gy * W + gx];
        if (!shelf_id) continue;
        if (!in_queue(e, shelf_id)) continue;
        shelf_delivered = 1;
        e->stat_deliveries++;
        /* 915-917: candidates in ascending shelf id, uniform draw, in-place replacement */
        int n_cand = 0;
        for (int s = 1; s <= e->S; s++) n_cand += !in_queue(e, s);
        if (n_cand > 0) {
            int pick = (int)np_bounded32(&e->rng, (uint32_t)(n_cand - 1));
            int new_id = 0;
            for (int s = 1; s <= e->S; s++)
                if (!in_queue(e, s) && pick-- == 0) { new_id = s; break; }
            for (int q = 0; q < c->request_queue_size; q++)
                if (e->queue[q] == shelf_id) { e->queue[q] = new_id; break; }
        } /* else: the reference raises ValueError (empty population); documented divergence */
        int agent_id = e->grid[gy * W + gx];
        int ai = agent_id > 0 ? agent_id - 1 : n - 1; /* python's rewards[-1] if nobody stands there */
        if (c->reward_type == ORC_GLOBAL) {
            for (int a = 0; a < n; a++) reward[a] += 1.0;
        } else if (c->reward_type == ORC_INDIVIDUAL) {
            reward[ai] += 1.0;
        } else {
            e->adeliv[ai] = 1;
            reward[ai] += 0.5;
        }
    }
    if (shelf_delivered) e->cur_inactive = 0; else e->cur_inactive++; /* 929-933 */
    e->cur_steps++;
    *done = (uint8_t)((c->max_inactivity_steps && e->cur_inactive >= c->max_inactivity_steps) ||
                      (c->max_steps && e->cur_steps >= c->max_steps)); /* 935-941 */
    if (obs) orc_obs(e, obs); /* 944 */
    return 0;
}

/* ---------------------------------------------------------------- state access */
void orc_get_state(const orc_env *e, int32_t *ax, int32_t *ay, int32_t *adir, int32_t *acarry, int32_t *adeliv,
                   int32_t *amsg, int32_t *sx, int32_t *sy, int32_t *queue, int32_t *counters, uint32_t *rng) {
    size_t n = (size_t)e->cfg.n_agents;
    if (ax) memcpy(ax, e->ax, 4 * n);
    if (ay) memcpy(ay, e->ay, 4 * n);
    if (adir) memcpy(adir, e->adir, 4 * n);
    if (acarry) memcpy(acarry, e->acarry, 4 * n);
    if (adeliv) memcpy(adeliv, e->adeliv, 4 * n);
    if (amsg && e->cfg.msg_bits) memcpy(amsg, e->amsg, 4 * n * (size_t)e->cfg.msg_bits);
    if (sx) memcpy(sx, e->sx, 4 * (size_t)e->S);
    if (sy) memcpy(sy, e->sy, 4 * (size_t)e->S);
    if (queue) memcpy(queue, e->queue, 4 * (size_t)e->cfg.request_queue_size);
    if (counters) { counters[0] = e->cur_steps; counters[1] = e->cur_inactive; }
    if (rng) np_export(&e->rng, rng);
}

void orc_set_state(orc_env *e, const int32_t *ax, const int32_t *ay, const int32_t *adir, const int32_t *acarry,
                   const int32_t *adeliv, const int32_t *amsg, const int32_t *sx, const int32_t *sy,
                   const int32_t *queue, const int32_t *counters, const uint32_t *rng) {
    size_t n = (size_t)e->cfg.n_agents;
    if (ax) memcpy(e->ax, ax, 4 * n);
    if (ay) memcpy(e->ay, ay, 4 * n);
    if (adir) memcpy(e->adir, adir, 4 * n);
    if (acarry) memcpy(e->acarry, acarry, 4 * n);
    if (adeliv) memcpy(e->adeliv, adeliv, 4 * n);
    if (amsg && e->cfg.msg_bits) memcpy(e->amsg, amsg, 4 * n * (size_t)e->cfg.msg_bits);
    if (sx) memcpy(e->sx, sx, 4 * (size_t)e->S);
    if (sy) memcpy(e->sy, sy, 4 * (size_t)e->S);
    if (queue) memcpy(e->queue, queue, 4 * (size_t)e->cfg.request_queue_size);
    if (counters) { e->cur_steps = counters[0]; e->cur_inactive = counters[1]; }
    if (rng) { np_import(&e->rng, rng); e->rng_seeded = 1; }
    recalc_grid(e); /* what the reference's tests do by hand: env._recalc_grid() */
}

void orc_get_grid(const orc_env *e, int32_t *grid) {
    memcpy(grid, e->grid, sizeof(int32_t) * 2 * (size_t)e->cfg.H * (size_t)e->cfg.W);
}

void orc_last_step_stats(const orc_env *e, int32_t *n_ties, int32_t *n_set_order_comps, int32_t *n_deliveries) {
    if (n_ties) *n_ties = e->stat_ties;
    if (n_set_order_comps) *n_set_order_comps = e->stat_setorder;
    if (n_deliveries) *n_deliveries = e->stat_deliveries;
}

/* ---------------------------------------------------------------- multi-threaded random rollout (CPU baseline) */
typedef struct {
    const orc_config *cfg;
    int env_lo, env_hi, n_steps;
    uint64_t seed;
    int64_t steps_done;
    double reward_sum;
} rollout_job;

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static void *rollout_worker(void *arg) {
    rollout_job *job = (rollout_job *)arg;
    const orc_config *c = job->cfg;
    int n = c->n_agents, mb = c->msg_bits;
    int32_t *actions = (int32_t *)calloc((size_t)n * (size_t)(1 + mb), sizeof(int32_t));
    double *reward = (double *)calloc((size_t)n, sizeof(double));
    for (int ei = job->env_lo; ei < job->env_hi; ei++) {
        orc_env *e = orc_create(c);
        float *obs = (float *)malloc(sizeof(float) * (size_t)n * (size_t)e->L);
        uint64_t as = job->seed * 0x9e3779b97f4a7c15ULL + (uint64_t)ei;
        orc_reset(e, (int64_t)(job->seed + (uint64_t)ei), obs);
        for (int t = 0; t < job->n_steps; t++) {
            for (int a = 0; a < n; a++) {
                actions[a * (1 + mb)] = (int32_t)(((splitmix64(&as) >> 32) * 5ULL) >> 32);
                for (int b = 0; b < mb; b++) actions[a * (1 + mb) + 1 + b] = (int32_t)(splitmix64(&as) & 1);
            }
            uint8_t done;
            orc_step(e, actions, obs, reward, &done);
            for (int a = 0; a < n; a++) job->reward_sum += reward[a];
            job->steps_done++;
            if (done) orc_reset(e, -1, obs);
        }
        free(obs);
        orc_destroy(e);
    }
    free(actions);
    free(reward);
    return NULL;
}

int64_t orc_rollout_random(const orc_config *cfg, int n_envs, int n_steps, uint64_t seed, int n_threads,
                           double *reward_sum) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_envs) n_threads = n_envs;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    rollout_job *jobs = (rollout_job *)calloc((size_t)n_threads, sizeof(rollout_job));
    for (int t = 0; t < n_threads; t++) {
        jobs[t].cfg = cfg;
        jobs[t].env_lo = (int)((int64_t)n_envs * t / n_threads);
        jobs[t].env_hi = (int)((int64_t)n_envs * (t + 1) / n_threads);
        jobs[t].n_steps = n_steps;
        jobs[t].seed = seed;
        pthread_create(&th[t], NULL, rollout_worker, &jobs[t]);
    }
    int64_t total = 0;
    double rs = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        total += jobs[t].steps_done;
        rs += jobs[t].reward_sum;
    }
    if (reward_sum) *reward_sum = rs;
    free(th);
    free(jobs);
    return total;
}

/* ---------------------------------------------------------------- batched oracle (test infrastructure)
 * E independent oracle worlds advanced in lock step with the product's batched step, so that the full-size parity tests can
 * compare EVERY world of a BASELINE-sized batch (65 536 worlds) instead of a sample.  Same per-world functions as above; the
 * worlds are split over `n_threads` pthreads per call. */
struct orc_batch {
    orc_config cfg;
    int64_t E;
    int n_threads;
    orc_env **env;
};

typedef struct {
    orc_batch *b;
    int64_t lo, hi;
    int mode;           /* 0 reset, 1 step */
    const uint64_t *seeds;
    const int32_t *actions;
    int auto_reset;
    float *obs, *reward;
    uint8_t *done;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    orc_batch *b = j->b;
    const int n = b->cfg.n_agents, A = 1 + b->cfg.msg_bits;
    double rew[MAX_AGENTS];
    for (int64_t e = j->lo; e < j->hi; e++) {
        orc_env *w = b->env[e];
        float *obs = j->obs ? j->obs + (size_t)e * n * w->L : NULL;
        float *scratch = NULL;
        if (!obs) obs = scratch = (float *)malloc(sizeof(float) * (size_t)n * (size_t)w->L);
        if (j->mode == 0) {
            orc_reset(w, j->seeds ? (int64_t)j->seeds[e] : -1, obs);
        } else {
            uint8_t d = 0;
            orc_step(w, j->actions + (size_t)e * n * A, obs, rew, &d);
            if (j->reward) for (int a = 0; a < n; a++) j->reward[(size_t)e * n + a] = (float)rew[a];
            if (j->done) j->done[e] = d;
            if (d && j->auto_reset) orc_reset(w, -1, obs); /* continues the world's generator, as the kernel's auto-reset */
        }
        free(scratch);
    }
    return NULL;
}

static void batch_run(orc_batch *b, batch_job proto) {
    int T = b->n_threads < 1 ? 1 : b->n_threads;
    if (T > b->E) T = (int)b->E;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * (size_t)T);
    for (int t = 0; t < T; t++) {
        jobs[t] = proto;
        jobs[t].lo = b->E * t / T;
        jobs[t].hi = b->E * (t + 1) / T;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

orc_batch *orc_batch_create(const orc_config *cfg, int64_t n_envs, int n_threads) {
    orc_batch *b = (orc_batch *)calloc(1, sizeof(orc_batch));
    b->cfg = *cfg;
    b->E = n_envs;
    b->n_threads = n_threads;
    b->env = (orc_env **)calloc((size_t)n_envs, sizeof(orc_env *));
    for (int64_t e = 0; e < n_envs; e++) b->env[e] = orc_create(cfg);
    return b;
}

void orc_batch_destroy(orc_batch *b) {
    if (!b) return;
    for (int64_t e = 0; e < b->E; e++) orc_destroy(b->env[e]);
    free(b->env);
    free(b);
}

void orc_batch_reset(orc_batch *b, const uint64_t *seeds, float *obs) {
    batch_job j;
    memset(&j, 0, sizeof(j));
    j.b = b; j.mode = 0; j.seeds = seeds; j.obs = obs;
    batch_run(b, j);
}

void orc_batch_step(orc_batch *b, const int32_t *actions, int auto_reset, float *obs, float *reward, uint8_t *done) {
    batch_job j;
    memset(&j, 0, sizeof(j));
    j.b = b; j.mode = 1; j.actions = actions; j.auto_reset = auto_reset; j.obs = obs; j.reward = reward; j.done = done;
    batch_run(b, j);
}

/* Lock-step random-action rollout of the whole batch: what a vector environment built on the reference does -- every step
 * advances ALL worlds once and leaves the stacked observations / rewards / done flags of the batch in [E][n][L] / [E][n] / [E]
 * arrays (written in full every step).  Threads are created once and meet at a barrier after every step; actions are generated
 * per world from a counter-based generator.  Returns env-steps done. */
typedef struct {
    orc_batch *b;
    int64_t lo, hi;
    int n_steps;
    uint64_t seed;
    float *obs, *reward;
    uint8_t *done;
    pthread_barrier_t *bar;
    double reward_sum;
} lockstep_job;

static void *lockstep_worker(void *arg) {
    lockstep_job *j = (lockstep_job *)arg;
    orc_batch *b = j->b;
    const int n = b->cfg.n_agents, mb = b->cfg.msg_bits, A = 1 + mb;
    int32_t actions[MAX_AGENTS * 17];
    double rew[MAX_AGENTS];
    for (int t = 0; t < j->n_steps; t++) {
        for (int64_t e = j->lo; e < j->hi; e++) {
            orc_env *w = b->env[e];
            uint64_t as = (j->seed + (uint64_t)t) * 0x9e3779b97f4a7c15ULL + (uint64_t)e * 0xbf58476d1ce4e5b9ULL;
            for (int a = 0; a < n; a++) {
                actions[a * A] = (int32_t)(((splitmix64(&as) >> 32) * 5ULL) >> 32);
                for (int k = 0; k < mb; k++) actions[a * A + 1 + k] = (int32_t)(splitmix64(&as) & 1);
            }
            float *obs = j->obs + (size_t)e * n * w->L;
            uint8_t d = 0;
            orc_step(w, actions, obs, rew, &d);
            for (int a = 0; a < n; a++) { j->reward[(size_t)e * n + a] = (float)rew[a]; j->reward_sum += rew[a]; }
            j->done[e] = d;
            if (d) orc_reset(w, -1, obs);
        }
        pthread_barrier_wait(j->bar);
    }
    return NULL;
}

int64_t orc_batch_rollout_lockstep(orc_batch *b, int n_steps, uint64_t seed, float *obs, float *reward, uint8_t *done,
                                   double *reward_sum) {
    int T = b->n_threads < 1 ? 1 : b->n_threads;
    if (T > b->E) T = (int)b->E;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)T);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    lockstep_job *jobs = (lockstep_job *)calloc((size_t)T, sizeof(lockstep_job));
    for (int t = 0; t < T; t++) {
        jobs[t].b = b; jobs[t].lo = b->E * t / T; jobs[t].hi = b->E * (t + 1) / T; jobs[t].n_steps = n_steps; jobs[t].seed = seed;
        jobs[t].obs = obs; jobs[t].reward = reward; jobs[t].done = done; jobs[t].bar = &bar;
        pthread_create(&th[t], NULL, lockstep_worker, &jobs[t]);
    }
    double rs = 0;
    for (int t = 0; t < T; t++) { pthread_join(th[t], NULL); rs += jobs[t].reward_sum; }
    if (reward_sum) *reward_sum = rs;
    pthread_barrier_destroy(&bar);
    free(th);
    free(jobs);
    return (int64_t)n_steps * b->E;
}
