/* rware_oracle.h -- TEST INFRASTRUCTURE: CPU oracle for the RWARE step path.
 *
 * This is a plain-C restatement of the reference algorithm (semitable/robotic-warehouse @ 96fbc64e,
 * rware/warehouse.py) for ONE warehouse world, written to be read side by side with the reference:
 * every function cites the reference file:line it follows.  It exists only to check the CUDA product
 * path (tests/, __graft_entry__.smoke(), and bench.py's cpu_baseline / --impl reference legs).  The
 * product package never imports, links or calls anything in oracle/.
 *
 * Parity status: PINNED.  The oracle is checked (tests/test_oracle_golden.py) against golden
 * trajectories generated in the build container from the *unmodified* reference
 * (tests/golden/make_golden.py), against the reference's own behavioural tests restated as fixtures,
 * against the SHA-256 known-answer vectors of SURVEY.md App. E, and lock-step against the unmodified
 * reference in the build container (oracle/refharness/diff_reference.py: 9 configurations x 30 000 steps
 * and 4 800 random configurations x 300 steps, every field and the full PCG64 state, 0 mismatches; logs
 * next to the script).  The collision resolver lives in a
 * third-party dependency of the reference that is not vendored (networkx, unpinned in setup.py:28,
 * 3.6.1 in this image) and its tie order depends on CPython 3.12 set iteration order; both are
 * restated here from their published sources and pinned by the same golden trajectories.
 */
#ifndef RWARE_ORACLE_H
#define RWARE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_NOOP = 0, ORC_FORWARD = 1, ORC_LEFT = 2, ORC_RIGHT = 3, ORC_TOGGLE_LOAD = 4 }; /* warehouse.py:31-36 */
enum { ORC_UP = 0, ORC_DOWN = 1, ORC_DIR_LEFT = 2, ORC_DIR_RIGHT = 3 };                    /* warehouse.py:39-43 */
enum { ORC_GLOBAL = 0, ORC_INDIVIDUAL = 1, ORC_TWO_STAGE = 2 };                           /* warehouse.py:46-49 */

typedef struct orc_config {
    int32_t H, W;                 /* grid_size = (H, W), warehouse.py:297-300 */
    int32_t n_agents, msg_bits, sensor_range, request_queue_size;
    int32_t max_inactivity_steps; /* 0 == None */
    int32_t max_steps;            /* 0 == None */
    int32_t reward_type;
    int32_t normalised_coordinates;
    int32_t n_goals;
    const uint8_t *highways;      /* [H*W] row-major [y][x], warehouse.py:308-326 */
    const int32_t *goals_xy;      /* [n_goals][2] = (x, y), warehouse.py:303-306 / 344-345 */
} orc_config;

typedef struct orc_env orc_env;

/* warehouse.py:294-326 (_make_layout_from_params): fills highways[H*W], goals_xy[4]; returns 0 */
int orc_layout_from_params(int shelf_columns, int shelf_rows, int column_height, int32_t *H, int32_t *W,
                           uint8_t *highways /* cap (>=H*W) */, int cap, int32_t *goals_xy /*[4]*/);

orc_env *orc_create(const orc_config *cfg); /* copies highways/goals */
void orc_destroy(orc_env *e);
int orc_n_shelves(const orc_env *e);
int orc_obs_len(const orc_env *e); /* warehouse.py:432-443 */

/* warehouse.py:962-964 seed(); warehouse.py:757-760 reset(seed=...) re-seeds, reset() continues */
void orc_seed(orc_env *e, uint64_t seed);
/* warehouse.py:757-802.  seed < 0: keep the current generator stream.  obs: [n][L] float32 or NULL */
void orc_reset(orc_env *e, int64_t seed, float *obs);
/* warehouse.py:804-946.  actions: [n][1+msg_bits]; obs [n][L] (may be NULL); reward [n] float64; done */
int orc_step(orc_env *e, const int32_t *actions, float *obs, double *reward, uint8_t *done);
/* warehouse.py:722-744 -> 598-674 for every agent */
void orc_obs(const orc_env *e, float *obs);

/* warehouse.py:527-596 (_make_img_obs) for every agent: out [n][n_layers][K][K] float32, K = 2*sensor_range+1.
 * layers: ImageLayer values (warehouse.py:59-70) 0 SHELVES 1 REQUESTS 2 AGENTS 3 AGENT_DIRECTION 4 AGENT_LOAD 5 GOALS
 * 6 ACCESSIBLE.  Layers 3 and 4 are written with transposed indices in the reference (552, 558: layer[ag.x, ag.y]) and
 * make it raise IndexError as soon as an agent stands at y >= W or x >= H; the oracle restates them literally and returns
 * -1 in exactly that case.  directional: rot90 towards the agent's heading (584-595).
 * features (IMAGE_DICT, warehouse.py:726-742): [n][6] = one-hot direction, on_highway, carrying; may be NULL. */
int orc_obs_image(const orc_env *e, const int32_t *layers, int n_layers, int directional, float *out, float *features);

/* warehouse.py:966-1040 (get_global_image) before the optional padding: out [n_layers][H][W] float32.  Same layer ids and
 * the same literal transposed writes for layers 3 and 4 (1004, 1009); returns -1 where the reference raises IndexError. */
int orc_global_image(const orc_env *e, const int32_t *layers, int n_layers, float *out);

/* plain-array state access (what the reference's tests poke through env.agents[i].x etc.).
 * agent arrays [n]; carry = shelf id (1-based) or 0; msg [n][msg_bits]; shelf arrays [S]; queue [R] ids;
 * counters[2] = {cur_steps, cur_inactive_steps}; rng[10] = np_export() words.  NULL pointers are skipped. */
void orc_get_state(const orc_env *e, int32_t *ax, int32_t *ay, int32_t *adir, int32_t *acarry, int32_t *adeliv,
                   int32_t *amsg, int32_t *sx, int32_t *sy, int32_t *queue, int32_t *counters, uint32_t *rng);
void orc_set_state(orc_env *e, const int32_t *ax, const int32_t *ay, const int32_t *adir, const int32_t *acarry,
                   const int32_t *adeliv, const int32_t *amsg, const int32_t *sx, const int32_t *sy,
                   const int32_t *queue, const int32_t *counters, const uint32_t *rng);
/* copy of self.grid [2][H][W] int32 (warehouse.py:749-755) */
void orc_get_grid(const orc_env *e, int32_t *grid);

/* diagnostics of the last orc_step: number of longest-path selections that were ties, and whether any
 * component was resolved in CPython-set order (SURVEY.md App. B regimes) */
void orc_last_step_stats(const orc_env *e, int32_t *n_ties, int32_t *n_set_order_comps, int32_t *n_deliveries);

/* exposed for unit tests of the CPython emulation */
uint64_t orc_py_tuple_hash(int64_t x, int64_t y);
/* iteration order of set(iterable) for `n` distinct (x,y) tuples inserted in the given order */
void orc_py_set_order(const int32_t *xy, int n, int32_t *out_index_order);

/* multi-threaded batch driver for the CPU baseline: steps `n_envs` independent envs `n_steps` times with
 * uniform random actions (own xorshift stream per env), resetting on done.  Returns total env-steps. */
int64_t orc_rollout_random(const orc_config *cfg, int n_envs, int n_steps, uint64_t seed, int n_threads,
                           double *reward_sum);

/* Batched form (test infrastructure): E independent worlds stepped in lock step on n_threads pthreads, so that full-size parity
 * tests can compare every world of a batch.  seeds [E] or NULL (continue the generators); obs [E][n][L] float32 or NULL;
 * reward [E][n] float32; done [E].  With auto_reset a world that is done is reset (continuing its generator) and obs holds the
 * first observation of the new episode, as the product's in-kernel auto-reset does. */
typedef struct orc_batch orc_batch;
orc_batch *orc_batch_create(const orc_config *cfg, int64_t n_envs, int n_threads);
void orc_batch_destroy(orc_batch *b);
void orc_batch_reset(orc_batch *b, const uint64_t *seeds, float *obs);
void orc_batch_step(orc_batch *b, const int32_t *actions, int auto_reset, float *obs, float *reward, uint8_t *done);
/* Lock-step random-action rollout of the whole batch (a vector environment on the CPU): every step advances all worlds once and
 * rewrites the stacked obs [E][n][L] / reward [E][n] / done [E] arrays; reset on done.  Returns env-steps done. */
int64_t orc_batch_rollout_lockstep(orc_batch *b, int n_steps, uint64_t seed, float *obs, float *reward, uint8_t *done,
                                   double *reward_sum);

#ifdef __cplusplus
}
#endif
#endif
