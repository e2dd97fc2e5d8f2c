/* rware_b200.h -- C ABI of the B200-native batched RWARE step path.
 *
 * This is the drop-in boundary.  In the reference the step path sits behind the Gymnasium `Env` protocol
 * of `rware.warehouse:Warehouse` (registered at rware/__init__.py:24-39).  The reference has no native
 * code and no FFI; the binding a maintainer adds is a ctypes stub that makes
 * `Warehouse.reset/step/_make_obs` forward to the entry points below.  Each entry point names the
 * reference interface it replaces (file:line in /root/reference/rware/warehouse.py unless noted).
 *
 * Conventions
 *   - every function returns 0 on success or a negative RW_E* code; rw_last_error() gives text;
 *   - no exceptions, no torch types, plain pointers and sizes only;
 *   - all device entry points are stream-ordered and asynchronous on `stream` (a cudaStream_t passed as
 *     void*, NULL = the legacy default stream); the *_host entry points synchronise before returning;
 *   - the caller owns every I/O buffer; the library owns the world state (a packed struct-of-arrays in HBM);
 *   - one handle is bound to one GPU; a handle is not thread-safe, different handles are independent;
 *   - `E` = num_envs, `n` = n_agents, `L` = rw_info.obs_len, `S` = rw_info.n_shelves,
 *     `R` = request_queue_size, `A` = 1 + msg_bits (ints per agent action).
 */
#ifndef RWARE_B200_H
#define RWARE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RW_ABI_VERSION 2

enum {
    RW_OK = 0,
    RW_EINVAL = -1,   /* bad argument / unsupported configuration */
    RW_ECUDA = -2,    /* a CUDA runtime call failed */
    RW_ENOMEM = -3,
    RW_ESTATE = -4    /* handle used in the wrong state (e.g. step before reset) */
};

/* Action / Direction / RewardType values are the reference's (warehouse.py:31-36, 39-43, 46-49). */
enum { RW_NOOP = 0, RW_FORWARD = 1, RW_LEFT = 2, RW_RIGHT = 3, RW_TOGGLE_LOAD = 4 };
enum { RW_UP = 0, RW_DOWN = 1, RW_DIR_LEFT = 2, RW_DIR_RIGHT = 3 };
enum { RW_REWARD_GLOBAL = 0, RW_REWARD_INDIVIDUAL = 1, RW_REWARD_TWO_STAGE = 2 };
/* ObservationType (warehouse.py:52-56) and ImageLayer (warehouse.py:59-70) values */
enum { RW_OBS_DICT = 0, RW_OBS_FLATTENED = 1, RW_OBS_IMAGE = 2, RW_OBS_IMAGE_DICT = 3 };
enum { RW_LAYER_SHELVES = 0, RW_LAYER_REQUESTS = 1, RW_LAYER_AGENTS = 2, RW_LAYER_AGENT_DIRECTION = 3, RW_LAYER_AGENT_LOAD = 4,
       RW_LAYER_GOALS = 5, RW_LAYER_ACCESSIBLE = 6 };

/* Constructor kwargs of Warehouse.__init__ (warehouse.py:146-170) after layout construction
 * (warehouse.py:294-350, done by the host-side facade): all pointers are HOST pointers, copied by rw_create. */
typedef struct rw_config {
    int32_t abi_version;          /* RW_ABI_VERSION */
    int32_t grid_h, grid_w;       /* grid_size, warehouse.py:297-300 / 337 */
    int32_t n_agents;             /* 1..32 */
    int32_t msg_bits;             /* 0..16 */
    int32_t sensor_range;         /* 0..8 */
    int32_t request_queue_size;   /* 0..n_shelves (equal to n_shelves: the reference raises on the first delivery) */
    int32_t max_inactivity_steps; /* 0 == None */
    int32_t max_steps;            /* 0 == None */
    int32_t reward_type;          /* RW_REWARD_* */
    int32_t normalised_coordinates;
    int32_t n_goals;              /* 1..16 */
    int32_t observation_type;     /* RW_OBS_FLATTENED (default path; RW_OBS_DICT is served as FLATTENED and unflattened by the
                                     host facade), RW_OBS_IMAGE or RW_OBS_IMAGE_DICT (warehouse.py:527-596, 726-742) */
    int32_t image_directional;    /* image_observation_directional (warehouse.py:584-595) */
    int32_t n_image_layers;       /* 0..7 */
    int32_t image_layers[8];      /* RW_LAYER_*; AGENT_DIRECTION / AGENT_LOAD are written as the reference writes them, with
                                     transposed indices (warehouse.py:552, 558); worlds where numpy raises IndexError for that
                                     are counted in rw_error_count once per observation pass */
    const uint8_t *highways;      /* [grid_h*grid_w], row-major [y][x], 1 = highway (warehouse.py:308-326) */
    const int32_t *goals_xy;      /* [n_goals][2] = (x, y) in the reference's order (warehouse.py:303-306, 344-345) */
} rw_config;

typedef struct rw_info {
    int64_t num_envs;
    int32_t n_agents, n_shelves, obs_len, act_width; /* act_width = 1 + msg_bits */
    int32_t request_queue_size, grid_h, grid_w, device;
    int32_t state_words_per_env;  /* packed int32 words per world in the HBM tile layout */
    int32_t envs_per_tile;        /* 32 */
    int32_t smem_bytes_per_warp, warps_per_cta;
    int64_t algorithmic_bytes_step; /* SURVEY.md section 8(d) B_step for this config */
    int64_t algorithmic_bytes_obs;  /* SURVEY.md section 8(d) B_obs */
    int64_t packed_bytes_step;      /* bytes this layout really moves per env-step (obs+2*state+act+rew+done) */
    int64_t packed_bytes_obs;
} rw_info;

/* Plain-array view of the world state: what the reference keeps in env.agents[i].{x,y,dir,carrying_shelf,
 * has_delivered,message}, env.shelfs[j].{x,y}, env.request_queue, env._cur_steps, env._cur_inactive_steps and
 * env.np_random (warehouse.py:82-93, 128-133, 249-250, 262-265).  All DEVICE pointers, int32 unless noted;
 * NULL members are skipped.  Shelf ids are 1-based (row-major rank of the home cell, warehouse.py:771-778);
 * carry == 0 means "not carrying". */
typedef struct rw_state_view {
    int32_t *agent_x, *agent_y, *agent_dir, *agent_carry, *agent_has_delivered; /* [E][n] */
    int32_t *agent_msg;   /* [E][n][msg_bits] */
    int32_t *shelf_x, *shelf_y; /* [E][S] */
    int32_t *queue;       /* [E][R] shelf ids */
    int32_t *cur_steps, *cur_inactive; /* [E] */
    uint32_t *rng;        /* [E][10]: numpy PCG64 state: state[4] lo..hi, inc[4], has_uint32, uinteger */
} rw_state_view;

typedef struct rw_handle rw_handle;

/* Warehouse.__init__ (warehouse.py:146-292) for `num_envs` independent worlds on CUDA device `device`. */
int rw_create(const rw_config *cfg, int64_t num_envs, int device, rw_handle **out);
int rw_destroy(rw_handle *h);
int rw_get_info(const rw_handle *h, rw_info *out);
const char *rw_last_error(const rw_handle *h); /* h may be NULL: error of the last failed rw_create */

/* Warehouse.seed(seed) / reset(seed=...) re-seeding (warehouse.py:962-964, 758-760): world e gets
 * Generator(PCG64(SeedSequence(seeds_host[e]))).  seeds_host: HOST [E].  mask_host: HOST [E] or NULL (all). */
int rw_seed(rw_handle *h, const uint64_t *seeds_host, const uint8_t *mask_host, void *stream);

/* Warehouse.reset() (warehouse.py:757-802) for the worlds with mask_dev[e] != 0 (NULL = all), continuing each
 * world's generator stream (call rw_seed first for reset(seed=...)).  obs_dev: [E][n][L] float32 or NULL;
 * all worlds' observations are (re)written. */
int rw_reset(rw_handle *h, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* Warehouse.step(actions) (warehouse.py:804-946): one fused kernel.  actions_dev [E][n][A] int32;
 * obs_dev [E][n][L] float32 (may be NULL: skip the observation gather; for the image types a row is the [C][K][K] image
 * followed, for IMAGE_DICT, by the 6 feature floats, L = rw_info.obs_len); reward_dev [E][n] float32;
 * done_dev [E] uint8 (`terminated`; `truncated` is always False, warehouse.py:942). */
int rw_step(rw_handle *h, const int32_t *actions_dev, float *obs_dev, float *reward_dev, uint8_t *done_dev,
            void *stream);

/* tuple(_make_obs(agent) for agent in agents) (warehouse.py:722-744 -> 598-674): the observation gather alone. */
int rw_obs(rw_handle *h, float *obs_dev, void *stream);

/* Warehouse.get_global_image(image_layers, recompute=True, pad_to_shape) (warehouse.py:966-1040) for every world:
 * out_dev [E][C'][H'][W'] float32, where (C', H', W') = pad_to_shape, or (n_layers, H, W) when pad_to_shape is NULL; padding is
 * the reference's (pad_before = diff / 2, constant 0; 1021-1039).  layers (host pointer): ImageLayer values 0 SHELVES 1 REQUESTS
 * 2 AGENTS 3 AGENT_DIRECTION 4 AGENT_LOAD 5 GOALS 6 ACCESSIBLE.  Layers 3 and 4 are written as the reference writes them,
 * transposed (`layer[ag.x, ag.y]`, 1004, 1009); where the reference raises IndexError for it (an agent with x >= H or y >= W)
 * the entry is skipped and the world is counted in rw_error_count.  rw_global_image_shape resolves (C', H', W'). */
int rw_global_image_shape(const rw_handle *h, int32_t n_layers, const int32_t *pad_to_shape /*[3] or NULL*/, int32_t *shape_out /*[3]*/);
int rw_global_image(rw_handle *h, const int32_t *layers, int32_t n_layers, const int32_t *pad_to_shape /*[3] or NULL*/,
                    float *out_dev, void *stream);

/* Warehouse.render() with render_mode="rgb_array" (warehouse.py:948-956 -> rendering.py:121-139) for the worlds
 * [first_env, first_env + n_envs): out_dev [n_envs][height][width][3] uint8 with (height, width) = (2 + 31 H, 1 + 31 W) as the
 * reference Viewer sizes its window (rendering.py:93-98).  Geometry and colours follow rendering.py:140-333 (grid, goal cells,
 * shelves, requested shelves, agents, loaded agents, heading line); the "G" glyph on goals is not drawn.  Pixel coverage is
 * stated in csrc/rw_views.cuh; no OpenGL frame could be recorded where this was written, so this view is not pixel-pinned. */
int rw_render_shape(const rw_handle *h, int32_t *height, int32_t *width);
int rw_render_rgb(rw_handle *h, int64_t first_env, int64_t n_envs, uint8_t *out_dev, void *stream);

/* When on, a world whose step returned done is reset inside the same kernel (as `reset()` without a seed would,
 * continuing its generator) and the returned observation is the first one of the new episode.  Default off
 * (the reference never auto-resets). */
int rw_set_auto_reset(rw_handle *h, int enabled);

/* State injection / extraction: what the reference's tests do through env.agents[i].x = ..., env._recalc_grid()
 * (tests/test_movement.py:14-47).  Also the checkpoint/restore hook. */
int rw_get_state(rw_handle *h, const rw_state_view *dev, void *stream);
int rw_set_state(rw_handle *h, const rw_state_view *dev, void *stream);

/* Count of invalid action values seen so far (the reference raises ValueError from Action(a), warehouse.py:811-814;
 * the batched kernel treats them as NOOP and counts them), plus the worlds whose transposed image layers (local or global) would have
 * raised IndexError (see rw_global_image).  Synchronises `stream`. */
int rw_error_count(rw_handle *h, int64_t *out, void *stream);

/* End-to-end convenience with HOST buffers (pinned memory recommended, see rw_host_alloc): copies the actions
 * to the device, runs rw_step, copies obs/reward/done back and synchronises.  Any output pointer may be NULL. */
int rw_step_host(rw_handle *h, const int32_t *actions_host, float *obs_host, float *reward_host, uint8_t *done_host);
int rw_reset_host(rw_handle *h, const uint64_t *seeds_host, float *obs_host);
/* Byte image of an observation tensor: out_dev[i] = (uint8_t)obs_dev[i] for the E*n*L values of obs_dev (a tensor written by
 * rw_step / rw_reset / rw_obs).  Exact, because every observation value is an integer in 0..255 unless normalised_coordinates
 * is set (then RW_EINVAL): a 4x smaller observation for consumers that embed or cast on the device.  (new; the reference's
 * observations are float32 arrays, warehouse.py:17-28) */
int rw_obs_bytes(rw_handle *h, const float *obs_dev, uint8_t *out_dev, void *stream);

/* Transport of rw_step_host.  The caller receives float32 rows in host memory (74 MB per step of 65 536 tiny-4ag worlds); a
 * plain device-to-host copy of them is PCIe bound.  Unless normalised_coordinates is set every value is a small integer, and
 * for batches of >= 2^20 values the rows travel compressed and a pool of host threads expands them into obs_host with
 * streaming stores: one BIT per value plus one byte per coordinate where everything else is a 0/1 flag (FLATTENED rows; image
 * rows without the AGENT_DIRECTION layer; obs_host must be 64-byte aligned), one BYTE per value otherwise.  The expansion is
 * bound by the host's DRAM write bandwidth, so the GPU's copy engine delivers a share of the rows as plain float32 at the same
 * time; the share follows the rates the call measures for the two, i.e. a host whose cores or memory bandwidth are taken by
 * other tenants drifts towards the plain copy by itself.  The buffers the caller gets back are bit-identical in every mode.
 *   rw_set_host_transport(h, mode)   0 = plain float32 copy, 1 = compressed (default), 2 = compressed, bytes only
 *   rw_set_host_dma_share(h, share)  fix the copy engine's share of the rows (0..1), negative = follow the measured rates
 * Environment: RWARE_B200_HOST_PACK=0|1|2, RWARE_B200_HOST_DMA_SHARE, RWARE_B200_HOST_THREADS, RWARE_B200_HOST_CHUNKS,
 * RWARE_B200_HOST_NUMA=0 (no placement), RWARE_B200_HOST_TRACE=1 (per-call timeline on stderr). */
int rw_set_host_transport(rw_handle *h, int mode);
int rw_set_host_dma_share(rw_handle *h, double share);
int rw_get_host_dma_share(const rw_handle *h, double *share);
/* The host half of the transport alone, with the same thread pool; each returns the number of threads used.  They need no GPU
 * and are exposed so that the expansion can be tested, and the host's write bandwidth measured, where there is none.
 *   rw_host_widen_u8      dst[i] = (float)src[i]
 *   rw_host_expand_bits   dst[row * row_len + k] = (float)wide[row * n_wide + k] for k < n_wide, else bit (row * row_len + k) of
 *                         `bits` (little-endian bytes) as 0.0f / 1.0f
 *   rw_host_stream_probe  streams `n` floats of a constant into dst `repeats` times; *gb_per_s = the best pass: what this
 *                         host's memory system takes from the pool when nothing has to be decoded (the transport's roofline) */
int rw_host_widen_u8(const uint8_t *src, float *dst, int64_t n);
int rw_host_expand_bits(const uint8_t *bits, const uint8_t *wide, float *dst, int64_t n_rows, int32_t row_len, int32_t n_wide);
int rw_host_stream_probe(float *dst, int64_t n, int32_t repeats, double *gb_per_s);
/* Page-locked host memory.  rw_host_alloc_near places the pages on the NUMA node the handle's GPU is attached to (first touch
 * by a thread confined to that node, then cudaHostRegister), so that neither the copy engine's writes nor the expansion cross
 * the socket interconnect; without NUMA information it is rw_host_alloc (= cudaHostAlloc).  rw_host_free takes either.
 * rw_host_numa_node / rw_gpu_numa_node report where a buffer / the GPU sits (-1 = unknown). */
int rw_host_alloc(void **out, int64_t bytes);
int rw_host_alloc_near(rw_handle *h, void **out, int64_t bytes);
int rw_host_free(void *p);
int rw_host_numa_node(const void *p);
int rw_gpu_numa_node(const rw_handle *h);

/* ---- Fused push all-gather (one process per GPU, NVLink / NVSwitch peers).  SURVEY.md section 8(e): the single optional
 * exchange of the path is an all-gather of obs / reward / done.  Instead of a separate collective, the step kernel's TMA
 * bulk stores write every observation chunk straight into this rank's slot of EVERY rank's gathered tensor.
 *   rw_gather_create    allocates this rank's gathered buffers (one cudaMalloc holding TWO slots of: obs [W*E][n][L] f32,
 *                       reward [W*E][n] f32, done [W*E] u8 for W = world_size ranks of E worlds each), returns the device
 *                       pointers of slot 0 plus a 64-byte CUDA IPC handle to hand to the other ranks (any transport:
 *                       torch.distributed, MPI...)
 *   rw_gather_open      maps the peers' buffers from their IPC handles ([W][64] bytes, own slot ignored)
 *   rw_step_gather      rw_step whose outputs land in slot (k & 1) of all ranks' gathered tensors, k = number of earlier
 *                       rw_step_gather calls on this handle (all ranks must call it the same number of times).  Returns the
 *                       slot index (0 / 1) on success, a negative RW_E* code on failure.  The caller then runs ONE
 *                       barrier-like collective on the same stream (e.g. a 4-byte all-reduce) and reads slot (k & 1).
 *   rw_gather_slot      device pointers of a slot of this rank's gathered tensors
 * Ordering contract (why two slots suffice): a rank that is ahead can at most be pushing step k+1 while this rank still reads
 * step k -- its step k+2 has to pass barrier k+1 first, which this rank only joins after its own rw_step_gather k+1, which is
 * stream-ordered after everything this rank enqueued on that stream to consume step k.  So: consume step k on the stream that is
 * passed to rw_step_gather (or make that stream wait for the consumer) before calling rw_step_gather k+1. */
int rw_gather_create(rw_handle *h, int world_size, int rank, float **obs_all, float **reward_all, uint8_t **done_all,
                     void *ipc_handle_out /* 64 bytes */);
int rw_gather_open(rw_handle *h, const void *ipc_handles /* [world_size][64] */);
int rw_step_gather(rw_handle *h, const int32_t *actions_dev, void *stream);
int rw_gather_slot(const rw_handle *h, int slot, float **obs_all, float **reward_all, uint8_t **done_all);

/* Diagnostics: when dev_ptr != NULL every warp of every CTA writes phase timestamps (globaltimer, ns) to
 * dev_ptr[(cta*32 + warp)*8 + k] (uint64).  NULL switches it off (default). */
int rw_debug_set_clock_buffer(rw_handle *h, void *dev_ptr);

/* Benchmark helper: write `bytes` (a multiple of 16, larger than the L2) to buf_dev with a kernel that uses the same
 * shared-memory carve-out as the step kernel, so that an L2 flush between timed launches does not also force an SM
 * shared-memory reconfiguration on the next launch. */
int rw_flush_l2(rw_handle *h, void *buf_dev, int64_t bytes, int value, void *stream);

/* Synthetic-rollout helper for benchmarks: fills actions_dev [E][n][A] with iid uniform actions over {0..4}
 * (message bits uniform over {0,1}) from a counter-based generator keyed by (seed, step). */
int rw_random_actions(rw_handle *h, int32_t *actions_dev, uint64_t seed, uint64_t step, void *stream);

#ifdef __cplusplus
}
#endif
#endif
