// rw_hostpack.h -- internal interface between rw_capi.cu and the host worker pool of the end-to-end transport (rw_hostpack.cpp).
#pragma once
#include <sched.h>
#include <stddef.h>
#include <stdint.h>

#define RW_HOST_BLOCK 16384  // floats per work block (64 KB of output): the unit a pool thread pulls
#define RW_HOST_MAX_NODES 8

enum { RW_HOSTJOB_WIDEN = 0, RW_HOSTJOB_BITS = 1, RW_HOSTJOB_FILL = 2 };

struct RwHostJob {
    int kind = RW_HOSTJOB_WIDEN;
    const uint8_t *src = nullptr;   // WIDEN: one byte per value.  BITS: bit j = "float j is non-zero" (little-endian bytes)
    const uint8_t *wide = nullptr;  // BITS: [n_rows][nwide] bytes, the values of the first nwide floats of every row
    float *dst = nullptr;           // n floats (BITS: 64-byte aligned for the vector forms, else scalar)
    size_t n = 0;                   // floats to produce
    int L = 1, nwide = 0;           // BITS: floats per row, leading non-flag values per row
    size_t n_rows = 0;              // BITS: rows covered by `wide`
    float fill_value = 0;           // FILL
    // the source arrives in `nchunks` pieces; piece c covers the blocks [chunk_end[c-1], chunk_end[c]) and a block may only
    // run once its piece has been announced (rw_hostpool_chunk_ready)
    int nchunks = 1;
    size_t chunk_end[8] = {0};
};

void rw_hostpool_prewake();                     // hint: a job follows shortly (wakes parked workers)
void rw_hostpool_begin(const RwHostJob &job);   // publishes the job (takes the process-wide job lock)
void rw_hostpool_chunk_ready(int chunks);       // pieces 0 .. chunks-1 of the source have arrived
void rw_hostpool_cancel();                      // the source will never arrive (CUDA error): blocks are skipped
void rw_hostpool_finish();                      // the caller works along, returns when every block is done (releases the lock)
int rw_hostpool_threads();
int rw_hostpool_isa();                          // 0 scalar, 1 AVX2, 2 AVX-512
bool rw_host_cpus_of_node(int node, cpu_set_t *set);
int rw_host_numa_nodes();                                          // NUMA nodes with CPUs (1 when unknown or RWARE_B200_HOST_NUMA=0)
void rw_host_nodes_of(void *const *pages, int count, int *status); // node of each page (-1 unknown)
