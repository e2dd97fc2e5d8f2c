// rw_launch.h -- internal interface between the C-ABI host code (rw_capi.cu) and the kernels (rw_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "rw_common.h"

enum RwMode { RW_MODE_STEP = 0, RW_MODE_OBS = 1, RW_MODE_RESET = 2 };

// one launch = the whole hot path for every world of the handle
cudaError_t rw_launch_main(RwMode mode, const RwParams &p, const int32_t *actions, float *obs, float *reward,
                           uint8_t *done, const uint8_t *mask, cudaStream_t stream, bool pdl_ok = true);
cudaError_t rw_launch_host_flag(uint32_t *counter, volatile uint32_t *host_flag, cudaStream_t stream);
cudaError_t rw_configure_kernels(const RwParams &p, int smem_bytes_per_cta);
void rw_force_dynamic_kernels(bool on);  // diagnostics/tests: bypass the static specialisations
cudaError_t rw_launch_get_state(const RwParams &p, const void *view /* rw_state_view */, cudaStream_t stream);
cudaError_t rw_launch_set_state(const RwParams &p, const void *view, cudaStream_t stream);
cudaError_t rw_launch_seed_scatter(const RwParams &p, const uint32_t *staging /*[10][E]*/, const uint8_t *mask,
                                   cudaStream_t stream);
cudaError_t rw_launch_random_actions(const RwParams &p, int32_t *actions, uint64_t seed, uint64_t step,
                                     cudaStream_t stream);
cudaError_t rw_launch_flush_l2(void *buf, size_t bytes, int value, int smem_bytes, cudaStream_t stream);

// whole-warehouse views (rw_views.cu)
struct RwViewSpec;
cudaError_t rw_launch_global_image(const RwParams &p, const RwViewSpec &spec, float *out, cudaStream_t stream);
cudaError_t rw_launch_render(const RwParams &p, uint8_t *out, int64_t env0, int64_t count, cudaStream_t stream);

// end-to-end transport (rw_step_host): device side (rw_kernels.cu); the host side is rw_hostpack.h
cudaError_t rw_launch_pack_u8(const float *in, uint8_t *out, size_t n, cudaStream_t stream);
cudaError_t rw_launch_pack_bits(const float *in, uint32_t *bits, uint8_t *wide, size_t f_lo, size_t f_hi, int L, int nwide,
                                size_t row_lo, size_t row_hi, cudaStream_t stream);
