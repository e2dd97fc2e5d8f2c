// rw_inst.cu -- instantiations of the fused kernel, compiled once per RW_INST_GROUP (0..6) so that the translation units
// build in parallel.  Each group exposes a selector that returns the kernel of the first configuration class of the
// group matching the parameters (or nullptr).
#include <cstdlib>

#include "rw_main_kernel.cuh"
#include "rw_static_configs.h"

#ifndef RW_INST_GROUP
#error "compile with -DRW_INST_GROUP=0..6"
#endif

#define RW_FULL(H_, W_, S_, R_, N_, RQ_)                                                                    \
    if (!force_dynamic && p.msg_bits == 0 && !p.coop && p.H == H_ && p.W == W_ && p.S == S_ && p.r == R_ && \
        p.n == N_ && p.R == RQ_)                                                                            \
        return rw_kernel_for_mode<RwStatic<H_, W_, S_, R_, 0, N_, RQ_>>(mode, small_cta, p.obs_bits_out != nullptr);
#define RW_GEOM(H_, W_, S_, R_)                                                                  \
    if (!force_dynamic && p.msg_bits == 0 && p.H == H_ && p.W == W_ && p.S == S_ && p.r == R_)   \
        return rw_kernel_for_mode<RwStatic<H_, W_, S_, R_, 0>>(mode, small_cta, p.obs_bits_out != nullptr);
#define RW_DYN() return rw_kernel_for_mode<RwDynamic>(mode, small_cta, p.obs_bits_out != nullptr);

#define RW_CAT2(a, b) a##b
#define RW_CAT(a, b) RW_CAT2(a, b)

rw_kernel_fn RW_CAT(rw_select_group_, RW_INST_GROUP)(RwMode mode, const RwParams &p, bool force_dynamic) {
    (void)mode; (void)p; (void)force_dynamic;
    static const bool allow_small = !(getenv("RWARE_B200_SMALL_CTA") && atoi(getenv("RWARE_B200_SMALL_CTA")) == 0);
    const bool small_cta = allow_small && p.warps_per_cta <= 7;  // CTA of at most 448 threads: the high-register build
    (void)small_cta;
#if RW_INST_GROUP == 0
    RW_FULL(11, 10, 32, 1, 2, 2) RW_FULL(11, 10, 32, 1, 4, 4)
#elif RW_INST_GROUP == 1
    RW_FULL(20, 10, 80, 1, 4, 4) RW_FULL(20, 16, 144, 1, 8, 4)
#elif RW_INST_GROUP == 2
    RW_FULL(29, 16, 224, 2, 16, 16) RW_GEOM(11, 10, 32, 1)
#elif RW_INST_GROUP == 3
    RW_GEOM(20, 10, 80, 1) RW_GEOM(20, 16, 144, 1)
#elif RW_INST_GROUP == 4
    RW_GEOM(29, 16, 224, 1) RW_GEOM(11, 10, 32, 2)
#elif RW_INST_GROUP == 5
    RW_GEOM(20, 10, 80, 2) RW_GEOM(20, 16, 144, 2)
#elif RW_INST_GROUP == 6
    RW_GEOM(29, 16, 224, 2) RW_DYN()
#endif
    return nullptr;
}
