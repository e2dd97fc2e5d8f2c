// gpu_quickcheck.cu -- TEST-ONLY: a few-hundred-millisecond parity check of the C ABI against the C oracle without Python, for
// when GPU time is too short to import torch.  Covers the paths added last in round 1: the AGENT_DIRECTION / AGENT_LOAD image
// layers (second-pass form) with their index-error count, and rw_obs_bytes on a tensor whose size is not a multiple of 4.
//   nvcc -O2 -o tests/ctest/_build/gpu_quickcheck tests/ctest/gpu_quickcheck.cu -Lrobotic-warehouse_b200/lib -lrware_b200 \
//        -Loracle/_build -lrware_oracle -Xlinker -rpath -Xlinker '$ORIGIN/../../../robotic-warehouse_b200/lib' ...
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/rware_b200.h"
#include "../../oracle/rware_oracle.h"

#define CHECK(x) do { int rc__ = (x); if (rc__) { printf("FAIL %s -> %d (%s)\n", #x, rc__, h ? rw_last_error(h) : rw_last_error(nullptr)); return 1; } } while (0)
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA FAIL %s: %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

int main() {
    rw_handle *h = nullptr;
    int32_t H, W, goals[4];
    std::vector<uint8_t> hw(1 << 12);
    if (orc_layout_from_params(3, 1, 8, &H, &W, hw.data(), (int)hw.size(), goals)) { printf("layout failed\n"); return 1; }
    const int E = 70, n = 5, r = 1, R = 3, K = 3, C = 3;
    const int32_t layers[3] = {3, 2, 4};  // AGENT_DIRECTION, AGENTS, AGENT_LOAD
    rw_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = RW_ABI_VERSION; cfg.grid_h = H; cfg.grid_w = W; cfg.n_agents = n; cfg.sensor_range = r;
    cfg.request_queue_size = R; cfg.max_steps = 40; cfg.reward_type = RW_REWARD_INDIVIDUAL; cfg.n_goals = 2;
    cfg.observation_type = RW_OBS_IMAGE; cfg.image_directional = 1; cfg.n_image_layers = C;
    for (int i = 0; i < C; i++) cfg.image_layers[i] = layers[i];
    cfg.highways = hw.data(); cfg.goals_xy = goals;
    CHECK(rw_create(&cfg, E, 0, &h));
    CHECK(rw_set_auto_reset(h, 1));
    rw_info info;
    CHECK(rw_get_info(h, &info));
    const int L = info.obs_len;
    if (L != C * K * K) { printf("FAIL obs_len %d\n", L); return 1; }
    orc_config oc;
    memset(&oc, 0, sizeof(oc));
    oc.H = H; oc.W = W; oc.n_agents = n; oc.sensor_range = r; oc.request_queue_size = R; oc.max_steps = 40; oc.reward_type = 1;
    oc.n_goals = 2; oc.highways = hw.data(); oc.goals_xy = goals;
    std::vector<orc_env *> orc(E);
    std::vector<uint64_t> seeds(E);
    for (int e = 0; e < E; e++) { orc[e] = orc_create(&oc); seeds[e] = 100 + e; orc_reset(orc[e], (int64_t)seeds[e], nullptr); }
    float *d_obs, *d_rew; uint8_t *d_done, *d_bytes; int32_t *d_act;
    CU(cudaMalloc(&d_obs, sizeof(float) * E * n * L)); CU(cudaMalloc(&d_rew, sizeof(float) * E * n));
    CU(cudaMalloc(&d_done, E)); CU(cudaMalloc(&d_bytes, (size_t)E * n * L)); CU(cudaMalloc(&d_act, sizeof(int32_t) * E * n));
    CHECK(rw_seed(h, seeds.data(), nullptr, nullptr));
    CHECK(rw_reset(h, nullptr, d_obs, nullptr));
    std::vector<float> obs((size_t)E * n * L), want((size_t)n * L), rew(E * n);
    std::vector<uint8_t> done(E), bytes((size_t)E * n * L);
    std::vector<int32_t> act(E * n);
    std::vector<double> orew(n);
    int64_t expect_err = 0, compared = 0;
    uint64_t rng = 88172645463325252ULL;
    for (int t = 0; t <= 45; t++) {
        CU(cudaMemcpy(obs.data(), d_obs, sizeof(float) * obs.size(), cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; e++) {
            int rc = orc_obs_image(orc[e], layers, C, 1, want.data(), nullptr);
            if (rc == -1) { expect_err++; continue; }   // the reference raises IndexError here
            if (rc) { printf("FAIL oracle rc %d\n", rc); return 1; }
            if (memcmp(want.data(), &obs[(size_t)e * n * L], sizeof(float) * n * L)) { printf("FAIL obs mismatch t=%d world=%d\n", t, e); return 1; }
            compared++;
        }
        int64_t got_err = -1;
        CHECK(rw_error_count(h, &got_err, nullptr));
        if (got_err != expect_err) { printf("FAIL error count t=%d: device %lld, oracle %lld\n", t, (long long)got_err, (long long)expect_err); return 1; }
        // byte image of the observations: E*n*L = 70*5*27 = 9450 values (not a multiple of 4: the tail path)
        CHECK(rw_obs_bytes(h, d_obs, d_bytes, nullptr));
        CU(cudaMemcpy(bytes.data(), d_bytes, bytes.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < bytes.size(); i++)
            if ((float)bytes[i] != obs[i]) { printf("FAIL obs_bytes at %zu\n", i); return 1; }
        for (int i = 0; i < E * n; i++) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; act[i] = (rng % 10 < 5) ? 1 : (int)(rng % 5); }
        CU(cudaMemcpy(d_act, act.data(), sizeof(int32_t) * act.size(), cudaMemcpyHostToDevice));
        CHECK(rw_step(h, d_act, d_obs, d_rew, d_done, nullptr));
        CU(cudaMemcpy(rew.data(), d_rew, sizeof(float) * rew.size(), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(done.data(), d_done, E, cudaMemcpyDeviceToHost));
        for (int e = 0; e < E; e++) {
            uint8_t od = 0;
            if (orc_step(orc[e], &act[e * n], nullptr, orew.data(), &od)) { printf("FAIL oracle step\n"); return 1; }
            for (int i = 0; i < n; i++) if ((float)orew[i] != rew[e * n + i]) { printf("FAIL reward t=%d world=%d\n", t, e); return 1; }
            if (od != done[e]) { printf("FAIL done t=%d world=%d\n", t, e); return 1; }
            if (od) orc_reset(orc[e], -1, nullptr);
        }
    }
    printf("gpu_quickcheck OK: %lld image observations compared bit-exactly, %lld index errors counted on both sides, obs_bytes exact\n",
           (long long)compared, (long long)expect_err);
    rw_destroy(h);
    return 0;
}
