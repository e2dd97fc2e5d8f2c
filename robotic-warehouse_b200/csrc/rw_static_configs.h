// rw_static_configs.h -- geometries compiled as static specialisations (shared by rw_kernels.cu and the test twin).
#pragma once
// (H, W, S, sensor_range): the four registered warehouse sizes (rware/__init__.py:7-12, column_height 8), n and R dynamic
#define RW_STATIC_CONFIGS(X) \
    X(11, 10, 32, 1) X(20, 10, 80, 1) X(20, 16, 144, 1) X(29, 16, 224, 1) \
    X(11, 10, 32, 2) X(20, 10, 80, 2) X(20, 16, 144, 2) X(29, 16, 224, 2)
// (H, W, S, sensor_range, n_agents, request_queue_size): the BASELINE.json configurations, everything static
#define RW_FULL_STATIC_CONFIGS(X) \
    X(11, 10, 32, 1, 2, 2) X(11, 10, 32, 1, 4, 4) X(20, 10, 80, 1, 4, 4) X(20, 16, 144, 1, 8, 4) X(29, 16, 224, 2, 16, 16)
