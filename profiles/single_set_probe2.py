"""Why is a rollout of ONE world set slower per launch than launches rotating over 16 sets?  Per-launch CUDA events in both
modes, with and without a 512 MiB L2-flushing write between launches, and SM clocks sampled under each load."""
import json, os, subprocess, sys, threading, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rware_b200
from bench import WORKLOADS, ClockSampler
name = sys.argv[1] if len(sys.argv) > 1 else "medium-8ag-hard"
env_id, over, E = WORKLOADS[name]
sets = [rware_b200.make_vec(env_id, E, auto_reset=True, **over) for _ in range(16)]
for k, v in enumerate(sets):
    v.reset(seed=k * E)
acts = [torch.empty((E, sets[0].n_agents, 1), dtype=torch.int32, device="cuda") for _ in range(8)]
for i, a in enumerate(acts):
    sets[0].backend.random_actions(a, 7, i)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = {"workload": name}
for label, nsets, do_flush in (("one_set", 1, False), ("rotating_16", 16, False), ("one_set_flushed", 1, True), ("rotating_16_flushed", 16, True)):
    for i in range(100):
        sets[i % nsets].step(acts[i % 8])
    torch.cuda.synchronize()
    N = 300
    a = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
    b = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
    for i in range(N):
        if do_flush:
            sets[0].backend.flush_l2(flush, i & 0xff)
        a[i].record()
        sets[i % nsets].step(acts[i % 8])
        b[i].record()
    torch.cuda.synchronize()
    us = sorted(x.elapsed_time(y) * 1e3 for x, y in zip(a, b))
    out[label + "_median_us_per_launch_events"] = round(us[N // 2], 2)
# clocks under the two free-running loads (no events)
for label, nsets in (("one_set", 1), ("rotating_16", 16)):
    s = ClockSampler(0)
    s.start()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 1.5:
        for i in range(500):
            sets[i % nsets].step(acts[i % 8])
        n += 500
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[label + "_free_running_us"] = round(dt / n * 1e6, 2)
    out[label + "_clocks"] = s.stop()
print(json.dumps(out))
