"""Per-launch time of ONE world set around the synchronous episode end (all worlds hit max_steps = 500 in the same launch and are
reset inside it): is the extra cost of a single-set rollout one expensive launch, or slower launches afterwards?"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rware_b200
from bench import WORKLOADS
for name in sys.argv[1:] or ["tiny-4ag", "medium-8ag-hard", "large-16ag-r2"]:
    env_id, over, E = WORKLOADS[name]
    v = rware_b200.make_vec(env_id, E, auto_reset=True, **over)
    v.reset(seed=0)
    acts = [torch.empty((E, v.n_agents, 1), dtype=torch.int32, device="cuda") for _ in range(8)]
    for i, a in enumerate(acts):
        v.backend.random_actions(a, 7, i)
    T = 1100
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
    ev[0].record()
    for t in range(T):
        v.step(acts[t % 8])
        ev[t + 1].record()
    torch.cuda.synchronize()
    us = [ev[t].elapsed_time(ev[t + 1]) * 1e3 for t in range(T)]
    avg = lambda a, b: round(sum(us[a:b]) / (b - a), 2)
    print(json.dumps({"workload": name, "avg_us_steps_100_400": avg(100, 400), "steps_495_505": [round(x, 1) for x in us[495:505]],
                      "avg_us_steps_520_900": avg(520, 900), "steps_995_1005": [round(x, 1) for x in us[995:1005]],
                      "avg_us_steps_1020_1100": avg(1020, 1100), "max_us": round(max(us), 1), "argmax": us.index(max(us))}), flush=True)
    v.close()
