"""How much does a rollout of ONE world set (step k+1 needs step k's state) lose against launches that rotate over several
independent sets?  us per launch for nsets in {1, 2, 4, 16}; run it with RWARE_B200_PDL=0/1.
   python profiles/single_set_probe.py [workload ...]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rware_b200
from bench import WORKLOADS
names = sys.argv[1:] or ["tiny-4ag", "medium-8ag-hard", "large-16ag-r2"]
for name in names:
    env_id, over, E = WORKLOADS[name]
    sets = [rware_b200.make_vec(env_id, E, auto_reset=True, **over) for _ in range(16)]
    for k, v in enumerate(sets):
        v.reset(seed=k * E)
    acts = [torch.empty((E, sets[0].n_agents, 1), dtype=torch.int32, device="cuda") for _ in range(8)]
    for i, a in enumerate(acts):
        sets[0].backend.random_actions(a, 7, i)
    out = {"workload": name, "pdl": os.environ.get("RWARE_B200_PDL", "1")}
    for nsets in (1, 2, 4, 16):
        for i in range(300):
            sets[i % nsets].step(acts[i % 8])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(400):
            sets[i % nsets].step(acts[i % 8])
        e1.record()
        torch.cuda.synchronize()
        out[f"us_per_launch_nsets{nsets}"] = round(e0.elapsed_time(e1) / 400 * 1e3, 2)
    # one set, observations not requested (state only): separates the observation writes from the rest
    for i in range(100):
        sets[0].step(acts[i % 8], compute_obs=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(400):
        sets[0].step(acts[i % 8], compute_obs=False)
    e1.record()
    torch.cuda.synchronize()
    out["us_per_launch_nsets1_no_obs"] = round(e0.elapsed_time(e1) / 400 * 1e3, 2)
    print(json.dumps(out), flush=True)
    for v in sets:
        v.close()
