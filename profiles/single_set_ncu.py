"""Launch pattern for an ncu comparison: medium-8ag-hard, 120 launches of ONE world set, then 120 launches rotating over 16 sets."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rware_b200
from bench import WORKLOADS
env_id, over, E = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "medium-8ag-hard"]
sets = [rware_b200.make_vec(env_id, E, auto_reset=True, **over) for _ in range(16)]
for k, v in enumerate(sets):
    v.reset(seed=k * E)                      # 16 reset launches (rw_main_kernel<2>)
acts = [torch.empty((E, sets[0].n_agents, 1), dtype=torch.int32, device="cuda") for _ in range(8)]
for i, a in enumerate(acts):
    sets[0].backend.random_actions(a, 7, i)
for i in range(120):
    sets[0].step(acts[i % 8])                # step launches 0..119: one set
for i in range(120):
    sets[i % 16].step(acts[i % 8])           # step launches 120..239: rotating
torch.cuda.synchronize()
