"""Prints where rw_step_host spends its time (RWARE_B200_HOST_TRACE=1): pool wake-up, per-chunk PCIe wait and widening."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rware_b200
env_id = sys.argv[1] if len(sys.argv) > 1 else "rware-tiny-4ag-v1"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
v = rware_b200.make_vec(env_id, E, auto_reset=True)
v.reset(seed=0)
n, L = v.n_agents, v.obs_len
h_act = torch.randint(0, 5, (E, n), dtype=torch.int32).pin_memory()
h_obs = torch.empty((E, n, L), dtype=torch.float32).pin_memory()
h_rew = torch.empty((E, n), dtype=torch.float32).pin_memory()
h_done = torch.empty((E,), dtype=torch.uint8).pin_memory()
for _ in range(6):
    v.backend.step_host(h_act, h_obs, h_rew, h_done)
