"""Tiny driver for `ncu` captures: reset + 6 step launches + 1 observation-gather launch of one workload, e.g.
    ncu --set full --clock-control none --import-source on -k regex:rw_main_kernel --launch-skip 6 --launch-count 2 \
        -o gpurun_out/prof_r1 python profiles/prof_step.py
captures the last step kernel and the stand-alone gather kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rware_b200

wl = sys.argv[1] if len(sys.argv) > 1 else "rware-tiny-4ag-v1"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
v = rware_b200.make_vec(wl, E, auto_reset=True)
v.reset(seed=0)
a = torch.empty((E, v.n_agents), dtype=torch.int32, device="cuda")
for i in range(6):
    v.backend.random_actions(a, 1, i)
    v.step(a)
torch.cuda.synchronize()
v.observations()
torch.cuda.synchronize()
