"""Per-warp phase timeline of the fused step kernel (globaltimer stamps written by the kernel itself).
stamps: 0 start, 1 after prologue+__syncthreads, 2 tile landed (mbarrier), 3 logic done, 4 pair barrier passed,
5 obs loop done, 6 last TMA store drained."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rware_b200

wl = sys.argv[1] if len(sys.argv) > 1 else "rware-tiny-4ag-v1"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
vec = rware_b200.make_vec(wl, E, auto_reset=True)
vec.reset(seed=0)
acts_ring = [torch.empty((E, vec.n_agents), dtype=torch.int32, device="cuda") for _ in range(8)]
for k, a in enumerate(acts_ring):
    vec.backend.random_actions(a, 1, k)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for k in range(120):  # fresh random actions every step (the same tensor again and again makes agents repeat one action: a much cheaper step)
    vec.step(acts_ring[k % 8])
acts = acts_ring[3]
ncta = 256
buf = torch.zeros((ncta * 32 * 8,), dtype=torch.int64, device="cuda")
vec.backend.set_clock_buffer(buf)
out = {}
for mode in ("warm", "flushed"):
    if mode == "flushed":
        flush.fill_(1)
    buf.zero_()
    torch.cuda.synchronize()
    vec.step(acts)
    torch.cuda.synchronize()
    t = buf.cpu().numpy().reshape(ncta, 32, 8).astype(np.float64)
    used = t[:, :, 0] > 0
    t0 = t[:, :, 0][used].min()
    rel = (t - t0) / 1e3  # us
    res = {}
    for role, name in ((0, "primary"), (1, "helper")):
        m = used.copy()
        m[:, (1 - role)::2] = False
        res[name] = {f"stamp{k}_us_mean": float(rel[:, :, k][m & (t[:, :, k] > 0)].mean()) if (m & (t[:, :, k] > 0)).any() else None
                     for k in range(7)}
        res[name].update({f"stamp{k}_us_max": float(rel[:, :, k][m & (t[:, :, k] > 0)].max()) if (m & (t[:, :, k] > 0)).any() else None
                          for k in (0, 3, 5, 6)})
    res["kernel_span_us"] = float(rel[:, :, 6][used & (t[:, :, 6] > 0)].max())
    res["ctas_seen"] = int(used.any(axis=1).sum())
    out[mode] = res
print(json.dumps(out, indent=1))
