"""End-to-end (rw_step_host) exploration on a GPU box: transport mode x copy-engine share x host threads.
   python profiles/hostexp/e2e_sweep.py [env_id] [E]      (RWARE_B200_HOST_THREADS etc. are read by the library at first use)
Prints one JSON line per setting: ms per step of 30 timed calls after 10 warm-up calls."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rware_b200

env_id = sys.argv[1] if len(sys.argv) > 1 else "rware-tiny-4ag-v1"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist = None
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
v = rware_b200.make_vec(env_id, E, auto_reset=True)
v.reset(seed=0)
n, L = v.n_agents, v.obs_len
be = v.backend
h_acts = [be.host_alloc((E, n), torch.int32) for _ in range(8)]   # fresh actions every step
for h in h_acts:
    h.copy_(torch.randint(0, 5, (E, n), dtype=torch.int32))
h_obs = be.host_alloc((E, n, L), torch.float32)
h_rew = be.host_alloc((E, n), torch.float32)
h_done = be.host_alloc((E,), torch.uint8)
gbs, threads = be.host_stream_probe(h_obs, 5)
base = {"env": env_id, "E": E, "rank": local_rank, "threads": threads, "host_stream_gbs": round(gbs, 1), "gpu_node": be.gpu_numa_node(),
        "obs_node": be.host_numa_node(h_obs)}
settings = [(0, None), (1, None), (1, 0.0), (1, 0.15), (1, 0.25), (1, 0.35), (1, 0.5), (2, None), (2, 0.0)]
if len(sys.argv) > 3:
    settings = [(1, None)]
for mode, share in settings:
    be.set_host_transport(mode)
    be.set_host_dma_share(share)
    for k in range(10):
        be.step_host(h_acts[k % 8], h_obs, h_rew, h_done)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(30):
        be.step_host(h_acts[k % 8], h_obs, h_rew, h_done)
    ms = (time.perf_counter() - t0) / 30 * 1e3
    if dist:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    print(json.dumps(dict(base, mode=mode, share=share, settled_share=round(be.host_dma_share(), 3), ms_per_step=round(ms, 4),
                          env_steps_per_s=round(E * int(os.environ.get("WORLD_SIZE", "1")) / ms * 1e3))), flush=True)
if os.environ.get("RWARE_B200_HOST_TRACE"):
    pass
if dist:
    dist.destroy_process_group()
