"""Multi-GPU check, run under torchrun on a box with >= 2 GPUs (tests/test_multigpu_gather.py spawns it when 2 GPUs are visible):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py
1. the fused push all-gather (TMA stores into peer-mapped buffers) returns exactly what the NCCL all-gather returns;
2. the same with a deliberately slow consumer on rank 0 and NOTHING else synchronising the ranks (the write-after-read hazard a
   single gathered buffer per rank would have: a fast rank's step k+1 landing in the tensors a slow rank still reads for step k);
3. timings per step: no gather / NCCL all-gather / fused push all-gather (device events, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rware_b200  # noqa: E402
from rware_b200.distributed import ShardedVecWarehouse  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    env_id = os.environ.get("MGPU_ENV_ID", "rware-tiny-4ag-v1")
    kw = rware_b200.env_kwargs(env_id)
    E_local = int(os.environ.get("MGPU_ENVS", "65536"))
    total = E_local * world
    a = ShardedVecWarehouse(total, auto_reset=True, device=dev, **kw)  # NCCL gather
    b = ShardedVecWarehouse(total, auto_reset=True, device=dev, **kw)  # fused push gather
    b.enable_fused_gather()
    a.reset(seed=3)
    b.reset(seed=3)
    acts = [torch.empty((E_local, kw["n_agents"]), dtype=torch.int32, device=dev) for _ in range(8)]
    for i, t in enumerate(acts):
        a.local.backend.random_actions(t, 100 + rank, i)
    ok = True
    for i in range(24):
        oa, ra, da = a.step(acts[i % 8], gather=True)
        ob, rb, db = b.step(acts[i % 8], gather=True)
        torch.cuda.synchronize()
        ok &= bool(torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db))
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    # 2. skewed consumers.  Reference trajectory first (NCCL gather, kept on the device), then the fused path ALONE: after
    # every step rank 0 stalls its stream for ~5 ms before it reads the gathered tensors, the other ranks read at once and race
    # ahead as far as the protocol lets them.
    E_skew, STEPS = int(os.environ.get("MGPU_SKEW_ENVS", "2048")), int(os.environ.get("MGPU_SKEW_STEPS", "200"))
    c = ShardedVecWarehouse(E_skew * world, auto_reset=True, device=dev, **kw)
    d = ShardedVecWarehouse(E_skew * world, auto_reset=True, device=dev, **kw)
    d.enable_fused_gather()
    c.reset(seed=11)
    d.reset(seed=11)
    sk_acts = [torch.empty((E_skew, kw["n_agents"]), dtype=torch.int32, device=dev) for _ in range(8)]
    for i, t in enumerate(sk_acts):
        c.local.backend.random_actions(t, 500 + rank, i)
    want = []
    for i in range(STEPS):
        o, r, dn = c.step(sk_acts[i % 8], gather=True)
        want.append((o.sum(dim=2, dtype=torch.float32).clone(), r.clone(), dn.clone(), o[:: max(1, E_skew // 64)].clone()))
    torch.cuda.synchronize()
    dist.barrier()
    got = []
    stall = int(os.environ.get("MGPU_SKEW_CYCLES", str(9_500_000)))  # ~5 ms at 1.9 GHz
    for i in range(STEPS):
        o, r, dn = d.step(sk_acts[i % 8], gather=True)
        if rank == 0:
            torch.cuda._sleep(stall)
        got.append((o.sum(dim=2, dtype=torch.float32).clone(), r.clone(), dn.clone(), o[:: max(1, E_skew // 64)].clone()))
    torch.cuda.synchronize()
    skew_ok = all(all(torch.equal(x, y) for x, y in zip(w, g)) for w, g in zip(want, got))
    sflag = torch.tensor([int(skew_ok)], device=dev)
    dist.all_reduce(sflag, op=dist.ReduceOp.MIN)

    def timed(fn, K=100):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(K):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / K], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    t_none = timed(lambda i: a.step(acts[i % 8]))
    t_nccl = timed(lambda i: a.step(acts[i % 8], gather=True))
    t_fused = timed(lambda i: b.step(acts[i % 8], gather=True))
    if rank == 0:
        obs_bytes = total * kw["n_agents"] * a.obs_len * 4
        print(json.dumps({"env_id": env_id, "world": world, "worlds_per_gpu": E_local, "fused_equals_nccl": bool(flag.item()),
                          "fused_equals_nccl_with_slow_consumer": bool(sflag.item()), "slow_consumer_steps": STEPS,
                          "ms_per_step": {"no_gather": t_none, "nccl_all_gather": t_nccl, "fused_push": t_fused},
                          "env_steps_per_s": {k: total / (v * 1e-3) for k, v in
                                              (("no_gather", t_none), ("nccl_all_gather", t_nccl), ("fused_push", t_fused))},
                          "gathered_obs_bytes_per_rank": obs_bytes}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
