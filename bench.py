#!/usr/bin/env python
"""bench.py -- throughput of the RWARE step path (BASELINE.json: env-steps/s on synthetic random-action rollouts).

    python bench.py --gpus 1 --steps 200 --warmup 20              # our CUDA path (default)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1  # CPU arm: the C oracle port on all host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                    # one rank per GPU, weak scaling

A "step" is ONE launch of the fused kernel (rw_step) over the whole batch of worlds of a rank: action decode, move
proposals, collision resolution, apply, deliveries/resample, counters, in-kernel auto-reset and the observation gather.
Workload (N=1): BASELINE.json configs[1] = rware-tiny-4ag-v1 x 65 536 worlds, actions iid uniform over {0..4}
(generated on the device beforehand, resident in HBM), episodes of 500 steps with auto-reset.
Timing (default `--l2 rotate`): the launches rotate over 16 independent sets of worlds (16 x 83 MB of state and
observation buffers = 1.3 GB >> the 126 MB L2, so every launch finds its inputs in HBM and its 74 MB of observations
evict the previous set's), K launches back to back inside ONE pair of CUDA events on the launching stream.
`--l2 flush` instead brackets every launch with its own event pair and writes 512 MiB between launches.  `value`
= worlds * steps / time, max over ranks.  `e2e` goes through the host-buffer C-ABI call (rw_step_host):
pinned host actions -> device, kernel, observations/rewards/done -> pinned host, inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name -> (env id, overrides, worlds per GPU)
    "tiny-4ag": ("rware-tiny-4ag-v1", {}, 65536),
    "tiny-2ag": ("rware-tiny-2ag-v1", {}, 65536),
    "small-4ag": ("rware-small-4ag-v1", {}, 32768),
    "medium-8ag-hard": ("rware-medium-8ag-hard-v1", {}, 32768),
    "large-16ag-r2": ("rware-large-16ag-v1", {"sensor_range": 2}, 16384),
    # widened rows: image observations gathered on the device
    "tiny-4ag-image": ("rware-tiny-4ag-v1", {"observation_type": 2}, 65536),
    "tiny-4ag-imagedict": ("rware-tiny-4ag-v1", {"observation_type": 3}, 65536),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(env_id, overrides, target_seconds=10.0, threads=None, n_worlds=None):
    """The oracle port (plain C, oracle/) on the host cores, random-action rollout of the same configuration, timed two ways:
    * lock step (the headline of this arm, `value`): the same batch of worlds as the GPU arm, every step advances ALL worlds
      once and rewrites the stacked observation / reward / done arrays of the batch in host memory -- what a vector
      environment built on the reference does, and exactly what the GPU arm's end-to-end call delivers per step;
    * independent worlds: every host thread runs its worlds' episodes back to back, one world at a time, nothing stacked
      (everything stays in that core's cache) -- an upper bound no batched consumer can use, kept for continuity with round 1."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle_py
    import rware_b200
    kw = rware_b200.env_kwargs(env_id, **{k: v for k, v in overrides.items() if k != "observation_type"})
    (H, W), highways, goals = rware_b200.layout_from_params(kw["shelf_columns"], kw["shelf_rows"], kw["column_height"])
    env = oracle_py.OracleEnv(highways, goals, kw["n_agents"], kw["msg_bits"], kw["sensor_range"], kw["request_queue_size"],
                              kw["max_inactivity_steps"], kw["max_steps"], kw["reward_type"].value)
    threads = threads or os.cpu_count() or 1
    quota = None
    try:  # the container's CFS quota, if any: `threads` host threads run, but only this many CPUs' worth of time is granted
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()
        quota = None if q == "max" else float(q) / float(period)
    except Exception:
        pass
    # independent worlds (round-1 protocol), a third of the time budget
    n_envs = threads * 8
    t0 = time.perf_counter()
    done, _ = oracle_py.rollout_random(env, n_envs, 200, seed=1, n_threads=threads)  # calibration
    rate = done / (time.perf_counter() - t0)
    steps = max(200, int(rate * target_seconds / 3 / n_envs))
    t0 = time.perf_counter()
    done, _ = oracle_py.rollout_random(env, n_envs, steps, seed=2, n_threads=threads)
    dt_ind = time.perf_counter() - t0
    independent = {"value": done / dt_ind, "sample": f"{n_envs} worlds x {steps} steps, one world at a time per thread, {dt_ind:.1f} s"}
    # lock step over the GPU arm's batch
    E = int(n_worlds or 65536)
    batch = oracle_py.OracleBatch(env, E, n_threads=threads, auto_reset=True)
    batch.reset(np.arange(E, dtype=np.uint64))
    t0 = time.perf_counter()
    done, _ = batch.rollout_lockstep(4, seed=1)  # calibration (also first touch of the stacked arrays)
    rate = done / (time.perf_counter() - t0)
    steps = max(4, int(rate * target_seconds * 2 / 3 / E))
    t0 = time.perf_counter()
    done, _ = batch.rollout_lockstep(steps, seed=2)
    dt = time.perf_counter() - t0
    batch.close()
    return {"value": done / dt, "unit": "env-steps/s", "cores": threads, "cgroup_cpu_quota": quota, "kind": "port",
            "sample": f"{E} worlds x {steps} lock-step random-action steps of {env_id} {overrides or ''} with reset on done, stacked "
                      f"float32 observations / rewards / done of the whole batch rewritten every step, {dt:.1f} s on {threads} host "
                      f"threads (C oracle port; the Python reference itself cannot travel to the GPU box)",
            "independent_worlds": independent, "seconds": dt + dt_ind}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    env_id, overrides, E = WORKLOADS[args.workload]
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(env_id, overrides, target_seconds=args.ref_seconds, n_worlds=(args.envs or E) * max(1, args.gpus))
        if i >= args.warmup:
            vals.append(cb)
    v = sum(c["value"] for c in vals) / len(vals)
    cb = dict(vals[-1])
    cb["value"] = v
    secs = sum(c["seconds"] for c in vals) / len(vals)
    print(json.dumps({
        "impl": "reference", "metric": "env-steps/s", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * secs, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{env_id} {overrides or ''}".strip() + f", {(args.envs or E) * max(1, args.gpus)} worlds (the GPU arm's whole batch at {args.gpus} GPU(s)), random-action rollout (uniform actions), "
                               "reset on done; CPU, lock step (a step of this arm = one bounded sample, see cpu_baseline.sample)",
                   "worlds_per_gpu": args.envs or E},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "independent_worlds")},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def e2e_record(world, E, n, A, L, steps, e2e_s, plain_s, d2h_ms, dma_share, compressed, bit_rows, host_gbs, host_threads, numa):
    """The `e2e` object of the bench line: throughput through rw_step_host with host buffers, its transport and its roofline."""
    rows = E * n
    obs_bytes = rows * L * 4
    small = E * n * 4 + E
    if compressed:
        rows_cpu = min(rows, int((1.0 - dma_share) * rows) & ~63)
        per_row = (L / 8.0 + 2) if bit_rows else L
        d2h = int(rows_cpu * per_row + (rows - rows_cpu) * L * 4) + small
        transport = (("bits: every 0/1 flag of a row crosses PCIe as one bit and each coordinate as one byte"
                      if bit_rows else "bytes: every value crosses PCIe as one byte")
                     + f"; the library's host threads ({host_threads}) rebuild the float32 rows in the caller's buffer with streaming "
                       f"stores while the GPU's copy engine delivers {dma_share:.2f} of the rows as plain float32 (share follows the "
                       "measured rates); results bit-identical to the plain copy")
    else:
        d2h, transport = obs_bytes + small, "plain float32 copy"
    pcie_gbs = (obs_bytes + small) / (d2h_ms * 1e-3) / 1e9
    ms = e2e_s / steps * 1e3
    # both writers (host pool, copy engine) end in this rank's host DRAM: the call cannot beat obs_bytes / (their summed rates)
    bound_ms = obs_bytes / ((host_gbs + pcie_gbs) * 1e9) * 1e3
    return {"value": world * E * steps / e2e_s, "unit": "env-steps/s", "h2d_bytes_per_step": E * n * A * 4,
            "d2h_bytes_per_step": d2h, "steps": steps,
            "api": "rw_step_host: page-locked host buffers in and out (int32 actions, float32 observations / rewards, uint8 done); "
                   "copies, kernel, host expansion and synchronisation inside the timed region",
            "transport": transport, "ms_per_step": ms, "copy_engine_share": dma_share if compressed else 1.0,
            "roofline": {"bound": "host DRAM writes of the float32 rows (host pool + copy engine)",
                         "float32_bytes_to_host_per_step_per_rank": obs_bytes,
                         "host_pool_stream_gbs": host_gbs, "host_pool_threads": host_threads, "pcie_d2h_gbs": pcie_gbs,
                         "bound_ms_per_step": bound_ms, "frac": bound_ms / ms,
                         "achieved_host_write_gbs_per_rank": obs_bytes / (ms * 1e-3) / 1e9, "numa": numa},
            "plain_float32_transport": {"value": world * E * steps / plain_s, "ms_per_step": plain_s / steps * 1e3,
                                        "d2h_bytes_per_step": obs_bytes + small, "bare_d2h_copy_ms": d2h_ms,
                                        "pcie_d2h_gbs": pcie_gbs, "frac_of_transport": d2h_ms / (plain_s / steps * 1e3)}}


def time_rotating(step_fn, steps, warmup, dist, dev):
    """K back-to-back calls inside one CUDA event pair after W warm-up calls; max over ranks; returns ms per call."""
    import torch
    for i in range(warmup):
        step_fn(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def make_actions(vec, E, n, A, dev, rank, p_forward, ring=8, seed=4242):
    import torch
    acts = [torch.empty((E, n, A), dtype=torch.int32, device=dev) for _ in range(ring)]
    for i, a in enumerate(acts):
        vec.backend.random_actions(a, seed + rank, i)
        if p_forward is not None:
            g = torch.Generator(device=dev).manual_seed(seed + 16 * rank + i)
            u = torch.rand((E, n), device=dev, generator=g)
            other = torch.tensor([0, 2, 3, 4], device=dev, dtype=torch.int32)[torch.randint(0, 4, (E, n), device=dev, generator=g)]
            a[..., 0] = torch.where(u < p_forward, torch.ones_like(other), other)
    return acts


def other_workloads_records(dev, rank, peak, steps, warmup):
    """N=1 sub-records: BASELINE config 4 (medium-8ag-hard x 32 768, the dense-collision stress) under uniform actions and under
    P(FORWARD)=0.68, and config 5's per-GPU shard.  Same rotate-over-world-sets protocol as the headline, timed
    over ONE FULL EPISODE of every set (500 steps after a 50-step warm-up): the cost of a step depends on how
    far the worlds are from their reset (shelves get carried away from their home cells), so a few steps after a reset would flatter it."""
    import rware_b200
    out = []
    for name, pf in (("medium-8ag-hard", None), ("medium-8ag-hard", 0.68), ("large-16ag-r2", None)):
        env_id, overrides, E = WORKLOADS[name]
        nsets = 8
        sets = [rware_b200.make_vec(env_id, E, auto_reset=True, device=dev, **overrides) for _ in range(nsets)]
        for k, v in enumerate(sets):
            v.reset(seed=(1000 + k) * E)
        v0 = sets[0]
        acts = make_actions(v0, E, v0.n_agents, 1 + v0.msg_bits, dev, rank, pf)
        episode = int(v0.cfg.max_steps or 500)
        ms = time_rotating(lambda i: sets[i % nsets].step(acts[(i // nsets + i % nsets) % len(acts)]), episode * nsets, 50 * nsets, None, dev)
        b_step = int(v0.info["algorithmic_bytes_step"])
        out.append({"workload": f"{env_id} {overrides or ''}".strip(), "worlds": E, "actions": "uniform" if pf is None else f"P(FORWARD)={pf}",
                    "kernel_us": ms * 1e3, "env_steps_per_s": E / (ms * 1e-3), "algorithmic_bytes_per_env_step": b_step,
                    "roofline_frac": b_step * E / (ms * 1e-3) / 1e9 / peak,
                    "protocol": f"rotating over {nsets} world sets, {episode} timed steps per set (one full episode) after 50 warm-up steps per set"})
        for v in sets:
            v.close()
        del sets
    return out


def peer_copy_gbs(dev, local_rank, world):
    """GB/s of a plain device-to-device copy into the next GPU of the box (the NVLink path the fused push uses); None if unavailable."""
    import torch
    try:
        peer = torch.device("cuda", (local_rank + 1) % world)
        src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        dst = torch.empty(256 << 20, dtype=torch.uint8, device=peer)
        best = 0.0
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize(dev)
            torch.cuda.synchronize(peer)
            best = max(best, src.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del src, dst
        return best
    except Exception:
        return None


def gathered_records(dist, dev, rank, local_rank, world, steps, warmup):
    """N>1 sub-records: the ONE collective of the path on BASELINE configs 3 and 5 -- worlds sharded over
    the ranks, obs / reward / done all-gathered every step -- as kernel + NCCL all_gather_into_tensor and as the fused push
    (the step kernel's TMA stores write every chunk into all ranks' gathered tensors over NVLink; one 4-byte all-reduce as barrier)."""
    import torch
    import rware_b200
    peer = peer_copy_gbs(dev, local_rank, world)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    out = []
    for name, label in (("small-4ag", "BASELINE config 3"), ("large-16ag-r2", "BASELINE config 5")):
        env_id, overrides, E = WORKLOADS[name]
        nsets = 2
        rec = {"config": f"{label}: {env_id} {overrides or ''}".strip() + f", {E * world} worlds over {world} GPUs, obs/reward/done all-gathered",
               "worlds_per_gpu": E}
        for mode in ("nccl", "fused"):
            sets = [rware_b200.make_vec(env_id, E, auto_reset=True, device=dev, **overrides) for _ in range(nsets)]
            for k, v in enumerate(sets):
                v.reset(seed=((7 + k) * world + rank) * E)
            v0 = sets[0]
            n, L = v0.n_agents, v0.obs_len
            acts = make_actions(v0, E, n, 1 + v0.msg_bits, dev, rank, None)
            if mode == "nccl":
                g = (torch.empty((world * E, n, L), dtype=torch.float32, device=dev),
                     torch.empty((world * E, n), dtype=torch.float32, device=dev),
                     torch.empty((world * E,), dtype=torch.uint8, device=dev))

                def step_fn(i):
                    obs, rew, done = sets[i % nsets].step(acts[(i // nsets + i % nsets) % len(acts)])
                    dist.all_gather_into_tensor(g[0], obs)
                    dist.all_gather_into_tensor(g[1], rew)
                    dist.all_gather_into_tensor(g[2], done)
            else:
                for v in sets:
                    _, handle = v.backend.gather_create(world, rank)
                    handles = [None] * world
                    dist.all_gather_object(handles, handle)
                    v.backend.gather_open(b"".join(handles))
                dist.barrier()

                def step_fn(i):
                    sets[i % nsets].backend.step_gather(acts[(i // nsets + i % nsets) % len(acts)])
                    dist.all_reduce(flag)
            ms = time_rotating(step_fn, max(steps, 50), max(30, warmup), dist, dev)
            rec[f"{mode}_ms"] = ms
            torch.cuda.synchronize()
            dist.barrier()
            for v in sets:
                v.close()
            del sets
            if mode == "nccl":
                del g
            torch.cuda.empty_cache()
        shard = E * (n * L * 4 + n * 4 + 1)
        rec.update({"env_steps_per_s": world * E / (rec["fused_ms"] * 1e-3), "env_steps_per_s_nccl": world * E / (rec["nccl_ms"] * 1e-3),
                    "gathered_bytes_per_rank": world * shard, "nvlink_egress_bytes_per_gpu": (world - 1) * shard,
                    "nvlink_gbs": (world - 1) * shard / (rec["fused_ms"] * 1e-3) / 1e9,
                    "peer_copy_peak_gbs": peer, "peer_copy_peak_source": "measured: 256 MiB torch copy into the next GPU" if peer else None,
                    "frac_of_peer_copy_peak": ((world - 1) * shard / (rec["fused_ms"] * 1e-3) / 1e9 / peer) if peer else None})
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tiny-4ag", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=None, help="worlds per GPU (default: the workload's)")
    ap.add_argument("--p-forward", type=float, default=None,
                    help="action distribution: P(FORWARD) = p, the other four actions share the rest (default: uniform); "
                         "0.68 is the dense-collision stress of SURVEY.md section 8(d)")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-warmup", type=int, default=8)
    ap.add_argument("--ref-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records (other_workloads at N=1, gathered configs 3 and 5 at N>1)")
    ap.add_argument("--gather", nargs="?", const="nccl", default=None, choices=["nccl", "fused"],
                    help="all-gather obs/reward/done across ranks every step: NCCL all_gather_into_tensor, or the fused "
                         "push (the step kernel's TMA stores write into every rank's gathered tensor over NVLink)")
    ap.add_argument("--streams", type=int, default=2,
                    help="additionally measure the same K launches dealt round-robin to this many CUDA streams "
                         "(independent world sets in flight concurrently; reported as `concurrent_streams`, not as `value`)")
    ap.add_argument("--l2", default="rotate", choices=["rotate", "flush", "none"],
                    help="cold-L2 protocol: rotate over 16 world sets (default), flush 512 MiB between launches, or none")
    ap.add_argument("--flush-with", default="library", choices=["library", "torch"],
                    help="L2 flush kernel: the library's (same shared-memory carve-out as the step kernel) or torch fill_")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import rware_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    env_id, overrides, E = WORKLOADS[args.workload]
    E = args.envs or E
    NSETS = 16 if args.l2 == "rotate" else 1
    sets = [rware_b200.make_vec(env_id, E, auto_reset=True, device=dev, **overrides) for _ in range(NSETS)]
    vec = sets[0]
    n, L, A = vec.n_agents, vec.obs_len, 1 + vec.msg_bits
    info = vec.info
    for k, v in enumerate(sets):  # world e of set k on rank r is seeded with (k*world + r)*E + e
        v.reset(seed=(k * world + rank) * E)

    # synthetic input: a ring of pre-generated uniform action tensors, resident in HBM
    RING = 16
    actions = [torch.empty((E, n, A), dtype=torch.int32, device=dev) for _ in range(RING)]
    for i, a in enumerate(actions):
        vec.backend.random_actions(a, 12345 + rank, i)
        if args.p_forward is not None:  # dense-collision stress: FORWARD with probability p
            g = torch.Generator(device=dev).manual_seed(777 + 16 * rank + i)
            u = torch.rand((E, n), device=dev, generator=g)
            other = torch.tensor([0, 2, 3, 4], device=dev, dtype=torch.int32)[torch.randint(0, 4, (E, n), device=dev, generator=g)]
            a[..., 0] = torch.where(u < args.p_forward, torch.ones_like(other), other)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    def do_flush(i):
        if args.flush_with == "torch":
            flush.fill_(i & 0xff)
        else:
            vec.backend.flush_l2(flush, i & 0xff)

    gathered = None
    fused = None
    if args.gather == "nccl" and world > 1:
        gathered = (torch.empty((world * E, n, L), dtype=torch.float32, device=dev),
                    torch.empty((world * E, n), dtype=torch.float32, device=dev),
                    torch.empty((world * E,), dtype=torch.uint8, device=dev))
    if args.gather == "fused" and world > 1:
        fused = []
        for v in sets:  # every world set gets its own peer-mapped gathered buffers
            tensors, handle = v.backend.gather_create(world, rank)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            v.backend.gather_open(b"".join(handles))
            fused.append(tensors)
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.barrier()

    def act_of(i):
        # every world set must see a DIFFERENT action tensor at each of its steps (with i % RING and RING == NSETS a set would be fed
        # the same tensor for ever: agents repeating one action, a much cheaper rollout than the random one this bench claims)
        return actions[(i // NSETS + i % NSETS) % RING]

    def one_step(i):
        if fused is not None:
            sets[i % NSETS].backend.step_gather(act_of(i))  # lands in slot (call count & 1) of every rank's tensors
            dist.all_reduce(flag)  # the barrier: all ranks' pushes have landed
            return
        obs, rew, done = sets[i % NSETS].step(act_of(i))
        if gathered is not None:
            dist.all_gather_into_tensor(gathered[0], obs)
            dist.all_gather_into_tensor(gathered[1], rew)
            dist.all_gather_into_tensor(gathered[2], done)

    for i in range(max(3, args.warmup)):
        one_step(i)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load = time.perf_counter()
    i_load = 0
    while time.perf_counter() - t_load < 0.6:  # ~0.6 s of the same load before the timed region (every rank: the gather is
        for _ in range(500):                    # a collective) so that the nvidia-smi clock samples are taken under load
            one_step(i_load)
            i_load += 1
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    if args.l2 == "flush":
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            do_flush(i)
            starts[i].record()
            one_step(args.warmup + i)
            stops[i].record()
        torch.cuda.synchronize()
        total_ms = float(sum(s.elapsed_time(e) for s, e in zip(starts, stops)))
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            one_step(args.warmup + i)
        ev1.record()
        torch.cuda.synchronize()
        total_ms = float(ev0.elapsed_time(ev1))
    if dist:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0

    # optional: the same K launches dealt round-robin to several streams, i.e. independent world sets stepped concurrently
    # (an asynchronous vector environment).  With CTAs of at most half an SM (RWARE_B200_WARPS_PER_CTA=7) the serial head
    # of one launch (state fetch + logic) overlaps the observation write phase of the other.
    multi = None
    if args.streams > 1 and args.l2 == "rotate" and fused is None and gathered is None:
        streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):  # first pass warms the streams up
            m0.record()
            for st in streams:
                st.wait_event(m0)
            for i in range(args.steps):
                with torch.cuda.stream(streams[i % args.streams]):
                    one_step(args.warmup + i)
            for st in streams:
                e = torch.cuda.Event()
                e.record(st)
                torch.cuda.current_stream(dev).wait_event(e)
            m1.record()
            torch.cuda.synchronize()
        multi_ms = float(m0.elapsed_time(m1))
        if dist:  # max over ranks, like the headline
            mt = torch.tensor([multi_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
            multi_ms = float(mt.item())
        multi = {"streams": args.streams, "ms_per_step": multi_ms / args.steps, "value": world * E * args.steps / (multi_ms * 1e-3),
                 "tiles_per_cta": int(info["warps_per_cta"]),
                 "note": "K launches over independent world sets, round-robin on the streams, one event pair around all of them"}

    # one world set stepped back to back (its 83 MB working set stays in the 126 MB L2): upper bound, for reference
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        vec.step(actions[i % RING])
    s1.record()
    torch.cuda.synchronize()
    warm_ms = s0.elapsed_time(s1)
    clocks = sampler.stop() if rank == 0 else None

    # the observation gather alone (rw_obs: state read + obs write), same flush + per-launch event protocol
    n_obs = min(args.steps, 64)
    if args.l2 == "flush":
        pairs = []
        for i in range(n_obs):
            do_flush(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            vec.observations()
            b.record()
            pairs.append((a, b))
        torch.cuda.synchronize()
        obs_kernel_ms = sum(a.elapsed_time(b) for a, b in pairs) / n_obs
    else:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(NSETS):
            sets[i].observations()
        a.record()
        for i in range(n_obs):
            sets[i % NSETS].observations()
        b.record()
        torch.cuda.synchronize()
        obs_kernel_ms = a.elapsed_time(b) / n_obs

    # end to end through the host-buffer C-ABI entry point: page-locked host buffers in and out (allocated next to the GPU
    # with the library's own allocator), H2D actions + kernel + D2H observations / rewards / done + sync inside the timed region
    e2e_steps = args.e2e_steps
    be = vec.backend
    h_acts = [be.host_alloc((E, n, A), torch.int32) for _ in range(8)]  # fresh actions every step, like the device loop
    for k, h in enumerate(h_acts):
        h.copy_(actions[k % RING].cpu())
    h_act = h_acts[0]
    h_obs = be.host_alloc((E, n, L), torch.float32)
    h_rew = be.host_alloc((E, n), torch.float32)
    h_done = be.host_alloc((E,), torch.uint8)
    def time_e2e(mode, repeats=1):
        be.set_host_transport(mode)
        for k in range(max(3, args.e2e_warmup)):  # also lets the copy-engine share settle
            be.step_host(h_acts[k % 8], h_obs, h_rew, h_done)
        runs = []
        for _ in range(repeats):
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(e2e_steps):
                be.step_host(h_acts[k % 8], h_obs, h_rew, h_done)  # synchronises internally
            runs.append(time.perf_counter() - t0)
        return sorted(runs)[len(runs) // 2], runs  # the median repeat (other tenants of the host share its memory system)
    e2e_plain_s, _ = time_e2e(0)   # float32 observations over PCIe
    e2e_s, e2e_runs = time_e2e(1, repeats=3)  # default transport: bits / bytes over PCIe, expanded by host threads, copy engine sharing the rows
    dma_share = be.host_dma_share()
    compressed = (not overrides.get("normalised_coordinates")) and E * n * L >= (1 << 20)
    bit_rows = compressed and (overrides.get("observation_type", 1) < 2)

    # transport rooflines of the end-to-end call: the same device->host bytes moved by a bare pinned-memory copy, and the rate at
    # which the library's host pool streams into the caller's buffer when there is nothing to decode
    d2h_ms = []
    for _ in range(5):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        h_obs.copy_(vec._obs, non_blocking=True)
        h_rew.copy_(vec._reward, non_blocking=True)
        h_done.copy_(vec._done, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        d2h_ms.append(c0.elapsed_time(c1))
    d2h_ms = min(d2h_ms)
    if dist:
        dist.barrier()
    host_stream_gbs, host_threads = be.host_stream_probe(h_obs, repeats=3)  # every rank at once, like the e2e loop
    numa = {"gpu_node": be.gpu_numa_node(), "obs_buffer_node": be.host_numa_node(h_obs)}

    # sub-records: free the headline's world sets first
    extras = {}
    if not args.no_extras and args.workload == "tiny-4ag" and args.gather is None:
        be.close()
        for v in sets[1:]:
            v.close()
        del actions, flush
        torch.cuda.empty_cache()
        peak_gbs, _ = load_peaks()
        if world == 1:
            extras["other_workloads"] = other_workloads_records(dev, rank, peak_gbs, min(args.steps, 100), args.warmup)
        else:
            extras["gathered"] = gathered_records(dist, dev, rank, local_rank, world, min(args.steps, 40), args.warmup)

    tot = torch.tensor([total_ms, warm_ms, e2e_s, obs_kernel_ms, e2e_plain_s], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, warm_ms, e2e_s, obs_kernel_ms, e2e_plain_s = (float(x) for x in tot.tolist())
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    kernel_ms = total_ms / args.steps  # per-GPU: one launch per step (max over ranks of the sums)
    b_step = int(info["algorithmic_bytes_step"])
    b_packed = int(info["packed_bytes_step"])
    achieved = b_step * E / (kernel_ms * 1e-3) / 1e9
    # dram__bytes_read + dram__bytes_write per launch: not measurable from inside this process -- taken from the committed ncu
    # capture of this same rotating loop (profiles/r2_step_traffic.json: one pass per launch, caches untouched, launches 60..75)
    traffic, traffic_source = None, None
    prof = os.path.join(ROOT, "profiles", "r2_step_traffic.json")
    if os.path.exists(prof) and args.workload == "tiny-4ag" and E == 65536 and args.l2 == "rotate":
        with open(prof) as f:
            rec = json.load(f)
        traffic = rec.get("dram_bytes_per_launch")
        traffic_source = ("profiles/r2_step_traffic.json: ncu --cache-control none, launches 60..75 of this bench loop (steady state, "
                          f"read {rec.get('dram_bytes_read_per_launch', 0) / 1e6:.1f} MB + write {rec.get('dram_bytes_write_per_launch', 0) / 1e6:.1f} MB "
                          "per launch); a file value, not measured in this run")
    cb = None
    if not args.no_cpu_baseline:
        cb = cpu_baseline(env_id, overrides, target_seconds=12.0, n_worlds=E)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "independent_worlds")}
    out = {
        "metric": "env-steps/s", "value": world * E * args.steps / (total_ms * 1e-3), "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": kernel_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32 (state) / float32 (obs)",
        "data": "synthetic",
        "config": {"workload": f"{env_id} {overrides or ''}".strip() + f", {E} worlds per GPU, random-action rollout"
                               + (f" with P(FORWARD)={args.p_forward}" if args.p_forward is not None else " (uniform actions)") + ", auto-reset",
                   "worlds_per_gpu": E, "n_agents": n, "obs_len": L, "parallelism": f"dp{world} (independent worlds, no exchange"
                   + (f", obs/reward/done all-gathered ({args.gather})" if args.gather and world > 1 else "") + ")",
                   "l2": {"rotate": f"inputs larger than L2: launches rotate over {NSETS} independent world sets "
                                    f"({NSETS} x {(int(info['packed_bytes_step']) * E) >> 20} MiB touched per launch)",
                          "flush": f"flushed between timed launches (512 MiB write, {args.flush_with} kernel)",
                          "none": "not flushed"}[args.l2],
                   "timing": {"rotate": "one CUDA event pair around the K back-to-back launches; max over ranks",
                              "flush": "per-launch CUDA event pairs, summed; max over ranks",
                              "none": "one CUDA event pair around the K back-to-back launches; max over ranks"}[args.l2]},
        "agent_steps_per_s": world * E * n * args.steps / (total_ms * 1e-3),
        "value_single_set_l2_resident": world * E * args.steps / (warm_ms * 1e-3),
        "sequential_rollout_of_one_world_set": {"value": world * E * args.steps / (warm_ms * 1e-3), "ms_per_step": warm_ms / args.steps,
                                                "note": "ONE world set stepped back to back on one stream (step k+1 needs step k's state: the "
                                                        "launches serialise, no overlap across launches; its 83 MB working set stays in the L2)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src, "kernel": "rw_main_kernel<STEP>",
                     "algorithmic_bytes_per_env_step": b_step, "layout_bytes_per_env_step": b_packed,
                     "achieved_layout_bytes": b_packed * E / (kernel_ms * 1e-3) / 1e9,
                     "kernel_us": kernel_ms * 1e3},
        "obs_gather": {"kernel": "rw_main_kernel<OBS>", "kernel_us": obs_kernel_ms * 1e3,
                       "algorithmic_bytes_per_env": int(info["algorithmic_bytes_obs"]),
                       "achieved_gbs": int(info["algorithmic_bytes_obs"]) * E / (obs_kernel_ms * 1e-3) / 1e9,
                       "frac_of_hbm_peak": int(info["algorithmic_bytes_obs"]) * E / (obs_kernel_ms * 1e-3) / 1e9 / peak,
                       "layout_bytes_per_env": int(info["packed_bytes_obs"])},
        "e2e": dict(e2e_record(world, E, n, A, L, e2e_steps, e2e_s, e2e_plain_s, d2h_ms, dma_share, compressed, bit_rows,
                               host_stream_gbs, host_threads, numa),
                    repeats_ms_per_step=[r / e2e_steps * 1e3 for r in e2e_runs],
                    repeat_rule=f"median of {len(e2e_runs)} repeats of {e2e_steps} steps (this rank's clock; other ranks start together)"),
        "gpu_launches": args.steps,
        "clocks": clocks,
        "wall_s_timed_region": t_wall,
    }
    if cb:
        out["cpu_baseline"] = cb
    if multi:
        out["concurrent_streams"] = multi
    out.update(extras)
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
