"""Multi-GPU data parallelism for the RWARE step path: one process per GPU, worlds sharded in contiguous blocks.

Worlds are independent (the reference has no cross-world state, rware/warehouse.py:804-946), so `step` needs NO
exchange between ranks.  World e of the global batch is seeded with `seed + e` whatever the number of ranks, so results
do not depend on the sharding.  The single optional collective is an all-gather of the observation / reward / done
tensors for callers that want one tensor of the whole batch (BASELINE.json north_star): `step(..., gather=True)`.

    torchrun --nproc-per-node 8 ... :
        env = ShardedVecWarehouse(262144, shelf_columns=3, column_height=8, shelf_rows=2, n_agents=4,
                                  request_queue_size=4, max_steps=500)
        obs = env.reset(seed=0)                    # local shard [E/G, n, L]
        obs, rew, done = env.step(local_actions, gather=True)   # gathered [E, n, L], [E, n], [E]
"""
import numpy as np
import torch
import torch.distributed as dist

from .vec import VecWarehouse


class ShardedVecWarehouse:
    def __init__(self, total_envs, *args, process_group=None, _backend_factory=None, **kwargs):
        if not dist.is_initialized():
            raise RuntimeError("ShardedVecWarehouse needs torch.distributed to be initialised (one process per GPU)")
        self.group = process_group
        self.rank = dist.get_rank(process_group)
        self.world_size = dist.get_world_size(process_group)
        self.total_envs = int(total_envs)
        # contiguous blocks; the first (total % G) ranks get one extra world
        base, extra = divmod(self.total_envs, self.world_size)
        self.local_envs = base + (1 if self.rank < extra else 0)
        self.offset = self.rank * base + min(self.rank, extra)
        if self.local_envs < 1:
            raise ValueError("fewer worlds than ranks")
        self.uniform = extra == 0
        self.local = VecWarehouse(self.local_envs, *args, _backend_factory=_backend_factory, **kwargs)
        self.device = self.local.device
        self.n_agents, self.obs_len = self.local.n_agents, self.local.obs_len
        self._gathered = None

    def reset(self, seed=None, mask=None):
        if seed is not None and np.isscalar(seed):
            seed = np.uint64(seed) + np.uint64(self.offset) + np.arange(self.local_envs, dtype=np.uint64)
        return self.local.reset(seed=seed, mask=mask)

    def _buffers(self):
        if self._gathered is None:
            E, n, L = self.total_envs, self.n_agents, self.obs_len
            self._gathered = (torch.empty((E, n, L), dtype=torch.float32, device=self.device),
                              torch.empty((E, n), dtype=torch.float32, device=self.device),
                              torch.empty((E,), dtype=torch.uint8, device=self.device))
        return self._gathered

    def gather(self, obs, reward, done):
        """One all-gather per tensor (NCCL all_gather_into_tensor on GPUs; list all_gather for ragged shards / gloo)."""
        g_obs, g_rew, g_done = self._buffers()
        for src, dst in ((obs, g_obs), (reward, g_rew), (done, g_done)):
            if self.uniform and src.is_cuda:
                dist.all_gather_into_tensor(dst, src, group=self.group)
            else:
                sizes = [self.total_envs // self.world_size + (1 if r < self.total_envs % self.world_size else 0)
                         for r in range(self.world_size)]
                parts = list(dst.split(sizes, dim=0))
                if self.uniform:
                    dist.all_gather(parts, src.contiguous(), group=self.group)
                else:  # ragged shards: pad to the largest shard
                    m = max(sizes)
                    pad = torch.zeros((m,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
                    pad[: src.shape[0]] = src
                    bufs = [torch.empty_like(pad) for _ in sizes]
                    dist.all_gather(bufs, pad, group=self.group)
                    for p, b, sz in zip(parts, bufs, sizes):
                        p.copy_(b[:sz])
        return g_obs, g_rew, g_done

    def enable_fused_gather(self):
        """Sets up the fused push all-gather (CUDA IPC peer mappings over NVLink): afterwards `step(..., gather=True)`
        lets the step kernel's TMA stores write every observation chunk into all ranks' gathered tensors, and the only
        collective left is a 4-byte all-reduce that acts as the barrier.  Needs uniform shards and one process per GPU.
        The gathered tensors are double-buffered: step k returns slot k & 1, valid until step k+2 is issued; consume them on
        the current stream before the next step (include/rware_b200.h, rw_step_gather)."""
        if not self.uniform:
            raise ValueError("fused gather needs total_envs divisible by the number of ranks")
        tensors, handle = self.local.backend.gather_create(self.world_size, self.rank)
        handles = [None] * self.world_size
        dist.all_gather_object(handles, handle, group=self.group)
        self.local.backend.gather_open(b"".join(handles))
        self._fused = tensors
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        dist.barrier(group=self.group)

    def step(self, actions, gather=False):
        if gather and getattr(self, "_fused", None) is not None:
            a = actions
            if not (isinstance(a, torch.Tensor) and a.dtype == torch.int32 and a.device == self.device and a.is_contiguous()):
                a = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
            slot = self.local.backend.step_gather(a)
            dist.all_reduce(self._flag, group=self.group)  # stream-ordered barrier: every rank's pushes have landed
            return self._fused[slot]
        obs, rew, done = self.local.step(actions)
        if gather:
            return self.gather(obs, rew, done)
        return obs, rew, done

    def get_state(self):
        return self.local.get_state()

    def close(self):
        self.local.close()
