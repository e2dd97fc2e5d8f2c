"""World-size-2 test of the sharded path on CPU (gloo + the host twin): sharding must not change results, and the
gathered tensors must equal a single-process run of the whole batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=4, request_queue_size=4, max_steps=25)
TOTAL, STEPS = 41, 60  # ragged: 21 + 20 worlds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _actions(t, total):
    return np.random.default_rng(1000 + t).integers(0, 5, size=(total, KW["n_agents"])).astype(np.int32)


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hostsim.backend import HostSimBackend
    from rware_b200.distributed import ShardedVecWarehouse
    env = ShardedVecWarehouse(TOTAL, auto_reset=True, _backend_factory=HostSimBackend, **KW)
    env.reset(seed=7)
    for t in range(STEPS):
        a = _actions(t, TOTAL)[env.offset: env.offset + env.local_envs]
        obs, rew, done = env.step(torch.from_numpy(a), gather=True)
    if rank == 0:
        torch.save((obs.clone(), rew.clone(), done.clone()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    from hostsim.backend import HostSimBackend
    import rware_b200
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g_obs, g_rew, g_done = torch.load(out)
    ref = rware_b200.VecWarehouse(TOTAL, auto_reset=True, _backend_factory=HostSimBackend, **KW)
    ref.reset(seed=7)
    for t in range(STEPS):
        obs, rew, done = ref.step(torch.from_numpy(_actions(t, TOTAL)))
    assert torch.equal(g_obs, obs) and torch.equal(g_rew, rew) and torch.equal(g_done, done)
    assert g_obs.shape == (TOTAL, KW["n_agents"], ref.obs_len)
