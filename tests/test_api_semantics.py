"""C-ABI / VecWarehouse semantics that the trajectory replays do not isolate: masked resets, state round trips,
stepping past `done` without auto-reset (the reference keeps stepping and keeps reporting done), determinism of seeds."""
import numpy as np
import pytest
import torch

import oracle_py
import rware_b200
from conftest import backend_factory, backend_params
from rware_b200.warehouse import Action, RewardType, Warehouse

KW = dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=3, request_queue_size=3, max_steps=12)


def _vec(backend, E, **extra):
    return rware_b200.VecWarehouse(E, _backend_factory=backend_factory(backend), **{**KW, **extra})


@pytest.mark.parametrize("backend", backend_params())
def test_masked_reset_only_touches_selected_worlds(backend):
    E = 70
    vec = _vec(backend, E)
    vec.reset(seed=1)
    rs = np.random.default_rng(0)
    for _ in range(5):
        vec.step(torch.from_numpy(rs.integers(0, 5, size=(E, 3)).astype(np.int32)))
    before = {k: v.clone() for k, v in vec.get_state().items()}
    mask = torch.zeros(E, dtype=torch.uint8)
    mask[[0, 31, 32, 69]] = 1
    obs = vec.reset(mask=mask).clone()
    after = vec.get_state()
    keep = ~mask.bool()
    for k in before:
        assert torch.equal(before[k].cpu()[keep], after[k].cpu()[keep]), k
    assert bool((after["cur_steps"].cpu()[mask.bool()] == 0).all()) and bool((before["cur_steps"].cpu()[keep] == 5).all())
    assert torch.equal(obs, vec.observations())  # the reset returns every world's current observation
    # the reset worlds continued their own generator streams: compare with the oracle
    for e in (0, 69):
        o = oracle_py.OracleEnv(vec.highways, vec.goals, 3, 0, 1, 3, None, 12, 1)
        o.reset(1 + e)
        o.set_state(rng=before["rng"][e].cpu().numpy().astype(np.int64).astype(np.uint32))
        assert np.array_equal(o.reset(), obs[e].cpu().numpy())


@pytest.mark.parametrize("backend", backend_params())
def test_state_round_trip_and_checkpoint_resume(backend):
    E = 40
    a, b = _vec(backend, E), _vec(backend, E)
    a.reset(seed=5)
    rs = np.random.default_rng(3)
    acts = [torch.from_numpy(rs.integers(0, 5, size=(E, 3)).astype(np.int32)) for _ in range(20)]
    for t in range(8):
        a.step(acts[t])
    b.set_state(**{k: v.cpu() for k, v in a.get_state().items()})  # "checkpoint" -> fresh handle
    sa, sb = a.get_state(), b.get_state()
    assert all(torch.equal(sa[k].cpu(), sb[k].cpu()) for k in sa)
    for t in range(8, 20):
        oa, ra, da = a.step(acts[t])
        ob, rb, db = b.step(acts[t])
        assert torch.equal(oa.cpu(), ob.cpu()) and torch.equal(ra.cpu(), rb.cpu()) and torch.equal(da.cpu(), db.cpu())


@pytest.mark.parametrize("backend", backend_params())
def test_stepping_past_done_without_auto_reset(backend):
    """warehouse.py:929-941: counters keep counting, done stays True, nothing resets."""
    vec = _vec(backend, 4)
    vec.reset(seed=2)
    o = oracle_py.OracleEnv(vec.highways, vec.goals, 3, 0, 1, 3, None, 12, 1)
    o.reset(2)
    rs = np.random.default_rng(8)
    for t in range(20):
        a = rs.integers(0, 5, size=3).astype(np.int32)
        obs, rew, done = vec.step(torch.from_numpy(np.broadcast_to(a, (4, 3)).copy()))
        oo, orw, od = o.step(a)
        assert bool(done[0]) == od == (t >= 11)
        assert np.array_equal(obs[0].cpu().numpy(), oo)
    assert int(vec.get_state()["cur_steps"][0]) == 20


@pytest.mark.parametrize("backend", backend_params())
def test_seed_array_and_scalar_forms(backend):
    a, b = _vec(backend, 6), _vec(backend, 6)
    oa = a.reset(seed=10).clone()
    ob = b.reset(seed=np.arange(10, 16, dtype=np.uint64)).clone()
    assert torch.equal(oa, ob)
    oc = b.reset(seed=np.full(6, 10, np.uint64))
    assert all(torch.equal(oc[0], oc[i]) for i in range(6)) and torch.equal(oc[0], oa[0])


def test_optional_registrations_resolve_like_the_reference():
    """image_registration() / full_registration() (rware/__init__.py:42-172): ids are unknown until enabled, then resolve to
    the kwargs the reference registers (tests/golden/registry_ids.json: all 912 image ids, 1 500 sampled full ids)."""
    import json
    import os
    from rware_b200 import registration as reg
    from rware_b200.warehouse import ObservationType, RewardType
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registry_ids.json")) as f:
        fx = json.load(f)
    cols = fx["columns"]
    saved = set(reg._enabled)
    reg._enabled.clear()
    try:
        some_img, some_full = next(iter(fx["image"])), next(iter(fx["full_sample"]))
        for id_ in (some_img, some_full):
            with pytest.raises(KeyError):
                reg.env_kwargs(id_)

        def check(table):
            for id_, row in table.items():
                want = dict(zip(cols, row))
                want["reward_type"] = RewardType(want["reward_type"])
                want["observation_type"] = ObservationType(want["observation_type"])
                assert reg.env_kwargs(id_) == want, id_
        rware_b200.image_registration()
        check(fx["image"])
        with pytest.raises(KeyError):
            reg.env_kwargs(some_full) if some_full not in fx["image"] and "h-" in some_full else (_ for _ in ()).throw(KeyError())
        rware_b200.full_registration()
        check(fx["full_sample"])
        for bad in ("rware-Nd-tiny-8h-2ag-v2", "rware-imgdict-2s-tiny-8h-2ag-v2", "rware-tiny-16h-2ag-v2", "rware-tiny-8h-20ag-v2",
                    "rware-5x3-8h-2ag-2req-indiv-v2", "rware-2x4-8h-2ag-2req-indiv-v2", "rware-2x3-8h-2ag-20req-indiv-v2"):
            with pytest.raises(KeyError):
                reg.env_kwargs(bad)
        assert reg.env_kwargs("rware-2s-large-8h-16ag-v2")["sensor_range"] == 2  # BASELINE.json configs[4]
    finally:
        reg._enabled.clear()
        reg._enabled.update(saved)


@pytest.mark.parametrize("backend", backend_params())
def test_observation_bytes_are_the_same_values(backend):
    """rw_obs_bytes / VecWarehouse.observation_bytes: the uint8 image of the float32 observations (exact integers 0..255)."""
    E = 77
    vec = rware_b200.VecWarehouse(E, shelf_columns=3, column_height=8, shelf_rows=1, n_agents=3, msg_bits=2, sensor_range=2,
                                  request_queue_size=3, max_steps=20, auto_reset=True, _backend_factory=backend_factory(backend))
    obs = vec.reset(seed=1)
    g = torch.Generator().manual_seed(0)
    for _ in range(12):
        b = vec.observation_bytes()
        assert b.dtype == torch.uint8 and b.shape == obs.shape and torch.equal(b.to(torch.float32), obs)
        acts = torch.cat([torch.randint(0, 5, (E, 3, 1), generator=g), torch.randint(0, 2, (E, 3, 2), generator=g)], dim=2).to(torch.int32)
        obs, _, _ = vec.step(acts.to(vec.device))
    out = torch.zeros_like(obs, dtype=torch.uint8)
    assert vec.observation_bytes(obs, out) is out and torch.equal(out.to(torch.float32), obs)
    norm = rware_b200.VecWarehouse(4, shelf_columns=3, column_height=8, shelf_rows=1, n_agents=2, normalised_coordinates=True,
                                   _backend_factory=backend_factory(backend))
    norm.reset(seed=0)
    with pytest.raises(RuntimeError):
        norm.observation_bytes()


@pytest.mark.parametrize("backend", backend_params())
def test_out_of_range_state_is_rejected_or_clamped(backend):
    """A coordinate outside the grid: the facade raises IndexError like the reference's _recalc_grid (rware/warehouse.py:749-755);
    the batched set_state clamps it into range and counts the world (rw_error_count) instead of letting later kernels index out
    of bounds."""
    env = Warehouse(3, 8, 1, 2, 0, 1, 2, None, None, RewardType.GLOBAL, _backend_factory=backend_factory(backend))
    env.reset(seed=4)
    env.agents[0].x = env.grid_size[1] + 3
    with pytest.raises(IndexError):
        env.step([Action.NOOP, Action.NOOP])
    vec = rware_b200.VecWarehouse(40, 3, 8, 1, 2, request_queue_size=2, _backend_factory=backend_factory(backend))
    vec.reset(seed=0)
    st = {k: v.clone() for k, v in vec.get_state().items()}
    H, W = vec.grid_size
    st["agent_x"][3, 0] = 200
    st["agent_y"][5, 1] = -4
    st["shelf_x"][7, 2] = W
    st["queue"][9, 0] = vec.n_shelves + 50
    st["agent_carry"][11, 1] = 8000
    before = vec.error_count()
    vec.set_state(**st)
    assert vec.error_count() - before == 5
    back = vec.get_state()
    assert int(back["agent_x"][3, 0]) == W - 1 and int(back["agent_y"][5, 1]) == 0 and int(back["shelf_x"][7, 2]) == W - 1
    assert int(back["queue"][9, 0]) == vec.n_shelves and int(back["agent_carry"][11, 1]) == vec.n_shelves
    obs, rew, done = vec.step(torch.ones((40, 2), dtype=torch.int32))  # and stepping such worlds is safe
    assert obs.shape == (40, 2, vec.obs_len)


@pytest.mark.gpu
def test_handles_of_different_batch_sizes_coexist():
    """The dynamic-shared-memory limit is a property of the kernel FUNCTION, shared by all handles of a process: creating a small
    batch (few tiles per CTA) after a large one must not lower it under the large one's feet."""
    big = rware_b200.make_vec("rware-tiny-4ag-v1", 8192)     # 2 tiles per CTA
    big.reset(seed=0)
    small = rware_b200.make_vec("rware-tiny-4ag-v1", 64)      # 1 tile per CTA, same kernel function
    small.reset(seed=0)
    a = torch.ones((8192, 4), dtype=torch.int32, device="cuda")
    obs, _, _ = big.step(a)
    obs2, _, _ = small.step(a[:64].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(obs[:64], obs2)
