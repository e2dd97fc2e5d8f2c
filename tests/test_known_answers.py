"""End-to-end known answers of the UNMODIFIED reference (tests/golden/known_answers.json = SURVEY.md App. E): registered
ids, `reset(seed=0)`, 500 steps of `default_rng(12345)` actions, SHA-256 over every float32 observation, float64 reward
and done flag.  No oracle involved: the product path alone has to reproduce the reference's hashes, which covers the
numpy-compatible reset sampler, the step, and the observation gather in one go.  Also: full-size batches (65 536 worlds)
checked through size-independent properties and against the oracle on every world."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import rware_b200
from conftest import backend_factory, backend_params

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
    KNOWN = json.load(f)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("key", sorted(KNOWN))
def test_reference_sha256_through_the_product_path(key, backend):
    rec = KNOWN[key]
    kw = rware_b200.env_kwargs(key.split("+")[0], **rec["extra"])
    E = 3  # three identical worlds: they must stay identical and all match the reference
    vec = rware_b200.VecWarehouse(E, _backend_factory=backend_factory(backend), **kw)
    vec.seed(np.zeros(E, np.uint64))
    obs = vec.reset().cpu().numpy()
    hs = [hashlib.sha256(obs[e].tobytes()) for e in range(E)]
    acts = np.random.default_rng(12345).integers(0, 5, size=(500, kw["n_agents"]))
    total = 0.0
    for t in range(500):
        a = torch.from_numpy(np.broadcast_to(acts[t].astype(np.int32), (E, kw["n_agents"])).copy())
        obs, rew, done = vec.step(a)
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy().astype(np.float64), done.cpu().numpy()
        total += rew[0].sum()
        for e in range(E):
            hs[e].update(obs[e].tobytes() + rew[e].tobytes() + bytes([int(done[e])]))
    assert [h.hexdigest() for h in hs] == [rec["sha256"]] * E
    assert total == rec["reward_sum"]
    st = vec.get_state()
    final = [[int(st["agent_x"][0, i]), int(st["agent_y"][0, i]), int(st["agent_dir"][0, i]), int(st["agent_carry"][0, i])]
             for i in range(kw["n_agents"])]
    assert final == rec["final_agents"]


@pytest.mark.gpu
def test_full_size_batch_properties_and_full_population_oracle_parity():
    """BASELINE.json configs[1]: rware-tiny-4ag x 65 536 worlds on one GPU.  EVERY world is compared with the oracle (a batched,
    multi-threaded instance of it) at every step -- a tile-boundary or CTA-tail bug anywhere in the batch fails this."""
    import oracle_py
    E, T = 65536, 40
    vec = rware_b200.make_vec("rware-tiny-4ag-v1", E, auto_reset=True)
    obs0 = vec.reset(seed=0).clone()
    orc = oracle_py.OracleBatch(oracle_py.OracleEnv(vec.highways, vec.goals, 4, 0, 1, 4, None, 500, 1), E, auto_reset=True)
    assert np.array_equal(orc.reset(np.arange(E, dtype=np.uint64)), obs0.cpu().numpy())
    g = torch.Generator(device="cuda").manual_seed(5)
    ref = rware_b200.make_vec("rware-tiny-4ag-v1", E, auto_reset=True)  # second instance: determinism
    ref.reset(seed=0)
    for t in range(T):
        a = torch.randint(0, 5, (E, 4), generator=g, device="cuda", dtype=torch.int32)
        obs, rew, done = vec.step(a)
        obs2, rew2, done2 = ref.step(a)
        assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(done, done2)
        # structural invariants of every observation row (App. C): one-hot direction blocks, 0/1 flags
        o = obs.view(E * 4, -1)
        assert bool(((o[:, 3:7].sum(1) == 1) & (o[:, 2] <= 1) & (o[:, 7] <= 1)).all())
        cells = o[:, 8:].view(E * 4, 9, 7)
        assert bool((cells[:, :, 1:5].sum(2) == 1).all()) and bool((cells[:, 4, 0] == 1).all())
        assert bool(((cells == 0) | (cells == 1)).all())
        assert bool((cells[:, :, 6] <= cells[:, :, 5]).all())  # requested implies shelf
        oo, orw, od = orc.step(a.cpu().numpy())
        oh, rh, dh = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        if not (np.array_equal(oo, oh) and np.array_equal(orw, rh) and np.array_equal(od, dh)):
            bad = np.nonzero((oo != oh).reshape(E, -1).any(1) | (orw != rh).any(1) | (od != dh))[0]
            raise AssertionError(f"step {t}: {bad.size} worlds differ from the oracle, first {bad[:8].tolist()} (tiles {(bad[:8] // 32).tolist()})")
    orc.close()
    st = vec.get_state()
    # no two agents of a world share a cell; carried shelves sit under their carriers
    cell = st["agent_y"] * vec.grid_size[1] + st["agent_x"]
    assert bool((torch.sort(cell, dim=1).values.diff(dim=1) != 0).all())
    carry = st["agent_carry"].long()
    has = carry > 0
    sx = torch.gather(st["shelf_x"], 1, (carry - 1).clamp(min=0))
    sy = torch.gather(st["shelf_y"], 1, (carry - 1).clamp(min=0))
    assert bool(((sx == st["agent_x"]) | ~has).all()) and bool(((sy == st["agent_y"]) | ~has).all())


@pytest.mark.gpu
@pytest.mark.parametrize("E,mode,share,kw", [
    (1000, 1, None, {}), (6000, 1, None, {}), (6000, 1, 0.0, {}), (6000, 1, 0.5, {}), (6000, 2, None, {}), (6000, 2, 0.3, {}),
    (6000, 0, None, {}), (5000, 1, None, {"normalised_coordinates": True}), (4099, 1, 0.37, {"msg_bits": 2}),
    (6007, 1, 0.2, {"observation_type": 2}), (6007, 1, 0.2, {"observation_type": 3, "image_observation_layers": (0, 1, 2, 3, 4, 5, 6)}),
    (3001, 1, None, {"sensor_range": 2})],
    ids=["small-batch-plain-copy", "bits-adaptive-share", "bits-all-host", "bits-half-dma", "bytes-adaptive-share", "bytes-fixed-share",
         "plain-transport", "normalised-falls-back", "bits-ragged-msg", "image-bits", "image-direction-layer-bytes", "bits-5x5-window"])
def test_host_buffer_entry_points_match_the_device_ones(E, mode, share, kw):
    """rw_reset_host / rw_step_host (page-locked host buffers, the end-to-end path of bench.py) against rw_reset / rw_step, for
    every transport: plain float32 copy, bits (one bit per flag + one byte per coordinate over PCIe, rows rebuilt by the host
    pool), bytes (one byte per value), each with the copy engine delivering a fixed or self-adjusting share of the rows as
    float32 at the same time.  The buffers the caller gets must be bit-identical in all of them."""
    a = rware_b200.make_vec("rware-small-4ag-v1", E, **kw)
    b = rware_b200.make_vec("rware-small-4ag-v1", E, **kw)
    b.backend.set_host_transport(mode)
    b.backend.set_host_dma_share(share)
    seeds = np.arange(E, dtype=np.uint64) + 9
    obs_d = a.reset(seed=seeds).cpu()
    h_obs = b.backend.host_alloc((E, 4, a.obs_len), torch.float32)  # on the GPU's NUMA node, 64-byte aligned
    h_rew = torch.empty((E, 4), dtype=torch.float32).pin_memory()
    h_done = torch.empty((E,), dtype=torch.uint8).pin_memory()
    b.backend.reset_host(seeds, h_obs)
    assert torch.equal(h_obs, obs_d)
    g = torch.Generator().manual_seed(1)
    for _ in range(25):
        acts = torch.cat([torch.randint(0, 5, (E, 4, 1), generator=g, dtype=torch.int32),
                          torch.randint(0, 2, (E, 4, a.msg_bits), generator=g, dtype=torch.int32)], dim=2).contiguous().pin_memory()
        obs_d, rew_d, done_d = a.step(acts.cuda())
        h_obs.fill_(-7.0)
        b.backend.step_host(acts, h_obs, h_rew, h_done)
        assert torch.equal(h_obs, obs_d.cpu()) and torch.equal(h_rew, rew_d.cpu()) and torch.equal(h_done, done_d.cpu())
    if share is None and mode and not kw.get("normalised_coordinates") and E >= 3000:
        assert 0.0 < b.backend.host_dma_share() <= 0.9
    b.close()


def _full_size_cases():
    from conftest import backend_params
    cases = []
    for bp in backend_params():
        backend = bp.values[0]
        for env_id, over, E_full, L_self in (("rware-small-4ag-v1", {}, 32768, 8), ("rware-medium-8ag-hard-v1", {}, 32768, 8),
                                            ("rware-large-16ag-v1", {"sensor_range": 2}, 16384, 8)):
            E = E_full if backend == "cuda" else 70  # the host twin runs the same test body on a few tiles
            cases.append(pytest.param(env_id, over, E, backend, marks=bp.marks, id=f"{env_id.split('-')[1]}-{E}-{backend}"))
    return cases


@pytest.mark.parametrize("env_id,over,E,backend", _full_size_cases())
def test_other_baseline_configs_at_full_size(env_id, over, E, backend):
    """BASELINE.json configs[2..4] at their per-GPU sizes (small-4ag x 32 768, medium-8ag-hard x 32 768, large-16ag r=2 x 16 384):
    determinism across two instances, structural invariants of every observation row and of the state, and oracle parity of
    EVERY world at every step (batched multi-threaded oracle)."""
    import oracle_py
    from conftest import backend_factory
    T = 25
    kw = rware_b200.env_kwargs(env_id, **over)
    n, r = kw["n_agents"], kw["sensor_range"]
    make = lambda: rware_b200.make_vec(env_id, E, auto_reset=True, _backend_factory=backend_factory(backend), **over)  # noqa: E731
    vec, twin = make(), make()
    dev = vec.device
    obs0 = vec.reset(seed=7).clone()
    twin.reset(seed=7)
    orc = oracle_py.OracleBatch(oracle_py.OracleEnv(vec.highways, vec.goals, n, 0, r, kw["request_queue_size"], None, kw["max_steps"],
                                                    kw["reward_type"].value), E, auto_reset=True)
    assert np.array_equal(orc.reset(7 + np.arange(E, dtype=np.uint64)), obs0.cpu().numpy())
    g = torch.Generator(device=dev).manual_seed(11)
    K = (2 * r + 1) ** 2
    for t in range(T):
        a = torch.randint(0, 5, (E, n), generator=g, device=dev, dtype=torch.int32)
        obs, rew, done = vec.step(a)
        obs2, rew2, done2 = twin.step(a)
        assert torch.equal(obs, obs2) and torch.equal(rew, rew2) and torch.equal(done, done2)
        o = obs.view(E * n, -1)
        assert bool(((o[:, 3:7].sum(1) == 1) & (o[:, 2] <= 1) & (o[:, 7] <= 1)).all())         # App. C: self block
        cells = o[:, 8:].view(E * n, K, 7)
        assert bool((cells[:, :, 1:5].sum(2) == 1).all()) and bool((cells[:, K // 2, 0] == 1).all())  # one-hot dirs; self in the centre
        assert bool(((cells == 0) | (cells == 1)).all()) and bool((cells[:, :, 6] <= cells[:, :, 5]).all())
        assert bool(((rew == 0) | (rew == 1)).all())                                           # INDIVIDUAL rewards
        oo, orw, od = orc.step(a.cpu().numpy())  # every world, not a sample
        oh, rh, dh = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        if not (np.array_equal(oo, oh) and np.array_equal(orw, rh) and np.array_equal(od, dh)):
            bad = np.nonzero((oo != oh).reshape(E, -1).any(1) | (orw != rh).any(1) | (od != dh))[0]
            raise AssertionError(f"step {t}: {bad.size} worlds differ from the oracle, first {bad[:8].tolist()} (tiles {(bad[:8] // 32).tolist()})")
    orc.close()
    st = vec.get_state()
    cell = st["agent_y"] * vec.grid_size[1] + st["agent_x"]
    assert bool((torch.sort(cell, dim=1).values.diff(dim=1) != 0).all())                        # no two agents share a cell
    carry = st["agent_carry"].long()
    has = carry > 0
    sx = torch.gather(st["shelf_x"], 1, (carry - 1).clamp(min=0))
    sy = torch.gather(st["shelf_y"], 1, (carry - 1).clamp(min=0))
    assert bool(((sx == st["agent_x"]) | ~has).all()) and bool(((sy == st["agent_y"]) | ~has).all())
    q = torch.sort(st["queue"], dim=1).values
    assert bool((q.diff(dim=1) != 0).all()) and bool((q >= 1).all()) and bool((q <= vec.n_shelves).all())  # queue: distinct shelf ids
    assert bool((st["cur_steps"] == T).all())
