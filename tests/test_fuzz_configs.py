"""Configuration fuzzing: random layouts (custom ASCII maps and parametric ones), agent counts, sensor ranges, message
bits, queue sizes, reward types and termination limits; every world of a small batch is compared with the C oracle on
every step (observations, rewards, done, positions, queue, generator state).  Exercises the dynamic kernel instantiation
and both step forms far away from the registered geometries."""
import numpy as np
import pytest
import torch

import oracle_py
import rware_b200
from conftest import backend_factory, backend_params


def random_case(rs):
    if rs.random() < 0.5:
        H, W = int(rs.integers(3, 9)), int(rs.integers(3, 11))
        while True:
            cells = rs.choice(list("x.."), size=(H, W))
            k = int(rs.integers(1, 4))
            for _ in range(k):
                cells[int(rs.integers(H)), int(rs.integers(W))] = "g"
            layout = "\n".join("".join(row) for row in cells)
            n_shelves = int((cells == "x").sum())
            if n_shelves >= 1 and (cells == "g").sum() >= 1:
                break
        kw = dict(layout=layout)
        n_cells = H * W
    else:
        cols, rows, height = int(rs.choice([1, 3, 5])), int(rs.integers(1, 3)), int(rs.integers(1, 5))
        (H, W), hw, _ = rware_b200.layout_from_params(cols, rows, height)
        n_shelves = int((hw == 0).sum())
        if n_shelves == 0:
            cols = 3
            (H, W), hw, _ = rware_b200.layout_from_params(cols, rows, height)
            n_shelves = int((hw == 0).sum())
        kw = dict(shelf_columns=cols, shelf_rows=rows, column_height=height)
        n_cells = H * W
    n = int(rs.integers(1, min(n_cells, 20) + 1))
    kw.update(n_agents=n, msg_bits=int(rs.choice([0, 0, 1, 3])), sensor_range=int(rs.integers(1, 4)),
              request_queue_size=int(rs.integers(0, min(n_shelves, 6) + 1)), reward_type=int(rs.integers(0, 3)),
              max_steps=int(rs.choice([0, 15, 40])) or None, max_inactivity_steps=int(rs.choice([0, 0, 10])) or None,
              normalised_coordinates=bool(rs.random() < 0.3))
    if kw["request_queue_size"] == n_shelves and n_shelves > 0:
        kw["request_queue_size"] -= 1  # a delivery with an empty candidate list raises in the reference
    return kw


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("coop", ["0", "1"], ids=["lane-per-world", "group-per-world"])
@pytest.mark.parametrize("seed", range(12))
def test_random_configuration_matches_oracle(seed, coop, backend, monkeypatch):
    monkeypatch.setenv("RWARE_B200_COOP", coop)
    rs = np.random.default_rng(1000 + seed)
    kw = random_case(rs)
    E, T = 9, 70
    vec = rware_b200.VecWarehouse(E, auto_reset=True, _backend_factory=backend_factory(backend), **kw)
    n, mb = kw["n_agents"], kw["msg_bits"]
    orcs = [oracle_py.OracleEnv(vec.highways, vec.goals, n, mb, kw["sensor_range"], kw["request_queue_size"],
                                kw["max_inactivity_steps"], kw["max_steps"], kw["reward_type"], kw["normalised_coordinates"])
            for _ in range(E)]
    obs = vec.reset(seed=7 * seed).cpu().numpy()
    for e, o in enumerate(orcs):
        assert np.array_equal(o.reset(7 * seed + e), obs[e]), kw
    pf = float(rs.choice([0.3, 0.6, 0.85]))
    probs = np.full(5, (1 - pf) / 4)
    probs[1] = pf
    for t in range(T):
        acts = np.concatenate([rs.choice(5, size=(E, n, 1), p=probs), rs.integers(0, 2, size=(E, n, mb))], axis=2).astype(np.int32)
        obs, rew, done = vec.step(torch.from_numpy(acts))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        st = {k: v.cpu().numpy() for k, v in vec.get_state().items()}
        for e, o in enumerate(orcs):
            oo, orw, od = o.step(acts[e])
            assert np.array_equal(orw.astype(np.float32), rew[e]) and od == bool(done[e]), (kw, t, e)
            if od:
                oo = o.reset()
            assert np.array_equal(oo, obs[e]), (kw, t, e)
            so = o.get_state()
            assert np.array_equal(so["agent_x"], st["agent_x"][e]) and np.array_equal(so["agent_y"], st["agent_y"][e])
            assert np.array_equal(so["queue"], st["queue"][e]) and np.array_equal(so["agent_carry"], st["agent_carry"][e])
            assert np.array_equal(so["rng"], st["rng"][e].astype(np.int64) & 0xFFFFFFFF)
        if t % 9 == 4:  # whole-warehouse view of the same states: random layer subset, random padding
            H, W = vec.grid_size
            layers = [int(v) for v in rs.choice(7, size=int(rs.integers(1, 6)))]
            pad = None if rs.random() < 0.5 else (len(layers) + int(rs.integers(0, 3)), H + int(rs.integers(0, 4)), W + int(rs.integers(0, 4)))
            got = vec.global_image(layers, pad).cpu().numpy()
            for e, o in enumerate(orcs):
                try:
                    want = o.global_image(layers, pad)
                except IndexError:  # transposed layer off the map: the reference raises, the batched call counts the world
                    continue
                assert np.array_equal(got[e], want), (kw, t, e, layers, pad)
