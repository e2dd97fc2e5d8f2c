"""Randomised differential test: product step logic (host twin / CUDA) vs the C oracle on batches of independent
worlds with dense traffic, so that collision ties (CPython-set order regime included), swaps, cycles, deliveries,
resamples, TWO_STAGE rewards and in-kernel auto-resets all occur.  Bit-exact on every field."""
import numpy as np
import pytest
import torch

import oracle_py
import rware_b200
from conftest import backend_factory, backend_params

CASES = [
    # name, kwargs, num_envs, steps, p_forward
    ("tiny4", dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=4, request_queue_size=4, max_steps=60), 40, 150, 0.5),
    ("dense12", dict(shelf_columns=3, column_height=1, shelf_rows=1, n_agents=12, request_queue_size=2, max_steps=50,
                     reward_type=2), 40, 150, 0.8),
    ("medium8_r2_msg", dict(shelf_columns=5, column_height=8, shelf_rows=2, n_agents=8, request_queue_size=4,
                            max_inactivity_steps=30, sensor_range=2, msg_bits=2, reward_type=0), 33, 100, 0.7),
    ("ring32", dict(layout="....\n.xx.\n.gx.\n....", n_agents=12, request_queue_size=1, max_steps=40), 64, 200, 0.9),
    ("n32", dict(shelf_columns=3, column_height=3, shelf_rows=1, n_agents=32, request_queue_size=5, max_steps=30), 32, 60, 0.85),
]


def _mk_oracles(vec, kw, E):
    return [oracle_py.OracleEnv(vec.highways, vec.goals, kw["n_agents"], kw.get("msg_bits", 0), kw.get("sensor_range", 1),
                                kw["request_queue_size"], kw.get("max_inactivity_steps"), kw.get("max_steps"),
                                kw.get("reward_type", 1)) for _ in range(E)]


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("coop", ["0", "1"], ids=["lane-per-world", "group-per-world"])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_random_rollout_matches_oracle(case, backend, coop, monkeypatch):
    monkeypatch.setenv("RWARE_B200_COOP", coop)
    name, kw, E, steps, pf = case
    vec = rware_b200.VecWarehouse(E, auto_reset=True, _backend_factory=backend_factory(backend), **kw)
    orcs = _mk_oracles(vec, kw, E)
    n, mb = kw["n_agents"], kw.get("msg_bits", 0)
    obs = vec.reset(seed=1000).cpu().numpy()
    for e, o in enumerate(orcs):
        assert np.array_equal(o.reset(1000 + e), obs[e])
    rs = np.random.default_rng(7)
    probs = np.full(5, (1 - pf) / 4)
    probs[1] = pf
    tot = dict(ties=0, set_order_components=0, deliveries=0, dones=0)
    for t in range(steps):
        acts = np.concatenate([rs.choice(5, size=(E, n, 1), p=probs), rs.integers(0, 2, size=(E, n, mb))], axis=2).astype(np.int32)
        obs, rew, done = vec.step(torch.from_numpy(acts))
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        st = {k: v.cpu().numpy() for k, v in vec.get_state().items()}
        for e, o in enumerate(orcs):
            oo, orw, od = o.step(acts[e])
            s = o.last_step_stats()
            for k in ("ties", "set_order_components", "deliveries"):
                tot[k] += s[k]
            assert np.array_equal(orw.astype(np.float32), rew[e]), (name, t, e, orw, rew[e])
            assert od == bool(done[e])
            if od:  # the batched path auto-resets in-kernel: the oracle continues its generator stream too
                oo = o.reset()
                tot["dones"] += 1
            assert np.array_equal(oo, obs[e]), (name, t, e)
            so = o.get_state()
            assert np.array_equal(so["agent_x"], st["agent_x"][e]) and np.array_equal(so["agent_y"], st["agent_y"][e])
            assert np.array_equal(so["agent_carry"], st["agent_carry"][e]) and np.array_equal(so["queue"], st["queue"][e])
            assert np.array_equal(so["shelf_x"], st["shelf_x"][e]) and np.array_equal(so["shelf_y"], st["shelf_y"][e])
            assert np.array_equal(so["rng"], st["rng"][e].astype(np.int64) & 0xFFFFFFFF)
    assert tot["dones"] > 0
    if name in ("dense12", "ring32", "n32"):
        assert tot["ties"] > 0 and tot["set_order_components"] > 0, tot


@pytest.mark.parametrize("backend", backend_params())
def test_invalid_actions_are_counted_and_treated_as_noop(backend):
    kw = dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=2, request_queue_size=2)
    vec = rware_b200.VecWarehouse(3, _backend_factory=backend_factory(backend), **kw)
    vec.reset(seed=3)
    before = {k: v.clone() for k, v in vec.get_state().items()}
    vec.step(torch.tensor([[7, -1], [0, 0], [5, 0]], dtype=torch.int32))
    after = vec.get_state()
    assert vec.backend.error_count() == 3
    assert torch.equal(before["agent_x"], after["agent_x"]) and torch.equal(before["agent_dir"], after["agent_dir"])
