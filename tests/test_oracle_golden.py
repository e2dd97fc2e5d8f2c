"""The oracle against the golden vectors recorded from the unmodified reference (CPU, no GPU needed)."""
import hashlib
import json
import os

import numpy as np
import pytest

import replay

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", replay.fixture_names())
def test_oracle_replays_reference_trajectory(name):
    fx = replay.load_fixture(name)
    assert replay.replay(replay.OracleDriver(fx), fx) == len(fx["actions"])


def test_oracle_known_answer_hashes():
    """registered ids, reset(seed=0), 500 steps of default_rng(12345) actions."""
    import oracle_py
    import rware_b200
    with open(os.path.join(HERE, "golden", "known_answers.json")) as f:
        ka = json.load(f)
    for key, rec in ka.items():
        kw = rware_b200.env_kwargs(key.split("+")[0], **rec["extra"])
        (H, W), highways, goals = rware_b200.layout_from_params(kw["shelf_columns"], kw["shelf_rows"], kw["column_height"])
        env = oracle_py.OracleEnv(highways, goals, kw["n_agents"], kw["msg_bits"], kw["sensor_range"], kw["request_queue_size"],
                                  kw["max_inactivity_steps"], kw["max_steps"], kw["reward_type"].value)
        obs = env.reset(0)
        h = hashlib.sha256(obs.tobytes())
        acts = np.random.default_rng(12345).integers(0, 5, size=(500, kw["n_agents"]))
        for t in range(500):
            obs, rew, done = env.step(acts[t])
            h.update(obs.tobytes() + rew.tobytes() + bytes([done]))
        assert h.hexdigest() == rec["sha256"], key


def test_layouts_match_reference():
    import oracle_py
    import rware_b200
    fx = np.load(os.path.join(HERE, "golden", "layouts.npz"))
    for key in [k[:-9] for k in fx.files if k.endswith("_highways")]:
        cols, rows, height = map(int, key.split("_"))
        (H, W), hw, goals = rware_b200.layout_from_params(cols, rows, height)
        assert np.array_equal(hw, fx[key + "_highways"]) and np.array_equal(np.asarray(goals), fx[key + "_goals"])
        H2, W2, hw2, goals2 = oracle_py.layout_from_params(cols, rows, height)
        assert (H2, W2) == (H, W) and np.array_equal(hw2, hw) and goals2 == goals
