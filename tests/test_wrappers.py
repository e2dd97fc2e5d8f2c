"""rware/utils/wrappers.py behaviours (reference tests/test_wrappers.py:36-68) on top of the drop-in facade."""
import numpy as np
import pytest
import torch

import rware_b200
from conftest import backend_factory, backend_params
from rware_b200.warehouse import RewardType, Warehouse
from rware_b200.wrappers import DictAgents, FlattenAgents, VecFlattenAgents


def _env(n, msg, backend):
    env = Warehouse(3, 8, 3, n, msg, 1, 5, None, None, RewardType.GLOBAL, _backend_factory=backend_factory(backend))
    env.reset(seed=1)
    return env


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("n,msg", [(1, 0), (2, 0), (2, 2)])
def test_flatten_agents(n, msg, backend):  # tests/test_wrappers.py:36-63
    env = FlattenAgents(_env(n, msg, backend))
    obs, _ = env.reset(seed=2)
    assert obs.ndim == 1 and obs.shape[0] == n * env.unwrapped._obs_length
    obs, rew, done, _, _ = env.step(env.action_space.sample())
    assert obs.ndim == 1 and np.shape(rew) == () and type(done) is bool


@pytest.mark.parametrize("backend", backend_params())
def test_dict_agents(backend):  # tests/test_wrappers.py:66-68
    env = DictAgents(_env(2, 0, backend))
    obs, info = env.reset(seed=3)
    assert sorted(obs) == ["agent_0", "agent_1"] and info == {}
    obs, rew, done, trunc, _ = env.step({"agent_0": 1, "agent_1": 2})
    assert set(obs) == set(rew) == set(done) == set(trunc) == {"agent_0", "agent_1"}


@pytest.mark.parametrize("backend", backend_params())
def test_vec_flatten_agents(backend):
    vec = rware_b200.VecWarehouse(5, 3, 8, 1, 4, request_queue_size=4, _backend_factory=backend_factory(backend))
    env = VecFlattenAgents(vec)
    obs = env.reset(seed=0)
    assert obs.shape == (5, 4 * vec.obs_len)
    obs, rew, done = env.step(torch.ones((5, 4), dtype=torch.int32))
    assert obs.shape == (5, 4 * vec.obs_len) and rew.shape == (5,) and done.shape == (5,)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", ["dict_msg2", "imagedict"])
def test_flatten_sa_observation_matches_the_reference_wrapper(name, backend):
    """FlattenSAObservation (rware/utils/wrappers.py:83-108) over DICT / IMAGE_DICT observations, against vectors recorded from
    the reference wrapper (tests/golden/flatten_sa.json)."""
    import json
    import os
    from rware_b200.warehouse import ObservationType, RewardType, Warehouse
    from rware_b200.wrappers import FlattenSAObservation
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flatten_sa.json")) as f:
        rec = json.load(f)[name]
    kw = dict(rec["kwargs"])
    kw["observation_type"] = ObservationType(kw["observation_type"])
    env = FlattenSAObservation(Warehouse(shelf_columns=3, column_height=8, shelf_rows=1, sensor_range=1, request_queue_size=2,
                                         max_inactivity_steps=None, max_steps=30, reward_type=RewardType.INDIVIDUAL,
                                         _backend_factory=backend_factory(backend), **kw))
    assert [s.shape[0] for s in env.observation_space] == rec["flatdim"]
    obs, _ = env.reset(seed=21)
    assert all(np.array_equal(np.asarray(o, np.float64), np.asarray(w)) for o, w in zip(obs, rec["obs"][0]))
    for t, a in enumerate(rec["actions"]):
        obs, _, _, _, _ = env.step(a)
        assert len(obs) == len(rec["obs"][t + 1])
        assert all(np.array_equal(np.asarray(o, np.float64), np.asarray(w)) for o, w in zip(obs, rec["obs"][t + 1])), t


def test_multi_agent_space_lists():  # rware/utils/spaces.py:4-32
    from rware_b200 import spaces
    ma = spaces.MultiAgentObservationSpace([spaces.Discrete(3), spaces.Box(0.0, 1.0, shape=(2,), dtype=np.float32)])
    sample = ma.sample()
    assert len(sample) == 2 and ma.contains(sample) and not ma.contains([5, np.zeros(2, np.float32)])
    acts = spaces.MultiAgentActionSpace([spaces.Discrete(5)] * 4)
    assert len(acts.sample()) == 4 and all(0 <= a < 5 for a in acts.sample())
