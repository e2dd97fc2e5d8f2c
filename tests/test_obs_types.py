"""Non-default observation types: IMAGE / IMAGE_DICT gathered on the device
(rware/warehouse.py:527-596, 726-742) and DICT unflattened on the host (676-720), against trajectories recorded from the
unmodified reference (tests/golden/obs_types.json) and, batched, against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

import oracle_py
import rware_b200
from conftest import backend_factory, backend_params
from rware_b200.warehouse import ImageLayer, ObservationType, RewardType, Warehouse

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "obs_types.json")) as f:
    FX = json.load(f)


def _make(kwargs, backend):
    kw = dict(kwargs)
    kw["reward_type"] = RewardType(kw["reward_type"])
    kw["observation_type"] = ObservationType(kw["observation_type"])
    if "image_observation_layers" in kw:
        kw["image_observation_layers"] = [ImageLayer(v) for v in kw["image_observation_layers"]]
    return Warehouse(_backend_factory=backend_factory(backend), **kw)


def _same_dict_obs(a, b):
    if isinstance(b, dict):
        return set(a.keys()) == set(b.keys()) and all(_same_dict_obs(a[k], b[k]) for k in b)
    if isinstance(b, (list, tuple)) and not isinstance(a, np.ndarray):
        return len(a) == len(b) and all(_same_dict_obs(x, y) for x, y in zip(a, b))
    if b is None:
        return a is None
    return np.array_equal(np.asarray(a), np.asarray(b))


def _oracle_image_case(name):
    rec = FX[name]
    k = rec["kwargs"]
    (H, W), hw, goals = rware_b200.layout_from_params(k["shelf_columns"], k["shelf_rows"], k["column_height"])
    orc = oracle_py.OracleEnv(hw, goals, k["n_agents"], k["msg_bits"], k["sensor_range"], k["request_queue_size"],
                              k["max_inactivity_steps"], k["max_steps"], k["reward_type"])
    layers = k.get("image_observation_layers", [0, 1, 2, 5, 6])
    return rec, k, orc, layers, k.get("image_observation_directional", True)


@pytest.mark.parametrize("name", [n for n in FX if n.startswith("image")])
def test_oracle_image_observations_match_the_reference(name):
    rec, k, orc, layers, directional = _oracle_image_case(name)
    orc.reset(rec["seed"])
    img, feat = orc.obs_image(layers, directional, with_features=True)
    assert np.array_equal(img, np.asarray(rec["image"][0], np.float32))
    for t, a in enumerate(rec["actions"]):
        _, _, done = orc.step(np.asarray(a))
        if done:
            orc.reset()
        img, feat = orc.obs_image(layers, directional, with_features=True)
        assert np.array_equal(img, np.asarray(rec["image"][t + 1], np.float32)), (name, t)
        if "features" in rec:
            assert np.array_equal(feat, np.asarray(rec["features"][t + 1], np.float32))


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", sorted(FX))
def test_facade_observation_types_match_the_reference(name, backend):
    rec = FX[name]
    env = _make(rec["kwargs"], backend)
    n, mb = rec["kwargs"]["n_agents"], rec["kwargs"]["msg_bits"]
    obs, info = env.reset(seed=rec["seed"])
    is_dict = rec["kwargs"]["observation_type"] == ObservationType.DICT.value

    def check(obs, t):
        if is_dict:
            assert _same_dict_obs(obs, rec["obs"][t]), (name, t)
        elif "features" in rec:
            assert all(np.array_equal(obs[i]["image"], np.asarray(rec["image"][t][i], np.float32)) for i in range(n)), (name, t)
            assert all(np.array_equal(obs[i]["features"], np.asarray(rec["features"][t][i], np.float32)) for i in range(n))
        else:
            assert all(np.array_equal(obs[i], np.asarray(rec["image"][t][i], np.float32)) for i in range(n)), (name, t)
        if not is_dict:
            assert env.observation_space.contains(obs)
    check(obs, 0)
    for t, a in enumerate(rec["actions"]):
        act = [list(r) for r in a] if mb else [r[0] for r in a]
        obs, rew, done, _, _ = env.step(act)
        if done:
            obs, _ = env.reset()
        check(obs, t + 1)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("otype,layers,directional,r", [(2, (0, 1, 2, 5, 6), True, 1), (3, (6, 5, 2), True, 2),
                                                        (2, (1, 0), False, 3)])
def test_batched_image_observations_match_the_oracle(otype, layers, directional, r, backend):
    E, n = 37, 5
    kw = dict(shelf_columns=3, column_height=8, shelf_rows=2, n_agents=n, request_queue_size=4, sensor_range=r, max_steps=40)
    vec = rware_b200.VecWarehouse(E, auto_reset=True, observation_type=otype, image_observation_layers=layers,
                                  image_observation_directional=directional, _backend_factory=backend_factory(backend), **kw)
    orcs = [oracle_py.OracleEnv(vec.highways, vec.goals, n, 0, r, 4, None, 40, 1) for _ in range(E)]
    obs = vec.reset(seed=50)
    for e, o in enumerate(orcs):
        o.reset(50 + e)
    rs = np.random.default_rng(1)
    for t in range(60):
        img, feat = vec.split_image_obs(obs)
        img = img.cpu().numpy()
        for e, o in enumerate(orcs):
            oi, of = o.obs_image(layers, directional, with_features=True)
            assert np.array_equal(oi, img[e]), (t, e)
            if otype == 3:
                assert np.array_equal(of, feat[e].cpu().numpy())
        acts = rs.choice(5, size=(E, n), p=[.1, .5, .15, .15, .1]).astype(np.int32)
        obs, rew, done = vec.step(torch.from_numpy(acts))
        for e, o in enumerate(orcs):
            _, _, od = o.step(acts[e])
            if od:
                o.reset()


def test_transposed_layers_are_rejected_by_the_cuda_abi():
    """AGENT_DIRECTION / AGENT_LOAD: the reference indexes them [x, y] (warehouse.py:552, 558) and raises IndexError
    off the diagonal; the C ABI refuses them instead of reproducing that."""
    import ctypes
    from rware_b200 import _capi
    lib = _capi.load_library()
    cfg = _capi.EnvConfig(np.zeros((4, 4), np.uint8), [(0, 0)], n_agents=1, observation_type=2, image_layers=(0, 3))
    h = ctypes.c_void_p()
    s = cfg.to_struct()
    assert lib.rw_create(ctypes.byref(s), 4, 0, ctypes.byref(h)) == -1
    assert b"transposed" in lib.rw_last_error(None)
