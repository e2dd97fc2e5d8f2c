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


TX = np.load(os.path.join(HERE, "golden", "image_transposed.npz"))
with open(os.path.join(HERE, "golden", "image_transposed.json")) as f:
    TXMETA = json.load(f)


def _tx_layout(k):
    if k.get("layout"):
        return rware_b200.layout_from_str(k["layout"])
    return rware_b200.layout_from_params(k["shelf_columns"], k["shelf_rows"], k["column_height"])


@pytest.mark.parametrize("name", sorted(TXMETA))
def test_oracle_transposed_image_layers_match_the_reference(name):
    """AGENT_DIRECTION / AGENT_LOAD in the local IMAGE observations: `layer[ag.x, ag.y]` (sic, warehouse.py:552, 558), including
    the steps where numpy raises IndexError in the reference (tests/golden/image_transposed.*)."""
    k = TXMETA[name]["kwargs"]
    _, hw, goals = _tx_layout(k)
    orc = oracle_py.OracleEnv(hw, goals, k["n_agents"], 0, k["sensor_range"], k["request_queue_size"], None, k["max_steps"],
                              k["reward_type"])
    imgs, raised, acts = TX[name + "_images"], TX[name + "_raised"], TX[name + "_actions"]
    layers, directional = k["image_observation_layers"], k["image_observation_directional"]

    def check(t):
        if raised[t]:
            with pytest.raises(IndexError):
                orc.obs_image(layers, directional)
        else:
            assert np.array_equal(orc.obs_image(layers, directional), imgs[t]), (name, t)
    orc.reset(TXMETA[name]["seed"])
    check(0)
    for t, a in enumerate(acts):
        _, _, done = orc.step(a.astype(np.int32))
        if done:
            orc.reset()
        check(t + 1)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", sorted(TXMETA))
def test_facade_transposed_image_layers_match_the_reference(name, backend):
    k = dict(TXMETA[name]["kwargs"])
    k["reward_type"] = RewardType(k["reward_type"])
    k["observation_type"] = ObservationType(k["observation_type"])
    k["image_observation_layers"] = [ImageLayer(v) for v in k["image_observation_layers"]]
    env = Warehouse(_backend_factory=backend_factory(backend), **k)
    imgs, raised, acts = TX[name + "_images"], TX[name + "_raised"], TX[name + "_actions"]

    if raised[0]:
        with pytest.raises(IndexError):
            env.reset(seed=TXMETA[name]["seed"])
    else:
        obs, _ = env.reset(seed=TXMETA[name]["seed"])
        assert np.array_equal(np.stack(obs), imgs[0])
    for t, a in enumerate(acts):
        # imgs[t + 1] / raised[t + 1]: what the caller continues from after step t (the reset observation when it ended the episode)
        steps_before = int(env._vec.get_state()["cur_steps"][0].item())
        try:
            obs, _, done, _, _ = env.step([int(v) for v in a])
            step_raised = False
        except IndexError:
            step_raised, done = True, steps_before + 1 >= k["max_steps"]  # a raising step has still advanced the state
        if done:
            try:
                obs, _ = env.reset()
                step_raised = False
            except IndexError:
                step_raised = True
        assert step_raised == bool(raised[t + 1]), (name, t)
        if not step_raised:
            assert np.array_equal(np.stack(obs), imgs[t + 1]), (name, t)
    assert raised.sum() > 0 or name != "tiny_raises"


@pytest.mark.parametrize("backend", backend_params())
def test_batched_transposed_image_layers_match_the_oracle_and_count_index_errors(backend):
    E, n, r = 33, 5, 1
    layers = (3, 2, 4)
    kw = dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=n, request_queue_size=4, sensor_range=r, max_steps=40)
    vec = rware_b200.VecWarehouse(E, auto_reset=True, observation_type=2, image_observation_layers=layers,
                                  image_observation_directional=True, _backend_factory=backend_factory(backend), **kw)
    orcs = [oracle_py.OracleEnv(vec.highways, vec.goals, n, 0, r, 4, None, 40, 1) for _ in range(E)]
    before = vec.error_count()
    obs = vec.reset(seed=3)
    for e, o in enumerate(orcs):
        o.reset(3 + e)
    rs = np.random.default_rng(2)
    n_bad = 0
    for t in range(50):
        img, _ = vec.split_image_obs(obs)
        img = img.cpu().numpy()
        for e, o in enumerate(orcs):
            try:
                want = o.obs_image(layers, True)
            except IndexError:
                n_bad += 1
                continue
            assert np.array_equal(want, img[e]), (t, e)
        assert vec.error_count() - before == n_bad, t
        acts = rs.choice(5, size=(E, n), p=[.05, .55, .15, .15, .1]).astype(np.int32)
        obs, rew, done = vec.step(torch.from_numpy(acts))
        for e, o in enumerate(orcs):
            _, _, od = o.step(acts[e])
            if od:
                o.reset()
    assert n_bad > 0
