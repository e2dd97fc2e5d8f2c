"""Whole-warehouse views: get_global_image (rware/warehouse.py:966-1040) against images recorded
from the unmodified reference (tests/golden/global_image.{npz,json}) and, batched, against the oracle; the rgb_array frame
(rware/rendering.py) against a numpy restatement of the Viewer's geometry (not pixel-pinned: no OpenGL where this was written)."""
import json
import math
import os

import numpy as np
import pytest
import torch

import oracle_py
import rware_b200
from conftest import backend_factory, backend_params
from rware_b200.warehouse import ImageLayer, RewardType, Warehouse

HERE = os.path.dirname(os.path.abspath(__file__))
FX = np.load(os.path.join(HERE, "golden", "global_image.npz"))
with open(os.path.join(HERE, "golden", "global_image.json")) as f:
    META = json.load(f)


def _layout_of(k):
    if k.get("layout"):
        return rware_b200.layout_from_str(k["layout"])
    return rware_b200.layout_from_params(k["shelf_columns"], k["shelf_rows"], k["column_height"])


@pytest.mark.parametrize("name", sorted(META))
def test_oracle_global_image_matches_the_reference(name):
    m = META[name]
    k = m["kwargs"]
    _, hw, goals = _layout_of(k)
    orc = oracle_py.OracleEnv(hw, goals, k["n_agents"], k["msg_bits"], k["sensor_range"], k["request_queue_size"],
                              k["max_inactivity_steps"], k["max_steps"], k["reward_type"])
    imgs, raised, acts = FX[name + "_images"], FX[name + "_raised"], FX[name + "_actions"]

    def check(t):
        if raised[t]:
            with pytest.raises(IndexError):
                orc.global_image(m["layers"], m["pad"])
        else:
            assert np.array_equal(orc.global_image(m["layers"], m["pad"]), imgs[t]), (name, t)
    orc.reset(m["seed"])
    check(0)
    for t, a in enumerate(acts):
        _, _, done = orc.step(a.astype(np.int32))
        if done:
            orc.reset()
        check(t + 1)
    assert raised.sum() > 0 or name != "transposed_tiny"


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", sorted(META))
def test_facade_global_image_matches_the_reference(name, backend):
    m = META[name]
    k = dict(m["kwargs"])
    k["reward_type"] = RewardType(k["reward_type"])
    env = Warehouse(_backend_factory=backend_factory(backend), **k)
    layers = [ImageLayer(v) for v in m["layers"]]
    imgs, raised, acts = FX[name + "_images"], FX[name + "_raised"], FX[name + "_actions"]

    def check(t):
        if raised[t]:
            with pytest.raises(IndexError):
                env.get_global_image(layers, recompute=True, pad_to_shape=m["pad"])
        else:
            got = env.get_global_image(layers, recompute=True, pad_to_shape=m["pad"])
            assert got.dtype == np.float32 and np.array_equal(got, imgs[t]), (name, t)
    env.reset(seed=m["seed"])
    check(0)
    for t, a in enumerate(acts):
        _, _, done, _, _ = env.step([int(v) for v in a])
        if done:
            env.reset()
        check(t + 1)
    cached = env.get_global_image(layers, recompute=False)  # warehouse.py:985: cached unless recompute
    assert cached is env.global_image


@pytest.mark.parametrize("backend", backend_params())
def test_batched_global_image_matches_the_oracle(backend):
    rs = np.random.default_rng(3)
    E, n = 40, 5
    kw = dict(shelf_columns=3, column_height=4, shelf_rows=2, n_agents=n, request_queue_size=4, max_steps=30)
    vec = rware_b200.VecWarehouse(E, auto_reset=True, _backend_factory=backend_factory(backend), **kw)
    H, W = vec.grid_size
    orcs = [oracle_py.OracleEnv(vec.highways, vec.goals, n, 0, 1, 4, None, 30, 1) for _ in range(E)]
    vec.reset(seed=11)
    for e, o in enumerate(orcs):
        o.reset(11 + e)
    cases = [((0, 5), None), ((6, 2, 1, 0, 5), None), ((1, 2), (4, H + 3, W + 2)), ((2,), (1, H, W + 5)), ((5, 5, 0, 0, 6, 6, 1, 2), None)]
    for t in range(45):
        acts = rs.choice(5, size=(E, n), p=[.05, .45, .15, .15, .2]).astype(np.int32)
        _, _, done = vec.step(torch.from_numpy(acts))
        for e, o in enumerate(orcs):
            _, _, d = o.step(acts[e])
            if d:
                o.reset()
        layers, pad = cases[t % len(cases)]
        got = vec.global_image(layers, pad).cpu().numpy()
        for e, o in enumerate(orcs):
            assert np.array_equal(got[e], o.global_image(layers, pad)), (t, e)
    # transposed layers on this non-square grid: worlds where the reference raises are counted, the rest match
    assert H != W
    before = vec.error_count()
    got = vec.global_image((3, 4, 2)).cpu().numpy()
    n_bad = 0
    for e, o in enumerate(orcs):
        try:
            want = o.global_image((3, 4, 2))
        except IndexError:
            n_bad += 1
            continue
        assert np.array_equal(got[e], want)
    assert vec.error_count() - before == n_bad
    with pytest.raises(ValueError):
        vec.global_image((0, 7))
    with pytest.raises(RuntimeError):
        vec.global_image((0, 5), (2, H - 1, W))


# ---------------------------------------------------------------- rgb_array frame
def numpy_frame(highways, goals, st):
    """rware/rendering.py:121-333 with plain numpy: fills in OpenGL coordinates (origin bottom-left), then flips (136)."""
    H, W = highways.shape
    fh, fw = 2 + 31 * H, 1 + 31 * W
    img = np.full((fh, fw, 3), 255, np.uint8)          # glClearColor(white) 122
    yy, xx = np.mgrid[0:fh, 0:fw]
    cxs, cys = xx + 0.5, yy + 0.5                     # pixel centres
    for r in range(H + 1):                            # horizontal grid lines 143-159: y = 31 r + 1 -> pixel row 31 r
        img[31 * r, 0:31 * W] = 0
    for c in range(W + 1):                            # vertical grid lines 162-177
        img[0:31 * H, 31 * c] = 0
    for gx, gy in goals:                              # goal quads 207-236
        gy = H - gy - 1
        m = (cxs >= 31 * gx + 1) & (cxs < 31 * (gx + 1)) & (cys >= 31 * gy + 1) & (cys < 31 * (gy + 1))
        img[m] = (60, 60, 60)
    queue = set(int(q) for q in st["queue"])
    for k, (sx, sy) in enumerate(zip(st["shelf_x"], st["shelf_y"])):   # shelf quads 180-205
        sy = H - sy - 1
        m = (cxs >= 31 * sx + 3) & (cxs < 31 * (sx + 1) - 2) & (cys >= 31 * sy + 3) & (cys < 31 * (sy + 1) - 2)
        img[m] = (0, 128, 128) if (k + 1) in queue else (72, 61, 139)
    for x, y, d, c in zip(st["agent_x"], st["agent_y"], st["agent_dir"], st["agent_carry"]):  # hexagons 262-296
        row = H - y - 1
        ccx, ccy = 31 * x + 16, 31 * row + 16
        inside = np.ones((fh, fw), bool)
        verts = [(ccx + 10 * math.cos(2 * math.pi * i / 6), ccy + 10 * math.sin(2 * math.pi * i / 6)) for i in range(6)]
        for i in range(6):
            (x0, y0), (x1, y1) = verts[i], verts[(i + 1) % 6]
            inside &= ((x1 - x0) * (cys - y0) - (y1 - y0) * (cxs - x0)) >= -1e-4
        img[inside] = (255, 0, 0) if c else (255, 140, 0)
    for x, y, d, c in zip(st["agent_x"], st["agent_y"], st["agent_dir"], st["agent_carry"]):  # heading lines 298-333
        row = H - y - 1
        ccx, ccy = 31 * x + 16, 31 * row + 16
        if d == 0:
            img[ccy:ccy + 10, ccx - 1] = 0
        elif d == 1:
            img[ccy - 10:ccy, ccx - 1] = 0
        elif d == 2:
            img[ccy - 1, ccx - 10:ccx] = 0
        else:
            img[ccy - 1, ccx:ccx + 10] = 0
    return img[::-1]


@pytest.mark.parametrize("backend", backend_params())
def test_rgb_frames_follow_the_viewer_geometry(backend):
    rs = np.random.default_rng(5)
    E, n = 35, 6
    kw = dict(shelf_columns=3, column_height=2, shelf_rows=2, n_agents=n, request_queue_size=3, max_steps=None)
    vec = rware_b200.VecWarehouse(E, _backend_factory=backend_factory(backend), **kw)
    vec.reset(seed=2)
    for t in range(25):
        vec.step(torch.from_numpy(rs.choice(5, size=(E, n), p=[.05, .45, .15, .15, .2]).astype(np.int32)))
    st = {k: v.cpu().numpy() for k, v in vec.get_state().items()}
    H, W = vec.grid_size
    frames = vec.render_rgb(30, 5).cpu().numpy()
    assert frames.shape == (5, 2 + 31 * H, 1 + 31 * W, 3) and frames.dtype == np.uint8
    for i in range(5):
        e = 30 + i
        want = numpy_frame(np.asarray(vec.highways), vec.goals, {k: v[e] for k, v in st.items()})
        assert np.array_equal(frames[i], want), f"world {e}: {np.argwhere((frames[i] != want).any(-1))[:5]}"
    assert (st["agent_carry"][30:35] > 0).any()  # loaded agents were drawn
    with pytest.raises(RuntimeError):
        vec.render_rgb(E - 1, 2)


@pytest.mark.parametrize("backend", backend_params())
def test_facade_render_rgb_array(backend):
    env = Warehouse(3, 8, 1, 2, 0, 1, 2, None, None, RewardType.INDIVIDUAL, render_mode="rgb_array",
                    _backend_factory=backend_factory(backend))
    env.reset(seed=0)
    frame = env.render()
    assert frame.shape == (2 + 31 * 11, 1 + 31 * 10, 3) and frame.dtype == np.uint8
    colours = {tuple(c) for c in frame.reshape(-1, 3)}
    assert {(255, 255, 255), (0, 0, 0), (72, 61, 139), (0, 128, 128), (255, 140, 0), (60, 60, 60)} <= colours
    env.agents[0].x, env.agents[0].y = 0, 0  # write-through entity state shows up in the next frame
    assert tuple(env.render()[2 + 15, 1 + 15 + 3]) in {(255, 140, 0), (0, 0, 0)}
    human = Warehouse(3, 8, 1, 2, 0, 1, 2, None, None, RewardType.INDIVIDUAL, _backend_factory=backend_factory(backend))
    human.reset(seed=0)
    with pytest.raises(NotImplementedError):
        human.render()
