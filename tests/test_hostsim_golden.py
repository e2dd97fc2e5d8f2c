"""The product's per-world logic (host twin of the CUDA kernel, same source) and the CUDA C-ABI path against the
golden trajectories of the unmodified reference.  hostsim runs everywhere; cuda is the parity test proper."""
import pytest

import replay
from conftest import backend_factory, backend_params


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", replay.fixture_names())
def test_replays_reference_trajectory(name, backend):
    fx = replay.load_fixture(name)
    # 37 worlds: two tiles, the second one ragged; every world must produce the reference trajectory
    drv = replay.VecDriver(fx, backend_factory(backend), num_envs=37)
    assert replay.replay(drv, fx) == len(fx["actions"])


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", ["tiny4_indiv", "large16_r2"])
def test_dynamic_instantiation_matches_too(name, backend, monkeypatch):
    """The registered geometries normally run static specialisations; force the fully dynamic kernel as well."""
    fx = replay.load_fixture(name)
    if backend == "hostsim":
        from hostsim import backend as hs
        hs._load().sim_force_dynamic(1)
    else:
        monkeypatch.setenv("RWARE_B200_FORCE_DYNAMIC", "1")
    try:
        drv = replay.VecDriver(fx, backend_factory(backend), num_envs=33)
        assert replay.replay(drv, fx, max_steps=400) == 400
    finally:
        if backend == "hostsim":
            hs._load().sim_force_dynamic(0)
        else:
            monkeypatch.setenv("RWARE_B200_FORCE_DYNAMIC", "0")
            replay.VecDriver(fx, backend_factory(backend), num_envs=1)  # resets the library-wide switch


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", ["tiny4_indiv", "tiny19_twostage", "custom_layout", "medium8_hard_twostage"])
@pytest.mark.parametrize("coop", ["1", "2"])
def test_warp_cooperative_step_matches_too(name, backend, coop, monkeypatch):
    """The default step advances one world per lane (rw_step_lane); the warp-cooperative form (rw_coop.cuh: a group of
    lanes per world, shfl/ballot/match_any resolver; "2" = with interleaved observation emission) must agree."""
    monkeypatch.setenv("RWARE_B200_COOP", coop)
    fx = replay.load_fixture(name)
    drv = replay.VecDriver(fx, backend_factory(backend), num_envs=35)
    assert replay.replay(drv, fx, max_steps=500) == 500


# the fully static configurations with an even agent count <= 8 step through the agent-split form by default
# (rw_logic.cuh, rw_pair_*: both warps of a tile's pair, half of the agents each)
PAIR_FIXTURES = ["tiny2_indiv", "tiny4_indiv", "small4_global", "medium8_hard_twostage"]


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("name", PAIR_FIXTURES)
def test_one_lane_step_matches_too(name, backend, monkeypatch):
    """RWARE_B200_PAIR=0 keeps the one-warp form of the step for those configurations; it must replay the reference too."""
    monkeypatch.setenv("RWARE_B200_PAIR", "0")
    fx = replay.load_fixture(name)
    drv = replay.VecDriver(fx, backend_factory(backend), num_envs=35)
    assert replay.replay(drv, fx) == len(fx["actions"])
    if backend == "hostsim":
        assert drv.vec.backend.pair_steps() == 0


@pytest.mark.parametrize("name", PAIR_FIXTURES)
def test_twin_takes_the_agent_split_step_by_default(name):
    """Guards the test above and the golden replays against silently testing one form only."""
    fx = replay.load_fixture(name)
    drv = replay.VecDriver(fx, backend_factory("hostsim"), num_envs=33)
    assert replay.replay(drv, fx, max_steps=50) == 50
    # (the fixtures inject a few degenerate states -- overlapping shelves -- which are stepped by the one-lane form)
    assert 0.9 * 33 * 50 <= drv.vec.backend.pair_steps() <= 33 * 50


@pytest.mark.parametrize("name", PAIR_FIXTURES)
def test_agent_split_phases_do_not_depend_on_the_order_of_the_two_parts(name, monkeypatch):
    """On the GPU the two warps of a pair run a phase concurrently.  The twin runs part 0 then part 1; RWARE_B200_PAIR=2 makes it
    run part 1 first.  Inside a phase the parts touch disjoint data, so the trajectory must be the reference's either way."""
    monkeypatch.setenv("RWARE_B200_PAIR", "2")
    fx = replay.load_fixture(name)
    drv = replay.VecDriver(fx, backend_factory("hostsim"), num_envs=33)
    assert replay.replay(drv, fx) == len(fx["actions"])
    assert drv.vec.backend.pair_steps() > 0


@pytest.mark.parametrize("backend", backend_params())
def test_worlds_with_two_agents_on_one_cell_take_the_one_lane_step(backend, monkeypatch):
    """An injected state with two agents on one cell is flagged like overlapping shelves and stepped by the one-lane form, so the
    agent-split step never sees two parts writing one cell; both forms must then agree world by world."""
    import numpy as np
    import torch
    import rware_b200
    outs = []
    for pair in ("1", "0"):
        monkeypatch.setenv("RWARE_B200_PAIR", pair)
        vec = rware_b200.make_vec("rware-tiny-4ag-v1", 64, _backend_factory=backend_factory(backend))
        vec.reset(seed=3)
        st = vec.get_state()
        ax, ay = st["agent_x"].clone(), st["agent_y"].clone()
        ax[::2, 3], ay[::2, 3] = ax[::2, 0], ay[::2, 0]   # every second world: agent 3 onto agent 0's cell
        vec.set_state(agent_x=ax, agent_y=ay)
        before = vec.backend.pair_steps() if backend == "hostsim" else None
        rs = np.random.default_rng(5)
        traj = []
        for t in range(6):
            a = torch.from_numpy(rs.integers(0, 5, size=(64, 4, 1)).astype(np.int32)).to(vec.device)
            o, r, d = vec.step(a)
            traj.append((o.cpu().clone(), r.cpu().clone(), d.cpu().clone()))
            if backend == "hostsim" and pair == "1" and t == 0:
                # the flagged half does not take the agent-split step (the flag goes away once the two agents have parted)
                assert vec.backend.pair_steps() - before == 32
        outs.append(traj)
        vec.close()
    for (o1, r1, d1), (o0, r0, d0) in zip(*outs):
        assert torch.equal(o1, o0) and torch.equal(r1, r0) and torch.equal(d1, d0)
