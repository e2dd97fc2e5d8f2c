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
