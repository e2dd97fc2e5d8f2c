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
