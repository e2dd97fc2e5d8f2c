"""The reference's OWN test-suite (/root/reference/tests: 31 movement, 20 env, 6 goals, 4 wrapper cases), run unchanged
against the drop-in facade.

The files are copied to a temporary directory where `rware` is an alias package of three one-line modules that re-export
rware_b200 (`rware.warehouse`, `rware.utils.wrappers`), exactly what a user who switches packages would get.  The only edit is
the one SURVEY.md section 0 documents: `parametrize("time_limit,", ...)` (tests/test_env.py:367) is a collection error under
pytest 9, so the stray comma is removed in the copy.  When gymnasium is absent, `gymnasium.spaces` resolves to the package's
own API-compatible spaces (the facade builds its spaces from the same classes, so `==` comparisons are meaningful).

Needs /root/reference (present in the build container, absent on the GPU box: skipped there -- the restated tables of
tests/test_reference_behaviour.py cover the same behaviours on the box, through the CUDA path).  With a GPU and the
reference both present the suite also runs against the CUDA C-ABI.
"""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

from conftest import backend_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"


def _stage(tmp_path, backend):
    work = tmp_path / "suite"
    shutil.copytree(REF_TESTS, work / "tests")
    env_py = work / "tests" / "test_env.py"
    env_py.write_text(env_py.read_text().replace('"time_limit,"', '"time_limit"'))
    alias = work / "rware"
    (alias / "utils").mkdir(parents=True)
    factory = ("from hostsim.backend import HostSimBackend as _Factory" if backend == "hostsim"
               else "from rware_b200._capi import CudaBackend as _Factory")
    (alias / "__init__.py").write_text("import rware_b200  # registers the env ids\n")
    (alias / "warehouse.py").write_text(textwrap.dedent(f"""\
        import functools
        from rware_b200.warehouse import *  # noqa: F401,F403
        from rware_b200.warehouse import Warehouse as _Warehouse
        {factory}


        class Warehouse(_Warehouse):
            def __init__(self, *args, **kwargs):
                kwargs.setdefault("_backend_factory", _Factory)
                super().__init__(*args, **kwargs)
        """))
    (alias / "utils" / "__init__.py").write_text("")
    (alias / "utils" / "wrappers.py").write_text("from rware_b200.wrappers import *  # noqa: F401,F403\n"
                                                 "from rware_b200.wrappers import DictAgents, FlattenAgents, FlattenSAObservation\n")
    (alias / "utils" / "spaces.py").write_text("from rware_b200.spaces import MultiAgentActionSpace, MultiAgentObservationSpace\n")
    try:
        import gymnasium  # noqa: F401
    except Exception:
        gym = work / "gymnasium"
        gym.mkdir()
        (gym / "__init__.py").write_text("from rware_b200 import spaces\n")
        (gym / "spaces.py").write_text("from rware_b200.spaces import *  # noqa: F401,F403\n"
                                       "from rware_b200.spaces import flatdim, flatten\n")
    return work


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="/root/reference is not mounted here")
@pytest.mark.parametrize("backend", backend_params())
def test_reference_suite_runs_unchanged(tmp_path, backend):
    work = _stage(tmp_path, backend)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(work), ROOT, os.path.join(ROOT, "tests"), env.get("PYTHONPATH", "")])
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", str(work / "tests")],
                         capture_output=True, text=True, env=env, cwd=str(work), timeout=1500)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "61 passed" in res.stdout, tail
