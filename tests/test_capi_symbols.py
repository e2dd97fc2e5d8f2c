"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/rware_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "rware_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rw_[a-z_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = _declared_functions()
    for must in ("rw_create", "rw_destroy", "rw_reset", "rw_step", "rw_obs", "rw_get_state", "rw_set_state", "rw_seed",
                 "rw_last_error", "rw_step_host"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from rware_b200 import _capi
    _capi.build_library()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/rware_b200.h but not exported"


def test_create_rejects_bad_configs_without_a_gpu_call():
    from rware_b200 import _capi
    lib = _capi.load_library()
    import numpy as np
    cfg = _capi.EnvConfig(np.zeros((3, 3), np.uint8), [(0, 0)], n_agents=40)
    h = ctypes.c_void_p()
    s = cfg.to_struct()
    assert lib.rw_create(ctypes.byref(s), 4, 0, ctypes.byref(h)) == -1
    assert b"n_agents" in lib.rw_last_error(None)


def test_product_path_raises_without_cuda():
    import torch
    import rware_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rware_b200.make("rware-tiny-2ag-v2")
