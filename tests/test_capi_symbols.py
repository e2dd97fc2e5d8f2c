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


def test_host_widening_of_the_packed_transport():
    """rw_host_widen_u8: the host half of rw_step_host's packed transport (one byte per observation value over PCIe, widened
    to float32 by a thread pool).  Needs no GPU: exact for every byte value, any length and alignment, repeated calls."""
    import numpy as np
    from rware_b200 import _capi
    lib = _capi.load_library()
    rs = np.random.default_rng(0)
    for n in (0, 1, 7, 31, 32, 33, 255, 256, 1000, 4097, 300_001, 2_000_003):
        for shift in (0, 1, 5):
            src = rs.integers(0, 256, size=n, dtype=np.uint8)
            if n >= 256:
                src[:256] = np.arange(256, dtype=np.uint8)
            buf = np.full(n + 16, -1.0, np.float32)
            dst = buf[shift:shift + n]
            threads = lib.rw_host_widen_u8(src.ctypes.data, dst.ctypes.data, n)
            assert threads >= 1
            assert np.array_equal(dst, src.astype(np.float32)), (n, shift)
            assert np.all(buf[:shift] == -1.0) and np.all(buf[shift + n:] == -1.0)  # nothing written outside


def test_host_widening_from_several_threads_and_after_fork():
    """The pool serves one widening at a time (callers queue on a mutex) and a forked child starts its own pool."""
    import os
    import threading
    import numpy as np
    from rware_b200 import _capi
    lib = _capi.load_library()
    n = 500_000
    srcs = [np.random.default_rng(k).integers(0, 256, size=n, dtype=np.uint8) for k in range(4)]
    dsts = [np.zeros(n, np.float32) for _ in range(4)]

    def work(k):
        for _ in range(5):
            lib.rw_host_widen_u8(srcs[k].ctypes.data, dsts[k].ctypes.data, n)
    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(4):
        assert np.array_equal(dsts[k], srcs[k].astype(np.float32))
    pid = os.fork()
    if pid == 0:  # child: the parent's worker threads do not exist here
        out = np.zeros(n, np.float32)
        lib.rw_host_widen_u8(srcs[0].ctypes.data, out.ctypes.data, n)
        os._exit(0 if np.array_equal(out, srcs[0].astype(np.float32)) else 1)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0


@pytest.mark.gpu
def test_c_abi_against_the_oracle_without_python():
    """tests/ctest/gpu_quickcheck: a plain CUDA host program that links librware_b200.so and the C oracle and compares IMAGE
    observations with the transposed agent layers, the index-error count, rewards, done flags and rw_obs_bytes over 46 steps."""
    import subprocess
    exe = os.path.join(ROOT, "tests", "ctest", "_build", "gpu_quickcheck")
    if not os.path.exists(exe):
        pytest.skip("tests/ctest/_build/gpu_quickcheck not built (__graft_entry__.build() builds it)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "gpu_quickcheck OK" in res.stdout, res.stdout + res.stderr
