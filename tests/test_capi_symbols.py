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


def _expand_bits_reference(rows, L, nwide, rs):
    """Random rows of the shape the bit transport carries: `nwide` leading byte values, then 0/1 flags."""
    import numpy as np
    obs = rs.integers(0, 2, size=(rows, L)).astype(np.float32)
    wide = rs.integers(0, 256, size=(rows, max(nwide, 1)), dtype=np.uint8)[:, :nwide]
    obs[:, :nwide] = wide
    flat = obs.reshape(-1)
    bits = np.packbits((flat != 0).astype(np.uint8), bitorder="little")
    bits = np.concatenate([bits, np.zeros(8, np.uint8)])  # the vector forms read whole 16-bit masks
    return flat, bits, np.ascontiguousarray(wide)


_EXPAND_CASES = "((1, 71, 2), (3, 15, 2), (64, 71, 2), (1000, 71, 2), (257, 183, 2), (5000, 16, 0), (4099, 45, 0), (20000, 71, 2), (70000, 8, 2))"


def test_host_expansion_of_the_bit_transport():
    """rw_host_expand_bits: the host half of rw_step_host's bit transport (one bit per 0/1 flag, one byte per coordinate over
    PCIe; float32 rows rebuilt by the thread pool).  Needs no GPU: exact for ragged sizes, row lengths that straddle the
    64-byte lines in every phase, aligned and unaligned destinations, repeated calls."""
    import numpy as np
    from rware_b200 import _capi
    lib = _capi.load_library()
    rs = np.random.default_rng(1)
    for rows, L, nwide in eval(_EXPAND_CASES):
        flat, bits, wide = _expand_bits_reference(rows, L, nwide, rs)
        for shift in (0, 16, 3):  # 64-byte aligned (vector form), 64-byte aligned at another line, unaligned (scalar form)
            raw = np.full(flat.size + 64, -1.0, np.float32)
            base = (-raw.ctypes.data // 4) % 16  # first 64-byte aligned float
            dst = raw[base + shift: base + shift + flat.size]
            threads = lib.rw_host_expand_bits(bits.ctypes.data, wide.ctypes.data if nwide else None, dst.ctypes.data, rows, L, nwide)
            assert threads >= 1
            assert np.array_equal(dst, flat), (rows, L, nwide, shift)
            assert np.all(raw[:base + shift] == -1.0) and np.all(raw[base + shift + flat.size:] == -1.0)


@pytest.mark.parametrize("isa", [0, 1])
def test_host_expansion_scalar_and_avx2_forms(isa):
    """The same check with the instruction set capped (RWARE_B200_HOST_ISA): the forms a host without AVX-512 / AVX2 runs."""
    import subprocess
    import sys
    code = f"""
import sys, numpy as np
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
from rware_b200 import _capi
from test_capi_symbols import _expand_bits_reference, _EXPAND_CASES
lib = _capi.load_library()
rs = np.random.default_rng(2)
for rows, L, nwide in eval(_EXPAND_CASES):
    flat, bits, wide = _expand_bits_reference(rows, L, nwide, rs)
    raw = np.full(flat.size + 64, -1.0, np.float32)
    base = (-raw.ctypes.data // 4) % 16
    dst = raw[base: base + flat.size]
    lib.rw_host_expand_bits(bits.ctypes.data, wide.ctypes.data if nwide else None, dst.ctypes.data, rows, L, nwide)
    assert np.array_equal(dst, flat), (rows, L, nwide)
    src = rs.integers(0, 256, size=flat.size, dtype=np.uint8)
    lib.rw_host_widen_u8(src.ctypes.data, dst.ctypes.data, src.size)
    assert np.array_equal(dst, src.astype(np.float32))
print("ok")
"""
    env = dict(os.environ, RWARE_B200_HOST_ISA=str(isa), RWARE_B200_HOST_THREADS="3")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout + res.stderr


def test_host_stream_probe_reports_a_bandwidth():
    import numpy as np
    from rware_b200 import _capi
    lib = _capi.load_library()
    buf = np.zeros(1 << 22, np.float32)
    gbs = ctypes.c_double()
    threads = lib.rw_host_stream_probe(buf.ctypes.data, buf.size, 2, ctypes.byref(gbs))
    assert threads >= 1 and gbs.value > 0.05 and np.all(buf == 2.0)


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
