"""TEST-ONLY backend: drives tests/hostsim/hostsim.cpp (the product's per-world logic compiled for the host) behind the
same interface as rware_b200._capi.CudaBackend, with CPU torch tensors.  Lets `pytest -m "not gpu"` exercise the
kernel logic, the packed layout and the Python facade without a GPU.  Never used by the product."""
import ctypes
import os
import subprocess

import numpy as np
import torch

from rware_b200._capi import STATE_FIELDS, RwStateView

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libhostsim.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "hostsim.cpp")
        csrc = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "robotic-warehouse_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in ("rw_logic.cuh", "rw_coop.cuh", "rw_views.cuh", "rw_common.h", "rw_host_layout.h", "rw_static_configs.h")]
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(d) for d in deps):
            os.makedirs(os.path.dirname(_LIB), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _LIB, src], check=True)
        L = ctypes.CDLL(_LIB)
        vp = ctypes.c_void_p
        L.sim_create.restype = vp
        L.sim_create.argtypes = [vp, ctypes.c_int64]
        L.sim_destroy.argtypes = [vp]
        L.sim_obs_len.argtypes = [vp]
        L.sim_n_shelves.argtypes = [vp]
        L.sim_set_auto_reset.argtypes = [vp, ctypes.c_int]
        L.sim_error_count.argtypes = [vp]
        L.sim_error_count.restype = ctypes.c_longlong
        L.sim_seed.argtypes = [vp, vp, vp]
        L.sim_run.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp]
        L.sim_force_dynamic.argtypes = [ctypes.c_int]
        L.sim_set_coop.argtypes = [vp, ctypes.c_int]
        L.sim_set_pair.argtypes = [vp, ctypes.c_int]
        L.sim_pair_steps.argtypes = [vp]
        L.sim_pair_steps.restype = ctypes.c_longlong
        L.sim_get_state.argtypes = [vp, ctypes.POINTER(RwStateView)]
        L.sim_set_state.argtypes = [vp, ctypes.POINTER(RwStateView)]
        L.sim_global_image.argtypes = [vp, vp, ctypes.c_int, vp, vp]
        L.sim_render_shape.argtypes = [vp, vp, vp]
        L.sim_render_rgb.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, vp]
        _lib = L
    return _lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class HostSimBackend:
    def __init__(self, cfg, num_envs, device=None):
        self.lib = _load()
        self.device = torch.device("cpu")
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self._struct = cfg.to_struct()
        self._h = self.lib.sim_create(ctypes.byref(self._struct), self.num_envs)
        self.info = dict(num_envs=self.num_envs, n_agents=cfg.n_agents, n_shelves=self.lib.sim_n_shelves(self._h),
                         obs_len=self.lib.sim_obs_len(self._h))
        self._ready = False
        if os.environ.get("RWARE_B200_COOP") is not None:
            self.lib.sim_set_coop(self._h, int(os.environ["RWARE_B200_COOP"]))

    def close(self):
        if self._h:
            self.lib.sim_destroy(self._h)
            self._h = None

    def seed(self, seeds, mask=None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.lib.sim_seed(self._h, seeds.ctypes.data, None if m is None else m.ctypes.data)

    def reset(self, mask, obs):
        self.lib.sim_run(self._h, 2, None, _ptr(obs), None, None, _ptr(mask))
        self._ready = True

    def step(self, actions, obs, reward, done):
        assert self._ready
        self.lib.sim_run(self._h, 0, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), None)

    def obs(self, obs):
        self.lib.sim_run(self._h, 1, None, _ptr(obs), None, None, None)

    def set_auto_reset(self, on):
        self.lib.sim_set_auto_reset(self._h, int(bool(on)))

    def _view(self, tensors):
        v = RwStateView()
        for k in STATE_FIELDS:
            t = tensors.get(k)
            setattr(v, k, None if t is None else t.data_ptr())
        return v

    def get_state(self, tensors):
        self.lib.sim_get_state(self._h, ctypes.byref(self._view(tensors)))

    def set_state(self, tensors):
        self.lib.sim_set_state(self._h, ctypes.byref(self._view(tensors)))
        self._ready = True

    def error_count(self):
        return int(self.lib.sim_error_count(self._h))

    def pair_steps(self):
        """World-steps taken through the agent-split form of the step (rw_logic.cuh, rw_pair_*)."""
        return int(self.lib.sim_pair_steps(self._h))

    def set_pair(self, on):
        self.lib.sim_set_pair(self._h, int(bool(on)))

    def global_image_shape(self, n_layers, pad_to_shape=None):
        H, W = self.cfg.grid_size
        shape = (n_layers, H, W) if pad_to_shape is None else tuple(int(v) for v in pad_to_shape)
        if len(shape) != 3 or any(a < b for a, b in zip(shape, (n_layers, H, W))):
            raise RuntimeError("pad_to_shape must be >= (n_layers, H, W) elementwise")
        return shape

    def global_image(self, layers, pad_to_shape, out):
        lay = np.ascontiguousarray(layers, np.int32)
        pad = None if pad_to_shape is None else np.ascontiguousarray(pad_to_shape, np.int32)
        rc = self.lib.sim_global_image(self._h, lay.ctypes.data, len(lay), None if pad is None else pad.ctypes.data, _ptr(out))
        if rc:
            raise RuntimeError("rware_b200 C-ABI call failed: unknown image layer type or bad pad_to_shape")

    def render_shape(self):
        h, w = ctypes.c_int32(), ctypes.c_int32()
        self.lib.sim_render_shape(self._h, ctypes.byref(h), ctypes.byref(w))
        return h.value, w.value

    def render_rgb(self, first_env, n_envs, out):
        if self.lib.sim_render_rgb(self._h, int(first_env), int(n_envs), _ptr(out)):
            raise RuntimeError("render_rgb: world range out of bounds")

    def obs_bytes(self, obs, out):
        if self.cfg.normalised_coordinates:
            raise RuntimeError("rw_obs_bytes: normalised coordinates are not integers")
        out.copy_(obs.to(torch.uint8))
