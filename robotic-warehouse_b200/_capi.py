"""ctypes binding of the C ABI (include/rware_b200.h) -> librware_b200.so, and the CUDA backend object the
Python facade drives.  PyTorch supplies device memory and streams; all compute is in the library's own kernels.

There is no CPU fallback: if the library is missing or CUDA is unavailable, construction raises.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# RWARE_B200_LIB_PATH: load another build of the library (A/B measurements of build-time options); default: the in-tree build
LIB_PATH = os.environ.get("RWARE_B200_LIB_PATH") or os.path.join(_PKG_DIR, "lib", "librware_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "rware_b200.h")
RW_ABI_VERSION = 2
RNG_WORDS = 10


class RwConfig(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "abi_version", "grid_h", "grid_w", "n_agents", "msg_bits", "sensor_range", "request_queue_size",
        "max_inactivity_steps", "max_steps", "reward_type", "normalised_coordinates", "n_goals", "observation_type",
        "image_directional", "n_image_layers")] + [("image_layers", ctypes.c_int32 * 8),
                                                   ("highways", ctypes.c_void_p), ("goals_xy", ctypes.c_void_p)]


class RwInfo(ctypes.Structure):
    _fields_ = [("num_envs", ctypes.c_int64)] + [(k, ctypes.c_int32) for k in (
        "n_agents", "n_shelves", "obs_len", "act_width", "request_queue_size", "grid_h", "grid_w", "device",
        "state_words_per_env", "envs_per_tile", "smem_bytes_per_warp", "warps_per_cta")] + [
        (k, ctypes.c_int64) for k in ("algorithmic_bytes_step", "algorithmic_bytes_obs", "packed_bytes_step",
                                      "packed_bytes_obs")]


STATE_FIELDS = ("agent_x", "agent_y", "agent_dir", "agent_carry", "agent_has_delivered", "agent_msg", "shelf_x",
                "shelf_y", "queue", "cur_steps", "cur_inactive", "rng")


class RwStateView(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in STATE_FIELDS]


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> lib/librware_b200.so (in-tree, travels with the repo)."""
    srcs = [os.path.join(_PKG_DIR, "csrc", f) for f in os.listdir(os.path.join(_PKG_DIR, "csrc"))] + [HEADER_PATH]
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    res = subprocess.run([os.path.join(_PKG_DIR, "build.sh")], capture_output=True, text=True)
    if verbose or res.returncode:
        print(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("building librware_b200.so failed")
    return LIB_PATH


_lib = None


def load_library():
    """Loads librware_b200.so (raises if it is missing: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
            f"or robotic-warehouse_b200/build.sh).  rware_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, u64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
    L.rw_create.argtypes = [ctypes.POINTER(RwConfig), i64, ci, ctypes.POINTER(vp)]
    L.rw_destroy.argtypes = [vp]
    L.rw_get_info.argtypes = [vp, ctypes.POINTER(RwInfo)]
    L.rw_last_error.argtypes = [vp]
    L.rw_last_error.restype = ctypes.c_char_p
    L.rw_seed.argtypes = [vp, vp, vp, vp]
    L.rw_reset.argtypes = [vp, vp, vp, vp]
    L.rw_step.argtypes = [vp, vp, vp, vp, vp, vp]
    L.rw_obs.argtypes = [vp, vp, vp]
    L.rw_set_auto_reset.argtypes = [vp, ci]
    L.rw_get_state.argtypes = [vp, ctypes.POINTER(RwStateView), vp]
    L.rw_set_state.argtypes = [vp, ctypes.POINTER(RwStateView), vp]
    L.rw_error_count.argtypes = [vp, ctypes.POINTER(i64), vp]
    L.rw_step_host.argtypes = [vp, vp, vp, vp, vp]
    L.rw_reset_host.argtypes = [vp, vp, vp]
    L.rw_host_alloc.argtypes = [ctypes.POINTER(vp), i64]
    L.rw_host_free.argtypes = [vp]
    L.rw_random_actions.argtypes = [vp, vp, u64, u64, vp]
    L.rw_debug_set_clock_buffer.argtypes = [vp, vp]
    L.rw_flush_l2.argtypes = [vp, vp, i64, ci, vp]
    L.rw_gather_create.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp), vp]
    L.rw_gather_open.argtypes = [vp, vp]
    L.rw_step_gather.argtypes = [vp, vp, vp]
    L.rw_gather_slot.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.rw_global_image_shape.argtypes = [vp, ctypes.c_int32, vp, vp]
    L.rw_global_image.argtypes = [vp, vp, ctypes.c_int32, vp, vp, vp]
    L.rw_render_shape.argtypes = [vp, vp, vp]
    L.rw_render_rgb.argtypes = [vp, i64, i64, vp, vp]
    L.rw_set_host_transport.argtypes = [vp, ci]
    L.rw_obs_bytes.argtypes = [vp, vp, vp, vp]
    L.rw_host_widen_u8.argtypes = [vp, vp, i64]
    L.rw_host_expand_bits.argtypes = [vp, vp, vp, i64, ctypes.c_int32, ctypes.c_int32]
    L.rw_host_stream_probe.argtypes = [vp, i64, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    L.rw_set_host_dma_share.argtypes = [vp, ctypes.c_double]
    L.rw_get_host_dma_share.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.rw_host_alloc_near.argtypes = [vp, ctypes.POINTER(vp), i64]
    L.rw_host_numa_node.argtypes = [vp]
    L.rw_gpu_numa_node.argtypes = [vp]
    for name in ("rw_create", "rw_destroy", "rw_get_info", "rw_seed", "rw_reset", "rw_step", "rw_obs", "rw_set_auto_reset",
                 "rw_get_state", "rw_set_state", "rw_error_count", "rw_step_host", "rw_reset_host", "rw_host_alloc",
                 "rw_host_free", "rw_random_actions", "rw_debug_set_clock_buffer", "rw_flush_l2", "rw_gather_create", "rw_gather_open", "rw_step_gather", "rw_gather_slot",
                 "rw_global_image_shape", "rw_global_image", "rw_render_shape", "rw_render_rgb", "rw_set_host_transport",
                 "rw_host_widen_u8", "rw_obs_bytes", "rw_host_expand_bits", "rw_host_stream_probe", "rw_set_host_dma_share",
                 "rw_get_host_dma_share", "rw_host_alloc_near", "rw_host_numa_node", "rw_gpu_numa_node"):
        getattr(L, name).restype = ci
    _lib = L
    return L


class EnvConfig:
    """Constructor kwargs of the reference's Warehouse (rware/warehouse.py:146-170) after layout construction."""

    def __init__(self, highways, goals, n_agents, msg_bits=0, sensor_range=1, request_queue_size=1,
                 max_inactivity_steps=None, max_steps=None, reward_type=1, normalised_coordinates=False,
                 observation_type=1, image_layers=(0, 1, 2, 5, 6), image_directional=True):
        self.highways = np.ascontiguousarray(highways, dtype=np.uint8)
        self.grid_size = tuple(int(v) for v in self.highways.shape)
        self.goals = [(int(x), int(y)) for x, y in goals]
        self.n_agents = int(n_agents)
        self.msg_bits = int(msg_bits)
        self.sensor_range = int(sensor_range)
        self.request_queue_size = int(request_queue_size)
        self.max_inactivity_steps = max_inactivity_steps
        self.max_steps = max_steps
        self.reward_type = int(getattr(reward_type, "value", reward_type))
        self.normalised_coordinates = bool(normalised_coordinates)
        self.n_shelves = int((self.highways == 0).sum())
        self.observation_type = int(getattr(observation_type, "value", observation_type))
        self.image_layers = [int(getattr(v, "value", v)) for v in image_layers]
        self.image_directional = bool(image_directional)
        K = 2 * self.sensor_range + 1
        if self.observation_type >= 2:  # IMAGE / IMAGE_DICT rows: [C][K][K] image (+ 6 features)
            self.obs_len = len(self.image_layers) * K * K + (6 if self.observation_type == 3 else 0)
        else:
            self.obs_len = 8 + K * K * (7 + self.msg_bits)  # rware/warehouse.py:432-443

    def to_struct(self):
        self._goals_arr = np.ascontiguousarray(np.asarray(self.goals, np.int32).reshape(-1, 2))
        H, W = self.grid_size
        return RwConfig(RW_ABI_VERSION, H, W, self.n_agents, self.msg_bits, self.sensor_range, self.request_queue_size,
                        int(self.max_inactivity_steps or 0), int(self.max_steps or 0), self.reward_type,
                        int(self.normalised_coordinates), len(self.goals), self.observation_type,
                        int(self.image_directional), len(self.image_layers) if self.observation_type >= 2 else 0,
                        (ctypes.c_int32 * 8)(*(self.image_layers + [0] * 8)[:8]), self.highways.ctypes.data,
                        self._goals_arr.ctypes.data)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class CudaBackend:
    """One rw_handle on one GPU.  All tensors passed in must live on `self.device` and be contiguous."""

    def __init__(self, cfg: EnvConfig, num_envs: int, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("rware_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = load_library()
        dev = torch.device("cuda") if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"rware_b200 runs on CUDA devices only (got {dev}); there is no CPU fallback")
        # 'cuda' without an index means the current device (under torchrun: the one set from LOCAL_RANK), not GPU 0
        self.device = torch.device("cuda", torch.cuda.current_device() if dev.index is None else dev.index)
        self.cfg = cfg
        self.num_envs = int(num_envs)
        self._struct = cfg.to_struct()
        h = ctypes.c_void_p()
        rc = self.lib.rw_create(ctypes.byref(self._struct), self.num_envs, self.device.index, ctypes.byref(h))
        if rc:
            raise RuntimeError(f"rw_create failed ({rc}): {self.lib.rw_last_error(None).decode()}")
        self._h = h
        info = RwInfo()
        self._check(self.lib.rw_get_info(self._h, ctypes.byref(info)))
        self.info = {k: getattr(info, k) for k, _ in RwInfo._fields_}

    def _check(self, rc):
        if rc:
            raise RuntimeError(f"rware_b200 C-ABI call failed ({rc}): {self.lib.rw_last_error(self._h).decode()}")

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            for ptr in getattr(self, "_host_blocks", []):  # tensors handed out by host_alloc() die with the backend
                self.lib.rw_host_free(ctypes.c_void_p(ptr))
            self._host_blocks = []
            self.lib.rw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seeds, mask=None):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        assert seeds.shape == (self.num_envs,)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._check(self.lib.rw_seed(self._h, seeds.ctypes.data, None if m is None else m.ctypes.data, self._stream()))

    def reset(self, mask, obs):
        self._check(self.lib.rw_reset(self._h, _ptr(mask), _ptr(obs), self._stream()))

    def step(self, actions, obs, reward, done):
        self._check(self.lib.rw_step(self._h, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), self._stream()))

    def obs(self, obs):
        self._check(self.lib.rw_obs(self._h, _ptr(obs), self._stream()))

    def set_auto_reset(self, on):
        self._check(self.lib.rw_set_auto_reset(self._h, int(bool(on))))

    def _view(self, tensors):
        v = RwStateView()
        for k in STATE_FIELDS:
            t = tensors.get(k)
            setattr(v, k, None if t is None else t.data_ptr())
        return v

    def get_state(self, tensors):
        self._check(self.lib.rw_get_state(self._h, ctypes.byref(self._view(tensors)), self._stream()))

    def set_state(self, tensors):
        self._check(self.lib.rw_set_state(self._h, ctypes.byref(self._view(tensors)), self._stream()))

    def error_count(self):
        out = ctypes.c_int64()
        self._check(self.lib.rw_error_count(self._h, ctypes.byref(out), self._stream()))
        return out.value

    # ---- whole-warehouse views
    def global_image_shape(self, n_layers, pad_to_shape=None):
        pad = None if pad_to_shape is None else (ctypes.c_int32 * 3)(*[int(v) for v in pad_to_shape])
        shape = (ctypes.c_int32 * 3)()
        if len(pad_to_shape or (0, 0, 0)) != 3 or self.lib.rw_global_image_shape(self._h, int(n_layers), pad, shape):
            raise RuntimeError("pad_to_shape must be >= (n_layers, H, W) elementwise (1..8 layers)")
        return tuple(shape)

    def global_image(self, layers, pad_to_shape, out):
        lay = (ctypes.c_int32 * len(layers))(*[int(v) for v in layers])
        pad = None if pad_to_shape is None else (ctypes.c_int32 * 3)(*[int(v) for v in pad_to_shape])
        self._check(self.lib.rw_global_image(self._h, lay, len(layers), pad, _ptr(out), self._stream()))

    def render_shape(self):
        h, w = ctypes.c_int32(), ctypes.c_int32()
        self._check(self.lib.rw_render_shape(self._h, ctypes.byref(h), ctypes.byref(w)))
        return h.value, w.value

    def render_rgb(self, first_env, n_envs, out):
        self._check(self.lib.rw_render_rgb(self._h, int(first_env), int(n_envs), _ptr(out), self._stream()))

    # ---- fused push all-gather (multi-GPU)
    def gather_create(self, world_size, rank):
        """Allocates this rank's gathered obs/reward/done buffers (two slots, written alternately); returns
        ([slot 0 tensors, slot 1 tensors], 64-byte IPC handle)."""
        po, pr, pd = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self._check(self.lib.rw_gather_create(self._h, world_size, rank, ctypes.byref(po), ctypes.byref(pr),
                                              ctypes.byref(pd), handle))
        E, n, L = world_size * self.num_envs, self.cfg.n_agents, self.cfg.obs_len

        def wrap(ptr, shape, typestr, dtype):
            class _Raw:  # __cuda_array_interface__ view of library-owned memory
                __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
            return torch.as_tensor(_Raw(), device=self.device).view(dtype).view(shape)
        slots = []
        for slot in range(2):
            self._check(self.lib.rw_gather_slot(self._h, slot, ctypes.byref(po), ctypes.byref(pr), ctypes.byref(pd)))
            slots.append((wrap(po.value, (E, n, L), "<f4", torch.float32), wrap(pr.value, (E, n), "<f4", torch.float32),
                          wrap(pd.value, (E,), "|u1", torch.uint8)))
        self._gather_slots = slots
        return slots, handle.raw

    def gather_open(self, handles: bytes):
        buf = ctypes.create_string_buffer(handles, len(handles))
        self._check(self.lib.rw_gather_open(self._h, buf))

    def step_gather(self, actions):
        """Returns the slot (0 / 1) of the gathered tensors this step is being pushed into."""
        slot = self.lib.rw_step_gather(self._h, _ptr(actions), self._stream())
        if slot < 0:
            self._check(slot)
        return slot

    def flush_l2(self, buf, value=0):
        self._check(self.lib.rw_flush_l2(self._h, _ptr(buf), buf.numel() * buf.element_size(), int(value), self._stream()))

    def set_clock_buffer(self, tensor):
        self._check(self.lib.rw_debug_set_clock_buffer(self._h, _ptr(tensor)))

    def random_actions(self, actions, seed, step):
        self._check(self.lib.rw_random_actions(self._h, _ptr(actions), int(seed), int(step), self._stream()))

    # end-to-end entry points with host (pinned) buffers
    def step_host(self, actions, obs, reward, done):
        self._check(self.lib.rw_step_host(self._h, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done)))

    def obs_bytes(self, obs, out):
        self._check(self.lib.rw_obs_bytes(self._h, _ptr(obs), _ptr(out), self._stream()))

    def set_host_transport(self, mode):
        """Transport of step_host: 0 / False = plain float32 copy, 1 / True = compressed (one bit per flag and one byte per
        coordinate over PCIe where the rows allow it, else one byte per value; expanded by the library's host threads while the
        copy engine delivers a share of the rows as plain float32), 2 = compressed, bytes only."""
        self._check(self.lib.rw_set_host_transport(self._h, int(mode)))

    def set_host_dma_share(self, share):
        """Fixes the share of the rows the copy engine delivers as float32 (0..1); None = follow the measured rates."""
        self._check(self.lib.rw_set_host_dma_share(self._h, -1.0 if share is None else float(share)))

    def host_dma_share(self):
        out = ctypes.c_double()
        self._check(self.lib.rw_get_host_dma_share(self._h, ctypes.byref(out)))
        return out.value

    def host_alloc(self, shape, dtype=torch.float32):
        """Page-locked host tensor on the NUMA node of this GPU (rw_host_alloc_near); freed with the backend."""
        shape = tuple(int(v) for v in shape)
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        ptr = ctypes.c_void_p()
        self._check(self.lib.rw_host_alloc_near(self._h, ctypes.byref(ptr), max(nbytes, 1)))
        if not hasattr(self, "_host_blocks"):
            self._host_blocks = []
        self._host_blocks.append(ptr.value)
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
        return torch.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).view(shape)

    def host_numa_node(self, tensor):
        return int(self.lib.rw_host_numa_node(ctypes.c_void_p(tensor.data_ptr())))

    def gpu_numa_node(self):
        return int(self.lib.rw_gpu_numa_node(self._h))

    def host_stream_probe(self, tensor, repeats=3):
        """GB/s the host pool streams into `tensor` (float32, page-locked) with non-temporal stores: the roofline of the
        compressed transport's expansion.  Overwrites the tensor."""
        out = ctypes.c_double()
        threads = self.lib.rw_host_stream_probe(ctypes.c_void_p(tensor.data_ptr()), tensor.numel(), int(repeats), ctypes.byref(out))
        if threads < 0:
            raise RuntimeError("rw_host_stream_probe failed")
        return out.value, threads

    def reset_host(self, seeds, obs):
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        self._check(self.lib.rw_reset_host(self._h, None if s is None else s.ctypes.data, _ptr(obs)))
