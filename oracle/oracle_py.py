"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- see oracle/rware_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (robotic-warehouse_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librware_oracle.so")


def build(force=False):
    """Compile oracle/_build/librware_oracle.so with gcc (a few hundred ms)."""
    srcs = [os.path.join(_HERE, f) for f in ("rware_oracle.c", "np_random.c", "rware_oracle.h", "np_random.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


class _Config(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "H", "W", "n_agents", "msg_bits", "sensor_range", "request_queue_size", "max_inactivity_steps",
        "max_steps", "reward_type", "normalised_coordinates", "n_goals")] + [
        ("highways", ctypes.c_void_p), ("goals_xy", ctypes.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(_Config)]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_n_shelves.argtypes = [ctypes.c_void_p]
        L.orc_obs_len.argtypes = [ctypes.c_void_p]
        L.orc_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        L.orc_reset.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.orc_step.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 4
        L.orc_step.restype = ctypes.c_int
        L.orc_obs.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_get_state.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 11
        L.orc_set_state.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 11
        L.orc_get_grid.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_obs_image.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_obs_image.restype = ctypes.c_int
        L.orc_global_image.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_global_image.restype = ctypes.c_int
        L.orc_last_step_stats.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 3
        L.orc_py_tuple_hash.restype = ctypes.c_uint64
        L.orc_py_tuple_hash.argtypes = [ctypes.c_int64, ctypes.c_int64]
        L.orc_py_set_order.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.orc_layout_from_params.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p]
        L.orc_rollout_random.restype = ctypes.c_int64
        L.orc_rollout_random.argtypes = [ctypes.POINTER(_Config), ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                         ctypes.c_int, ctypes.c_void_p]
        L.orc_batch_create.restype = ctypes.c_void_p
        L.orc_batch_create.argtypes = [ctypes.POINTER(_Config), ctypes.c_int64, ctypes.c_int]
        L.orc_batch_destroy.argtypes = [ctypes.c_void_p]
        L.orc_batch_reset.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_batch_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_batch_rollout_lockstep.restype = ctypes.c_int64
        L.orc_batch_rollout_lockstep.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64] + [ctypes.c_void_p] * 4
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def layout_from_params(shelf_columns, shelf_rows, column_height):
    """(H, W, highways[H,W] uint8, goals [(x,y),(x,y)]) per warehouse.py:294-326."""
    H = ctypes.c_int32()
    W = ctypes.c_int32()
    cap = 1 << 16
    hw = np.zeros(cap, np.uint8)
    goals = np.zeros(4, np.int32)
    rc = lib().orc_layout_from_params(shelf_columns, shelf_rows, column_height, ctypes.byref(H), ctypes.byref(W),
                                      _p(hw), cap, _p(goals))
    if rc:
        raise ValueError(f"orc_layout_from_params failed: {rc}")
    return H.value, W.value, hw[: H.value * W.value].reshape(H.value, W.value).copy(), [
        (int(goals[0]), int(goals[1])), (int(goals[2]), int(goals[3]))]


class OracleEnv:
    """One reference-semantics warehouse world on the CPU.  State arrays use the reference's conventions."""

    def __init__(self, highways, goals, n_agents, msg_bits=0, sensor_range=1, request_queue_size=1,
                 max_inactivity_steps=None, max_steps=None, reward_type=1, normalised_coordinates=False):
        self.highways = np.ascontiguousarray(highways, np.uint8)
        self.H, self.W = self.highways.shape
        self.goals = np.ascontiguousarray(np.asarray(goals, np.int32).reshape(-1, 2))
        self.n = n_agents
        self.msg_bits = msg_bits
        self.R = request_queue_size
        self._cfg = _Config(self.H, self.W, n_agents, msg_bits, sensor_range, request_queue_size,
                            int(max_inactivity_steps or 0), int(max_steps or 0), int(reward_type),
                            int(bool(normalised_coordinates)), len(self.goals),
                            self.highways.ctypes.data, self.goals.ctypes.data)
        self._h = lib().orc_create(ctypes.byref(self._cfg))
        if not self._h:
            raise ValueError("orc_create failed")
        self.S = lib().orc_n_shelves(self._h)
        self.L = lib().orc_obs_len(self._h)

    @property
    def config_struct(self):
        return self._cfg

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def seed(self, seed):
        lib().orc_seed(self._h, int(seed))

    def reset(self, seed=None):
        obs = np.zeros((self.n, self.L), np.float32)
        lib().orc_reset(self._h, -1 if seed is None else int(seed), _p(obs))
        return obs

    def step(self, actions):
        act = np.ascontiguousarray(np.asarray(actions, np.int32).reshape(self.n, 1 + self.msg_bits))
        obs = np.zeros((self.n, self.L), np.float32)
        rew = np.zeros(self.n, np.float64)
        done = np.zeros(1, np.uint8)
        rc = lib().orc_step(self._h, _p(act), _p(obs), _p(rew), _p(done))
        if rc:
            raise ValueError(f"orc_step failed: {rc}")
        return obs, rew, bool(done[0])

    def obs(self):
        obs = np.zeros((self.n, self.L), np.float32)
        lib().orc_obs(self._h, _p(obs))
        return obs

    def obs_image(self, layers=(0, 1, 2, 5, 6), directional=True, with_features=False):
        """IMAGE observations [n, C, K, K] (and IMAGE_DICT features [n, 6]); IndexError where the reference raises."""
        lay = np.ascontiguousarray(layers, np.int32)
        K = 2 * self._cfg.sensor_range + 1
        out = np.zeros((self.n, len(lay), K, K), np.float32)
        feat = np.zeros((self.n, 6), np.float32)
        rc = lib().orc_obs_image(self._h, _p(lay), len(lay), int(directional), _p(out), _p(feat))
        if rc == -1:
            raise IndexError("AGENT_DIRECTION / AGENT_LOAD layer: transposed index out of bounds (as in the reference)")
        if rc:
            raise ValueError(f"orc_obs_image failed: {rc}")
        return (out, feat) if with_features else out

    def global_image(self, layers=(0, 5), pad_to_shape=None):
        """get_global_image (warehouse.py:966-1040): [C, H, W] float32, zero-padded to pad_to_shape as the reference does."""
        lay = np.ascontiguousarray(layers, np.int32)
        out = np.zeros((len(lay), self.H, self.W), np.float32)
        rc = lib().orc_global_image(self._h, _p(lay), len(lay), _p(out))
        if rc == -1:
            raise IndexError("AGENT_DIRECTION / AGENT_LOAD layer: transposed index out of bounds (as in the reference)")
        if rc:
            raise ValueError(f"Unknown image layer type (orc_global_image: {rc})")
        if pad_to_shape is not None:  # 1021-1039
            diff = [p - g for p, g in zip(pad_to_shape, out.shape)]
            assert all(d >= 0 for d in diff)
            out = np.pad(out, tuple((d // 2, d // 2 if d % 2 == 0 else d // 2 + 1) for d in diff))
        return out

    def get_state(self):
        n, S, R, mb = self.n, self.S, self.R, self.msg_bits
        st = dict(agent_x=np.zeros(n, np.int32), agent_y=np.zeros(n, np.int32), agent_dir=np.zeros(n, np.int32),
                  agent_carry=np.zeros(n, np.int32), agent_deliv=np.zeros(n, np.int32),
                  agent_msg=np.zeros((n, max(mb, 1)), np.int32)[:, :mb].copy(),
                  shelf_x=np.zeros(S, np.int32), shelf_y=np.zeros(S, np.int32), queue=np.zeros(R, np.int32),
                  counters=np.zeros(2, np.int32), rng=np.zeros(10, np.uint32))
        amsg = np.zeros((n, max(mb, 1)), np.int32)
        lib().orc_get_state(self._h, _p(st["agent_x"]), _p(st["agent_y"]), _p(st["agent_dir"]), _p(st["agent_carry"]),
                            _p(st["agent_deliv"]), _p(amsg) if mb else None, _p(st["shelf_x"]), _p(st["shelf_y"]),
                            _p(st["queue"]), _p(st["counters"]), _p(st["rng"]))
        if mb:
            st["agent_msg"] = amsg.reshape(-1)[: n * mb].reshape(n, mb).copy()
        return st

    def set_state(self, **st):
        def arr(k, dt=np.int32):
            v = st.get(k)
            return None if v is None else np.ascontiguousarray(v, dt)
        keep = [arr("agent_x"), arr("agent_y"), arr("agent_dir"), arr("agent_carry"), arr("agent_deliv"),
                arr("agent_msg") if self.msg_bits else None, arr("shelf_x"), arr("shelf_y"), arr("queue"),
                arr("counters"), arr("rng", np.uint32)]
        lib().orc_set_state(self._h, *[_p(a) for a in keep])

    def grid(self):
        g = np.zeros((2, self.H, self.W), np.int32)
        lib().orc_get_grid(self._h, _p(g))
        return g

    def last_step_stats(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        lib().orc_last_step_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(ties=a.value, set_order_components=b.value, deliveries=c.value)


def rollout_random(env: OracleEnv, n_envs, n_steps, seed=0, n_threads=1):
    """Multi-threaded random-action rollout; returns (env_steps_done, reward_sum)."""
    rs = ctypes.c_double()
    total = lib().orc_rollout_random(ctypes.byref(env.config_struct), n_envs, n_steps, seed, n_threads, ctypes.byref(rs))
    return int(total), rs.value


class OracleBatch:
    """E independent oracle worlds stepped in lock step on `n_threads` host threads (orc_batch_*): the checker of the
    full-size parity tests, where every world of a BASELINE-sized batch is compared, not a sample."""

    def __init__(self, env: OracleEnv, n_envs, n_threads=None, auto_reset=True):
        self.proto = env  # keeps the config struct (and the arrays it points to) alive
        self.E, self.n, self.L = int(n_envs), env.n, env.L
        self.A = 1 + env.msg_bits
        self.auto_reset = bool(auto_reset)
        self._h = lib().orc_batch_create(ctypes.byref(env.config_struct), self.E, int(n_threads or os.cpu_count() or 1))
        self.obs = np.zeros((self.E, self.n, self.L), np.float32)
        self.reward = np.zeros((self.E, self.n), np.float32)
        self.done = np.zeros((self.E,), np.uint8)

    def reset(self, seeds=None):
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        lib().orc_batch_reset(self._h, None if s is None else _p(s), _p(self.obs))
        return self.obs

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.E, self.n, self.A)
        lib().orc_batch_step(self._h, _p(a), int(self.auto_reset), _p(self.obs), _p(self.reward), _p(self.done))
        return self.obs, self.reward, self.done

    def rollout_lockstep(self, n_steps, seed=0):
        """Random-action rollout in lock step (every step rewrites the stacked obs / reward / done arrays of the whole batch);
        returns (env_steps_done, reward_sum)."""
        rs = ctypes.c_double()
        total = lib().orc_batch_rollout_lockstep(self._h, int(n_steps), int(seed), _p(self.obs), _p(self.reward), _p(self.done),
                                                 ctypes.byref(rs))
        return int(total), rs.value

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
