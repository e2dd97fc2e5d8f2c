"""Replay harness: drives an implementation through a golden trajectory recorded from the unmodified reference
(tests/golden/make_golden.py) and asserts bit-exact agreement on every field after every step."""
import glob
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

STATE_KEYS = ("agent_x", "agent_y", "agent_dir", "agent_carry", "agent_deliv", "agent_msg", "shelf_x", "shelf_y",
              "queue", "counters", "rng")


def fixture_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(HERE, "golden", "traj_*.npz")))


def load_fixture(name):
    fx = dict(np.load(os.path.join(HERE, "golden", f"traj_{name}.npz")))
    fx["kwargs"] = json.loads(str(fx["kwargs"]))
    return fx


class OracleDriver:
    """The C oracle (one world)."""

    def __init__(self, fx):
        import oracle_py
        k = fx["kwargs"]
        self.env = oracle_py.OracleEnv(fx["highways"], fx["goals"], k["n_agents"], k["msg_bits"], k["sensor_range"],
                                       k["request_queue_size"], k["max_inactivity_steps"], k["max_steps"], k["reward_type"],
                                       k.get("normalised_coordinates", False))

    def reset(self, seed):
        return self.env.reset(seed)

    def step(self, actions):
        return self.env.step(actions)

    def get_state(self):
        return self.env.get_state()

    def set_state(self, st):
        self.env.set_state(**st)


class VecDriver:
    """VecWarehouse with B identical worlds (all seeded alike) on a given backend; checks that all worlds agree."""

    def __init__(self, fx, backend_factory, num_envs):
        import rware_b200
        k = dict(fx["kwargs"])
        self.B = num_envs
        self.vec = rware_b200.VecWarehouse(num_envs, _backend_factory=backend_factory, **k)
        self.n, self.mb = k["n_agents"], k["msg_bits"]

    def _same(self, t):
        a = t.cpu().numpy()
        assert (a == a[:1]).all(), "worlds of the batch diverged"
        return a[0]

    def reset(self, seed):
        if seed is not None:
            self.vec.seed(np.full(self.B, seed, np.uint64))
        return self._same(self.vec.reset())

    def step(self, actions):
        import torch
        a = np.broadcast_to(np.asarray(actions, np.int32).reshape(1, self.n, 1 + self.mb), (self.B, self.n, 1 + self.mb))
        obs, rew, done = self.vec.step(torch.from_numpy(np.ascontiguousarray(a)))
        return self._same(obs), self._same(rew).astype(np.float64), bool(self._same(done))

    def get_state(self):
        st = {k: self._same(v) for k, v in self.vec.get_state().items()}
        out = dict(agent_x=st["agent_x"], agent_y=st["agent_y"], agent_dir=st["agent_dir"], agent_carry=st["agent_carry"],
                   agent_deliv=st["agent_has_delivered"], shelf_x=st["shelf_x"], shelf_y=st["shelf_y"], queue=st["queue"],
                   counters=np.array([st["cur_steps"], st["cur_inactive"]]), rng=st["rng"].astype(np.int64) & 0xFFFFFFFF)
        out["agent_msg"] = st.get("agent_msg", np.zeros((self.n, 0), np.int32))
        return out

    def set_state(self, st):
        f = dict(agent_x=st["agent_x"], agent_y=st["agent_y"], agent_dir=st["agent_dir"], agent_carry=st["agent_carry"],
                 agent_has_delivered=st["agent_deliv"], shelf_x=st["shelf_x"], shelf_y=st["shelf_y"], queue=st["queue"],
                 cur_steps=st["counters"][0:1], cur_inactive=st["counters"][1:2])
        if self.mb:
            f["agent_msg"] = st["agent_msg"]
        rep = {k: np.broadcast_to(np.asarray(v, np.int64)[None], (self.B,) + np.asarray(v).shape).reshape(
            (self.B,) + (np.asarray(v).shape if k not in ("cur_steps", "cur_inactive") else ())) for k, v in f.items()}
        self.vec.set_state(**rep)


def _cmp_state(tag, fx_state, got):
    for k in STATE_KEYS:
        exp = np.asarray(fx_state[k]).astype(np.int64)
        g = np.asarray(got[k]).astype(np.int64)
        if k == "rng":
            exp, g = exp & 0xFFFFFFFF, g & 0xFFFFFFFF
        assert exp.shape == g.shape or exp.size == g.size == 0, f"{tag}: {k} shape {exp.shape} vs {g.shape}"
        assert np.array_equal(exp.reshape(-1), g.reshape(-1)), f"{tag}: field {k} differs\n exp={exp}\n got={g}"


def replay(driver, fx, max_steps=None):
    """Returns the number of steps replayed.  Raises AssertionError at the first difference."""
    T = len(fx["actions"]) if max_steps is None else min(max_steps, len(fx["actions"]))
    obs = driver.reset(int(fx["seed"]))
    assert np.array_equal(obs, fx["obs0"]), "reset observation differs"
    _cmp_state("reset", {k: fx[f"st0_{k}"] for k in STATE_KEYS}, driver.get_state())
    inj = {int(s): i for i, s in enumerate(fx["inject_steps"])}
    keep = set(int(t) for t in fx["obs_keep_steps"])
    resets = {int(t): int(s) for t, s in fx["resets"]}
    for t in range(T):
        if t in inj:
            i = inj[t]
            driver.set_state({k: fx[f"inj{i}_{k}"] for k in STATE_KEYS if k != "rng"})
        obs, rew, done = driver.step(fx["actions"][t].astype(np.int32))
        sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(obs, np.float32).tobytes()).digest(), np.uint8)
        if t in keep:
            assert np.array_equal(obs, fx[f"obs_{t}"]), f"step {t}: observation differs"
        assert np.array_equal(sha, fx["obs_sha"][t]), f"step {t}: observation hash differs"
        assert np.array_equal(np.asarray(rew, np.float64), fx["rewards"][t]), f"step {t}: rewards {rew} vs {fx['rewards'][t]}"
        assert bool(done) == bool(fx["dones"][t]), f"step {t}: done differs"
        _cmp_state(f"step {t}", {k: fx[f"st_{k}"][t] for k in STATE_KEYS}, driver.get_state())
        if done:
            s = resets[t]
            obs = driver.reset(None if s < 0 else s)
            i = inj[-(t + 1)]
            assert np.array_equal(obs, fx[f"obs_{-(t + 1)}"]), f"reset after step {t}: observation differs"
            _cmp_state(f"reset after step {t}", {k: fx[f"inj{i}_{k}"] for k in STATE_KEYS}, driver.get_state())
    return T
