"""Helpers to drive the UNMODIFIED reference (/root/reference/rware) in the build container.

TEST INFRASTRUCTURE, build container only: /root/reference does not exist on the GPU box.  Used by
diff_reference.py (oracle pinning) and tests/golden/make_golden.py (fixture generation).
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("RWARE_REFERENCE_ROOT", "/root/reference")


def load_reference():
    """Import the reference `rware` package through the gymnasium stand-in.  Returns (gymnasium, rware.warehouse)."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "rware")):
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for p in (os.path.join(_HERE, "gymnasium_standin"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gymnasium
    import rware  # noqa: F401  (registers the ids)
    import rware.warehouse as wh
    assert os.path.realpath(wh.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return gymnasium, wh


def rng_words(gen):
    """numpy Generator(PCG64) state -> the 10 uint32 words of oracle/np_random.h:np_export."""
    st = gen.bit_generator.state
    s, inc = st["state"]["state"], st["state"]["inc"]
    w = [(s >> (32 * i)) & 0xFFFFFFFF for i in range(4)] + [(inc >> (32 * i)) & 0xFFFFFFFF for i in range(4)]
    w += [int(st["has_uint32"]), int(st["uinteger"])]
    return np.array(w, np.uint32)


def ref_state(env):
    """Reference env -> plain arrays in the oracle's conventions (ids 1-based, carry 0 = none)."""
    n = env.n_agents
    st = dict(
        agent_x=np.array([int(a.x) for a in env.agents], np.int32),
        agent_y=np.array([int(a.y) for a in env.agents], np.int32),
        agent_dir=np.array([a.dir.value for a in env.agents], np.int32),
        agent_carry=np.array([a.carrying_shelf.id if a.carrying_shelf else 0 for a in env.agents], np.int32),
        agent_deliv=np.array([int(bool(a.has_delivered)) for a in env.agents], np.int32),
        agent_msg=np.array([np.asarray(a.message, np.int32) for a in env.agents], np.int32).reshape(n, env.msg_bits),
        shelf_x=np.array([int(s.x) for s in env.shelfs], np.int32),
        shelf_y=np.array([int(s.y) for s in env.shelfs], np.int32),
        queue=np.array([s.id for s in env.request_queue], np.int32),
        counters=np.array([env._cur_steps, env._cur_inactive_steps], np.int32),
        rng=rng_words(env.np_random),
    )
    return st


def ref_set_state(env, wh, st):
    """Write oracle-convention arrays into a reference env (what the reference's tests do by hand)."""
    for i, a in enumerate(env.agents):
        a.x, a.y = int(st["agent_x"][i]), int(st["agent_y"][i])
        a.dir = wh.Direction(int(st["agent_dir"][i]))
        c = int(st["agent_carry"][i])
        a.carrying_shelf = env.shelfs[c - 1] if c else None
        a.has_delivered = bool(st["agent_deliv"][i])
    for j, s in enumerate(env.shelfs):
        s.x, s.y = int(st["shelf_x"][j]), int(st["shelf_y"][j])
    env.request_queue = [env.shelfs[int(q) - 1] for q in st["queue"]]
    env._cur_steps, env._cur_inactive_steps = int(st["counters"][0]), int(st["counters"][1])
    env._recalc_grid()
