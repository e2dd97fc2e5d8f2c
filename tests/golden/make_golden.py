"""Generates tests/golden/*.npz from the UNMODIFIED reference (build container only; /root/reference is not on the
GPU box).  Run:  python tests/golden/make_golden.py

Each fixture is a whole trajectory of one reference world: constructor kwargs, reset seed, the action stream, the
state injections ("teleports" that make deliveries happen under random actions), and per step the complete world
state, rewards, done, the generator state and a SHA-256 of the float32 observations.  The replay harness
(tests/replay.py) drives the oracle, the host twin and the CUDA path through the same trajectory and demands
bit-exact agreement on every field.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refharness"))
import refenv  # noqa: E402

STATE_KEYS = ("agent_x", "agent_y", "agent_dir", "agent_carry", "agent_deliv", "agent_msg", "shelf_x", "shelf_y",
              "queue", "counters", "rng")


def kw(cols, height, rows, n, msg, sr, rq, inact, maxs, rt, **extra):
    d = dict(shelf_columns=cols, column_height=height, shelf_rows=rows, n_agents=n, msg_bits=msg, sensor_range=sr,
             request_queue_size=rq, max_inactivity_steps=inact, max_steps=maxs, reward_type=rt)
    d.update(extra)
    return d


def record(wh, name, kwargs, steps, p_forward, teleport_every, seed, reseed_on_done=True):
    env = wh.Warehouse(**{**kwargs, "reward_type": wh.RewardType(kwargs["reward_type"])})
    rs = np.random.default_rng(seed + 4242)
    n, mb = kwargs["n_agents"], kwargs["msg_bits"]
    obs, _ = env.reset(seed=seed)
    obs0 = np.stack(obs).astype(np.float32)
    st0 = refenv.ref_state(env)
    probs = np.full(5, (1 - p_forward) / 4)
    probs[1] = p_forward
    actions = np.zeros((steps, n, 1 + mb), np.int8)
    states = {k: [] for k in STATE_KEYS}
    rewards, dones, obs_sha, obs_keep = [], [], [], {}
    inject_step, inject_state = [], []
    resets = []  # (step index after which reset happened, seed or -1)
    for t in range(steps):
        if teleport_every and t % teleport_every == teleport_every - 1:
            st = refenv.ref_state(env)
            a = int(rs.integers(n))
            gx, gy = env.goals[int(rs.integers(len(env.goals)))]
            free = (gy > 0 and not any(st["agent_x"][i] == gx and st["agent_y"][i] in (gy - 1, gy) for i in range(n) if i != a)
                    and len(st["queue"]) > 0)
            if free:
                sid = int(st["queue"][int(rs.integers(len(st["queue"])))])
                if not any(st["agent_carry"][i] == sid for i in range(n) if i != a):
                    st["agent_carry"][a] = sid
                    st["agent_x"][a], st["agent_y"][a], st["agent_dir"][a] = gx, gy - 1, 1
                    st["shelf_x"][sid - 1], st["shelf_y"][sid - 1] = gx, gy - 1
                    refenv.ref_set_state(env, wh, st)
                    inject_step.append(t)
                    inject_state.append({k: np.array(v) for k, v in st.items() if k != "rng"})
                    actions[t, a, 0] = -1  # marker: forced FORWARD below
        acts = rs.choice(5, size=n, p=probs).astype(np.int8)
        forced = np.where(actions[t, :, 0] == -1)[0]
        acts[forced] = 1
        actions[t, :, 0] = acts
        if mb:
            actions[t, :, 1:] = rs.integers(0, 2, size=(n, mb))
            ref_actions = [list(map(int, row)) for row in actions[t]]
        else:
            ref_actions = [int(a) for a in acts]
        obs, rew, done, _, _ = env.step(ref_actions)
        o = np.stack(obs).astype(np.float32)
        obs_sha.append(np.frombuffer(hashlib.sha256(o.tobytes()).digest(), np.uint8))
        if t < 8 or t % 97 == 0:
            obs_keep[t] = o
        st = refenv.ref_state(env)
        for k in STATE_KEYS:
            states[k].append(st[k])
        rewards.append(np.asarray(rew, np.float64))
        dones.append(bool(done))
        if done:
            s2 = int(rs.integers(1 << 30)) if (reseed_on_done and rs.random() < 0.5) else -1
            obs, _ = env.reset(seed=None if s2 < 0 else s2)
            resets.append((t, s2))
            st = refenv.ref_state(env)
            # state right after the reset, stored under a separate key family
            inject_step.append(-(t + 1))  # negative: "expected state after reset following step t"
            inject_state.append({k: np.array(v) for k, v in st.items()})
            obs_keep[-(t + 1)] = np.stack(obs).astype(np.float32)
    out = dict(
        kwargs=json.dumps({k: (v if not hasattr(v, "value") else v.value) for k, v in kwargs.items()}),
        seed=seed, actions=actions, obs0=obs0, rewards=np.stack(rewards), dones=np.array(dones),
        obs_sha=np.stack(obs_sha), obs_keep_steps=np.array(sorted(obs_keep)), resets=np.array(resets, np.int64).reshape(-1, 2),
        highways=np.asarray(env.highways, np.uint8), goals=np.asarray(env.goals, np.int32),
        inject_steps=np.array(inject_step, np.int64))
    for t in sorted(obs_keep):
        out[f"obs_{t}"] = obs_keep[t]
    for k in STATE_KEYS:
        out[f"st0_{k}"] = st0[k]
        out[f"st_{k}"] = np.stack(states[k]).astype(np.uint32 if k == "rng" else np.int16)
    for i, stt in enumerate(inject_state):
        for k, v in stt.items():
            out[f"inj{i}_{k}"] = v.astype(np.uint32 if k == "rng" else np.int16)
    path = os.path.join(HERE, f"traj_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {steps} steps, {len(inject_step)} injections/resets, {sum(map(float, np.stack(rewards).sum(1) > 0))} rewarded steps, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def known_answers(gym):
    """SHA-256 over obs0 then per step obs || rewards(f64) || done, registered ids, seed 0."""
    out = {}
    for id_, extra in [("rware-tiny-2ag-v2", {}), ("rware-tiny-4ag-v2", {}), ("rware-small-4ag-v2", {}),
                       ("rware-medium-8ag-hard-v2", {}), ("rware-large-16ag-v2", {"sensor_range": 2})]:
        env = gym.make(id_, **extra)
        obs, _ = env.reset(seed=0)
        h = hashlib.sha256(np.stack(obs).astype(np.float32).tobytes())
        n = env.n_agents
        acts = np.random.default_rng(12345).integers(0, 5, size=(500, n))
        total = 0.0
        for t in range(500):
            obs, rew, done, _, _ = env.step([int(a) for a in acts[t]])
            total += float(np.sum(rew))
            h.update(np.stack(obs).astype(np.float32).tobytes() + np.asarray(rew, np.float64).tobytes() + bytes([done]))
        st = refenv.ref_state(env)
        out[id_ + ("" if not extra else "+sensor_range=2")] = dict(
            sha256=h.hexdigest(), extra=extra, reward_sum=total,
            final_agents=[[int(st["agent_x"][i]), int(st["agent_y"][i]), int(st["agent_dir"][i]), int(st["agent_carry"][i])]
                          for i in range(n)])
        print(id_, h.hexdigest())
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(out, f, indent=1)


def layouts(wh):
    """Highway bitmaps / goals / grid sizes of the registered sizes plus a few odd ones (warehouse.py:294-350)."""
    out = {}
    for cols, rows, height in [(3, 1, 8), (3, 2, 8), (5, 2, 8), (5, 3, 8), (1, 1, 1), (7, 4, 3), (9, 1, 15), (3, 3, 8)]:
        env = wh.Warehouse(cols, height, rows, 1, 0, 1, 1, None, None, wh.RewardType.GLOBAL)
        out[f"{cols}_{rows}_{height}_highways"] = np.asarray(env.highways, np.uint8)
        out[f"{cols}_{rows}_{height}_goals"] = np.asarray(env.goals, np.int32)
    np.savez_compressed(os.path.join(HERE, "layouts.npz"), **out)


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, np.generic):
        return o.item()
    return o


def record_obs_types(wh):
    """Non-default observation types of the reference (warehouse.py:352-427, 527-596, 676-742) on short trajectories."""
    OT, IL = wh.ObservationType, wh.ImageLayer
    cases = {
        "image_default": dict(n=4, msg=0, r=1, ot=OT.IMAGE, extra={}),
        "image_nd_subset": dict(n=3, msg=0, r=2, ot=OT.IMAGE, extra=dict(image_observation_directional=False,
                                image_observation_layers=[IL.GOALS, IL.SHELVES, IL.ACCESSIBLE])),
        "imagedict_r2": dict(n=5, msg=0, r=2, ot=OT.IMAGE_DICT, extra={}),
        "dict_msg2": dict(n=3, msg=2, r=1, ot=OT.DICT, extra={}),
        "dict_norm": dict(n=2, msg=0, r=1, ot=OT.DICT, extra=dict(normalised_coordinates=True)),
    }
    out = {}
    for name, cs in cases.items():
        kwargs = dict(shelf_columns=3, column_height=8, shelf_rows=1, n_agents=cs["n"], msg_bits=cs["msg"],
                      sensor_range=cs["r"], request_queue_size=cs["n"], max_inactivity_steps=None, max_steps=50,
                      reward_type=wh.RewardType.INDIVIDUAL, observation_type=cs["ot"], **cs["extra"])
        env = wh.Warehouse(**kwargs)
        rs = np.random.default_rng(len(name))
        T = 80
        acts = np.concatenate([rs.choice(5, size=(T, cs["n"], 1), p=[.1, .5, .15, .15, .1]),
                               rs.integers(0, 2, size=(T, cs["n"], cs["msg"]))], axis=2).astype(np.int8)
        snap = (lambda o: _jsonable(o)) if cs["ot"] == OT.DICT else (lambda o: o)  # DICT obs alias agent.message: copy now
        obs, _ = env.reset(seed=77)
        seq = [snap(obs)]
        for t in range(T):
            a = [list(map(int, r)) for r in acts[t]] if cs["msg"] else [int(r[0]) for r in acts[t]]
            obs, rew, done, _, _ = env.step(a)
            seq.append(snap(obs))
            if done:
                obs, _ = env.reset()
                seq[-1] = snap(obs)  # the observation the caller continues from
        kw_json = {k: (v.value if hasattr(v, "value") else ([x.value for x in v] if isinstance(v, list) else v))
                   for k, v in kwargs.items()}
        rec = dict(kwargs=kw_json, seed=77, actions=acts.tolist())
        if cs["ot"] == OT.DICT:
            rec["obs"] = _jsonable(seq)
        elif cs["ot"] == OT.IMAGE:
            rec["image"] = np.stack([np.stack(o) for o in seq]).astype(np.float32).tolist()
        else:
            rec["image"] = np.stack([np.stack([a["image"] for a in o]) for o in seq]).astype(np.float32).tolist()
            rec["features"] = np.stack([np.stack([a["features"] for a in o]) for o in seq]).astype(np.float32).tolist()
        out[name] = rec
        print("obs-type fixture", name, len(seq), "observations")
    with open(os.path.join(HERE, "obs_types.json"), "w") as f:
        json.dump(out, f)


def record_global_images(wh):
    """get_global_image (warehouse.py:966-1040) of the unmodified reference along short trajectories: every layer, the padding
    rule, and the IndexError of the transposed AGENT_DIRECTION / AGENT_LOAD layers on a non-square grid."""
    IL = wh.ImageLayer
    cases = {
        "default_tiny": dict(kw=dict(shelf_columns=3, column_height=8, shelf_rows=1), n=4, layers=[IL.SHELVES, IL.GOALS], pad=None),
        "all_plain_tiny_padded": dict(kw=dict(shelf_columns=3, column_height=8, shelf_rows=1), n=5,
                                      layers=[IL.ACCESSIBLE, IL.AGENTS, IL.REQUESTS, IL.SHELVES, IL.GOALS], pad=(7, 14, 13)),
        "all_layers_square": dict(kw=dict(shelf_columns=1, column_height=1, shelf_rows=1,
                                          layout="..x..x.\n.xx.xx.\n.......\nx..g..x\n.x...x.\n..xgx..\n......."), n=6,
                                  layers=[IL.SHELVES, IL.REQUESTS, IL.AGENTS, IL.AGENT_DIRECTION, IL.AGENT_LOAD, IL.GOALS, IL.ACCESSIBLE],
                                  pad=(7, 8, 10)),
        "transposed_tiny": dict(kw=dict(shelf_columns=3, column_height=8, shelf_rows=1), n=6,
                                layers=[IL.AGENT_DIRECTION, IL.AGENT_LOAD, IL.AGENTS], pad=None),
    }
    out, meta = {}, {}
    for name, cs in cases.items():
        n = cs["n"]
        kwargs = dict(n_agents=n, msg_bits=0, sensor_range=1, request_queue_size=3, max_inactivity_steps=None, max_steps=60,
                      reward_type=wh.RewardType.INDIVIDUAL, **cs["kw"])
        env = wh.Warehouse(**kwargs)
        rs = np.random.default_rng(len(name))
        T = 120
        acts = rs.choice(5, size=(T, n), p=[.05, .5, .1, .1, .25]).astype(np.int8)
        imgs, raised = [], []

        def snap():
            try:
                imgs.append(env.get_global_image(image_layers=cs["layers"], recompute=True, pad_to_shape=cs["pad"]).astype(np.float32))
                raised.append(0)
            except IndexError:
                imgs.append(np.zeros_like(imgs[-1]) if imgs else None)
                raised.append(1)
        env.reset(seed=5)
        snap()
        for t in range(T):
            _, _, done, _, _ = env.step([int(a) for a in acts[t]])
            if done:
                env.reset()
            snap()
        first = next(i for i in imgs if i is not None)
        imgs = [np.zeros_like(first) if i is None else i for i in imgs]
        out[name + "_images"] = np.stack(imgs)
        out[name + "_raised"] = np.asarray(raised, np.uint8)
        out[name + "_actions"] = acts
        meta[name] = dict(kwargs={k: (v.value if hasattr(v, "value") else v) for k, v in kwargs.items()}, seed=5,
                          layers=[l.value for l in cs["layers"]], pad=cs["pad"])
        print("global-image fixture", name, out[name + "_images"].shape, "raised", int(sum(raised)))
    np.savez_compressed(os.path.join(HERE, "global_image.npz"), **out)
    with open(os.path.join(HERE, "global_image.json"), "w") as f:
        json.dump(meta, f)


def record_image_transposed(wh):
    """IMAGE observations with the AGENT_DIRECTION / AGENT_LOAD layers (warehouse.py:548-558: `layer[ag.x, ag.y]`, sic): a
    square map where the transposed writes always land inside the layer, and rware-tiny (11 x 10) where numpy raises IndexError
    as soon as an agent reaches the bottom row."""
    OT, IL = wh.ObservationType, wh.ImageLayer
    cases = {
        "square_directional": dict(kw=dict(shelf_columns=1, column_height=1, shelf_rows=1,
                                           layout="..x..x.\n.xx.xx.\n.......\nx..g..x\n.x...x.\n..xgx..\n......."), n=6, r=1,
                                   directional=True, layers=[IL.AGENT_DIRECTION, IL.SHELVES, IL.AGENT_LOAD, IL.AGENTS]),
        "square_r2_plain": dict(kw=dict(shelf_columns=1, column_height=1, shelf_rows=1,
                                        layout="..x..x.\n.xx.xx.\n.......\nx..g..x\n.x...x.\n..xgx..\n......."), n=4, r=2,
                                directional=False, layers=[IL.AGENT_LOAD, IL.AGENT_DIRECTION]),
        "tiny_raises": dict(kw=dict(shelf_columns=3, column_height=8, shelf_rows=1), n=5, r=1, directional=True,
                            layers=[IL.AGENTS, IL.AGENT_DIRECTION, IL.AGENT_LOAD]),
    }
    out, meta = {}, {}
    for name, cs in cases.items():
        n = cs["n"]
        kwargs = dict(n_agents=n, msg_bits=0, sensor_range=cs["r"], request_queue_size=3, max_inactivity_steps=None, max_steps=70,
                      reward_type=wh.RewardType.INDIVIDUAL, observation_type=OT.IMAGE, image_observation_layers=cs["layers"],
                      image_observation_directional=cs["directional"], **cs["kw"])
        env = wh.Warehouse(**kwargs)
        rs = np.random.default_rng(40 + len(name))
        T = 140
        acts = rs.choice(5, size=(T, n), p=[.05, .5, .1, .1, .25]).astype(np.int8)
        K = 2 * cs["r"] + 1
        imgs, raised = [], []

        def keep(fn):
            try:
                obs = fn()
                imgs.append(np.stack(obs).astype(np.float32))
                raised.append(0)
            except IndexError:
                imgs.append(np.zeros((n, len(cs["layers"]), K, K), np.float32))
                raised.append(1)
        keep(lambda: env.reset(seed=9)[0])
        for t in range(T):
            done_box = []

            def one():
                o, _, d, _, _ = env.step([int(a) for a in acts[t]])
                done_box.append(d)
                return o
            keep(one)
            # a raising step has still advanced the state (the exception comes from the observation code at the end of step)
            if (done_box and done_box[0]) or (not done_box and env._cur_steps >= env.max_steps):
                keep(lambda: env.reset()[0])
                imgs.pop(-2)
                raised.pop(-2)
        out[name + "_images"] = np.stack(imgs)
        out[name + "_raised"] = np.asarray(raised, np.uint8)
        out[name + "_actions"] = acts
        meta[name] = dict(kwargs={k: (v.value if hasattr(v, "value") else ([x.value for x in v] if isinstance(v, list) else v))
                                  for k, v in kwargs.items()}, seed=9)
        print("transposed-image fixture", name, out[name + "_images"].shape, "raised", int(sum(raised)))
    np.savez_compressed(os.path.join(HERE, "image_transposed.npz"), **out)
    with open(os.path.join(HERE, "image_transposed.json"), "w") as f:
        json.dump(meta, f)


def record_flatten_sa(wh):
    """rware.utils.wrappers.FlattenSAObservation (wrappers.py:83-108) over DICT and IMAGE_DICT observations.  Note: the flatten
    rule itself is gymnasium's; in this container it is the stand-in's restatement of it (oracle/refharness/gymnasium_standin)."""
    from rware.utils import wrappers as W
    OT = wh.ObservationType
    out = {}
    for name, kw in (("dict_msg2", dict(n_agents=3, msg_bits=2, observation_type=OT.DICT)),
                     ("imagedict", dict(n_agents=2, msg_bits=0, observation_type=OT.IMAGE_DICT))):
        env = W.FlattenSAObservation(wh.Warehouse(shelf_columns=3, column_height=8, shelf_rows=1, sensor_range=1, request_queue_size=2,
                                                  max_inactivity_steps=None, max_steps=30, reward_type=wh.RewardType.INDIVIDUAL, **kw))
        rs = np.random.default_rng(len(name))
        n, mb = kw["n_agents"], kw["msg_bits"]
        obs, _ = env.reset(seed=21)
        seq = [[np.asarray(o, np.float64).tolist() for o in obs]]
        acts = []
        for t in range(25):
            a = [[int(rs.integers(5))] + [int(b) for b in rs.integers(0, 2, size=mb)] for _ in range(n)] if mb else \
                [int(rs.integers(5)) for _ in range(n)]
            acts.append(a)
            obs, _, done, _, _ = env.step(a)
            seq.append([np.asarray(o, np.float64).tolist() for o in obs])
        out[name] = dict(kwargs={k: (v.value if hasattr(v, "value") else v) for k, v in kw.items()}, actions=acts, obs=seq,
                         flatdim=[int(s.shape[0]) for s in env.observation_space])
        print("flatten-sa fixture", name, len(seq), out[name]["flatdim"])
    with open(os.path.join(HERE, "flatten_sa.json"), "w") as f:
        json.dump(out, f)


def record_registry_ids(gym):
    """id -> kwargs of the reference's optional registrations (rware/__init__.py:42-172): every image id and a random sample
    of the 5.25 million full_registration ids (enumerating them takes ~30 s)."""
    import rware
    base = set(gym.registry)
    rware.image_registration()
    img_ids = sorted(set(gym.registry) - base)
    rware.full_registration()
    full_ids = sorted(set(gym.registry) - base - set(img_ids))
    rs = np.random.default_rng(0)
    sample = [full_ids[i] for i in rs.choice(len(full_ids), 1500, replace=False)]
    cols = ["column_height", "shelf_rows", "shelf_columns", "n_agents", "msg_bits", "sensor_range", "request_queue_size",
            "max_inactivity_steps", "max_steps", "reward_type", "observation_type", "image_observation_directional"]

    def row(kw):
        return [(kw[c].value if hasattr(kw[c], "value") else kw[c]) for c in cols]
    out = dict(n_image=len(img_ids), n_full=len(full_ids), columns=cols, image={i: row(gym.registry[i][1]) for i in img_ids},
               full_sample={i: row(gym.registry[i][1]) for i in sample})
    with open(os.path.join(HERE, "registry_ids.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("registry fixture:", len(img_ids), "image ids,", len(sample), "of", len(full_ids), "full ids")


def main():
    gym, wh = refenv.load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "image_transposed":
        record_image_transposed(wh)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "flatten_sa":
        record_flatten_sa(wh)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "registry":
        record_registry_ids(gym)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "obs_types":
        record_obs_types(wh)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "global_image":
        record_global_images(wh)
        return
    G, I, T = 0, 1, 2
    record(wh, "tiny2_indiv", kw(3, 8, 1, 2, 0, 1, 2, None, 500, I), 700, 0.3, 9, seed=0)
    record(wh, "tiny4_indiv", kw(3, 8, 1, 4, 0, 1, 4, None, 500, I), 1100, 0.3, 7, seed=1)
    record(wh, "small4_global", kw(3, 8, 2, 4, 0, 1, 4, 40, 500, G), 700, 0.5, 5, seed=2)
    record(wh, "medium8_hard_twostage", kw(5, 8, 2, 8, 0, 1, 4, 60, 500, T), 800, 0.68, 5, seed=3)
    record(wh, "large16_r2", kw(5, 8, 3, 16, 0, 2, 16, None, 500, I), 700, 0.68, 9, seed=4)
    record(wh, "tiny19_twostage", kw(3, 8, 1, 19, 0, 1, 9, 40, 300, T), 900, 0.68, 5, seed=5)
    record(wh, "tiny10_msg3_norm", kw(3, 8, 1, 10, 3, 1, 5, None, 200, G, normalised_coordinates=True), 500, 0.6, 6, seed=6)
    record(wh, "dense8_h2", kw(3, 2, 1, 8, 0, 1, 3, None, 200, I), 700, 0.75, 4, seed=7)
    record(wh, "custom_layout", dict(kw(1, 1, 1, 5, 1, 2, 3, 30, 150, T),
                                     layout="..x..x..\n.xx..xx.\n........\n.x.gg.x.\n..x..x.."), 500, 0.6, 3, seed=8)
    record(wh, "tiny1_hard_r0", kw(3, 8, 1, 1, 0, 1, 0, None, 100, I), 250, 0.5, 0, seed=9)
    known_answers(gym)
    layouts(wh)
    record_obs_types(wh)
    record_global_images(wh)
    record_image_transposed(wh)
    record_flatten_sa(wh)
    record_registry_ids(gym)


if __name__ == "__main__":
    main()
