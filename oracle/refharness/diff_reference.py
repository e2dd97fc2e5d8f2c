"""Pin the C oracle against the UNMODIFIED reference, lock-step, in the build container.

    python oracle/refharness/diff_reference.py [--steps N] [--quick]

For every config: reset(seed) on both, then N steps with identical actions; after every step compare the
float32 observations, float64 rewards, done flag, every agent/shelf/queue/counter field and the full
PCG64 generator state.  Every `teleport_every` steps an agent carrying a requested shelf is placed next to a
goal (the same mutation on both sides) so deliveries, resamples and TWO_STAGE rewards actually occur.
Exit code 0 iff there is no mismatch.  (SURVEY.md App. G restated as an executable check.)
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_py  # noqa: E402
import refenv  # noqa: E402


def compare(tag, ref, orc_state, keys=None):
    bad = []
    for k in keys or ref.keys():
        if not np.array_equal(np.asarray(ref[k]), np.asarray(orc_state[k])):
            bad.append(k)
    if bad:
        for k in bad:
            print(f"MISMATCH {tag} field {k}:\n  ref={ref[k]}\n  orc={orc_state[k]}")
    return not bad


def run_config(gym, wh, name, kwargs, steps, p_forward, teleport_every, seed):
    env = wh.Warehouse(**kwargs)
    orc = oracle_py.OracleEnv(env.highways, env.goals, kwargs["n_agents"], kwargs["msg_bits"], kwargs["sensor_range"],
                              kwargs["request_queue_size"], kwargs["max_inactivity_steps"], kwargs["max_steps"],
                              kwargs["reward_type"].value, kwargs.get("normalised_coordinates", False))
    rs = np.random.default_rng(seed + 99)
    obs_r, _ = env.reset(seed=seed)
    obs_o = orc.reset(seed=seed)
    ok = np.array_equal(np.stack(obs_r), obs_o) and compare(f"{name} reset", refenv.ref_state(env), orc.get_state())
    n, mb = kwargs["n_agents"], kwargs["msg_bits"]
    stats = dict(ties=0, set_order_components=0, deliveries=0, resets=0)
    probs = np.full(5, (1 - p_forward) / 4)
    probs[1] = p_forward
    for t in range(steps):
        if not ok:
            break
        if teleport_every and kwargs["request_queue_size"] and t % teleport_every == teleport_every - 1:
            st = refenv.ref_state(env)
            a = int(rs.integers(n))
            gx, gy = env.goals[int(rs.integers(len(env.goals)))]
            if gy > 0 and not any(st["agent_x"][i] == gx and st["agent_y"][i] == gy - 1 for i in range(n) if i != a) \
                    and not any(st["agent_x"][i] == gx and st["agent_y"][i] == gy for i in range(n)):
                old = st["agent_carry"][a]
                sid = int(st["queue"][int(rs.integers(len(st["queue"])))])
                holders = [i for i in range(n) if st["agent_carry"][i] == sid and i != a]
                if not holders:
                    if old and old != sid:  # put the old shelf down where it is; may overlap, harmless for the check
                        pass
                    st["agent_carry"][a] = sid
                    st["agent_x"][a], st["agent_y"][a] = gx, gy - 1
                    st["agent_dir"][a] = 1  # DOWN
                    st["shelf_x"][sid - 1], st["shelf_y"][sid - 1] = gx, gy - 1
                    if old and old != sid:
                        # the previously carried shelf stays behind at its current cell
                        pass
                    refenv.ref_set_state(env, wh, st)
                    orc.set_state(**{k: v for k, v in st.items() if k != "rng"})
                    forced = a
                else:
                    forced = None
            else:
                forced = None
        else:
            forced = None
        acts = rs.choice(5, size=n, p=probs)
        if forced is not None:
            acts[forced] = 1
        if mb:
            full = np.concatenate([acts[:, None], rs.integers(0, 2, size=(n, mb))], axis=1)
            ref_actions = [list(map(int, row)) for row in full]
            orc_actions = full
        else:
            ref_actions = [int(a) for a in acts]
            orc_actions = acts
        obs_r, rew_r, done_r, _, _ = env.step(ref_actions)
        obs_o, rew_o, done_o = orc.step(orc_actions)
        s = orc.last_step_stats()
        for k in ("ties", "set_order_components", "deliveries"):
            stats[k] += s[k]
        ok = (np.array_equal(np.stack(obs_r).astype(np.float32), obs_o)
              and np.array_equal(np.asarray(rew_r, np.float64), rew_o) and bool(done_r) == done_o
              and compare(f"{name} step {t}", refenv.ref_state(env), orc.get_state()))
        if not ok:
            print(f"  first mismatch at step {t}: obs_eq={np.array_equal(np.stack(obs_r).astype(np.float32), obs_o)} "
                  f"rew {rew_r} vs {rew_o} done {done_r} vs {done_o} actions {acts}")
            break
        if done_r:
            if rs.random() < 0.5:
                obs_r, _ = env.reset()
                obs_o = orc.reset()
            else:
                s2 = int(rs.integers(1 << 30))
                obs_r, _ = env.reset(seed=s2)
                obs_o = orc.reset(seed=s2)
            stats["resets"] += 1
            ok = np.array_equal(np.stack(obs_r), obs_o) and compare(f"{name} re-reset", refenv.ref_state(env), orc.get_state())
    return ok, stats


def random_config(rs, RT):
    """Random custom ASCII maps / parametric layouts with random agent count, sensor range, message bits, queue size,
    reward type and limits (same distribution as tests/test_fuzz_configs.py, which checks the device code against the oracle)."""
    if rs.random() < 0.5:
        H, W = int(rs.integers(3, 9)), int(rs.integers(3, 11))
        while True:
            cells = rs.choice(list("x.."), size=(H, W))
            for _ in range(int(rs.integers(1, 4))):
                cells[int(rs.integers(H)), int(rs.integers(W))] = "g"
            if (cells == "x").sum() >= 2:
                break
        kw = dict(layout="\n".join("".join(r) for r in cells), shelf_columns=1, column_height=1, shelf_rows=1)
        n_shelves, n_cells = int((cells == "x").sum()), H * W
    else:
        cols, rows, height = int(rs.choice([1, 3, 5])), int(rs.integers(1, 3)), int(rs.integers(1, 5))
        kw = dict(shelf_columns=cols, shelf_rows=rows, column_height=height)
        _, _, hw, _ = oracle_py.layout_from_params(cols, rows, height)
        if int((hw == 0).sum()) < 2:
            cols = 3
            kw["shelf_columns"] = 3
            _, _, hw, _ = oracle_py.layout_from_params(cols, rows, height)
        n_shelves, n_cells = int((hw == 0).sum()), hw.size
    n = int(rs.integers(1, min(n_cells, 20) + 1))
    rq = int(rs.integers(0, min(n_shelves - 1, 6) + 1))  # rq == n_shelves makes the reference raise on a delivery
    kw.update(n_agents=n, msg_bits=int(rs.choice([0, 0, 1, 3])), sensor_range=int(rs.integers(1, 4)), request_queue_size=rq,
              reward_type=RT(int(rs.integers(0, 3))), max_steps=int(rs.choice([0, 15, 40])) or None,
              max_inactivity_steps=int(rs.choice([0, 0, 10])) or None, normalised_coordinates=bool(rs.random() < 0.3))
    return kw, float(rs.choice([0.3, 0.6, 0.85]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--fuzz", type=int, default=0, help="additionally run this many random configurations")
    ap.add_argument("--fuzz-steps", type=int, default=300)
    ap.add_argument("--fuzz-seed", type=int, default=0)
    args = ap.parse_args()
    gym, wh = refenv.load_reference()
    RT = wh.RewardType

    def kw(cols, height, rows, n, msg, sr, rq, inact, maxs, rt, **extra):
        d = dict(shelf_columns=cols, column_height=height, shelf_rows=rows, n_agents=n, msg_bits=msg, sensor_range=sr,
                 request_queue_size=rq, max_inactivity_steps=inact, max_steps=maxs, reward_type=rt)
        d.update(extra)
        return d
    configs = [
        ("tiny-2ag", kw(3, 8, 1, 2, 0, 1, 2, None, 500, RT.INDIVIDUAL), 0.2, 7),
        ("tiny-4ag", kw(3, 8, 1, 4, 0, 1, 4, None, 500, RT.INDIVIDUAL), 0.2, 7),
        ("small-4ag-global", kw(3, 8, 2, 4, 0, 1, 4, 40, 500, RT.GLOBAL), 0.5, 5),
        ("medium-8ag-hard-twostage", kw(5, 8, 2, 8, 0, 1, 4, 60, 500, RT.TWO_STAGE), 0.68, 5),
        ("large-16ag-r2", kw(5, 8, 3, 16, 0, 2, 16, None, 500, RT.INDIVIDUAL), 0.68, 9),
        ("tiny-19ag-twostage", kw(3, 8, 1, 19, 0, 1, 9, 40, 300, RT.TWO_STAGE), 0.68, 5),
        ("tiny-10ag-msg3-norm", kw(3, 8, 1, 10, 3, 1, 5, None, 200, RT.GLOBAL, normalised_coordinates=True), 0.6, 6),
        ("1x3-h2-8ag-dense", kw(3, 2, 1, 8, 0, 1, 3, None, 200, RT.INDIVIDUAL), 0.75, 4),
        ("custom-layout", dict(kw(1, 1, 1, 5, 1, 2, 3, 30, 150, RT.TWO_STAGE),
                               layout="..x..x..\n.xx..xx.\n........\n.x.gg.x.\n..x..x.."), 0.6, 3),
    ]
    t0 = time.time()
    all_ok = True
    for name, kwargs, pf, tele in configs:
        steps = args.steps // (4 if args.quick else 1)
        ok, stats = run_config(gym, wh, name, kwargs, steps, pf, tele, seed=hash(name) % 1000)
        print(f"{name:28s} steps={steps:6d} {'OK ' if ok else 'FAIL'} {stats}")
        all_ok &= ok
    rs = np.random.default_rng(args.fuzz_seed)
    n_fuzz_ok = 0
    for i in range(args.fuzz):
        kwargs, pf = random_config(rs, RT)
        ok, stats = run_config(gym, wh, f"fuzz{i}", kwargs, args.fuzz_steps, pf, int(rs.choice([0, 3, 6])), seed=int(rs.integers(1 << 20)))
        n_fuzz_ok += ok
        if not ok:
            print(f"fuzz{i} FAIL {kwargs}")
        all_ok &= ok
    if args.fuzz:
        print(f"fuzz: {n_fuzz_ok}/{args.fuzz} random configurations x {args.fuzz_steps} steps OK (seed {args.fuzz_seed})")
    print(f"total {time.time() - t0:.1f}s -> {'ALL OK' if all_ok else 'MISMATCH'}")
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
