"""The reference's behavioural tests (tests/test_movement.py, tests/test_goals.py, parts of tests/test_env.py of
semitable/robotic-warehouse), restated as data tables and driven through the drop-in `Warehouse` facade with the
reference's own fixture idiom: poke `env.agents[i].x/.y/.dir/.carrying_shelf`, `env.shelfs[j].x/.y`,
`env.request_queue[k]`, call `env._recalc_grid()`, step with `Action` members, read the attributes back.

Runs on the host twin everywhere and on the CUDA C-ABI path on GPU boxes.
"""
import numpy as np
import pytest

from conftest import backend_factory, backend_params
from rware_b200.warehouse import Action, Direction, RewardType, Warehouse

U, D, L, R = Direction.UP, Direction.DOWN, Direction.LEFT, Direction.RIGHT
F, N, TL, TR, T = Action.FORWARD, Action.NOOP, Action.LEFT, Action.RIGHT, Action.TOGGLE_LOAD


def make_env(n_agents, backend, reward=RewardType.GLOBAL, **kw):
    # the 10 x 29 grid every reference movement test uses: Warehouse(3, 8, 3, n, 0, 1, 5, None, None, GLOBAL)
    args = dict(shelf_columns=3, column_height=8, shelf_rows=3, n_agents=n_agents, msg_bits=0, sensor_range=1,
                request_queue_size=5, max_inactivity_steps=None, max_steps=None, reward_type=reward)
    args.update(kw)
    env = Warehouse(_backend_factory=backend_factory(backend), **args)
    env.reset(seed=11)
    return env


def place(env, spec):
    """spec: list of (x, y, dir, carries) per agent; carries=True puts shelf i under agent i and loads it."""
    for i, (x, y, d, carries) in enumerate(spec):
        a = env.agents[i]
        a.x, a.y, a.dir = x, y, d
        if carries:
            env.shelfs[i].x, env.shelfs[i].y = x, y
            a.carrying_shelf = env.shelfs[i]
    env._recalc_grid()


# name, agents [(x, y, dir, carries)], actions, expected agent (x, y) after one step  (reference tests/test_movement.py)
ONE_STEP = [
    ("move_down", [(4, 25, D, 0)], [F], [(4, 26)]),                                   # :50-60
    ("move_up", [(4, 25, U, 0)], [F], [(4, 24)]),                                     # :63-73
    ("move_left", [(4, 25, L, 0)], [F], [(3, 25)]),                                   # :76-86
    ("move_right", [(4, 25, R, 0)], [F], [(5, 25)]),                                  # :89-99
    ("wall_up", [(4, 0, U, 0)], [F], [(4, 0)]),                                       # :116-126
    ("wall_down", [(4, 28, D, 0)], [F], [(4, 28)]),                                   # :129-139
    ("wall_left", [(0, 25, L, 0)], [F], [(0, 25)]),                                   # :142-152
    ("wall_right", [(9, 25, R, 0)], [F], [(9, 25)]),                                  # :155-165
    ("head_on", [(4, 25, R, 0), (5, 25, L, 0)], [F, F], [(4, 25), (5, 25)]),          # :168-185
    ("head_on_one_loaded", [(4, 25, R, 1), (5, 25, L, 0)], [F, F], [(4, 25), (5, 25)]),   # :188-206
    ("head_on_both_loaded", [(4, 25, R, 1), (5, 25, L, 1)], [F, F], [(4, 25), (5, 25)]),  # :209-228
    ("head_on_loaded_left", [(3, 25, L, 1), (2, 25, R, 0)], [F, F], [(3, 25), (2, 25)]),  # :231-245
    ("chain_follows", [(3, 25, R, 0), (4, 25, R, 0)], [F, F], [(4, 25), (5, 25)]),        # :248-264
    ("chain_blocked_by_wall", [(8, 25, R, 0), (9, 25, R, 0)], [F, F], [(8, 25), (9, 25)]),  # :267-283
    ("longest_chain_wins", [(3, 25, R, 0), (4, 25, R, 0), (5, 26, U, 0)], [F, F, F], [(4, 25), (5, 25), (5, 26)]),  # :286-309
    ("four_cycle_rotates", [(3, 25, R, 0), (4, 25, U, 0), (4, 24, L, 0), (3, 24, D, 0)], [F] * 4,
     [(4, 25), (4, 24), (3, 24), (3, 25)]),                                                                         # :312-343
    ("cycle_and_outsider", [(3, 25, R, 0), (4, 25, U, 0), (4, 24, L, 0), (3, 24, D, 0), (5, 24, L, 0)], [F] * 5,
     [(4, 25), (4, 24), (3, 24), (3, 25), (5, 24)]),                                                                # :346-386
    ("carry_moves_shelf", [(4, 25, D, 1)], [F], [(4, 26)]),                           # :501-516
    ("carrier_blocked_by_shelf", [(3, 25, L, 1)], [F], [(3, 25)]),                    # :519-534
    ("two_carriers_chain", [(3, 25, R, 1), (4, 25, R, 1)], [F, F], [(4, 25), (5, 25)]),   # :537-562
]


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("case", ONE_STEP, ids=[c[0] for c in ONE_STEP])
def test_one_step_movement(case, backend):
    name, agents, actions, expected = case
    env = make_env(len(agents), backend)
    place(env, agents)
    env.step(actions)
    for i, (x, y) in enumerate(expected):
        assert (env.agents[i].x, env.agents[i].y) == (x, y), (name, i)
        if agents[i][3]:  # a carried shelf travels with (or stays with) its carrier
            assert (env.shelfs[i].x, env.shelfs[i].y) == (x, y)


@pytest.mark.parametrize("backend", backend_params())
def test_moving_under_shelves_when_unloaded(backend):  # tests/test_movement.py:102-113
    env = make_env(1, backend)
    place(env, [(0, 25, R, 0)])
    for i in range(10):
        env.step([F])
        assert (env.agents[0].x, env.agents[0].y) == (min(i + 1, 9), 25)


TURNS = [(U, TR, R), (R, TR, D), (D, TR, L), (L, TR, U), (U, TL, L), (R, TL, U), (D, TL, R), (L, TL, D)]  # :389-498


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("start,action,end", TURNS)
def test_turns(start, action, end, backend):
    env = make_env(1, backend)
    place(env, [(4, 25, start, 0)])
    env.step([action])
    assert (env.agents[0].x, env.agents[0].y, env.agents[0].dir) == (4, 25, end)


@pytest.mark.parametrize("backend", backend_params())
def test_pickup_carry_and_unload_rules(backend):  # tests/test_movement.py:565-620
    env = make_env(1, backend)
    place(env, [(3, 25, L, 0)])
    env.step([F])
    env.step([T])
    shelf = env.agents[0].carrying_shelf
    assert shelf is not None and (shelf.x, shelf.y) == (2, 25)
    env.step([TL])
    env.step([TL])
    env.step([F])
    assert (env.agents[0].x, env.agents[0].y, shelf.x, shelf.y) == (3, 25, 3, 25)
    env.step([F])
    assert (env.agents[0].x, shelf.x) == (4, 4)
    env.step([T])  # cannot unload on a highway
    assert env.agents[0].carrying_shelf is shelf
    env.step([F])
    assert (env.agents[0].x, env.agents[0].y, shelf.x, shelf.y) == (5, 25, 5, 25)
    env.step([T])  # (5, 25) is part of the queue corridor in front of the goals: still a highway
    assert env.agents[0].carrying_shelf is shelf
    for act in (TL, TL, F, F, F):  # back to the shelf's empty home cell (2, 25)
        env.step([act])
    assert (env.agents[0].x, shelf.x, shelf.y) == (2, 2, 25)
    env.step([T])  # a shelf cell: unloading is allowed
    assert env.agents[0].carrying_shelf is None
    env.step([F])  # unloaded agents walk under standing shelves
    assert (env.agents[0].x, shelf.x) == (1, 2)


def goal_env(backend, n_agents, reward):  # tests/test_goals.py:14-96
    env = make_env(n_agents, backend, reward)
    env.agents[0].x, env.agents[0].y, env.agents[0].dir = 4, 27, D
    env.shelfs[0].x, env.shelfs[0].y = 4, 27
    env.agents[0].carrying_shelf = env.shelfs[0]
    if n_agents > 1:
        env.agents[1].x, env.agents[1].y = 3, 3
    env.request_queue[0] = env.shelfs[0]
    env._recalc_grid()
    return env


@pytest.mark.parametrize("backend", backend_params())
def test_goal_cells_and_grid_size(backend):  # tests/test_goals.py:99-101, tests/test_env.py:42-68
    env = make_env(1, backend)
    assert env.goals == [(4, 28), (5, 28)] and env.grid_size == (29, 10)
    for cols, height, rows in [(3, 1, 1), (3, 8, 1), (5, 8, 2), (5, 3, 3), (9, 15, 4)]:
        e = Warehouse(cols, height, rows, 1, 0, 1, 1, None, None, RewardType.GLOBAL, _backend_factory=backend_factory(backend))
        assert e.grid_size == ((height + 1) * rows + 2, 3 * cols + 1)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("reward,n_agents,expected", [
    (RewardType.GLOBAL, 1, [1.0]), (RewardType.GLOBAL, 2, [1.0, 1.0]), (RewardType.INDIVIDUAL, 2, [1.0, 0.0]),
    (RewardType.TWO_STAGE, 2, [0.5, 0.0])])
def test_delivery_rewards(reward, n_agents, expected, backend):  # tests/test_goals.py:104-137, 156-168
    env = goal_env(backend, n_agents, reward)
    assert env.request_queue[0] is env.shelfs[0]
    _, rewards, _, _, _ = env.step([F] + [N] * (n_agents - 1))
    assert (env.agents[0].x, env.agents[0].y) == (4, 28)
    assert env.request_queue[0] is not env.shelfs[0]
    assert env.shelfs[0] not in env.request_queue
    assert [float(r) for r in rewards] == expected and all(isinstance(r, np.float64) for r in rewards)


@pytest.mark.parametrize("backend", backend_params())
def test_walking_away_from_the_goal_gives_nothing(backend):  # tests/test_goals.py:140-153
    env = goal_env(backend, 1, RewardType.GLOBAL)
    for act in (TL, TL):
        assert env.step([act])[1][0] == 0.0
    _, rewards, _, _, _ = env.step([F])
    assert (env.agents[0].x, env.agents[0].y) == (4, 26) and rewards[0] == 0.0
    assert env.request_queue[0] is env.shelfs[0]


@pytest.mark.parametrize("backend", backend_params())
def test_two_stage_second_half_on_unload(backend):  # tests/test_goals.py:156-184
    env = goal_env(backend, 2, RewardType.TWO_STAGE)
    assert [float(r) for r in env.step([F, N])[1]] == [0.5, 0.0]
    env.agents[0].x, env.agents[0].y = 1, 1
    env.shelfs[0].x, env.shelfs[0].y = 1, 1
    env._recalc_grid()
    assert [float(r) for r in env.step([T, N])[1]] == [0.5, 0.0]
    assert env.agents[0].carrying_shelf is None and env.agents[0].has_delivered is False
    assert [float(r) for r in env.step([T, N])[1]] == [0.0, 0.0]  # re-loading earns nothing
    assert [float(r) for r in env.step([T, N])[1]] == [0.0, 0.0]


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("limit", [1, 100, 200])
def test_time_limit(limit, backend):  # tests/test_env.py:367-388
    env = make_env(2, backend, max_steps=limit)
    for _ in range(limit - 1):
        assert env.step([N, N])[2] is False
    _, _, done, truncated, info = env.step([N, N])
    assert done is True and truncated is False and info == {}


@pytest.mark.parametrize("backend", backend_params())
def test_inactivity_limit_resets_on_delivery(backend):  # tests/test_env.py:336-364, 391-403
    env = goal_env(backend, 1, RewardType.GLOBAL)
    env.max_inactivity_steps = 3  # (constructor value is what the device uses; rebuild with it)
    env = Warehouse(3, 8, 3, 1, 0, 1, 5, 3, None, RewardType.GLOBAL, _backend_factory=backend_factory(backend))
    env.reset(seed=5)
    env.agents[0].x, env.agents[0].y, env.agents[0].dir = 4, 26, D
    env.shelfs[0].x, env.shelfs[0].y = 4, 26
    env.agents[0].carrying_shelf = env.shelfs[0]
    env.request_queue[0] = env.shelfs[0]
    env._recalc_grid()
    assert env.step([F])[2] is False            # inactive 1
    assert env.step([F])[2] is False            # delivers at (4, 28): counter back to 0
    assert env._cur_inactive_steps == 0
    assert env.step([N])[2] is False
    assert env.step([N])[2] is False
    assert env.step([N])[2] is True             # 3 steps without a delivery


@pytest.mark.parametrize("backend", backend_params())
def test_action_and_observation_spaces(backend):  # tests/test_env.py:71-146, 283-333
    from rware_b200 import spaces
    env = make_env(2, backend)
    assert env.action_space == spaces.Tuple(2 * (spaces.Discrete(len(Action)),))
    env = make_env(2, backend, msg_bits=2)
    assert env.action_space == spaces.Tuple(2 * (spaces.MultiDiscrete([len(Action), 2, 2]),))
    env = make_env(3, backend, sensor_range=2)
    obs, info = env.reset(seed=3)
    assert len(obs) == 3 and all(o.shape == (8 + 25 * 7,) and o.dtype == np.float32 for o in obs) and info == {}
    for _ in range(30):
        obs, rew, done, trunc, info = env.step(env.action_space.sample())
        assert env.observation_space.contains(obs) and len(rew) == 3


@pytest.mark.parametrize("backend", backend_params())
def test_seed_reproducibility(backend):  # tests/test_env.py:515-600
    def rollout(seed):
        env = make_env(3, backend)
        env.seed(seed)
        out = []
        for _ in range(3):
            obs, _ = env.reset()
            out.append((np.stack(obs).copy(), env.grid.copy(), [s.id for s in env.request_queue],
                        [(a.x, a.y, a.dir) for a in env.agents]))
        return out
    a, b, c = rollout(42), rollout(42), rollout(43)
    for x, y in zip(a, b):
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) and x[2] == y[2] and x[3] == y[3]
    assert any(not np.array_equal(x[0], y[0]) for x, y in zip(a, c))


@pytest.mark.parametrize("backend", backend_params())
def test_bad_inputs_raise_like_the_reference(backend):  # rware/warehouse.py:807, 811-814
    env = make_env(2, backend)
    with pytest.raises(AssertionError):
        env.step([F])
    with pytest.raises(ValueError):
        env.step([7, 0])


@pytest.mark.parametrize("backend", backend_params())
def test_registered_ids(backend):  # rware/__init__.py:16-39 (+ the -v1 spellings of BASELINE.json)
    import rware_b200
    assert len([k for k in rware_b200.registry if k.endswith("-v2")]) == 228
    kw = rware_b200.env_kwargs("rware-medium-8ag-hard-v1")
    assert (kw["shelf_rows"], kw["shelf_columns"], kw["n_agents"], kw["request_queue_size"]) == (2, 5, 8, 4)
    assert kw == rware_b200.env_kwargs("rware-medium-8ag-hard-v2")
    env = Warehouse(_backend_factory=backend_factory(backend), **rware_b200.env_kwargs("rware-tiny-2ag-v2", sensor_range=3))
    obs, _ = env.reset(seed=0)
    assert obs[0].shape == (8 + 49 * 7,)
