"""`Warehouse`: the reference's single-world Gymnasium environment surface on top of the CUDA step path.

Drop-in for `rware.warehouse:Warehouse` (rware/warehouse.py:140-964) for the default FLATTENED observation type:
same constructor signature, `reset(seed, options) -> (obs tuple, info)`, `step(actions) -> (obs tuple, rewards list,
done, truncated, info)`, `seed()`, and the attributes the reference's tests read and *write*
(`env.agents[i].x/.y/.dir/.carrying_shelf`, `env.shelfs[j].x/.y`, `env.request_queue`, `env._recalc_grid()`,
`env.grid`, ...).  It is a batch of one world living on the GPU: before every step the host-side entity objects are
written through to the device (rw_set_state), afterwards read back (rw_get_state).  Use VecWarehouse for throughput.
"""
from enum import Enum
from typing import List, Optional

import numpy as np
import torch

from . import spaces
from .vec import VecWarehouse

_COLLISION_LAYERS = 2
_LAYER_AGENTS = 0
_LAYER_SHELFS = 1


class Action(Enum):  # rware/warehouse.py:31-36
    NOOP = 0
    FORWARD = 1
    LEFT = 2
    RIGHT = 3
    TOGGLE_LOAD = 4


class Direction(Enum):  # rware/warehouse.py:39-43
    UP = 0
    DOWN = 1
    LEFT = 2
    RIGHT = 3


class RewardType(Enum):  # rware/warehouse.py:46-49
    GLOBAL = 0
    INDIVIDUAL = 1
    TWO_STAGE = 2


class ObservationType(Enum):  # rware/warehouse.py:52-56
    DICT = 0
    FLATTENED = 1
    IMAGE = 2
    IMAGE_DICT = 3


class ImageLayer(Enum):  # rware/warehouse.py:59-70
    SHELVES = 0
    REQUESTS = 1
    AGENTS = 2
    AGENT_DIRECTION = 3
    AGENT_LOAD = 4
    GOALS = 5
    ACCESSIBLE = 6


class Entity:  # rware/warehouse.py:73-79
    def __init__(self, id_: int, x: int, y: int):
        self.id = id_
        self.prev_x = None
        self.prev_y = None
        self.x = x
        self.y = y


class Agent(Entity):  # rware/warehouse.py:82-125 (ids are per-environment here, not class-level counters)
    def __init__(self, id_: int, x: int, y: int, dir_: Direction, msg_bits: int):
        super().__init__(id_, x, y)
        self.dir = dir_
        self.message = np.zeros(msg_bits)
        self.req_action: Optional[Action] = None
        self.carrying_shelf: Optional["Shelf"] = None
        self.canceled_action = None
        self.has_delivered = False

    @property
    def loaded(self):
        return self.carrying_shelf is not None

    @property
    def collision_layers(self):
        return (_LAYER_AGENTS, _LAYER_SHELFS) if self.loaded else (_LAYER_AGENTS,)


class Shelf(Entity):  # rware/warehouse.py:128-137
    @property
    def collision_layers(self):
        return (_LAYER_SHELFS,)


class Warehouse:
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}
    spec = None

    def __init__(self, shelf_columns: int, column_height: int, shelf_rows: int, n_agents: int, msg_bits: int,
                 sensor_range: int, request_queue_size: int, max_inactivity_steps: Optional[int],
                 max_steps: Optional[int], reward_type: RewardType, layout: Optional[str] = None,
                 observation_type: ObservationType = ObservationType.FLATTENED, image_observation_layers=None,
                 image_observation_directional: bool = True, normalised_coordinates: bool = False,
                 render_mode: str = "human", device=None, _backend_factory=None):
        if image_observation_layers is None:  # warehouse.py:160-166
            image_observation_layers = [ImageLayer.SHELVES, ImageLayer.REQUESTS, ImageLayer.AGENTS, ImageLayer.GOALS,
                                        ImageLayer.ACCESSIBLE]
        self.observation_type = ObservationType(getattr(observation_type, "value", observation_type))
        self.image_observation_layers = [ImageLayer(getattr(v, "value", v)) for v in image_observation_layers]
        self.image_observation_directional = image_observation_directional
        # DICT observations carry the same information as FLATTENED rows: the device produces the rows, the host
        # unflattens them (warehouse.py:676-720).  IMAGE / IMAGE_DICT rows come from the image gather on the device.
        dev_type = ObservationType.FLATTENED if self.observation_type == ObservationType.DICT else self.observation_type
        self._vec = VecWarehouse(1, shelf_columns, column_height, shelf_rows, n_agents, msg_bits, sensor_range,
                                 request_queue_size, max_inactivity_steps, max_steps, reward_type, layout,
                                 normalised_coordinates, device=device, _backend_factory=_backend_factory,
                                 observation_type=dev_type.value,
                                 image_observation_layers=[v.value for v in self.image_observation_layers],
                                 image_observation_directional=image_observation_directional)
        v = self._vec
        self.grid_size = v.grid_size
        self.highways = v.highways
        self.goals = list(v.goals)
        if not layout:
            self.column_height = column_height
        self.grid = np.zeros((_COLLISION_LAYERS, *self.grid_size), dtype=np.int32)
        self.n_agents = n_agents
        self.msg_bits = msg_bits
        self.sensor_range = sensor_range
        self.max_inactivity_steps = max_inactivity_steps
        self.reward_type = reward_type
        self.reward_range = (0, 1)
        self._cur_inactive_steps = None
        self._cur_steps = 0
        self.max_steps = max_steps
        self.normalised_coordinates = normalised_coordinates
        sa = [len(Action), *msg_bits * (2,)]  # rware/warehouse.py:255-260
        sa_space = spaces.Discrete(sa[0]) if len(sa) == 1 else spaces.MultiDiscrete(sa)
        self.action_space = spaces.Tuple(tuple(n_agents * [sa_space]))
        self.request_queue_size = request_queue_size
        self.request_queue: List[Shelf] = []
        self.agents: List[Agent] = []
        self.shelfs: List[Shelf] = []
        self.fast_obs = self.observation_type == ObservationType.FLATTENED
        self.image_obs = self.observation_type == ObservationType.IMAGE
        self.image_dict_obs = self.observation_type == ObservationType.IMAGE_DICT
        # rware/warehouse.py:432-443, 505-522
        self._obs_bits_for_self = 4 + len(Direction)
        self._obs_bits_per_agent = 1 + len(Direction) + msg_bits
        self._obs_bits_per_shelf = 2
        self._obs_bits_for_requests = 2
        self._obs_sensor_locations = (1 + 2 * sensor_range) ** 2
        self._obs_length = (self._obs_bits_for_self + self._obs_sensor_locations * self._obs_bits_per_agent
                            + self._obs_sensor_locations * self._obs_bits_per_shelf)
        self.observation_space = self._make_observation_space(self.observation_type)
        self.global_image = None
        self.renderer = None
        self.render_mode = render_mode
        self._seeded = False

    @property
    def unwrapped(self):
        return self

    # ------------------------------------------------------------------ host <-> device state
    def _is_highway(self, x: int, y: int) -> bool:
        return self.highways[y, x]

    def _recalc_grid(self):  # rware/warehouse.py:749-755
        self.grid[:] = 0
        for s in self.shelfs:
            self.grid[_LAYER_SHELFS, s.y, s.x] = s.id
        for a in self.agents:
            self.grid[_LAYER_AGENTS, a.y, a.x] = a.id

    def _push(self):
        """host entity objects -> device world (what the reference keeps in Python objects)."""
        n = self.n_agents
        shelf_ids = {id(s): s.id for s in self.shelfs}
        H, W = self.grid_size
        for ent in (*self.agents, *self.shelfs):  # the reference's _recalc_grid (749-755) raises IndexError for these
            if not (0 <= int(ent.x) < W and 0 <= int(ent.y) < H):
                axis, bad = (2, int(ent.x)) if not 0 <= int(ent.x) < W else (1, int(ent.y))
                raise IndexError(f"index {bad} is out of bounds for axis {axis} with size {W if axis == 2 else H}")
        st = dict(
            agent_x=[[int(a.x) for a in self.agents]], agent_y=[[int(a.y) for a in self.agents]],
            agent_dir=[[Direction(a.dir).value for a in self.agents]],
            agent_carry=[[shelf_ids[id(a.carrying_shelf)] if a.carrying_shelf is not None else 0 for a in self.agents]],
            agent_has_delivered=[[int(bool(a.has_delivered)) for a in self.agents]],
            shelf_x=[[int(s.x) for s in self.shelfs]], shelf_y=[[int(s.y) for s in self.shelfs]],
            queue=np.asarray([[shelf_ids[id(s)] for s in self.request_queue]], np.int32).reshape(1, self.request_queue_size),
            cur_steps=[int(self._cur_steps)], cur_inactive=[int(self._cur_inactive_steps or 0)])
        if self.msg_bits:
            st["agent_msg"] = np.asarray([[np.asarray(a.message, np.int32) for a in self.agents]], np.int32).reshape(
                1, n, self.msg_bits)
        self._vec.set_state(**st)

    def _pull(self, create=False):
        """device world -> host entity objects (identity of Agent/Shelf objects is preserved)."""
        st = {k: v[0].cpu().numpy() for k, v in self._vec.get_state().items()}
        if create:
            self.shelfs = [Shelf(j + 1, int(st["shelf_x"][j]), int(st["shelf_y"][j])) for j in range(self._vec.n_shelves)]
            self.agents = [Agent(i + 1, int(st["agent_x"][i]), int(st["agent_y"][i]), Direction(int(st["agent_dir"][i])),
                                 self.msg_bits) for i in range(self.n_agents)]
        for j, s in enumerate(self.shelfs):
            s.x, s.y = int(st["shelf_x"][j]), int(st["shelf_y"][j])
        for i, a in enumerate(self.agents):
            a.x, a.y = int(st["agent_x"][i]), int(st["agent_y"][i])
            a.dir = Direction(int(st["agent_dir"][i]))
            c = int(st["agent_carry"][i])
            a.carrying_shelf = self.shelfs[c - 1] if c else None
            a.has_delivered = bool(st["agent_has_delivered"][i])
            if self.msg_bits:
                a.message[:] = st["agent_msg"][i]
        self.request_queue = [self.shelfs[int(q) - 1] for q in st["queue"]]
        self._cur_steps = int(st["cur_steps"])
        self._cur_inactive_steps = int(st["cur_inactive"])
        self._recalc_grid()

    def _use_slow_obs(self):
        """rware/warehouse.py:429-503: switch the non-image observations to the nested DICT form; returns its space."""
        self.fast_obs = False
        return self._make_observation_space(ObservationType.DICT)

    def _use_fast_obs(self):
        """rware/warehouse.py:505-522: switch the non-image observations to the flattened vectors; returns their space."""
        if self.fast_obs:
            return self.observation_space
        self.fast_obs = True
        return self._make_observation_space(ObservationType.FLATTENED)

    def _make_observation_space(self, observation_type):
        n, K = self.n_agents, 1 + 2 * self.sensor_range
        inf = float("inf")
        if observation_type == ObservationType.FLATTENED:  # warehouse.py:505-522
            return spaces.Tuple(tuple(spaces.Box(low=-inf, high=inf, shape=(self._obs_length,), dtype=np.float32)
                                      for _ in range(n)))
        if observation_type in (ObservationType.IMAGE, ObservationType.IMAGE_DICT):  # warehouse.py:352-427
            C = len(self.image_observation_layers)
            hi = np.ones((C, K, K), np.float32)
            for i, layer in enumerate(self.image_observation_layers):
                if layer == ImageLayer.AGENT_DIRECTION:
                    hi[i] *= 4
            image = spaces.Box(np.zeros((C, K, K), np.float32), hi, dtype=np.float32)
            if observation_type == ObservationType.IMAGE:
                return spaces.Tuple(tuple([image] * n))
            feats = spaces.Box(low=-inf, high=inf, shape=(6,), dtype=np.float32)
            return spaces.Tuple(tuple(spaces.Dict({"image": image, "features": feats}) for _ in range(n)))
        # DICT, warehouse.py:429-503
        from collections import OrderedDict
        hi = np.ones(2) if self.normalised_coordinates else np.ones(2) * max(self.grid_size)
        loc = spaces.Box(low=np.zeros(2), high=hi, shape=(2,), dtype=np.float32 if self.normalised_coordinates else np.int32)
        self_space = spaces.Dict(OrderedDict([("location", loc), ("carrying_shelf", spaces.MultiBinary(1)),
                                              ("direction", spaces.Discrete(4)), ("on_highway", spaces.MultiBinary(1))]))
        per = OrderedDict([("has_agent", spaces.MultiBinary(1)), ("direction", spaces.Discrete(4))])
        if self.msg_bits > 0:
            per["local_message"] = spaces.MultiBinary(self.msg_bits)
        per.update([("has_shelf", spaces.MultiBinary(1)), ("shelf_requested", spaces.MultiBinary(1))])
        return spaces.Tuple(tuple(
            spaces.Dict(OrderedDict([("self", self_space),
                                     ("sensors", spaces.Tuple(self._obs_sensor_locations * (spaces.Dict(per),)))]))
            for _ in range(n)))

    def _unflatten(self, row, agent):
        """FLATTENED row -> the nested DICT observation of warehouse.py:676-720."""
        mb, per = self.msg_bits, 7 + self.msg_bits
        ax, ay = (agent.x / (self.grid_size[1] - 1), agent.y / (self.grid_size[0] - 1)) if self.normalised_coordinates \
            else (agent.x, agent.y)
        obs = {"self": {"location": np.array([ax, ay], dtype=np.int32), "carrying_shelf": [int(row[2])],
                        "direction": int(np.argmax(row[3:7])), "on_highway": [int(row[7])]}}
        sensors = []
        for c in range(self._obs_sensor_locations):
            cell = row[8 + c * per: 8 + (c + 1) * per]
            has_agent = int(cell[0])
            msg = None
            if mb > 0:
                msg = np.asarray(cell[5:5 + mb], dtype=np.float64) if has_agent else mb * [0]
            sensors.append({"has_agent": [has_agent], "direction": int(np.argmax(cell[1:5])) if has_agent else 0,
                            "local_message": msg, "has_shelf": [int(cell[5 + mb])], "shelf_requested": [int(cell[6 + mb])]})
        obs["sensors"] = tuple(sensors)
        return obs

    def _obs_tuple(self, obs_tensor):
        o = obs_tensor[0].cpu().numpy()
        if not (self.image_obs or self.image_dict_obs):  # warehouse.py:722-744: `fast_obs` picks the form (598-674 / 676-720)
            if self.fast_obs:
                return tuple(o[i].copy() for i in range(self.n_agents))
            return tuple(self._unflatten(o[i], a) for i, a in enumerate(self.agents))
        K, C = 1 + 2 * self.sensor_range, len(self.image_observation_layers)
        for layer in self.image_observation_layers:  # layer[ag.x, ag.y] (sic, 552 / 558): numpy's IndexError, raised like there
            if layer in (ImageLayer.AGENT_DIRECTION, ImageLayer.AGENT_LOAD):
                for ag in self.agents:
                    if (layer == ImageLayer.AGENT_DIRECTION or ag.carrying_shelf is not None) and (
                            ag.x >= self.grid_size[0] or ag.y >= self.grid_size[1]):
                        axis, bad = (0, ag.x) if ag.x >= self.grid_size[0] else (1, ag.y)
                        raise IndexError(f"index {bad} is out of bounds for axis {axis} with size {self.grid_size[axis]}")
        imgs = [o[i, : C * K * K].reshape(C, K, K).copy() for i in range(self.n_agents)]
        if self.image_obs:
            return tuple(imgs)
        return tuple({"image": imgs[i], "features": o[i, C * K * K:].copy()} for i in range(self.n_agents))

    # ------------------------------------------------------------------ Gymnasium surface
    def seed(self, seed=None):  # rware/warehouse.py:962-964
        if seed is not None:
            self._vec.seed(int(seed))
            self._seeded = True

    def reset(self, seed=None, options=None):  # rware/warehouse.py:757-802
        if seed is not None:
            self.seed(seed)
        elif not self._seeded:  # the reference draws OS entropy for an unseeded generator
            self.seed(int.from_bytes(np.random.SeedSequence().generate_state(2).tobytes(), "little") >> 1)
        obs = self._vec.reset()
        self._pull(create=True)
        return self._obs_tuple(obs), self._get_info()

    def step(self, actions):  # rware/warehouse.py:804-946
        assert len(actions) == len(self.agents)
        acts = np.zeros((1, self.n_agents, 1 + self.msg_bits), np.int32)
        for i, (agent, action) in enumerate(zip(self.agents, actions)):
            if self.msg_bits > 0:
                agent.req_action = Action(getattr(action[0], "value", action[0]))  # ValueError on bad values, as the reference
                acts[0, i, 0] = agent.req_action.value
                acts[0, i, 1:] = np.asarray(action[1:], np.int32)
            else:
                agent.req_action = action if isinstance(action, Action) else Action(int(action))
                acts[0, i, 0] = agent.req_action.value
            agent.prev_x, agent.prev_y = agent.x, agent.y
        self._push()
        obs, reward, done = self._vec.step(torch.from_numpy(acts))
        rewards = [np.float64(r) for r in reward[0].cpu().numpy()]
        done = bool(done[0].item())
        self._pull()
        new_obs = self._obs_tuple(obs)
        return new_obs, rewards, done, False, self._get_info()

    def _make_obs(self, agent):  # rware/warehouse.py:722-744 (FLATTENED)
        self._push()
        return self._obs_tuple(self._vec.observations())[agent.id - 1]

    def _get_info(self):
        return {}

    def render(self):
        """rware/warehouse.py:948-956.  render_mode="rgb_array" returns the frame rasterised on the device from the world
        state (uint8 [2 + 31 H, 1 + 31 W, 3], the reference Viewer's window size); "human" needs a window system and is not
        part of the accelerated path."""
        if self.render_mode != "rgb_array":
            raise NotImplementedError('render_mode="human" opens a pyglet window in the reference; use render_mode="rgb_array"')
        self._push()
        return self._vec.render_rgb(0, 1)[0].cpu().numpy()

    def get_global_image(self, image_layers=(ImageLayer.SHELVES, ImageLayer.GOALS), recompute=False, pad_to_shape=None):
        """rware/warehouse.py:966-1040: stacked layers of the whole grid, computed on the device; cached in
        `self.global_image` until `recompute=True`, as in the reference."""
        if recompute or self.global_image is None:
            layers = [ImageLayer(getattr(v, "value", v)) for v in image_layers]
            for layer in layers:  # layer[ag.x, ag.y] (sic, 1004 / 1009): same IndexError as numpy raises in the reference
                if layer in (ImageLayer.AGENT_DIRECTION, ImageLayer.AGENT_LOAD):
                    for ag in self.agents:
                        if (layer == ImageLayer.AGENT_DIRECTION or ag.carrying_shelf is not None) and (
                                ag.x >= self.grid_size[0] or ag.y >= self.grid_size[1]):
                            bad_axis, bad = (0, ag.x) if ag.x >= self.grid_size[0] else (1, ag.y)
                            raise IndexError(f"index {bad} is out of bounds for axis {bad_axis} with size {self.grid_size[bad_axis]}")
            if pad_to_shape is not None:
                assert all(p >= g for p, g in zip(pad_to_shape, (len(layers), *self.grid_size)))  # 1027
            self._push()
            self.global_image = self._vec.global_image(layers, pad_to_shape)[0].cpu().numpy()
        return self.global_image

    def close(self):
        self._vec.close()
