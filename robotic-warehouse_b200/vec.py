"""VecWarehouse: E independent RWARE worlds advanced by one fused CUDA kernel per step.

The batched counterpart of the reference's `Warehouse` (rware/warehouse.py:140-946): same constructor kwargs, same
per-world semantics (bit-exact observations / rewards / done, and the same numpy generator streams per world), but
tensors in and out:

    env = VecWarehouse(65536, shelf_columns=3, column_height=8, shelf_rows=1, n_agents=4, request_queue_size=4,
                       max_steps=500)
    obs = env.reset(seed=0)                     # float32 [E, n, L]; world e is seeded with seed + e
    obs, reward, done = env.step(actions)       # actions int32 [E, n] (or [E, n, 1 + msg_bits])

The returned tensors are views of buffers owned by the object and are overwritten by the next call.
"""
import numpy as np
import torch

from . import layout as _layout
from ._capi import STATE_FIELDS, CudaBackend, EnvConfig


class VecWarehouse:
    def __init__(self, num_envs, shelf_columns=None, column_height=None, shelf_rows=None, n_agents=1, msg_bits=0,
                 sensor_range=1, request_queue_size=1, max_inactivity_steps=None, max_steps=None, reward_type=1,
                 layout=None, normalised_coordinates=False, device=None, auto_reset=False, _backend_factory=None,
                 observation_type=1, image_observation_layers=(0, 1, 2, 5, 6), image_observation_directional=True):
        if not layout:
            self.grid_size, self.highways, self.goals = _layout.layout_from_params(shelf_columns, shelf_rows, column_height)
            self.column_height = column_height
        else:
            self.grid_size, self.highways, self.goals = _layout.layout_from_str(layout)
        self.cfg = EnvConfig(self.highways, self.goals, n_agents, msg_bits, sensor_range, request_queue_size,
                             max_inactivity_steps, max_steps, reward_type, normalised_coordinates,
                             observation_type, image_observation_layers, image_observation_directional)
        self.num_envs = int(num_envs)
        self.n_agents, self.msg_bits, self.sensor_range = self.cfg.n_agents, self.cfg.msg_bits, self.cfg.sensor_range
        self.request_queue_size = self.cfg.request_queue_size
        self.n_shelves, self.obs_len = self.cfg.n_shelves, self.cfg.obs_len
        factory = _backend_factory or CudaBackend  # tests may inject the host twin; the product path is CUDA only
        self.backend = factory(self.cfg, self.num_envs, device)
        self.device = self.backend.device
        self.info = self.backend.info
        E, n, L = self.num_envs, self.n_agents, self.obs_len
        self._obs = torch.empty((E, n, L), dtype=torch.float32, device=self.device)
        self._reward = torch.empty((E, n), dtype=torch.float32, device=self.device)
        self._done = torch.empty((E,), dtype=torch.uint8, device=self.device)
        self.auto_reset = bool(auto_reset)
        self.backend.set_auto_reset(self.auto_reset)

    def split_image_obs(self, obs):
        """For the image observation types: (image [E, n, C, K, K], features [E, n, 6] or None) views of an obs tensor."""
        c = self.cfg
        assert c.observation_type >= 2
        K = 2 * c.sensor_range + 1
        C = len(c.image_layers)
        img = obs[..., : C * K * K].reshape(*obs.shape[:-1], C, K, K)
        return img, (obs[..., C * K * K:] if c.observation_type == 3 else None)

    # ------------------------------------------------------------------ seeding / reset
    def seed(self, seed, mask=None):
        """World e gets Generator(PCG64(SeedSequence(seed + e))) (int seed) or SeedSequence(seed[e]) (array)."""
        if np.isscalar(seed):
            seeds = (np.uint64(seed) + np.arange(self.num_envs, dtype=np.uint64))
        else:
            seeds = np.asarray(seed, dtype=np.uint64)
        self.backend.seed(seeds, mask)

    def reset(self, seed=None, mask=None):
        """Warehouse.reset for the worlds selected by `mask` (uint8/bool tensor [E], default all)."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        if seed is not None:
            self.seed(seed, None if m is None else m.cpu().numpy())
        self.backend.reset(m, self._obs)
        return self._obs

    # ------------------------------------------------------------------ step
    def step(self, actions, compute_obs=True):
        a = actions
        if not (isinstance(a, torch.Tensor) and a.dtype == torch.int32 and a.device == self.device and a.is_contiguous()):
            a = torch.as_tensor(np.asarray(actions) if not isinstance(actions, torch.Tensor) else actions,
                                device=self.device).to(torch.int32).contiguous()
        expect = self.num_envs * self.n_agents * (1 + self.msg_bits)
        assert a.numel() == expect, f"actions must have {expect} elements, got {a.numel()}"
        self.backend.step(a, self._obs if compute_obs else None, self._reward, self._done)
        return (self._obs if compute_obs else None), self._reward, self._done

    def observations(self):
        self.backend.obs(self._obs)
        return self._obs

    def observation_bytes(self, obs=None, out=None):
        """uint8 image of an observation tensor (default: the last one returned): the same values, one byte each.  Exact because
        observation values are integers in 0..255 unless `normalised_coordinates` is set (then this raises)."""
        obs = self._obs if obs is None else obs
        if out is None:
            out = torch.empty(obs.shape, dtype=torch.uint8, device=self.device)
        assert obs.is_contiguous() and out.is_contiguous() and obs.numel() == self._obs.numel() == out.numel()
        self.backend.obs_bytes(obs, out)
        return out

    # ------------------------------------------------------------------ whole-warehouse views
    def global_image(self, image_layers=(0, 5), pad_to_shape=None, out=None):
        """Warehouse.get_global_image(image_layers, recompute=True, pad_to_shape) (rware/warehouse.py:966-1040) of every world:
        float32 [E, C', H', W'] on the device (input of a centralised critic).  Layers AGENT_DIRECTION / AGENT_LOAD are written
        transposed exactly as the reference writes them; worlds where the reference would raise IndexError for that are
        counted in `error_count()`."""
        layers = [int(getattr(v, "value", v)) for v in image_layers]
        if not layers or any(v < 0 or v > 6 for v in layers):
            raise ValueError(f"Unknown image layer type: {image_layers}")  # warehouse.py:1018-1019
        shape = self.backend.global_image_shape(len(layers), pad_to_shape)
        if out is None:
            out = torch.empty((self.num_envs, *shape), dtype=torch.float32, device=self.device)
        assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (self.num_envs, *shape)
        self.backend.global_image(layers, pad_to_shape, out)
        return out

    def render_rgb(self, first_env=0, n_envs=1, out=None):
        """`render()` of render_mode="rgb_array" (rware/warehouse.py:948-956, rware/rendering.py) for the worlds
        [first_env, first_env + n_envs): uint8 [n_envs, 2 + 31 H, 1 + 31 W, 3] on the device."""
        h, w = self.backend.render_shape()
        if out is None:
            out = torch.empty((n_envs, h, w, 3), dtype=torch.uint8, device=self.device)
        assert out.is_contiguous() and out.dtype == torch.uint8 and tuple(out.shape) == (n_envs, h, w, 3)
        self.backend.render_rgb(first_env, n_envs, out)
        return out

    def error_count(self):
        """Invalid action values seen so far (the reference raises ValueError) + transposed-layer index errors."""
        return self.backend.error_count()

    # ------------------------------------------------------------------ state access (checkpoint / test injection)
    def _state_shapes(self):
        E, n, S, R, mb = self.num_envs, self.n_agents, self.n_shelves, self.request_queue_size, self.msg_bits
        shapes = {k: (E, n) for k in ("agent_x", "agent_y", "agent_dir", "agent_carry", "agent_has_delivered")}
        shapes.update(agent_msg=(E, n, mb), shelf_x=(E, S), shelf_y=(E, S), queue=(E, R), cur_steps=(E,),
                      cur_inactive=(E,), rng=(E, 10))
        return shapes

    def get_state(self):
        shapes = self._state_shapes()
        out = {}
        for k in STATE_FIELDS:
            if k == "agent_msg" and self.msg_bits == 0:
                continue
            out[k] = torch.zeros(shapes[k], dtype=torch.int32, device=self.device)
        self.backend.get_state(out)
        return out

    def set_state(self, **fields):
        shapes = self._state_shapes()
        tens = {}
        for k, v in fields.items():
            if k not in shapes:
                raise KeyError(k)
            if k == "agent_msg" and self.msg_bits == 0:
                continue
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v, device=self.device)
            t = t.to(torch.int32).contiguous()
            assert tuple(t.shape) == shapes[k], f"{k}: expected {shapes[k]}, got {tuple(t.shape)}"
            tens[k] = t
        self.backend.set_state(tens)
        self._keepalive = tens

    def close(self):
        self.backend.close()
