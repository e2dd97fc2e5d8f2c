"""Small workload that goes through every mode of the fused kernel (RESET, STEP with auto-reset, OBS), both step forms
(one lane per world / warp-cooperative, RWARE_B200_COOP=0|1), the dynamic and the static instantiations, image observations,
state injection, the views and the host transport -- meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/sanitize_workload.py
    compute-sanitizer --tool racecheck python profiles/sanitize_workload.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rware_b200

E = int(os.environ.get("SAN_ENVS", "300"))   # 9.4 tiles: ragged last tile
rs = np.random.default_rng(0)
for env_id, over in (("rware-tiny-4ag-v1", {}), ("rware-tiny-4ag-v1", {"max_steps": 7}), ("rware-small-4ag-v1", {"msg_bits": 2, "sensor_range": 2}),
                     ("rware-medium-8ag-hard-v1", {"observation_type": 2}),
                     ("rware-tiny-2ag-v1", {"observation_type": 3, "image_observation_layers": (0, 1, 2, 3, 4, 5, 6)})):
    v = rware_b200.make_vec(env_id, E, auto_reset=True, **over)
    v.reset(seed=1)
    A = 1 + v.msg_bits
    for t in range(12):
        acts = rs.integers(0, 5 if A == 1 else 2, size=(E, v.n_agents, A)).astype(np.int32)
        acts[..., 0] = rs.integers(0, 5, size=(E, v.n_agents))
        if rs.random() < 0.6:
            acts[..., 0] = np.where(rs.random((E, v.n_agents)) < 0.7, 1, acts[..., 0])  # collisions
        obs, rew, done = v.step(torch.from_numpy(acts).cuda())
    v.observations()
    st = v.get_state()
    v.set_state(**{k: t for k, t in st.items() if k != "rng"})
    v.step(torch.from_numpy(acts).cuda())
    v.reset(mask=(torch.arange(E) % 3 == 0))
    v.global_image()
    v.render_rgb(0, 2)
    torch.cuda.synchronize()
    print("ok", env_id, over, "errors", v.error_count(), flush=True)
    v.close()
# host transport (bits + copy engine share) on a batch large enough to take the compressed path
v = rware_b200.make_vec("rware-small-4ag-v1", 4000)
v.reset(seed=2)
be = v.backend
h_act = torch.randint(0, 5, (4000, 4), dtype=torch.int32).pin_memory()
h_obs = be.host_alloc((4000, 4, v.obs_len)); h_rew = be.host_alloc((4000, 4)); h_done = be.host_alloc((4000,), torch.uint8)
for _ in range(3):
    be.step_host(h_act, h_obs, h_rew, h_done)
print("ok host transport", float(h_obs.sum()), flush=True)
