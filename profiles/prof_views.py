"""Times the whole-warehouse view kernels (csrc/rw_views.cu) on one GPU: batched get_global_image and rgb frames.
    python profiles/prof_views.py [--envs 65536] > profiles/r1_views.json
Roofline: HBM bytes = output tensor written + packed state read, against MEASURED_PEAKS.json hbm_gbs."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rware_b200  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    args = ap.parse_args()
    peak = 6574.8
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    out = {"peak_hbm_gbs": peak, "cases": []}
    for env_id in ("rware-tiny-4ag-v2", "rware-medium-8ag-hard-v2"):
        E = args.envs
        vec = rware_b200.make_vec(env_id, E, auto_reset=True)
        vec.reset(seed=0)
        acts = torch.randint(0, 5, (E, vec.n_agents), dtype=torch.int32, device="cuda")
        for _ in range(5):
            vec.step(acts)
        state_bytes = vec.info["state_words_per_env"] * 4
        H, W = vec.grid_size
        # rotate over several output buffers larger than L2 in total so stores reach HBM
        for layers, pad in (((0, 5), None), ((0, 1, 2, 5, 6), None), ((0, 1, 2, 3, 4, 5, 6), (7, max(H, W), max(H, W)))):
            shape = vec.backend.global_image_shape(len(layers), pad)
            nbuf = max(2, int(400e6 // (E * shape[0] * shape[1] * shape[2] * 4)) + 1)
            bufs = [torch.empty((E, *shape), dtype=torch.float32, device="cuda") for _ in range(nbuf)]
            k = [0]

            def run():
                vec.global_image(layers, pad, out=bufs[k[0] % nbuf])
                k[0] += 1
            us = timed(run)
            nbytes = E * (shape[0] * shape[1] * shape[2] * 4 + state_bytes)
            out["cases"].append(dict(kernel="rw_global_image", env=env_id, worlds=E, layers=list(layers), pad_to_shape=pad,
                                     us_per_launch=round(us, 2), bytes_per_launch=nbytes, gbs=round(nbytes / us / 1e3, 1),
                                     frac_of_peak=round(nbytes / us / 1e3 / peak, 3), out_buffers_rotated=nbuf))
        n_frames = 1024
        h, w = vec.backend.render_shape()
        bufs = [torch.empty((n_frames, h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        k = [0]

        def run2():
            vec.render_rgb(0, n_frames, out=bufs[k[0] % 2])
            k[0] += 1
        us = timed(run2, reps=10)
        nbytes = n_frames * h * w * 3
        out["cases"].append(dict(kernel="rw_render_rgb", env=env_id, frames=n_frames, frame_shape=[h, w, 3], us_per_launch=round(us, 2),
                                 bytes_per_launch=nbytes, gbs=round(nbytes / us / 1e3, 1), frac_of_peak=round(nbytes / us / 1e3 / peak, 3)))
        vec.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
