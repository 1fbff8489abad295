#!/usr/bin/env python3
"""Per-warp schedule profile of the path kernel for one BASELINE scene (gdpt_shader_set_warp_profile):
when each warp started / ended, how many scheduler iterations it spent per phase, paths started.
Prints a summary; --out saves the raw table (npy)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, _lib, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="demo")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--soup-tris", type=int, default=1_000_000)
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--tune", action="append", default=[], metavar="NAME=N")
ap.add_argument("--out", default="")
args = ap.parse_args()

sc = {"demo": scenes.demo_scene, "cornell32": scenes.cornell32, "instanced": scenes.instanced_grid,
      "soup": lambda: scenes.triangle_soup(args.soup_tris)}[args.scene]()
grp = scenes.populate(sc)
cam = PathTracingCamera()
cam.fov = sc.fov
cam.geometry_group = grp
cam.denoising_mode = 2
cam.set_window_size(args.width, args.height)
cam.set_global_transform(sc.camera_transform12)
cam.set_max_depth(args.depth)
if args.variant >= 0:
    cam.set_variant(args.variant)
for kv in args.tune:
    cam.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
cam.init()
_lib.check(_lib.cuda.gdpt_shader_set_warp_profile(cam.main_shader, 1), cam.device, "set_warp_profile")
for f in range(args.frames):
    cam.render_device_only()
    st = cam.stats()
buf = np.zeros(8 * 65536, np.uint64)
n = _lib.cuda.gdpt_shader_read_warp_profile(cam.main_shader, buf.ctypes.data_as(ctypes.c_void_p), buf.size)
assert n > 0, n
t = buf[:8 * n].reshape(n, 8).astype(np.int64)
t = t[t[:, 1] > 0]
lanes = t[:, 2:7] >> 32          # (unused: no schedule packs lane counts into the high words any more)
idle_iters = t[:, 7] >> 32
t[:, 2:8] &= 0xFFFFFFFF
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
iters = t[:, 2:7]
tot = iters.sum(axis=1)
dur = end - start
print(json.dumps({
    "warps": int(len(t)), "k1_ms": round(st["k1_ms"], 4), "rays": int(st["rays"]),
    "kernel_span_us": round(float(end.max()), 1),
    "warp_end_us_pcts(10,50,90,99,100)": [round(float(x), 1) for x in np.percentile(end, [10, 50, 90, 99, 100])],
    "iters_per_warp_pcts(10,50,90,99,100)": [int(x) for x in np.percentile(tot, [10, 50, 90, 99, 100])],
    "iters_total": int(tot.sum()), "by_phase_I_L_T_F_E": [int(x) for x in iters.sum(axis=0)],
    "ns_per_iter_median": round(float(np.median(dur * 1e3 / np.maximum(tot, 1))), 1),
    "ns_per_iter_slowest_warps": round(float(np.mean((dur * 1e3 / np.maximum(tot, 1))[np.argsort(end)[-32:]])), 1),
    "paths_started_pcts(10,50,90,100)": [int(x) for x in np.percentile(t[:, 7], [10, 50, 90, 100])],
    "busy_fraction": round(float(dur.sum() / (len(t) * end.max())), 3),
    "lanes_per_iteration_I_L_T_F_E": [round(float(lanes[:, k].sum() / max(iters[:, k].sum(), 1)), 1) for k in range(5)],
    "idle_iterations": int(idle_iters.sum()),
}))
if args.out:
    np.save(args.out, t)
