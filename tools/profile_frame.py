#!/usr/bin/env python3
"""Render a few frames of one BASELINE scene and print per-stage device times.  Meant to be
wrapped by ncu (see profiles/README.md); never a source of benchmark numbers under a profiler."""
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
ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--soup-tris", type=int, default=1_000_000)
ap.add_argument("--variant", type=int, default=-1)
ap.add_argument("--tune", action="append", default=[], metavar="NAME=N")
ap.add_argument("--mode", type=int, default=0, help="0 progressive, 2 none")
args = ap.parse_args()

sc = {"demo": scenes.demo_scene, "cornell32": scenes.cornell32, "instanced": scenes.instanced_grid,
      "soup": lambda: scenes.triangle_soup(args.soup_tris)}[args.scene]()
grp = scenes.populate(sc)
cam = PathTracingCamera()
cam.fov = sc.fov
cam.geometry_group = grp
cam.denoising_mode = args.mode
cam.set_window_size(args.width, args.height)
cam.set_global_transform(sc.camera_transform12)
cam.set_max_depth(args.depth)
if args.variant >= 0:
    cam.set_variant(args.variant)
for kv in args.tune:
    cam.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
cam.init()
_lib.cuda.gdpt_shader_set_stage_timing(cam.main_shader, 1)
buf = (ctypes.c_float * 64)()
for f in range(args.frames):
    cam.render_device_only()
    st = cam.stats()
    n = _lib.cuda.gdpt_shader_get_stage_times(cam.main_shader, buf, 64)
    print(json.dumps({"frame": f + 1, "rays": st["rays"], "k1_ms": round(st["k1_ms"], 4), "k2_ms": round(st["k2_ms"], 4),
                      "mrays_s": round(st["rays"] / max(st["k1_ms"] + st["k2_ms"], 1e-9) / 1e3, 1),
                      "stage_ms": [round(x, 4) for x in buf[:n]]}), flush=True)
