#!/usr/bin/env python3
"""Frames in flight: ms per frame of render_begin/render_wait on the 1080p demo frame for a few tuning settings
(probe for DESIGN.md; bench.py's e2e leg is the reported number)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, scenes

def run(tune, depth_in_flight=3, frames=200):
    sc = scenes.demo_scene(); grp = scenes.populate(sc)
    cam = PathTracingCamera(); cam.fov = sc.fov; cam.geometry_group = grp; cam.denoising_mode = 0
    cam.set_window_size(1920, 1080); cam.set_global_transform(sc.camera_transform12); cam.set_max_depth(8)
    for k, v in tune.items():
        cam.set_tuning(k, v)
    cam.init()
    for _ in range(5):
        cam.render()
    best = 1e9
    for rep in range(3):
        infl = 0
        t = time.perf_counter()
        for i in range(frames):
            cam.render_begin(); infl += 1
            if infl == depth_in_flight:
                cam.render_wait(); infl -= 1
        while infl:
            cam.render_wait(); infl -= 1
        best = min(best, (time.perf_counter() - t) / frames * 1e3)
    return best

for name, tune in (("default (half grids when pipelined)", {}), ("full grids", {"BLOCKS_PER_SM": 4}), ("one block per SM", {"BLOCKS_PER_SM": 1})):
    for d in (2, 3, 4):
        print(f"{name:36s} frames in flight {d}: {run(tune, d):.4f} ms/frame", flush=True)
