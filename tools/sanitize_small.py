#!/usr/bin/env python3
"""Small frames of every scene through the default path (pooled kernel) -- meant to run under compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, scenes  # noqa: E402

for make, W, H, depth in ((scenes.cornell32, 96, 64, 4), (scenes.demo_scene, 128, 72, 8),
                          (lambda: scenes.instanced_grid(3, 800, seed=3), 96, 54, 5)):
    sc = make()
    grp = scenes.populate(sc)
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.PROGRESSIVE_RENDERING
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth)
    cam.init()
    for _ in range(2):
        cam.render()
    cam.render_begin(); cam.render_wait()
    print(sc.name, cam.stats()["rays"], "rays ok", flush=True)
