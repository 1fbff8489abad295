#!/usr/bin/env python3
"""Where the end-to-end frame time goes: back-to-back device-only frames, pipelined frames (render_begin/wait) and
blocking frames of the demo scene, wall-clock per frame over many frames (no L2 flush, no per-frame Python work
beyond the calls themselves)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, scenes  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
DEPTH = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sc = scenes.demo_scene()
grp = scenes.populate(sc)
cam = PathTracingCamera()
cam.fov = sc.fov
cam.geometry_group = grp
cam.denoising_mode = PathTracingCamera.PROGRESSIVE_RENDERING
cam.set_window_size(1920, 1080)
cam.set_global_transform(sc.camera_transform12)
cam.set_max_depth(8)
cam.init()
for _ in range(5):
    cam.render()
out = {}
cam.synchronize()
t0 = time.perf_counter()
for _ in range(N):
    cam.render_device_only()
cam.synchronize()
out["device_only_ms"] = (time.perf_counter() - t0) / N * 1e3
t0 = time.perf_counter()
for _ in range(N):
    cam.render()
out["blocking_ms"] = (time.perf_counter() - t0) / N * 1e3
for _ in range(3):
    cam.render_begin(); cam.render_wait()
t0 = time.perf_counter()
inflight = 0
host_begin = host_wait = 0.0
for _ in range(N):
    a = time.perf_counter()
    cam.render_begin()
    host_begin += time.perf_counter() - a
    inflight += 1
    if inflight == DEPTH:
        a = time.perf_counter()
        cam.render_wait()
        host_wait += time.perf_counter() - a
        inflight -= 1
while inflight:
    cam.render_wait(); inflight -= 1
out["pipelined_ms"] = (time.perf_counter() - t0) / N * 1e3
out["pipelined_host_begin_ms"] = host_begin / N * 1e3
out["pipelined_host_wait_ms"] = host_wait / N * 1e3


def pipelined(tag, set_index=False, touch=False):
    t0 = time.perf_counter()
    inflight = 0
    acc = 0
    for i in range(N):
        if set_index:
            cam.set_frame_index(1000 + i)
        cam.render_begin()
        inflight += 1
        if inflight == DEPTH:
            img, fst = cam.render_wait()
            if touch:
                acc += int(img[0, 0, 3])
            inflight -= 1
    while inflight:
        cam.render_wait(); inflight -= 1
    out[tag] = (time.perf_counter() - t0) / N * 1e3


pipelined("pipe_plain_ms")
pipelined("pipe_set_index_ms", set_index=True)
pipelined("pipe_touch_ms", touch=True)
from gdpathtracing_b200 import _lib
_lib.cuda.gdpt_shader_set_stage_timing(cam.main_shader, 1)
pipelined("pipe_stage_timing_ms")
pipelined("pipe_all_ms", set_index=True, touch=True)
print(json.dumps(out))
