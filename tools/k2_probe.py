#!/usr/bin/env python3
"""K2 alone (gdpt_progressive_accumulate): CUDA-event time per launch at 1080p and 4K with L2 flushed between launches,
as GB/s of the 36 B/pixel SURVEY 8d charges it (probe for DESIGN.md; bench.py's roofline_accumulate is the reported number)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, multigpu, scenes  # noqa: E402

sc = scenes.cornell32(); grp = scenes.populate(sc)
cam = PathTracingCamera(); cam.fov = sc.fov; cam.geometry_group = grp; cam.denoising_mode = PathTracingCamera.NONE
cam.set_window_size(64, 64); cam.set_global_transform(sc.camera_transform12); cam.set_max_depth(2); cam.init()
k2 = multigpu.cuda_k2(cam)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(cam.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for name, (W, H) in (("1080p", (1920, 1080)), ("4k", (3840, 2160))):
    raw = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev)
    raw[: H // 8] = 0  # black rows: zero numerators
    screen = torch.empty_like(raw)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    times = []
    with torch.cuda.stream(stream):
        for it in range(40):
            flush.fill_(it & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); k2(raw, screen, acc, it + 1); b.record(stream)
            times.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in times[8:])
    med = ms[len(ms) // 2]
    out[name] = {"us_median": round(med * 1e3, 2), "us_min": round(ms[0] * 1e3, 2), "gb_s_at_36B_px": round(W * H * 36 / med / 1e6, 1)}
print(json.dumps(out))
