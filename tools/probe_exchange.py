#!/usr/bin/env python3
"""Where the time of SampleIndexAccumulator.add/present goes (run under torchrun, 2+ GPUs); not a benchmark."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")))
from gdpathtracing_b200 import PathTracingCamera, multigpu, scenes
rank, local, world = multigpu.init_process_group()
dev = torch.device("cuda", local)
H, W, n = 1080, 1920, 20
sc = scenes.cornell32(); grp = scenes.populate(sc)
cam = PathTracingCamera(); cam.fov = sc.fov; cam.geometry_group = grp; cam.denoising_mode = PathTracingCamera.NONE
cam.set_window_size(64, 64); cam.set_global_transform(sc.camera_transform12); cam.set_cuda_device(local); cam.init()
stream = torch.cuda.ExternalStream(cam.stream())
frames = torch.randint(0, 255, (n, H, W, 4), dtype=torch.uint8, device=dev)
for rep in range(3):
    with torch.cuda.stream(stream):
        acc = multigpu.SampleIndexAccumulator(H, W, rank, world, dev, multigpu.cuda_k2(cam))
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        acc.add(frames)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        acc.present()
        torch.cuda.synchronize(); t2 = time.perf_counter()
    if rank == 0:
        print(f"rep {rep}: add {1e3 * (t1 - t0):.2f} ms  present {1e3 * (t2 - t1):.2f} ms  ({n} frames, {world} ranks)", flush=True)
    if hasattr(acc, "timing") and rank == 0:
        print(acc.timing)
