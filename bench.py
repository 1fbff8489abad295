#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric: Mrays/s and ms/frame @1080p, 1 spp/frame, depth 8 on the
project demo scene (config C2), N B200s of one node, next to the host-CPU reference.

A "step" is one frame: camera upload, K1 (ray generation .. bounce loop: a camera-ray
classification kernel and a persistent path kernel) and K2 (progressive accumulation + ACES).  A "ray" is one ray_trace() call of the
reference (main.glsl:352): a primary ray or one bounce segment; counted exactly on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference ...                           the reference arm: the CPU
        restatement of the reference's shader loop (oracle/, kind "port") on all host threads

Prints ONE JSON line (rank 0).  Multi-GPU (torchrun): sample-index partition, each rank renders
frame_index = step*N + rank + 1 of the same scene, one NCCL sum-reduce of the accumulations at
the end of the timed region (weak scaling; `--partition rows` = row-band strong scaling with a
per-step all-gather).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

LANE_OPS_BOX, LANE_OPS_TRI, LANE_OPS_TLAS_LEAF = 22, 55, 45
JSON_OUT = sys.stdout


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=float(p["hbm_gbs"]), sm_max_mhz=float(p.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


def ncu_traffic(kernel, args):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the newest committed
    `ncu --set full` capture under profiles/ (taken on the default workload, so only quoted for it)."""
    if (args.scene, args.width, args.height, args.depth) != ("demo", 1920, 1080, 8):
        return None
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(REPO, "profiles", f"r*_ncu_{kernel}_raw.csv"))):
        vals = {}
        for row in open(path):
            parts = row.strip().split(",")
            if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(parts[1])
                if scale:
                    vals[parts[0]] = float(parts[2]) * scale
        if len(vals) == 2:
            best = {"bytes": sum(vals.values()), "source": os.path.relpath(path, REPO)}
    return best


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(args):
    from gdpathtracing_b200 import scenes
    if args.scene == "demo":
        sc = scenes.demo_scene()
    elif args.scene == "cornell32":
        sc = scenes.cornell32()
    elif args.scene == "soup":
        sc = scenes.triangle_soup(args.soup_tris)
    else:
        sc = scenes.instanced_grid()
    return sc, scenes.populate(sc)


def workload_name(args):
    return {"demo": "C2 project demo scene (Cornell room + 2 Suzanne + Gobot, albedo textures)",
            "cornell32": "C1 synthetic Cornell box (32 tris)", "soup": f"C3 {args.soup_tris}-triangle soup",
            "instanced": "C4 10M-triangle instanced grid"}[args.scene] + f" {args.width}x{args.height}, 1 spp/frame, depth {args.depth}"


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from gdpathtracing_b200 import nodes
    from oracle import oracle
    sc, grp = build_scene(args)
    build_s = grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    W, H = args.width, args.height
    threads = oracle.hardware_threads()

    def frame(idx, row_step):
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, idx))
        t0 = time.perf_counter()
        r = oracle.path_trace(osc, W, H, cam, max_depth=args.depth, threads=threads, row_step=row_step)
        screen, acc = r["rgba8"], np.zeros((H, W, 4), np.float32)
        if row_step == 1:
            oracle.progressive(screen, acc, 1)
        return time.perf_counter() - t0, r["stats"]["rays"]

    t_probe, _ = frame(1, 8)  # every 8th row: sizes the bounded sample
    budget = 150.0
    est_full = t_probe * 8.0 * (args.steps + args.warmup)
    row_step = 1 if est_full <= budget else int(min(64, np.ceil(est_full / budget)))
    for w in range(args.warmup):
        frame(w + 1, row_step)
    total_t, total_rays = 0.0, 0
    for s in range(args.steps):
        t, rays = frame(args.warmup + s + 1, row_step)
        total_t += t; total_rays += rays
    mrays = total_rays / total_t / 1e6
    ref_build = None
    if oracle.ref_available():
        t0 = time.perf_counter()
        h, _ = oracle.reference_build(sc)
        oracle.ref().refbvh_free(h)
        ref_build = time.perf_counter() - t0
    sample = f"every {row_step}-th row of each {W}x{H} frame" if row_step > 1 else f"full {W}x{H} frames"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_t / args.steps * 1e3 * row_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "what": "CPU restatement of main.glsl/brdfs.glsl/progressive_rendering.glsl "
                   "(the reference's GPU half cannot run without Godot+Vulkan); ms_per_step is scaled to a full frame"},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample,
                         "bvh_build_s": build_s, "reference_bvh_build_s": ref_build},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------ our arm
def stage_kinds(n, depth):
    """Kernel name of each timed K1 stage, from the number of stages the backend reports:
    1 = single path kernel, 2 = camera-ray classification + path kernel, 2*depth = wavefront."""
    if n == 1:
        return ["k_path"]
    if n == 2:
        return ["k_primary_cull", "k_path"]
    return ["k_trace<primary>" if i == 0 else ("k_shade" if i % 2 == 1 else "k_trace<bounce>") for i in range(n)]


def run_ours(args, rank, local, world):
    import torch
    from gdpathtracing_b200 import PathTracingCamera, _lib, multigpu
    from gdpathtracing_b200._lib import cuda

    peaks = measured_peaks()
    torch.cuda.set_device(local)
    dev_t = torch.device("cuda", local)
    sc, grp = build_scene(args)
    W, H, D = args.width, args.height, args.depth
    rows_mode = world > 1 and args.partition == "rows"

    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.PROGRESSIVE_RENDERING
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(D)
    cam.set_cuda_device(local)
    if rows_mode:
        cam.set_shard(rank, world, args.band)
    if args.variant >= 0:
        cam.set_variant(args.variant)
    for kv in args.tune:
        k, v = kv.split("=")
        cam.set_tuning(k, int(v))
    cam.init()
    build_s = None  # GeometryGroup3D.build ran inside init(); timed separately below on rank 0
    # per-kernel split: measured on a few untimed frames after the timed region (the events it records between the K1
    # kernels cost ~60 us per frame, so they stay out of every timed leg)

    def set_index(step):
        # frame_index is incremented by render(); make step s use index s*world + rank + 1 (sample partition)
        cam.set_frame_index(step * world + rank if not rows_mode else step)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t) if args.l2_flush else None

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        cam.synchronize()

    out_ptr, out_size = cam.device_pointer("output")
    frame_t = multigpu.as_tensor(out_ptr, (H, W, 4), torch.uint8, dev_t)
    # row bands: the presented frame assembles itself in every rank's image (K2 writes its rows to the peers over
    # NVLink, frames separated by a one-element all-reduce) unless --present gather asks for the NCCL all-gather
    peer_frame = multigpu.PeerFrame(cam, rank, world) if (rows_mode and args.present == "peer") else None

    def present():
        if peer_frame is not None:
            peer_frame.barrier()
        else:
            multigpu.gather_row_bands(frame_t, args.band, rank, world)

    # ---- warm-up
    for w in range(args.warmup):
        set_index(w)
        cam.render_device_only()
        cam.synchronize()
        if rows_mode:
            present()
            torch.cuda.synchronize()
    stage_ms = np.zeros(64)
    n_stage = 0
    launches_per_frame = 0
    buf = (ctypes.c_float * 64)()

    # ---- device-timed leg: inputs resident, nothing leaves the GPU
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    dev_ms, rays_total, k2_ms_total, gather_ms, k1_ms_total = 0.0, 0, 0.0, 0.0, 0.0
    for s in range(args.steps):
        if flush is not None:
            flush.fill_(s & 0xFF)
            torch.cuda.synchronize()
        set_index(args.warmup + s)
        cam.render_device_only()
        st = cam.stats()  # blocks until the frame is done; k1_ms/k2_ms are CUDA-event times on the launching stream
        dev_ms += st["k1_ms"] + st["k2_ms"]
        k2_ms_total += st["k2_ms"]
        k1_ms_total += st["k1_ms"]
        rays_total += st["rays"]
        launches_per_frame = st["kernel_launches"] + 1  # K1 kernels + K2
        if rows_mode:
            ts = peer_frame._stream if peer_frame is not None else torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            present()
            e1.record(ts); torch.cuda.synchronize(); cam.synchronize()
            gather_ms += e0.elapsed_time(e1)
    if world > 1 and not rows_mode:
        acc_ptr, _ = cam.device_pointer("accum")
        acc_t = multigpu.as_tensor(acc_ptr, (H, W, 4), torch.float32, dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        multigpu.reduce_accumulations(acc_t, dst=0)
        e1.record(); torch.cuda.synchronize()
        gather_ms += e0.elapsed_time(e1)
    # ---- per-kernel split of K1 (untimed frames, stage events on)
    _lib.check(cuda.gdpt_shader_set_stage_timing(cam.main_shader, 1), cam.device, "set_stage_timing")
    split_frames = min(args.steps, 8)
    for s in range(split_frames):
        set_index(args.warmup + s)
        cam.render_device_only()
        cam.stats()
        n = cuda.gdpt_shader_get_stage_times(cam.main_shader, buf, 64)
        stage_ms[:n] += np.array(buf[:n])
        n_stage = n
        if rows_mode:
            present()
            torch.cuda.synchronize(); cam.synchronize()
    _lib.check(cuda.gdpt_shader_set_stage_timing(cam.main_shader, 0), cam.device, "set_stage_timing")
    present_check = None
    if peer_frame is not None:
        # the image every rank holds now must be what the NCCL all-gather of the bands assembles
        whole = frame_t.clone()
        gathered = multigpu.gather_row_bands(frame_t, args.band, rank, world)
        same = torch.tensor([1 if torch.equal(whole, gathered) else 0], dtype=torch.int32, device=dev_t)
        torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
        present_check = bool(same.item())
        if not present_check:
            raise AssertionError("peer-written frame differs from the all-gathered frame")
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    dev_ms += gather_ms

    # ---- end-to-end leg: host camera block in, RGBA8 frame back in (pinned) host memory, every step.
    # (a) the reference's blocking render(): one frame at a time; (b) the pipelined form of the same call
    # (render_begin / render_wait, up to three frames in flight: read-back and the next frame's kernels overlap the tail of a frame).
    # Every step of both does its own H2D camera upload and its own full-frame D2H.
    barrier()
    sync_rays = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        set_index(args.warmup + args.steps + s)
        cam.render()  # gdpt_render_frame: H2D camera, K1, K2, D2H frame; blocking
        sync_rays += cam.stats()["rays"]
    barrier()
    sync_s = time.perf_counter() - t0

    for rep in range(2):  # untimed: first use of the pipelined path (every frame slot allocates its buffers and stream once)
        for s in range(3):
            set_index(3 * rep + s)
            cam.render_begin()
        for s in range(3):
            cam.render_wait()
    barrier()
    e2e_rays, in_flight, touched = 0, 0, 0
    own_row = rank * args.band if rows_mode else 0  # a row this rank renders (row bands: rank r owns rows r*band ..)
    t0 = time.perf_counter()
    for s in range(args.steps):
        set_index(args.warmup + 2 * args.steps + s)
        cam.render_begin()
        in_flight += 1
        if in_flight == 3:
            img, fst = cam.render_wait()
            e2e_rays += fst["rays"]; touched += int(img[own_row, 0, 3]); in_flight -= 1
    while in_flight:
        img, fst = cam.render_wait()
        e2e_rays += fst["rays"]; touched += int(img[own_row, 0, 3]); in_flight -= 1
    barrier()
    e2e_s = time.perf_counter() - t0
    assert touched == 255 * args.steps, "a pipelined frame came back without pixels"

    if peer_frame is not None:
        peer_frame.close()
    # ---- reduce over ranks: slowest rank's time, everybody's rays
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, wall_ms, sync_s], dtype=torch.float64, device=dev_t)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        r = torch.tensor([rays_total, e2e_rays, sync_rays], dtype=torch.int64, device=dev_t)
        torch.distributed.all_reduce(r, op=torch.distributed.ReduceOp.SUM)
        dev_ms, e2e_s, wall_ms, sync_s = [float(x) for x in t.tolist()]
        rays_total, e2e_rays, sync_rays = [int(x) for x in r.tolist()]
    if rank != 0:
        return

    # ---- algorithmic work of one representative frame (instrumented kernels, untimed)
    work = trace_work(sc, grp, args, local)
    stage_avg = stage_ms[:n_stage] / min(args.steps, 8)
    kinds = stage_kinds(n_stage, D)
    by_kernel = {}
    for i in range(n_stage):
        k = by_kernel.setdefault(kinds[i], {"ms": 0.0, "launches": 0, "lane_ops": 0.0})
        k["ms"] += stage_avg[i]; k["launches"] += 1
        if n_stage > 2 and i % 2 == 0:
            k["lane_ops"] += work["lane_ops_per_segment"][i // 2]
    if n_stage <= 2:
        # the traversal of every segment runs inside k_path; k_primary_cull (if any) does the TLAS walk of the camera
        # rays that hit nothing.  The frame's algorithmic work is charged to the pair and timed over both.
        top = "k_path"
        # duration: CUDA events around K1 in the timed region (the split above only apportions it)
        tk = {"ms": k1_ms_total / args.steps, "launches": 1, "lane_ops": float(sum(work["lane_ops_per_segment"]))}
        top_label = "k_path" if n_stage == 1 else "k_path (+ k_primary_cull, timed together)"
    else:
        top = max((k for k in by_kernel if by_kernel[k]["lane_ops"] > 0), key=lambda k: by_kernel[k]["ms"])
        tk = by_kernel[top]
        top_label = top
    peak_lane_ops = 148 * 4 * 32 * peaks["sm_max_mhz"] * 1e6
    achieved = tk["lane_ops"] / (tk["ms"] * 1e-3) if tk["ms"] > 0 else 0.0
    k2_avg_ms = k2_ms_total / args.steps
    k2_bytes = 40.0 * W * H
    roofline = {"kernel": top_label, "bound": "issue", "achieved": achieved / 1e12, "peak": peak_lane_ops / 1e12, "unit": "Tlane-op/s",
                "frac": achieved / peak_lane_ops, "traffic": None, "launches_per_step": tk["launches"],
                "algorithmic_bytes": work["algorithmic_bytes"],
                "avg_launch_ms": tk["ms"] / tk["launches"], "share_of_step": (by_kernel[top]["ms"] / max(stage_avg.sum(), 1e-9)) * tk["ms"] / max(tk["ms"] + k2_avg_ms, 1e-9)
                if n_stage <= 2 else by_kernel[top]["ms"] / max(stage_avg.sum() + k2_avg_ms, 1e-9),
                "peak_source": f"148 SM x 4 schedulers x 32 lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)",
                "work_model": "22*box_tests + 55*tri_tests + 45*tlas_leaf_visits lane-ops per ray (SURVEY 8d) of the REFERENCE traversal "
                              "(no culling), counted by the instrumented kernels on one frame; the timed kernels skip the part of it "
                              "that tight-box culling proves fruitless"}
    for rf, kern in ((roofline, "k_path"),):
        tr = ncu_traffic(kern, args)
        if tr:
            rf["traffic"], rf["traffic_unit"], rf["traffic_source"] = tr["bytes"], "bytes of DRAM per launch", tr["source"]
    roofline_k2 = {"kernel": "k_progressive", "bound": "hbm", "achieved": k2_bytes / (k2_avg_ms * 1e-3) / 1e9 if k2_avg_ms > 0 else 0.0,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (k2_bytes / (k2_avg_ms * 1e-3) / 1e9) / peaks["hbm_gbs"] if k2_avg_ms > 0 else 0.0,
                   "traffic": None, "avg_launch_ms": k2_avg_ms, "bytes_per_pixel": 40, "peak_source": peaks["source"] + " hbm_gbs"}
    tr = ncu_traffic("k_progressive", args)
    if tr:
        roofline_k2["traffic"], roofline_k2["traffic_source"] = tr["bytes"], tr["source"]

    cpu = cpu_baseline(sc, grp, args) if (world == 1 and args.cpu_baseline) else None
    value = rays_total / (dev_ms * 1e-3) / 1e6
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if rows_mode else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "partition": (("row bands, bands written to every peer's image by K2 over NVLink + "
                   "one-element all-reduce per step; equal to the NCCL all-gather of the bands: " + str(present_check)
                   if peer_frame is not None else "row bands + all-gather per step") if rows_mode else
                   ("sample index, one accumulation sum-reduce at the end" if world > 1 else "single GPU")),
                   "rays_per_step": rays_total / args.steps / world, "frames_per_step": world if not rows_mode else 1,
                   "l2": "256 MiB write between steps (outside the per-step events)" if args.l2_flush else "no flush",
                   "timing": "CUDA events around K1 and K2 on the launching stream, summed over steps, max over ranks",
                   "wall_ms_per_step_incl_flush_and_sync": wall_ms / args.steps,
                   "stage_ms_note": "split measured on untimed frames with an event between the K1 kernels",
                   "stage_ms": {f"{i}:{kinds[i]}": round(float(stage_avg[i]), 5) for i in range(n_stage)},
                   "kernels": {k: {"ms_per_step": v["ms"], "launches": v["launches"]} for k, v in by_kernel.items()},
                   "work_per_frame": work["totals"]},
        "clocks": clocks,
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": 160 + 12, "d2h_bytes_per_step": W * H * 4,
                "api": "PathTracingCamera.render_begin()/render_wait() -> gdpt_render_frame_begin/_wait: host camera block in, "
                       "pinned host RGBA8 frame out, every step; three frames in flight on two streams",
                "blocking_render": {"value": sync_rays / sync_s / 1e6, "ms_per_step": sync_s / args.steps * 1e3,
                                    "api": "PathTracingCamera.render() -> gdpt_render_frame, one frame at a time"}},
        "gpu_launches": int(launches_per_frame * args.steps * 3),  # device-timed leg + the two end-to-end legs
        "roofline": roofline, "roofline_accumulate": roofline_k2,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), file=JSON_OUT, flush=True)


def trace_work(sc, grp, args, local):
    """Per-segment algorithmic work of one frame from the instrumented (GDPT_TRACE) kernels."""
    from gdpathtracing_b200 import PathTracingCamera
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.NONE
    cam.set_window_size(args.width, args.height)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(args.depth)
    cam.set_cuda_device(local)
    cam.set_trace(args.depth, 0)
    cam.set_frame_index(args.warmup)
    cam.init()
    cam.render_device_only()
    st = cam.stats()
    per_seg = []
    for s in range(args.depth):
        t = cam.read_trace(s)
        live = t["hit"] != 0xFFFFFFFF
        per_seg.append(float(LANE_OPS_BOX * t["box_tests"][live].astype(np.float64).sum()
                             + LANE_OPS_TRI * t["tri_tests"][live].astype(np.float64).sum()
                             + LANE_OPS_TLAS_LEAF * t["tlas_leaves"][live].astype(np.float64).sum()))
    totals = {k: int(st[k]) for k in ("rays", "primary_hits", "node_pops", "box_tests", "tri_tests", "tlas_leaves", "max_stack")}
    del cam
    # SURVEY 8d secondary figure: 144 B per internal pop (= per pair of box tests), 48 B per leaf pop, 48 B per triangle test,
    # 208 B per TLAS leaf, 324 B per shaded hit -- bytes the reference traversal of these rays asks the memory system for
    internal = totals["box_tests"] // 2
    leaves = max(totals["node_pops"] - internal, 0)
    hits = max(totals["rays"] - args.width * args.height, 0)  # every bounce ray was spawned by one shaded hit (lower bound)
    alg_bytes = 144.0 * internal + 48.0 * leaves + 48.0 * totals["tri_tests"] + 208.0 * totals["tlas_leaves"] + 324.0 * hits
    return {"lane_ops_per_segment": per_seg, "totals": totals, "algorithmic_bytes": alg_bytes}


def cpu_baseline(sc, grp, args):
    """The oracle (kind "port") on the box's host threads, bounded to ~10-30 s; BVH build timed separately."""
    from gdpathtracing_b200 import nodes
    from oracle import oracle
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    threads = oracle.hardware_threads()
    W, H = args.width, args.height
    t_total, rays, frames = 0.0, 0, 0
    while t_total < 10.0 and frames < 16:
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, args.warmup + frames + 1))
        t0 = time.perf_counter()
        r = oracle.path_trace(osc, W, H, cam, max_depth=args.depth, threads=threads)
        screen, acc = r["rgba8"], np.zeros((H, W, 4), np.float32)
        oracle.progressive(screen, acc, 1)
        t_total += time.perf_counter() - t0
        rays += r["stats"]["rays"]; frames += 1
    build_s = {}
    for th in (1, 0):  # scene build (GeometryGroup3D.build): upstream's single thread, then the multi-threaded top + subtrees
        grp.build_threads = th
        t0 = time.perf_counter()
        grp.build()
        build_s[th] = time.perf_counter() - t0
    return {"value": rays / t_total / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{frames} full {W}x{H} frames of the same workload", "ms_per_frame": t_total / frames * 1e3,
            "bvh_build_s_single_thread": build_s[1], "bvh_build_s_all_threads": build_s[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="demo", choices=["demo", "cornell32", "soup", "instanced"])
    ap.add_argument("--soup-tris", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--partition", default="sample", choices=["sample", "rows"])
    ap.add_argument("--band", type=int, default=8)
    ap.add_argument("--present", default="peer", choices=["peer", "gather"],
                    help="row bands: how the presented frame is assembled (peer writes fused into K2, or NCCL all-gather)")
    ap.add_argument("--variant", type=int, default=-1, help="A/B: kernel schedule (include/gdpt.h GDPT_VARIANT); -1 = backend default")
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=N", help="A/B: scheduling knob (GDPT_TUNE_<NAME>)")
    ap.add_argument("--no-l2-flush", dest="l2_flush", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false",
                    help="A/B runs of kernel variants only: skip the CPU leg (the official line always carries it)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # The contract is ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version there, build tools
    # their command lines) are sent to stderr for the whole run; the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global JSON_OUT
    JSON_OUT = os.fdopen(json_fd, "w")

    import __graft_entry__ as entry
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            entry.build()
            run_reference(args, rank, world)
        return
    if world > 1:
        # process group first (plain torch), so the other ranks can wait for rank 0's build
        import torch
        torch.cuda.set_device(local)
        torch.distributed.init_process_group(backend="nccl", rank=rank, world_size=world)
    if rank == 0:
        entry.build()
    if world > 1:
        torch.distributed.barrier()
    run_ours(args, rank, local, world)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
