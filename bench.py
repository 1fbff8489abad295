#!/usr/bin/env python3
"""bench.py -- BASELINE.json metric: Mrays/s and ms/frame @1080p, 1 spp/frame, depth 8 on the
project demo scene (config C2), N B200s of one node, next to the host-CPU reference.

A "step" is one frame: camera upload, K1 (ray generation .. bounce loop: a camera-ray
classification kernel and a persistent path kernel) and K2 (progressive accumulation + ACES).  A "ray" is one ray_trace() call of the
reference (main.glsl:352): a primary ray or one bounce segment; counted exactly on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference ...                           the reference arm: the CPU
        reference's own shader text compiled as C++ (oracle/_ref, kind "reference"; our restatement of it,
        kind "port", where that library is absent) on all host threads

Prints ONE JSON line (rank 0).  Multi-GPU (torchrun): sample-index partition, each rank renders
frame_index = step*N + rank + 1 of the same scene; the row blocks of those frames are exchanged (NCCL
over NVLink) and accumulated in frame order by their owners inside the timed region (weak scaling;
`--partition rows` = row-band strong scaling, presentation fused into K2 as peer writes).  The line also
carries `extra.c5`: BASELINE config C5 (4K instanced scene accumulated to 256 spp, presented every 64)
in both partitions.
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
def reference_impl():
    """What plays the reference on the host: its own shader text compiled as C++ (oracle/_ref/libgdpt_refshader.so,
    built by oracle/Makefile where /root/reference exists and shipped with the snapshot) -> kind "reference";
    our C++ restatement of it (oracle/pt_oracle.cpp) where that library is missing -> kind "port"."""
    from oracle import oracle
    return ("reference", "reference") if oracle.ref_shader_available() else ("restatement", "port")


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
    impl, kind = reference_impl()

    def frame(idx, row_step, which=impl):
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, idx))
        t0 = time.perf_counter()
        r = oracle.path_trace(osc, W, H, cam, max_depth=args.depth, threads=threads, row_step=row_step, impl=which, counters=False)
        screen, acc = r["rgba8"], np.zeros((H, W, 4), np.float32)
        if row_step == 1:
            oracle.progressive(screen, acc, 1, impl=which)
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
    port_mrays = None
    if impl == "reference":  # the restatement beside it, on two frames of the same sample
        pt, pr = 0.0, 0
        for s in range(2):
            t, rays = frame(args.warmup + s + 1, row_step, "restatement")
            pt += t; pr += rays
        port_mrays = pr / pt / 1e6
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
        "config": {"workload": workload_name(args), "partition": "host CPU, all threads"},
        "note": ("the reference's own shader text (main.glsl + brdfs.glsl + progressive_rendering.glsl) compiled as C++ and run on the "
                 "host cores" if kind == "reference" else "C++ restatement of the reference's shaders (the compiled shader text is "
                 "not on this box)") + "; its GPU half cannot run without Godot + Vulkan; ms_per_step is scaled to a full frame",
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample,
                         "restatement_value": port_mrays, "scene_build_s": build_s, "reference_bvh_build_s": ref_build},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------ our arm
IN_FLIGHT = 4  # GDPT_MAX_FRAMES_IN_FLIGHT (include/gdpt.h): frames the end-to-end leg keeps in flight


SCHEDULES = {
    2: ("k_path<all pixels>", "the reference's visiting order on the reference arrays, no culling (trace / DEBUG_STEPS mode)"),
    3: ("k_primary_cull + k_path<survivors>", "the reference's visiting order on the reference arrays, tight-box culling"),
    6: ("k_primary_cull + k_path_pool", "closest-hit search over our own four-wide SAH tables + proof that the reference traversal "
        "returns the same record; paths pooled per warp"),
}


def make_camera(sc, grp, args, local, mode, W=None, H=None, depth=None, variant=None, shard=None, count_work=False, trace=0):
    from gdpathtracing_b200 import PathTracingCamera
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = mode
    cam.set_window_size(W or args.width, H or args.height)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth or args.depth)
    cam.set_cuda_device(local)
    if shard:
        cam.set_shard(*shard)
    v = args.variant if variant is None else variant
    if v >= 0:
        cam.set_variant(v)
    for kv in args.tune:
        k, val = kv.split("=")
        cam.set_tuning(k, int(val))
    if count_work:
        cam.set_count_work(True)
    if trace:
        cam.set_trace(trace, 0)
    cam.init()
    return cam


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
    sample_mode = world > 1 and not rows_mode

    # N = 1: the reference's own sequence, K1 + K2 per frame.  Row bands: the same on this rank's rows.  Sample index:
    # K1 per frame; the accumulation happens after the exchange (multigpu.SampleIndexAccumulator), inside the timed region.
    cam = make_camera(sc, grp, args, local, PathTracingCamera.NONE if sample_mode else PathTracingCamera.PROGRESSIVE_RENDERING,
                      shard=(rank, world, args.band) if rows_mode else None)
    backend_stream = torch.cuda.ExternalStream(cam.stream())

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
    # NVLink, frames separated by a two-step handshake) unless --present gather asks for the NCCL all-gather
    peer_frame = multigpu.PeerFrame(cam, rank, world) if (rows_mode and args.present == "peer") else None

    def present():
        if peer_frame is not None:
            peer_frame.barrier()         # frame N is complete in every rank's image
            peer_frame.frame_consumed()  # ... and nobody overwrites it before everybody is done with it
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
    # sample index: every rank keeps its frames in device memory the other ranks can read over NVLink (CUDA IPC); the owner
    # of a row block accumulates straight from the source rank's copy (multigpu.PeerFrameStore: the exchange is K2's loads)
    store = multigpu.PeerFrameStore(cam, args.steps, H, W, rank, world, dev_t) if sample_mode else None
    frames_kept = store.frames if sample_mode else None
    if sample_mode:
        for _ in range(2):  # warm-up of the path at the size the timed region uses (peer mappings, NCCL channels, allocator blocks)
            warm = multigpu.SampleIndexAccumulator(H, W, rank, world, dev_t, multigpu.cuda_k2_at(cam))
            with torch.cuda.stream(backend_stream):
                warm.add_from_store(store)
                warm.present()
            torch.cuda.synchronize()
            del warm
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
    dev_ms, rays_total, k2_ms_total, gather_ms, k1_ms_total, retraced = 0.0, 0, 0.0, 0.0, 0.0, 0
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
        retraced += st["retraced"]
        launches_per_frame = st["kernel_launches"] + (0 if sample_mode else 1)  # K1 kernels + K2
        if sample_mode:
            with torch.cuda.stream(backend_stream):
                frames_kept[s].copy_(frame_t)  # this rank's frame of the step, kept for the accumulation below (8 MB, device to device)
        if rows_mode:
            ts = backend_stream if peer_frame is not None else torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts)
            present()
            e1.record(ts); torch.cuda.synchronize(); cam.synchronize()
            gather_ms += e0.elapsed_time(e1)
    exchange = None
    if sample_mode:
        # the partition's exchange step, inside the timed region: every owner runs K2 over its row block of every rank's
        # frames, in frame order, reading the block from the source rank's memory over NVLink; two one-element all-reduces
        # fence the reads; all-gather of the presented blocks
        acc = multigpu.SampleIndexAccumulator(H, W, rank, world, dev_t, multigpu.cuda_k2_at(cam))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(backend_stream):
            e0.record(backend_stream)
            acc.add_from_store(store)
            presented = acc.present()
            e1.record(backend_stream)
        torch.cuda.synchronize(); cam.synchronize()
        gather_ms += e0.elapsed_time(e1)
        launches_per_frame += world  # K2 launches of the accumulation, per step
        exchange = {"bytes_read_from_peers_per_rank": int(acc.bytes_exchanged), "frames_accumulated": int(acc.frames_done),
                    "what": "K2 per frame in frame order on the owner of each row block, reading the block from the source rank's "
                            "memory over NVLink (CUDA IPC) + two one-element all-reduces + all-gather of the presented blocks"}
        del presented
        frames_kept = None
        store.close()
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
    # (render_begin / render_wait, up to four frames in flight: read-back and the next frames' kernels overlap the tail of a frame).
    # Every step of both does its own H2D camera upload and its own full-frame D2H.
    if peer_frame is not None:
        peer_frame.close()  # the end-to-end legs run without the frame handshake: no peer may write this rank's image in them
        peer_frame = None
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
        for s in range(IN_FLIGHT):
            set_index(IN_FLIGHT * rep + s)
            cam.render_begin()
        for s in range(IN_FLIGHT):
            cam.render_wait()
    barrier()
    e2e_rays, in_flight, touched = 0, 0, 0
    own_row = rank * args.band if rows_mode else 0  # a row this rank renders (row bands: rank r owns rows r*band ..)
    t0 = time.perf_counter()
    for s in range(args.steps):
        set_index(args.warmup + 2 * args.steps + s)
        cam.render_begin()
        in_flight += 1
        if in_flight == IN_FLIGHT:
            img, fst = cam.render_wait()
            e2e_rays += fst["rays"]; touched += int(img[own_row, 0, 3]); in_flight -= 1
    while in_flight:
        img, fst = cam.render_wait()
        e2e_rays += fst["rays"]; touched += int(img[own_row, 0, 3]); in_flight -= 1
    barrier()
    e2e_s = time.perf_counter() - t0
    assert touched == 255 * args.steps, "a pipelined frame came back without pixels"
    schedule = int(cuda.gdpt_shader_get_schedule(cam.main_shader))

    # ---- reduce over ranks: slowest rank's time, everybody's rays
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, wall_ms, sync_s, gather_ms], dtype=torch.float64, device=dev_t)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        r = torch.tensor([rays_total, e2e_rays, sync_rays, retraced], dtype=torch.int64, device=dev_t)
        torch.distributed.all_reduce(r, op=torch.distributed.ReduceOp.SUM)
        dev_ms, e2e_s, wall_ms, sync_s, gather_ms = [float(x) for x in t.tolist()]
        rays_total, e2e_rays, sync_rays, retraced = [int(x) for x in r.tolist()]

    # ---- work of one representative frame of THIS rank's share: what the timed kernel executes itself (its counting
    # instantiation, untimed) and what the reference traversal would execute on the same rays (trace mode, untimed)
    shard = (rank, world, args.band) if rows_mode else None
    own = own_work(sc, grp, args, local, shard)
    ref_work = trace_work(sc, grp, args, local, shard)
    # ---- the reference's visiting order (schedule 3) on the same frames, beside the headline
    s3 = None
    if rank == 0 and args.schedule3 and schedule != 3:
        c3 = make_camera(sc, grp, args, local, PathTracingCamera.PROGRESSIVE_RENDERING, variant=3, shard=shard)
        t3, r3 = 0.0, 0
        for s in range(3 + min(args.steps, 8)):
            c3.set_frame_index(args.warmup + s - 3 if s >= 3 else s)
            c3.render_device_only()
            st = c3.stats()
            if s >= 3:
                t3 += st["k1_ms"] + st["k2_ms"]; r3 += st["rays"]
        s3 = {"value": r3 / (t3 * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": t3 / min(args.steps, 8),
              "kernels": SCHEDULES[3][0], "traversal": SCHEDULES[3][1]}
        del c3
    c5 = run_c5(args, rank, local, world, dev_t) if args.c5 else None
    if rank != 0:
        return

    stage_avg = stage_ms[:n_stage] / split_frames
    kinds = stage_kinds(n_stage)
    k1_avg = k1_ms_total / args.steps
    # duration of the path kernel: CUDA events around K1 in the timed region, apportioned by the untimed stage split
    path_share = float(stage_avg[-1] / max(stage_avg.sum(), 1e-9)) if n_stage else 1.0
    path_ms = k1_avg * path_share
    k2_avg_ms = k2_ms_total / args.steps
    peak_lane_ops = 148 * 4 * 32 * peaks["sm_max_mhz"] * 1e6
    own_lane_ops = LANE_OPS_BOX * own["box_tests"] + LANE_OPS_TRI * own["tri_tests"] + LANE_OPS_TLAS_LEAF * own["inst_entries"]
    ref_lane_ops = float(sum(ref_work["lane_ops_per_segment"]))
    achieved = own_lane_ops / (path_ms * 1e-3) if path_ms > 0 else 0.0
    step_ms = dev_ms / args.steps
    roofline = {"kernel": kinds[-1], "bound": "issue", "achieved": achieved / 1e12, "peak": peak_lane_ops / 1e12, "unit": "Tlane-op/s",
                "frac": achieved / peak_lane_ops, "traffic": None, "launches_per_step": 1,
                "avg_launch_ms": path_ms, "share_of_step": path_ms / max(step_ms, 1e-9),
                "own_work_per_launch": own,
                "reference_work_ratio": ref_lane_ops / own_lane_ops if own_lane_ops > 0 else None,
                "algorithmic_bytes": own["algorithmic_bytes"],
                "peak_source": f"148 SM x 4 schedulers x 32 lanes x {peaks['sm_max_mhz']:.0f} MHz ({peaks['source']} sm_max_mhz)",
                "work_model": "lane-ops the TIMED kernel executes itself, counted by its own counting instantiation on one frame of this "
                              "rank's share: 22 per slab test (4 per four-wide node step) + 55 per triangle test + 45 per instance entry "
                              "(unit costs of SURVEY 8d); reference_work_ratio = the same formula over the reference traversal of the "
                              "same rays (trace mode) divided by it"}
    tr = ncu_traffic(kinds[-1].split("<")[0], args)
    if tr:
        roofline["traffic"], roofline["traffic_unit"], roofline["traffic_source"] = tr["bytes"], "bytes of DRAM per launch", tr["source"]
    k2_bytes = 36.0 * W * H * (ref_work["local_rows"] / H)
    roofline_k2 = None
    if k2_avg_ms > 0:
        roofline_k2 = {"kernel": "k_progressive", "bound": "hbm", "achieved": k2_bytes / (k2_avg_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                       "unit": "GB/s", "frac": (k2_bytes / (k2_avg_ms * 1e-3) / 1e9) / peaks["hbm_gbs"], "traffic": None,
                       "avg_launch_ms": k2_avg_ms, "bytes_per_pixel": 36, "peak_source": peaks["source"] + " hbm_gbs",
                       "work_model": "36 B per pixel (SURVEY 8d): 16 R + 16 W accumulation, 4 W screen; the 4 B raw read rides on top"}
        tr = ncu_traffic("k_progressive", args)
        if tr:
            roofline_k2["traffic"], roofline_k2["traffic_source"] = tr["bytes"], tr["source"]

    cpu = cpu_baseline(sc, grp, args) if (world == 1 and args.cpu_baseline) else None
    value = rays_total / (dev_ms * 1e-3) / 1e6
    path_rays = own["proofs"]  # every ray the path kernel answers ends in one verdict
    if rows_mode:
        partition = ("row bands; K2 writes its rows into every peer's image over NVLink, frames separated by two one-element "
                     "all-reduces; equal to the NCCL all-gather of the bands: " + str(present_check)) if args.present == "peer" \
            else "row bands + NCCL all-gather per step"
    elif sample_mode:
        partition = ("sample index: every rank renders its own frames; row blocks exchanged (NCCL send/recv over NVLink) and accumulated "
                     "in frame order by their owners, presented frame all-gathered -- inside the timed region, bit-identical to one GPU")
    else:
        partition = "single GPU"
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if rows_mode else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "partition": partition,
                   "schedule": schedule, "kernels": SCHEDULES[schedule][0], "traversal": SCHEDULES[schedule][1]
                   + f"; {retraced / args.steps / world:.1f} of {rays_total / args.steps / world:.0f} rays per frame re-traced in reference order",
                   "tuning": args.tune, "rays_per_step": rays_total / args.steps / world, "frames_per_step": world if not rows_mode else 1,
                   "rays_in_path_kernel_per_frame": path_rays,
                   "mrays_s_over_rays_that_enter_an_instance": path_rays / (path_ms * 1e-3) / 1e6 if path_ms > 0 else None,
                   "gather_ms": gather_ms, "exchange": exchange,
                   "reference_visiting_order": s3,
                   "l2": "256 MiB write between steps (outside the per-step events)" if args.l2_flush else "no flush",
                   "timing": "CUDA events around K1 and K2 on the launching stream, summed over steps, max over ranks; gather_ms (events "
                             "on the same stream) is part of it",
                   "wall_ms_per_step_incl_flush_and_sync": wall_ms / args.steps,
                   "stage_ms_note": "split measured on untimed frames with an event between the K1 kernels",
                   "stage_ms": {f"{i}:{kinds[i]}": round(float(stage_avg[i]), 5) for i in range(n_stage)},
                   "k2_ms": k2_avg_ms,
                   "reference_work_per_frame": ref_work["totals"]},
        "clocks": clocks,
        "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_s / args.steps * 1e3,
                "h2d_bytes_per_step": 160 + 12, "d2h_bytes_per_step": W * H * 4,
                "api": "PathTracingCamera.render_begin()/render_wait() -> gdpt_render_frame_begin/_wait: host camera block in, "
                       "pinned host RGBA8 frame out, every step; four frames in flight on four streams",
                "blocking_render": {"value": sync_rays / sync_s / 1e6, "ms_per_step": sync_s / args.steps * 1e3,
                                    "api": "PathTracingCamera.render() -> gdpt_render_frame, one frame at a time"}},
        "gpu_launches": int(launches_per_frame * args.steps * 3),  # device-timed leg + the two end-to-end legs
        "roofline": roofline, "roofline_accumulate": roofline_k2,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    if c5:
        line["extra"] = {"c5": c5}
    print(json.dumps(line), file=JSON_OUT, flush=True)


def restart_accumulation(cam, sc):
    """One frame from a moved camera: the next frame at the scene's pose is frame_count 1 again
    (progressive_rendering.cpp:53-60 resets the count when the camera transform changes)."""
    moved = np.array(sc.camera_transform12, np.float32).copy()
    moved[9] += 1.0
    cam.set_global_transform(moved)
    cam.render_device_only()
    cam.synchronize()
    cam.set_global_transform(sc.camera_transform12)


def stage_kinds(n):
    return ["k_path"] if n == 1 else ["k_primary_cull", "k_path"]


def own_work(sc, grp, args, local, shard):
    """What the timed path kernel executes on one frame: its counting instantiation (GDPT_COUNT_WORK), untimed."""
    from gdpathtracing_b200 import PathTracingCamera
    cam = make_camera(sc, grp, args, local, PathTracingCamera.NONE, shard=shard, count_work=True)
    cam.set_frame_index(args.warmup)
    cam.render_device_only()
    st = cam.stats()
    del cam
    own = {"node_steps": int(st["own_node_steps"]), "box_tests": int(st["own_box_tests"]), "tri_tests": int(st["own_tri_tests"]),
           "inst_entries": int(st["own_inst_entries"]), "proofs": int(st["own_proofs"]), "retraced": int(st["retraced"])}
    # bytes the search asks the memory system for: 128 B per four-wide node, 48 B per triangle, 112 B per instance record,
    # 324 B per shaded hit (triangle data 80 + instance 176 + material 64 + texel 4)
    own["algorithmic_bytes"] = 128.0 * own["node_steps"] + 48.0 * own["tri_tests"] + 112.0 * own["inst_entries"] + 324.0 * max(st["rays"] - own["proofs"], 0)
    return own


def trace_work(sc, grp, args, local, shard):
    """Per-segment algorithmic work of the REFERENCE traversal of one frame, from the instrumented (GDPT_TRACE) kernels."""
    from gdpathtracing_b200 import PathTracingCamera
    cam = make_camera(sc, grp, args, local, PathTracingCamera.NONE, shard=shard, trace=args.depth, variant=-1)
    cam.set_frame_index(args.warmup)
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
    local_rows = args.height if not shard else int(sum(1 for y in range(args.height) if (y // shard[2]) % shard[1] == shard[0]))
    del cam
    return {"lane_ops_per_segment": per_seg, "totals": totals, "local_rows": local_rows}


def run_c5(args, rank, local, world, dev_t):
    """BASELINE config C5: the C4 scene (1 000 instances of a 10 000-triangle BLAS) at 3840x2160, depth 8, accumulated to
    `--c5-spp` samples per pixel, a presented frame every `--c5-present` samples INSIDE the timed region, in both
    partitions.  N = 1 is the single-GPU run of the same code; the driver's scaling run gives the ratio."""
    import torch
    from gdpathtracing_b200 import PathTracingCamera, multigpu, scenes
    W, H, D = 3840, 2160, 8
    spp, every = args.c5_spp, args.c5_present
    assert spp % world == 0 and every % world == 0 and spp % every == 0
    sc = scenes.instanced_grid()
    grp = scenes.populate(sc)
    out = {"workload": f"C5: C4 scene (10M instanced triangles) {W}x{H}, depth {D}, {spp} spp, presented every {every} spp",
           "spp": spp, "present_every": every}

    def timed(fn):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return (time.perf_counter() - t0) * 1e3, res

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev_t)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.int64, device=dev_t)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
        return int(t.item())

    first_presented = {}
    # ---------------- sample index
    cam = make_camera(sc, grp, args, local, PathTracingCamera.NONE, W=W, H=H, depth=D)
    stream = torch.cuda.ExternalStream(cam.stream())
    ptr, _ = cam.device_pointer("output")
    frame_t = multigpu.as_tensor(ptr, (H, W, 4), torch.uint8, dev_t)
    batch = every // world
    store = multigpu.PeerFrameStore(cam, batch, H, W, rank, world, dev_t)  # this rank's frames of a batch, readable by every rank
    kept = store.frames
    for w in range(3):  # warm-up: kernels, exchange, K2
        cam.set_frame_index(w)
        cam.render_device_only()
    cam.synchronize()
    for _ in range(2):  # at the size the timed region uses: peer mappings, NCCL channels and the allocator's blocks are set up on first use
        warm = multigpu.SampleIndexAccumulator(H, W, rank, world, dev_t, multigpu.cuda_k2_at(cam))
        with torch.cuda.stream(stream):
            warm.add_from_store(store); warm.present()
        torch.cuda.synchronize()
        del warm

    def sample_index():
        acc = multigpu.SampleIndexAccumulator(H, W, rank, world, dev_t, multigpu.cuda_k2_at(cam))
        rays, k1_ms, ex_ms = 0, 0.0, 0.0
        for b in range(spp // every):
            for j in range(batch):
                cam.set_frame_index((b * batch + j) * world + rank)  # render() increments: global index (b*batch+j)*world + rank + 1
                cam.render_device_only()
                st = cam.stats()
                rays += st["rays"]; k1_ms += st["k1_ms"]
                with torch.cuda.stream(stream):
                    kept[j].copy_(frame_t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                acc.add_from_store(store)
                shown = acc.present()
                e1.record(stream)
            torch.cuda.synchronize()
            ex_ms += e0.elapsed_time(e1)
            if b == 0:
                first_presented["sample_index"] = shown.clone()
        return rays, k1_ms, ex_ms, acc.bytes_exchanged

    ms, (rays, k1_ms, ex_ms, sent) = timed(sample_index)
    rays_all = reduce_sum(rays)
    out["sample_index"] = {"ms_total": ms, "ms_per_spp": ms / spp, "mrays_s": rays_all / (ms * 1e-3) / 1e6,
                           "k1_ms_slowest_rank": reduce_max(k1_ms), "exchange_accumulate_present_ms_slowest_rank": reduce_max(ex_ms),
                           "bytes_read_over_nvlink_per_rank": int(sent),
                           "collective": "none on the data path: K2 reads each frame's row block from the source rank's memory over NVLink "
                                         "(CUDA IPC); two one-element NCCL all-reduces per batch + all-gather of the presented blocks"}
    del kept
    store.close()
    del cam

    # ---------------- row bands
    cam = make_camera(sc, grp, args, local, PathTracingCamera.PROGRESSIVE_RENDERING, W=W, H=H, depth=D, shard=(rank, world, args.band))
    ptr, _ = cam.device_pointer("output")
    frame_t = multigpu.as_tensor(ptr, (H, W, 4), torch.uint8, dev_t)
    peer = multigpu.PeerFrame(cam, rank, world)
    stream = torch.cuda.ExternalStream(cam.stream())
    for w in range(3):
        cam.set_frame_index(w)
        cam.render_device_only()
        cam.synchronize()
        peer.barrier(); peer.frame_consumed()
    restart_accumulation(cam, sc)
    peer.barrier(); peer.frame_consumed()
    torch.cuda.synchronize()

    def rows():
        rays, k_ms, hs_ms = 0, 0.0, 0.0
        for f in range(spp):
            cam.set_frame_index(f)
            cam.render_device_only()  # frame_count = f + 1 follows the host policy (camera at rest)
            st = cam.stats()
            rays += st["rays"]; k_ms += st["k1_ms"] + st["k2_ms"]
            if (f + 1) % every == 0:  # a presented frame: every rank's image is complete after the handshake
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                peer.barrier()
                if f + 1 == every:
                    with torch.cuda.stream(stream):
                        first_presented["rows"] = frame_t.clone()
                peer.frame_consumed()
                e1.record(stream)
                torch.cuda.synchronize(); cam.synchronize()
                hs_ms += e0.elapsed_time(e1)
        return rays, k_ms, hs_ms

    ms, (rays, k_ms, hs_ms) = timed(rows)
    rays_all = reduce_sum(rays)
    out["rows"] = {"ms_total": ms, "ms_per_spp": ms / spp, "mrays_s": rays_all / (ms * 1e-3) / 1e6, "band_rows": args.band,
                   "k1_k2_ms_slowest_rank": reduce_max(k_ms), "handshake_ms_slowest_rank": reduce_max(hs_ms),
                   "bytes_written_to_peers_per_rank_per_frame": int(W * H * 4 // world * (world - 1)),
                   "collective": "none on the data path: K2 stores its rows into every peer's image over NVLink (CUDA IPC); two "
                                 "one-element NCCL all-reduces per presented frame"}
    peer.close()
    del cam

    # ---------------- equality with ONE GPU accumulating the same frames sequentially (rank 0, untimed): the first presented frame
    equal = {}
    if rank == 0:
        ref = make_camera(sc, grp, args, local, PathTracingCamera.PROGRESSIVE_RENDERING, W=W, H=H, depth=D)
        t0 = time.perf_counter()
        for f in range(every):
            ref.set_frame_index(f)
            ref.render_device_only()
        ref.synchronize()
        single_ms = (time.perf_counter() - t0) * 1e3 / every
        rptr, _ = ref.device_pointer("output")
        want = multigpu.as_tensor(rptr, (H, W, 4), torch.uint8, dev_t)
        for k, img in first_presented.items():
            equal[k] = bool(torch.equal(img, want))
        out["single_gpu_sequential"] = {"ms_per_spp": single_ms, "frames": every}
        del ref
    if world > 1:
        torch.distributed.barrier()
    out["first_presented_frame_equals_single_gpu"] = equal
    return out


def cpu_baseline(sc, grp, args):
    """The reference on the box's host threads, bounded to ~10-30 s: its own shader text compiled as C++ where that
    library travelled with the snapshot (kind "reference"), else our restatement (kind "port"); scene build timed apart."""
    from gdpathtracing_b200 import nodes
    from oracle import oracle
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    threads = oracle.hardware_threads()
    impl, kind = reference_impl()
    W, H = args.width, args.height

    def run(which, budget_s, most):
        t_total, rays, frames = 0.0, 0, 0
        while t_total < budget_s and frames < most:
            cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, args.warmup + frames + 1))
            t0 = time.perf_counter()
            r = oracle.path_trace(osc, W, H, cam, max_depth=args.depth, threads=threads, impl=which, counters=False)
            screen, acc = r["rgba8"], np.zeros((H, W, 4), np.float32)
            oracle.progressive(screen, acc, 1, impl=which)
            t_total += time.perf_counter() - t0
            rays += r["stats"]["rays"]; frames += 1
        return rays / t_total / 1e6, t_total / frames * 1e3, frames

    mrays, ms_frame, frames = run(impl, 10.0, 16)
    port = run("restatement", 3.0, 4)[0] if impl == "reference" else None
    saved = grp.build_threads
    build = {}
    for th in (1, 0):  # GeometryGroup3D.build: material conversion + texture resize + BLAS/TLAS build; upstream's single thread, then all
        grp.build_threads = th
        t0 = time.perf_counter()
        grp.build()
        build[th] = time.perf_counter() - t0
    grp.build_threads = saved
    ref_bvh = None
    if oracle.ref_available():  # the reference's own bvh.cpp on the same meshes (BLAS + TLAS only: compare with nothing above)
        t0 = time.perf_counter()
        h, _ = oracle.reference_build(sc)
        oracle.ref().refbvh_free(h)
        ref_bvh = time.perf_counter() - t0
    return {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": kind,
            "sample": f"{frames} full {W}x{H} frames of the same workload", "ms_per_frame": ms_frame, "restatement_value": port,
            "scene_build_s": {"what": "GeometryGroup3D.build of our host layer: material conversion, texture resize to the array "
                                      "resolution, BLAS + TLAS build", "one_thread": build[1], "all_threads": build[0]},
            "reference_bvh_cpp_build_s": ref_bvh}


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
    ap.add_argument("--no-c5", dest="c5", action="store_false", help="skip the C5 block (4K instanced, accumulate to --c5-spp)")
    ap.add_argument("--c5-spp", type=int, default=256)
    ap.add_argument("--c5-present", type=int, default=64)
    ap.add_argument("--no-schedule3", dest="schedule3", action="store_false", help="skip the reference-visiting-order comparison value")
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
