"""GPU tier: the CUDA kernels, driven through the C-ABI (libgdpt_cuda.so) by the host twins,
against the oracle.  Bit-exact for hit ids, node-visit order, t/u/v, work counters, RGBA8
frames, depth and the accumulation buffer (north_star asks for t within 1e-5 relative; we hold
bit equality, which implies it)."""
import ctypes

import numpy as np
import pytest

from conftest import records_equal
from gdpathtracing_b200 import PathTracingCamera, _lib, nodes, scenes
from oracle import oracle

pytestmark = pytest.mark.gpu


def make_camera(sc, grp, W, H, depth, mode=PathTracingCamera.NONE, trace=0, visits=0, frame_index=0, fused=True,
                debug=False, shard=None):
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = mode
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth)
    cam.set_frame_index(frame_index)
    cam.set_fused_frame(fused)
    if trace:
        cam.set_trace(trace, visits)
    if debug:
        cam.set_debug_steps(True)
    if shard:
        cam.set_shard(*shard)
    cam.init()
    return cam


def oracle_scene(grp):
    return oracle.Scene(grp.buffers(), grp.texture_layers())


CASES = [
    ("cornell32", lambda: scenes.cornell32(), 256, 256, 4, 4),            # BASELINE config C1 at full size
    ("demo", lambda: scenes.demo_scene(), 480, 270, 8, 4),
    ("soup", lambda: scenes.triangle_soup(20000, seed=1), 160, 90, 2, 2),
    ("instanced", lambda: scenes.instanced_grid(3, 800, seed=3), 192, 108, 5, 3),
]


@pytest.mark.parametrize("name,make,W,H,depth,segs", CASES, ids=[c[0] for c in CASES])
def test_trace_parity(name, make, W, H, depth, segs):
    """Hit primitive ids, instance ids, t/u/v bits, visit-order hash + the first 48 visited
    node ids, per-ray work counters, ray count, frame and depth: all identical to the oracle."""
    sc = make()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, W, H, depth, trace=segs, visits=48)
    frame = cam.render().copy()
    stats = cam.stats()
    ref = oracle.path_trace(oracle_scene(grp), W, H, bytes(cam.camera_block()), max_depth=depth, trace_segments=segs,
                            visits_per_ray=48)
    assert ref["stats"]["primary_hits"] > 0
    assert stats["rays"] == ref["stats"]["rays"]
    assert stats["primary_hits"] == ref["stats"]["primary_hits"]
    for k in ("node_pops", "box_tests", "tri_tests", "tlas_leaves", "max_stack"):
        assert stats[k] == ref["stats"][k], k
    for s in range(segs):
        ok, why = records_equal(cam.read_trace(s), ref["trace"][s])
        assert ok, f"segment {s}: {why}"
    assert np.array_equal(cam.read_visits(), ref["visits"]), "node-visit order differs"
    assert np.array_equal(frame, ref["rgba8"])
    assert np.array_equal(cam.read_image("depth").view(np.uint32), ref["depth"].view(np.uint32))


@pytest.mark.parametrize("name,make,W,H,depth,segs", CASES, ids=[c[0] for c in CASES])
def test_culled_traversal_gives_identical_hits(name, make, W, H, depth, segs):
    """Tight-box culling (the default when rendering) only shortens the visit list: per-segment hit
    records (hit, triangle, instance, front, t/u/v bits), frame and depth equal the oracle's full traversal."""
    sc = make()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, W, H, depth, trace=segs)
    cam2 = PathTracingCamera()  # same, but instrumented kernels with culling forced on
    cam2.fov = sc.fov
    cam2.geometry_group = grp
    cam2.denoising_mode = PathTracingCamera.NONE
    cam2.set_window_size(W, H)
    cam2.set_global_transform(sc.camera_transform12)
    cam2.set_max_depth(depth)
    cam2.set_trace(segs, 0)
    cam2.set_cull(1)
    cam2.init()
    frame = cam2.render().copy()
    st = cam2.stats()
    ref = oracle.path_trace(oracle_scene(grp), W, H, bytes(cam2.camera_block()), max_depth=depth, trace_segments=segs)
    assert st["rays"] == ref["stats"]["rays"] and st["primary_hits"] == ref["stats"]["primary_hits"]
    assert st["node_pops"] <= ref["stats"]["node_pops"] and st["tri_tests"] <= ref["stats"]["tri_tests"]
    for s in range(segs):
        a, b = cam2.read_trace(s), ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"])
        live = b["hit"] != 0xFFFFFFFF
        for f in ("triangle", "blas", "front", "t", "u", "v"):
            x, y = a[f][live], b[f][live]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            assert np.array_equal(x, y), f"segment {s} field {f}"
    assert np.array_equal(frame, ref["rgba8"])
    assert np.array_equal(cam2.read_image("depth").view(np.uint32), ref["depth"].view(np.uint32))
    del cam


@pytest.mark.parametrize("variant", [6, 3], ids=["closest_hit_pooled_paths", "reference_order_culled"])
@pytest.mark.parametrize("name,make,W,H,depth,segs", CASES, ids=[c[0] for c in CASES])
def test_rendering_kernels_record_the_reference_hits(name, make, W, H, depth, segs, variant):
    """The kernels that are timed (schedule 6: closest-hit search + proof + exact re-trace, pt_fast.cuh, paths pooled in
    shared memory; schedule 3: culled reference order) write their own hit records: hit, triangle, instance, front and the
    t/u/v bits of every segment equal the oracle's full reference traversal, as do frame, depth and ray count."""
    sc = make()
    grp = scenes.populate(sc)
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.NONE
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth)
    cam.set_record_hits(depth)
    cam.set_variant(variant)
    cam.init()
    frame = cam.render().copy()
    st = cam.stats()
    ref = oracle.path_trace(oracle_scene(grp), W, H, bytes(cam.camera_block()), max_depth=depth, trace_segments=depth)
    assert st["rays"] == ref["stats"]["rays"] and st["primary_hits"] == ref["stats"]["primary_hits"]
    for s in range(depth):
        a, b = cam.read_trace(s), ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"]), f"segment {s}: hit flags"
        live = b["hit"] != 0xFFFFFFFF
        for f in ("triangle", "blas", "front", "t", "u", "v"):
            x, y = a[f][live], b[f][live]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            assert np.array_equal(x, y), f"segment {s} field {f}: {(x != y).sum()} differ"
    assert np.array_equal(frame, ref["rgba8"])
    assert np.array_equal(cam.read_image("depth").view(np.uint32), ref["depth"].view(np.uint32))
    if variant == 6:
        assert st["retraced"] * 20 < st["rays"], "the proof should fail for a small minority of rays only"
        print(f"{name}: {st['retraced']} of {st['rays']} rays re-traced in reference order")
    else:
        assert st["retraced"] == 0


@pytest.mark.parametrize("variant", [6, 3], ids=["pooled_paths", "reference_order"])
@pytest.mark.parametrize("W,H,depth", [(8, 4, 3), (37, 19, 1), (5, 3, 8), (64, 36, 32)], ids=["one_tile", "ragged_depth1", "sub_tile", "max_depth32"])
def test_rendering_kernels_on_edge_sizes(W, H, depth, variant):
    """Fewer pixels than one warp's pool, image sizes that are not multiples of the 8x4 tile, a single segment and the
    backend's maximum depth: frame, depth image, ray count and every recorded hit equal the oracle's."""
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.NONE
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth)
    segs = min(depth, 4)
    cam.set_record_hits(segs)
    cam.set_variant(variant)
    cam.init()
    for _ in range(2):  # the second frame runs with cost classes from the first
        frame = cam.render().copy()
        st = cam.stats()
        ref = oracle.path_trace(oracle_scene(grp), W, H, bytes(cam.camera_block()), max_depth=depth, trace_segments=segs)
        assert st["rays"] == ref["stats"]["rays"]
        for s in range(segs):
            a, b = cam.read_trace(s), ref["trace"][s]
            assert np.array_equal(a["hit"], b["hit"])
            live = b["hit"] != 0xFFFFFFFF
            for f in ("triangle", "blas", "front", "t", "u", "v"):
                assert np.array_equal(a[f][live].view(np.uint32), b[f][live].view(np.uint32)), f"segment {s} field {f}"
        assert np.array_equal(frame, ref["rgba8"])
        assert np.array_equal(cam.read_image("depth").view(np.uint32), ref["depth"].view(np.uint32))


def test_fast_kernels_equal_traced_kernels():
    """The un-instrumented instantiation (the one that is timed) produces the same frame."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    a = make_camera(sc, grp, 480, 270, 8).render().copy()
    ref = oracle.path_trace(oracle_scene(grp), 480, 270,
                            bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, 480, 270, 1)), max_depth=8)
    assert np.array_equal(a, ref["rgba8"])


def test_progressive_frames_and_both_call_sequences():
    """Five accumulated frames: fused gdpt_render_frame == the reference's dispatch-by-dispatch sequence
    (compute(main); compute(progressive); get_image) == oracle K1 + K2, including the accumulation buffer."""
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    W = H = 256
    fused = make_camera(sc, grp, W, H, 4, mode=PathTracingCamera.PROGRESSIVE_RENDERING, fused=True)
    split = make_camera(sc, grp, W, H, 4, mode=PathTracingCamera.PROGRESSIVE_RENDERING, fused=False)
    osc = oracle_scene(grp)
    acc = np.zeros((H, W, 4), np.float32)
    for f in range(1, 6):
        a = fused.render().copy()
        b = split.render().copy()
        assert fused.camera_block().frame_index == f and fused.last_frame_count() == f == split.last_frame_count()
        ref = oracle.path_trace(osc, W, H, bytes(fused.camera_block()), max_depth=4)
        screen = ref["rgba8"].copy()
        oracle.progressive(screen, acc, f)
        assert np.array_equal(a, screen), f"fused frame {f}"
        assert np.array_equal(b, screen), f"split frame {f}"
    assert np.array_equal(fused.read_image("accum").view(np.uint32), acc.view(np.uint32))
    assert np.array_equal(split.read_image("accum").view(np.uint32), acc.view(np.uint32))


def test_frame_count_resets_when_camera_moves():
    """progressive_rendering.cpp:53-60 (and quirk Q13: a camera sitting at the identity pose starts at 2)."""
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, 64, 64, 2, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    cam.render(); cam.render()
    assert cam.last_frame_count() == 2
    moved = sc.camera_transform12.copy()
    moved[9] += 0.5
    cam.set_global_transform(moved)
    cam.render()
    assert cam.last_frame_count() == 1
    ident = make_camera(scenes.SceneDesc("x", fov=60.0), grp, 64, 64, 2, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    ident.render()
    assert ident.last_frame_count() == 2


def test_debug_steps_define():
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, 320, 180, 5, debug=True)
    frame = cam.render().copy()
    ref = oracle.path_trace(oracle_scene(grp), 320, 180, bytes(cam.camera_block()), debug_steps=True)
    assert np.array_equal(frame, ref["rgba8"]) and frame[..., 0].max() > 0


def test_full_size_demo_frame_sampled_against_oracle():
    """BASELINE config C2 at full size (1920x1080, depth 8): exact ray count, exact frame on a band
    of rows the oracle renders, determinism (same frame_index twice -> identical bytes)."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    W, H = 1920, 1080
    cam = make_camera(sc, grp, W, H, 8)
    a = cam.render().copy()
    rays = cam.stats()["rays"]
    ref = oracle.path_trace(oracle_scene(grp), W, H, bytes(cam.camera_block()), max_depth=8)
    assert rays == ref["stats"]["rays"]
    assert np.array_equal(a, ref["rgba8"])
    cam.set_frame_index(0)
    assert np.array_equal(cam.render(), a), "same frame_index must reproduce the same bytes"


@pytest.mark.parametrize("name,make,W,H,depth,row_step", [
    ("C3_soup_1M", lambda: scenes.triangle_soup(1_000_000), 1920, 1080, 2, 120),
    ("C4_instanced_10M_4K", lambda: scenes.instanced_grid(), 3840, 2160, 8, 240),
], ids=["C3_soup_1M", "C4_instanced_10M_4K"])
def test_full_size_stress_configs_sampled_against_oracle(name, make, W, H, depth, row_step):
    """BASELINE configs C3 (1 M-triangle soup, 1080p) and C4 (10 M instanced triangles, 4K, depth 8) at full size: the rows
    the oracle renders (every row_step-th) are bit-identical in the GPU frame and depth image, two frames in a row
    (the second one scheduled by the cost classes of the first)."""
    sc = make()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, W, H, depth)
    osc = oracle_scene(grp)
    for frame in range(2):
        img = cam.render().copy()
        dep = cam.read_image("depth")
        ref = oracle.path_trace(osc, W, H, bytes(cam.camera_block()), max_depth=depth, row_step=row_step)
        rows = np.arange(0, H, row_step)
        assert np.array_equal(img[rows], ref["rgba8"][rows]), f"{name} frame {frame}: sampled rows differ"
        assert np.array_equal(dep[rows].view(np.uint32), ref["depth"][rows].view(np.uint32)), f"{name} frame {frame}: depth differs"
        assert ref["rgba8"][rows].any()
    assert cam.stats()["rays"] >= W * H


def test_row_sharding_reassembles_the_unsharded_frame():
    """Tile/row-band sharding (multi-GPU partition) is exact: the union of the parts' rows is the full frame,
    rows outside a part are untouched, ray counts add up."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    W, H, parts, band = 480, 270, 3, 8
    full_cam = make_camera(sc, grp, W, H, 6, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    full = [full_cam.render().copy() for _ in range(3)][-1]
    total_rays = full_cam.stats()["rays"]
    assembled = np.zeros_like(full)
    rays = 0
    for p in range(parts):
        cam = make_camera(sc, grp, W, H, 6, mode=PathTracingCamera.PROGRESSIVE_RENDERING, shard=(p, parts, band))
        img = [cam.render().copy() for _ in range(3)][-1]
        rays += cam.stats()["rays"]
        own = ((np.arange(H) // band) % parts) == p
        assert not img[~own].any(), "a shard wrote rows it does not own"
        assembled[own] = img[own]
    assert np.array_equal(assembled, full)
    assert rays == total_rays


def test_peer_screens_assemble_the_frame_without_a_gather():
    """Row-band presentation fused into the accumulate/tone-map kernel (gdpt_shader_set_peer_screens): every part also
    writes its rows into the other parts' images, so each image ends up as the whole, unsharded frame.  Here the parts are
    three cameras on one GPU (plain device pointers); across processes the same pointers come from CUDA IPC (bench.py
    --partition rows checks that path against the NCCL all-gather)."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    W, H, parts, band = 480, 270, 3, 8
    full_cam = make_camera(sc, grp, W, H, 6, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    cams = [make_camera(sc, grp, W, H, 6, mode=PathTracingCamera.PROGRESSIVE_RENDERING, shard=(p, parts, band)) for p in range(parts)]
    ptrs = [c.device_pointer("output")[0] for c in cams]
    for p, c in enumerate(cams):
        c.set_peer_screens([q for i, q in enumerate(ptrs) if i != p])
    for frame in range(3):
        full = full_cam.render().copy()
        for c in cams:
            c.render_device_only()
        for c in cams:
            c.synchronize()
        for p, c in enumerate(cams):
            assert np.array_equal(c.read_image("output"), full), f"frame {frame}: image of part {p} is not the whole frame"
    cams[0].set_peer_screens([])
    cams[0].render_device_only(); cams[0].synchronize()
    own = ((np.arange(H) // band) % parts) == 0
    img0, img1 = cams[0].read_image("output"), cams[1].read_image("output")
    assert np.array_equal(img1[own], full[own]) and not np.array_equal(img0[own], full[own]), "cleared peers are still written"


def test_gdcs_style_error_behaviour():
    """compute() before finish_create_uniforms is refused (gdcs.cpp:239-240,258-273); unknown shaders and bad
    binding tables are reported through the status code + gdpt_last_error, never by crashing."""
    dev = ctypes.c_void_p()
    _lib.check(_lib.cuda.gdpt_device_create(0, ctypes.byref(dev)), None, "device_create")
    sh = ctypes.c_void_p()
    assert _lib.cuda.gdpt_shader_create(dev, b"res://nope.glsl", None, 0, ctypes.byref(sh)) == -5
    _lib.check(_lib.cuda.gdpt_shader_create(dev, b"res://addons/jar_path_tracing/src/shaders/main.glsl", None, 0, ctypes.byref(sh)),
               dev, "shader_create")
    assert _lib.cuda.gdpt_shader_check_ready(sh) == 0
    assert _lib.cuda.gdpt_shader_compute(sh, 1, 1, 1) == -4
    assert _lib.cuda.gdpt_shader_finish_create_uniforms(sh) == -6
    assert b"main.glsl needs" in _lib.cuda.gdpt_last_error(dev)
    _lib.cuda.gdpt_shader_destroy(sh)
    _lib.cuda.gdpt_device_destroy(dev)


def test_storage_buffer_round_trip():
    dev = ctypes.c_void_p()
    _lib.check(_lib.cuda.gdpt_device_create(0, ctypes.byref(dev)), None, "device_create")
    sh = ctypes.c_void_p()
    _lib.check(_lib.cuda.gdpt_shader_create(dev, b"progressive_rendering.glsl", None, 0, ctypes.byref(sh)), dev, "create")
    data = np.arange(64, dtype=np.uint32)
    rid = _lib.cuda.gdpt_shader_create_storage_buffer_uniform(sh, data.ctypes.data_as(ctypes.c_void_p), data.nbytes, 0, 0)
    assert rid != 0
    upd = data[::-1].copy()
    _lib.check(_lib.cuda.gdpt_shader_update_storage_buffer_uniform(sh, rid, upd.ctypes.data_as(ctypes.c_void_p), upd.nbytes), dev, "update")
    back = np.zeros_like(data)
    _lib.check(_lib.cuda.gdpt_shader_get_storage_buffer_uniform(sh, rid, back.ctypes.data_as(ctypes.c_void_p), back.nbytes), dev, "get")
    assert np.array_equal(back, upd)
    _lib.cuda.gdpt_shader_destroy(sh)
    _lib.cuda.gdpt_device_destroy(dev)


@pytest.mark.gpu
def test_pipelined_frames_equal_blocking_frames():
    """gdpt_render_frame_begin/_wait (four frames in flight on four streams, consecutive frames overlapping each other and
    the read-back) returns the very bytes and ray counts of the blocking render(), frame after frame."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    a = make_camera(sc, grp, 320, 180, 5, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    b = make_camera(sc, grp, 320, 180, 5, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    want, want_rays = [], []
    for _ in range(8):
        want.append(a.render().copy()); want_rays.append(a.stats()["rays"])
    got, got_rays = [], []
    b.render_begin(); b.render_begin(); b.render_begin(); b.render_begin()
    with pytest.raises(_lib.GdptError):
        b.render_begin()  # a fifth frame in flight (GDPT_MAX_FRAMES_IN_FLIGHT = 4) is refused, loudly
    for i in range(8):
        img, st = b.render_wait()
        got.append(img.copy()); got_rays.append(st["rays"])
        if i + 4 < 8:
            b.render_begin()
    with pytest.raises(_lib.GdptError):
        b.render_wait()
    assert got_rays == want_rays
    for i in range(8):
        assert np.array_equal(got[i], want[i]), f"pipelined frame {i} differs"
    assert np.array_equal(b.read_image("accum").view(np.uint32), a.read_image("accum").view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [PathTracingCamera.NONE, PathTracingCamera.PROGRESSIVE_RENDERING], ids=["none", "progressive"])
def test_frames_in_flight_through_the_c_abi_with_depth_read_back(mode):
    """gdpt_render_frame_begin / _wait called directly (the host twin never asks for depth): three overlapped frames
    with colour AND depth read back equal the blocking frames bit for bit, and the bound images hold the last frame."""
    from gdpathtracing_b200._lib import cuda
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    W, H, depth, frames = 320, 180, 5, 7
    a = make_camera(sc, grp, W, H, depth, mode=mode)
    b = make_camera(sc, grp, W, H, depth, mode=mode)
    want = []
    for f in range(frames):
        img = a.render().copy()
        want.append((img, a.read_image("depth").copy(), bytes(a.camera_block()), a.last_frame_count()))
    b.render()  # creates the post-process shader exactly like the first frame of `a` did
    assert np.array_equal(b.output_image(), want[0][0])
    n = W * H
    slots = [(cuda.gdpt_host_alloc(n * 4), cuda.gdpt_host_alloc(n * 4)) for _ in range(3)]
    prog = b.progressive_shader if mode == PathTracingCamera.PROGRESSIVE_RENDERING else None
    try:
        got = []
        def begin(f):
            cam = _lib.Camera.from_buffer_copy(want[f][2])
            rgba, dep = slots[f % 3]
            _lib.check(cuda.gdpt_render_frame_begin(b.main_shader, prog, ctypes.byref(cam), mode, want[f][3], rgba, dep), b.device, "begin")
        def wait(f):
            st = _lib.FrameStats()
            _lib.check(cuda.gdpt_render_frame_wait(b.main_shader, ctypes.byref(st)), b.device, "wait")
            rgba, dep = slots[f % 3]
            got.append((np.ctypeslib.as_array(ctypes.cast(rgba, ctypes.POINTER(ctypes.c_uint8)), (H, W, 4)).copy(),
                        np.ctypeslib.as_array(ctypes.cast(dep, ctypes.POINTER(ctypes.c_float)), (H, W)).copy()))
        begin(1); begin(2); begin(3)
        for f in range(1, frames):
            wait(f)
            if f + 3 < frames:
                begin(f + 3)
        for f in range(1, frames):
            assert np.array_equal(got[f - 1][0], want[f][0]), f"frame {f}: colour differs"
            assert np.array_equal(got[f - 1][1].view(np.uint32), want[f][1].view(np.uint32)), f"frame {f}: depth differs"
        assert np.array_equal(b.read_image("output"), want[-1][0])
        assert np.array_equal(b.read_image("depth").view(np.uint32), want[-1][1].view(np.uint32))
    finally:
        b.synchronize()
        for rgba, dep in slots:
            cuda.gdpt_host_free(rgba); cuda.gdpt_host_free(dep)


def test_sample_index_accumulation_from_a_frame_store_equals_progressive_rendering():
    """multigpu.PeerFrameStore + SampleIndexAccumulator.add_from_store (one rank: the store is this GPU's own memory; across
    ranks the same addresses are CUDA IPC mappings, checked by bench.py's first_presented_frame_equals_single_gpu): K2 run
    from the kept frames, in frame order, equals the PROGRESSIVE_RENDERING camera frame for frame and the oracle's K2."""
    import torch
    from gdpathtracing_b200 import multigpu
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    W, H, n = 192, 128, 5
    ref_cam = make_camera(sc, grp, W, H, 4, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    want = [ref_cam.render().copy() for _ in range(n)]
    cam = make_camera(sc, grp, W, H, 4, mode=PathTracingCamera.NONE)
    dev = torch.device("cuda", 0)
    store = multigpu.PeerFrameStore(cam, n, H, W, 0, 1, dev)
    stream = torch.cuda.ExternalStream(cam.stream())
    ptr, _ = cam.device_pointer("output")
    frame_t = multigpu.as_tensor(ptr, (H, W, 4), torch.uint8, dev)
    for f in range(n):
        cam.set_frame_index(f)
        cam.render_device_only()
        with torch.cuda.stream(stream):
            store.frames[f].copy_(frame_t)
    acc = multigpu.SampleIndexAccumulator(H, W, 0, 1, dev, multigpu.cuda_k2_at(cam))
    with torch.cuda.stream(stream):
        assert acc.add_from_store(store, 3) == 3
        third = acc.present().clone()
    torch.cuda.synchronize()
    assert np.array_equal(third.cpu().numpy(), want[2])
    acc2 = multigpu.SampleIndexAccumulator(H, W, 0, 1, dev, multigpu.cuda_k2_at(cam))
    with torch.cuda.stream(stream):
        acc2.add_from_store(store)
        last = acc2.present().clone()
    torch.cuda.synchronize()
    assert np.array_equal(last.cpu().numpy(), want[-1])
    assert np.array_equal(acc2.accum.cpu().numpy().view(np.uint32), ref_cam.read_image("accum").view(np.uint32))
    store.close()
