"""GPU tier: the CUDA kernels against the REFERENCE'S OWN SHADER TEXT (oracle/_ref/libgdpt_refshader.so, which
travels to the GPU box prebuilt) and against the committed digests of its outputs
(tests/golden/ref_shader_hashes.json).  Both the reference-order trace kernels and the kernels that are timed."""
import json

import numpy as np
import pytest

import test_ref_shader as t
from gdpathtracing_b200 import PathTracingCamera, scenes
from oracle import oracle

pytestmark = pytest.mark.gpu


def gpu_camera(sc, grp, W, H, depth, frame, trace):
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = PathTracingCamera.NONE
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(depth)
    cam.set_frame_index(frame - 1)  # render() increments before the dispatch (path_tracing_camera.cpp:199)
    if trace:
        cam.set_trace(min(t.SEGS, depth), t.VISITS)
    cam.init()
    return cam


@pytest.mark.parametrize("name,make,W,H,depth,frame", t.CASES, ids=[c[0] for c in t.CASES])
def test_trace_kernels_equal_the_committed_reference_shader_digests(name, make, W, H, depth, frame):
    """Reference-order kernels (trace mode): every digest the reference's shader text produced -- frame, depth,
    node-visit order, per-segment hit ids, counters, visit hashes -- is reproduced by the GPU."""
    golden = json.load(open(t.GOLDEN))["frames"][name]
    sc = make()
    grp = scenes.populate(sc)
    cam = gpu_camera(sc, grp, W, H, depth, frame, trace=True)
    frame_px = cam.render().copy()
    st = cam.stats()
    tr = np.stack([cam.read_trace(s) for s in range(min(t.SEGS, depth))])
    if depth < t.SEGS:  # segments beyond the path length do not exist: the marker the CPU checkers use
        pad = np.zeros((t.SEGS - depth,) + tr.shape[1:], tr.dtype)
        pad["hit"] = 0xFFFFFFFF
        tr = np.concatenate([tr, pad])
    got = {"rgba8": t.sha(frame_px), "depth": t.sha(cam.read_image("depth")), "visits": t.sha(cam.read_visits()),
           "rays": st["rays"], "primary_hits": st["primary_hits"]}
    for k in ("node_pops", "box_tests", "tri_tests", "tlas_leaves"):
        got[k] = st[k]
    live = tr["hit"] != 0xFFFFFFFF
    for f in t.LOG_FIELDS:
        got["trace_" + f] = t.sha(np.where(live, tr[f], 0))
    for k, v in got.items():
        assert golden[k] == v, k


@pytest.mark.parametrize("name,make,W,H,depth,frame", t.CASES, ids=[c[0] for c in t.CASES])
def test_timed_kernels_equal_the_reference_shader_text(name, make, W, H, depth, frame):
    """The kernels that are benchmarked (default schedule): frame, depth and ray count equal the reference shader's --
    run here when the prebuilt library travelled with the snapshot, else its committed digests."""
    sc = make()
    grp = scenes.populate(sc)
    cam = gpu_camera(sc, grp, W, H, depth, frame, trace=False)
    frame_px = cam.render().copy()
    depth_px = cam.read_image("depth")
    st = cam.stats()
    golden = json.load(open(t.GOLDEN))["frames"][name]
    assert t.sha(frame_px) == golden["rgba8"] and t.sha(depth_px) == golden["depth"] and st["rays"] == golden["rays"]
    if oracle.ref_shader_available():
        ref = oracle.path_trace(oracle.Scene(grp.buffers(), grp.texture_layers()), W, H, bytes(cam.camera_block()),
                                max_depth=depth, impl="reference")
        assert np.array_equal(frame_px, ref["rgba8"])
        assert np.array_equal(depth_px.view(np.uint32), ref["depth"].view(np.uint32))
        assert st["rays"] == ref["stats"]["rays"] and st["primary_hits"] == ref["stats"]["primary_hits"]
