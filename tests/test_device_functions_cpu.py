"""CPU tier: the backend's per-ray device functions (compiled for the host by tests/devcheck,
see its header) against the oracle, bit for bit: hit records, node-visit order, per-ray work
counters, colour, depth, and the K2 accumulate/tone-map arithmetic.  The GPU tier
(test_gpu_parity.py) repeats these checks through the C-ABI on the real kernels."""
import ctypes

import numpy as np
import pytest

from conftest import ptr, records_equal
from gdpathtracing_b200 import nodes, scenes
from oracle import oracle

CASES = [
    ("cornell32", lambda: scenes.cornell32(), 256, 256, 4, 4, 1),          # BASELINE config C1, full size
    ("demo", lambda: scenes.demo_scene(), 240, 135, 8, 3, 1),
    ("demo_frame7", lambda: scenes.demo_scene(), 160, 90, 5, 2, 7),
    ("soup", lambda: scenes.triangle_soup(20000, seed=1), 96, 54, 2, 2, 1),
    ("instanced", lambda: scenes.instanced_grid(3, 800, seed=3), 128, 72, 5, 2, 3),
]


def run_devcheck(devcheck, osc, cam_bytes, W, H, depth, segs, vis, debug=False):
    params = np.zeros(9, np.uint32)
    params[4], params[5] = W, H
    cam = np.frombuffer(cam_bytes, np.uint8).copy()
    out = np.zeros((H, W, 4), np.uint8)
    dep = np.zeros((H, W), np.float32)
    tr = np.zeros((segs, H * W), dtype=oracle.TRACE_DTYPE)
    vs = np.full((H * W, vis), 0xFFFFFFFF, np.uint32)
    rays = np.zeros(1, np.uint64)
    rc = devcheck.devcheck_path_trace(ctypes.byref(osc.c), ptr(params), ptr(cam), depth, 1 if debug else 0, 0, H, ptr(out),
                                      ptr(dep), ptr(tr), segs, ptr(vs), vis, ptr(rays))
    assert rc == 0
    return out, dep, tr, vs, int(rays[0])


@pytest.mark.parametrize("stepwise", [0, 1, 2], ids=["node_steps", "triangle_steps", "compact_steps"])
@pytest.mark.parametrize("name,make,W,H,depth,segs,frame", CASES, ids=[c[0] for c in CASES])
def test_device_functions_match_oracle(devcheck, name, make, W, H, depth, segs, frame, stepwise):
    devcheck.devcheck_set_stepwise(stepwise)
    sc = make()
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame))
    ref = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs, visits_per_ray=48)
    out, dep, tr, vs, rays = run_devcheck(devcheck, osc, cam, W, H, depth, segs, 48)
    assert ref["stats"]["stack_overflow"] == 0
    assert rays == ref["stats"]["rays"]
    assert ref["stats"]["primary_hits"] > 0, "scene/camera produced no hits: not a test"
    for s in range(segs):
        ok, why = records_equal(tr[s], ref["trace"][s])
        assert ok, f"segment {s}: {why}"
    assert np.array_equal(vs, ref["visits"]), "node-visit order differs"
    assert np.array_equal(out, ref["rgba8"])
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))


HIT_FIELDS = ("hit", "triangle", "blas", "front", "t", "u", "v")


@pytest.mark.parametrize("stepwise", [0, 1, 2], ids=["node_steps", "triangle_steps", "compact_steps"])
@pytest.mark.parametrize("name,make,W,H,depth,segs,frame", CASES, ids=[c[0] for c in CASES])
def test_tight_box_culling_changes_no_result(devcheck, name, make, W, H, depth, segs, frame, stepwise):
    """Culling (pt_scene.cuh) may only shorten the visit list: hit records, ray counts, colour and depth
    stay bit-identical to the oracle's full reference traversal, and it must actually save work."""
    devcheck.devcheck_set_stepwise(stepwise)
    devcheck.devcheck_set_cull(1)
    try:
        sc = make()
        grp = scenes.populate(sc)
        grp.build()
        osc = oracle.Scene(grp.buffers(), grp.texture_layers())
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame))
        ref = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs)
        out, dep, tr, _, rays = run_devcheck(devcheck, osc, cam, W, H, depth, segs, 1)
    finally:
        devcheck.devcheck_set_cull(0)
    assert rays == ref["stats"]["rays"]
    pops_ref = pops_cull = 0
    for s in range(segs):
        a, b = tr[s], ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"])
        live = b["hit"] != 0xFFFFFFFF
        for f in HIT_FIELDS:
            x, y = a[f][live], b[f][live]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            assert np.array_equal(x, y), f"segment {s} field {f}"
        assert (a["tri_tests"][live] <= b["tri_tests"][live]).all() and (a["node_pops"][live] <= b["node_pops"][live]).all()
        pops_ref += int(b["node_pops"][live].sum()); pops_cull += int(a["node_pops"][live].sum())
    assert np.array_equal(out, ref["rgba8"])
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))
    if name != "cornell32":  # every Cornell-32 BLAS is a single leaf: nothing to cull below the TLAS
        assert pops_cull < pops_ref
    print(f"{name}: node pops {pops_ref} -> {pops_cull}")


def test_debug_steps_mode(devcheck):
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, 160, 90, 1))
    ref = oracle.path_trace(osc, 160, 90, cam, debug_steps=True)
    out, dep, _, _, _ = run_devcheck(devcheck, osc, cam, 160, 90, 5, 1, 1, debug=True)
    assert np.array_equal(out, ref["rgba8"]) and ref["rgba8"][..., 0].max() > 0
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))


def test_progressive_accumulation_arithmetic(devcheck):
    """K2 (progressive_rendering.glsl:28-46) over 5 frames incl. the frame_count == 1 reset."""
    rng = np.random.default_rng(3)
    a_acc = np.zeros((40, 64, 4), np.float32)
    b_acc = rng.random((40, 64, 4), dtype=np.float32)  # garbage that frame_count == 1 must ignore
    a_acc[:] = b_acc
    for fc in (1, 2, 3, 4, 5):
        frame = rng.integers(0, 256, (40, 64, 4), dtype=np.uint8)
        a, b = frame.copy(), frame.copy()
        oracle.progressive(a, a_acc, fc)
        devcheck.devcheck_progressive(ptr(b), ptr(b_acc), 64, 40, fc)
        assert np.array_equal(a, b) and np.array_equal(a_acc.view(np.uint32), b_acc.view(np.uint32))
        assert (a[..., 3] == 255).all()


def test_tie_on_equal_t_last_tested_triangle_wins(devcheck):
    """Quirk Q8 (main.glsl:247): `t > hit.t` rejects, so an equal-t duplicate tested later replaces the hit."""
    sc = scenes.SceneDesc("dup", camera_transform12=scenes.transform12(None, (0, 0, 5)), fov=40.0)
    sc.materials = [dict()]
    sc.default_material = 0
    tri = np.array([[-1, -1, 0], [0, 1, 0], [1, -1, 0]], np.float32)  # clockwise seen from +z
    p = np.concatenate([tri, tri])
    n = np.tile(np.array([0, 0, 1], np.float32), (6, 1))
    sc.meshes = [[{"positions": p, "normals": n, "uvs": np.zeros((6, 2), np.float32), "indices": np.arange(6, dtype=np.int32)}]]
    sc.instances = [dict(mesh=0)]
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, 32, 32, 1))
    ref = oracle.path_trace(osc, 32, 32, cam, max_depth=1, trace_segments=1)
    hits = ref["trace"][0][ref["trace"][0]["hit"] == 1]
    assert len(hits) > 0 and (hits["triangle"] == 1).all()
    _, _, tr, _, _ = run_devcheck(devcheck, osc, cam, 32, 32, 1, 1, 1)
    ok, why = records_equal(tr[0], ref["trace"][0])
    assert ok, why


def fast_counts(devcheck):
    c = np.zeros(3, np.uint64)
    devcheck.devcheck_fast_counts(ptr(c))
    return dict(rays=int(c[0]), retraced=int(c[1]), ties=int(c[2]))


FAST_CASES = CASES + [
    ("demo_wide", lambda: scenes.demo_scene(), 320, 180, 8, 8, 3),
    ("soup_dense", lambda: scenes.triangle_soup(60000, seed=5), 96, 54, 3, 3, 2),
]


@pytest.mark.parametrize("tables", [1, 2], ids=["two_wide", "four_wide"])
@pytest.mark.parametrize("name,make,W,H,depth,segs,frame", FAST_CASES, ids=[c[0] for c in FAST_CASES])
def test_closest_hit_search_with_proof_equals_reference_traversal(devcheck, name, make, W, H, depth, segs, frame, tables):
    """pt_fast.cuh: an order-free closest-hit search over our own BVH (two-wide, or its four-wide collapse) plus the proof
    that the reference reaches that triangle (exact re-trace where the proof fails) returns the reference's hit records,
    bit for bit."""
    devcheck.devcheck_set_fast(tables)
    try:
        sc = make()
        grp = scenes.populate(sc)
        grp.build()
        osc = oracle.Scene(grp.buffers(), grp.texture_layers())
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame))
        ref = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs)
        out, dep, tr, _, rays = run_devcheck(devcheck, osc, cam, W, H, depth, segs, 1)
        counts = fast_counts(devcheck)
    finally:
        devcheck.devcheck_set_fast(0)
    assert rays == ref["stats"]["rays"]
    assert counts["rays"] == rays, "the closest-hit tables were not used"
    for s in range(segs):
        a, b = tr[s], ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"])
        live = b["hit"] != 0xFFFFFFFF
        for f in HIT_FIELDS:
            x, y = a[f][live], b[f][live]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), y.view(np.uint32)
            assert np.array_equal(x, y), f"segment {s} field {f}"
    assert np.array_equal(out, ref["rgba8"])
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))
    print(f"{name}: {counts}")


def test_closest_hit_search_sends_ties_to_the_exact_traversal(devcheck):
    """Equal-t duplicates (quirk Q8): the search must notice the tie and fall back, not pick a winner."""
    sc = scenes.SceneDesc("dup", camera_transform12=scenes.transform12(None, (0, 0, 5)), fov=40.0)
    sc.materials = [dict()]
    sc.default_material = 0
    tri = np.array([[-1, -1, 0], [0, 1, 0], [1, -1, 0]], np.float32)
    p = np.concatenate([tri, tri])
    n = np.tile(np.array([0, 0, 1], np.float32), (6, 1))
    sc.meshes = [[{"positions": p, "normals": n, "uvs": np.zeros((6, 2), np.float32), "indices": np.arange(6, dtype=np.int32)}]]
    sc.instances = [dict(mesh=0)]
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, 32, 32, 1))
    ref = oracle.path_trace(osc, 32, 32, cam, max_depth=1, trace_segments=1)
    devcheck.devcheck_set_fast(1)
    try:
        _, _, tr, _, _ = run_devcheck(devcheck, osc, cam, 32, 32, 1, 1, 1)
        counts = fast_counts(devcheck)
    finally:
        devcheck.devcheck_set_fast(0)
    hits = ref["trace"][0]["hit"] == 1
    # every hit is a tie; a camera ray with a zero direction component skips the search and is counted with them
    assert 0 < int(hits.sum()) <= counts["ties"] <= int(hits.sum()) + 4
    for f in HIT_FIELDS:
        assert np.array_equal(tr[0][f][hits].view(np.uint32), ref["trace"][0][f][hits].view(np.uint32)), f


def test_closest_hit_tables_do_not_depend_on_the_thread_count(devcheck):
    """fast_bvh.h builds large trees with a thread-pooled top and independent subtrees placed in depth-first order:
    one, three and all threads give byte-identical two-wide, triangle and four-wide tables."""
    sc = scenes.triangle_soup(150_000, seed=4)
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    devcheck.devcheck_fast_layout_digest.restype = ctypes.c_int
    digests = []
    for threads in (1, 3, 0):
        d = np.zeros(4, np.uint64)
        assert devcheck.devcheck_fast_layout_digest(ctypes.byref(osc.c), threads, ptr(d)) == 0
        digests.append(d.tolist())
    assert digests[0] == digests[1] == digests[2]


@pytest.mark.parametrize("name,make", [(c[0], c[1]) for c in FAST_CASES], ids=[c[0] for c in FAST_CASES])
def test_four_wide_tables_are_a_partition_with_true_boxes(devcheck, name, make):
    """fast_bvh.h Collapse: below every BLAS root each triangle copy is reached exactly once, every triangle lies inside
    every box on its chain (the intersection of the ancestors' boxes), nodes have two to four children."""
    sc = make()
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    st = np.zeros(3, np.uint64)
    devcheck.devcheck_fast4_structure.restype = ctypes.c_int
    rc = devcheck.devcheck_fast4_structure(ctypes.byref(osc.c), ptr(st))
    assert rc == 0, f"four-wide tables violate invariant {rc}"
    nodes_walked, children, deepest = (int(x) for x in st)
    if nodes_walked:
        assert children / nodes_walked > 2.5, "the collapse should fill most nodes beyond two children"
    print(f"{name}: {nodes_walked} nodes, {children / max(nodes_walked, 1):.2f} children per node, depth {deepest}")
