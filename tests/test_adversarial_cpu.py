"""CPU tier of the adversarial cases (tests/adversarial.py): (1) the restatement against the reference's own shader text on
exactly these inputs, so the oracle is pinned where the GPU tier leans on it hardest; (2) the closest-hit search + proof of
pt_fast.cuh (compiled for the host) against the oracle, with the number of rays it hands to the exact traversal -- the
number the GPU tier must reproduce (test_gpu_adversarial.py)."""
import numpy as np
import pytest

import adversarial
from conftest import records_equal
from oracle import oracle
from test_device_functions_cpu import HIT_FIELDS, fast_counts, run_devcheck

SEGS = 4
pin = pytest.mark.skipif(not oracle.ref_shader_available(), reason="oracle/_ref/libgdpt_refshader.so not built")


def build(make):
    from gdpathtracing_b200 import scenes
    sc = make()
    grp = scenes.populate(sc)
    grp.build()
    return sc, grp, oracle.Scene(grp.buffers(), grp.texture_layers())


@pin
@pytest.mark.parametrize("name,make,W,H,depth,zero_axes", adversarial.CASES, ids=adversarial.IDS)
def test_restatement_equals_reference_shader_text_on_adversarial_input(name, make, W, H, depth, zero_axes):
    sc, _, osc = build(make)
    cam = bytes(adversarial.camera_for(sc, W, H, 3, zero_axes))
    segs = min(SEGS, depth)
    a = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs, visits_per_ray=32, radiance=True, impl="restatement")
    b = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs, visits_per_ray=32, radiance=True, impl="reference")
    assert b["stats"]["primary_hits"] > 0, "not a test: no camera ray hits anything"
    assert np.array_equal(a["rgba8"], b["rgba8"])
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
    assert np.array_equal(a["radiance"].view(np.uint32), b["radiance"].view(np.uint32))
    for k in ("rays", "primary_hits", "node_pops", "box_tests", "tri_tests", "tlas_leaves"):
        assert a["stats"][k] == b["stats"][k], k
    live = b["trace"]["hit"] != 0xFFFFFFFF
    for f in ("hit", "triangle", "blas", "node_pops", "box_tests", "tri_tests", "tlas_leaves", "visit_hash_lo", "visit_hash_hi"):
        assert np.array_equal(a["trace"][f][live], b["trace"][f][live]), f
    assert np.array_equal(a["visits"], b["visits"])


@pytest.mark.parametrize("tables", [1, 2], ids=["two_wide", "four_wide"])
@pytest.mark.parametrize("name,make,W,H,depth,zero_axes", adversarial.CASES, ids=adversarial.IDS)
def test_closest_hit_search_on_adversarial_input(devcheck, name, make, W, H, depth, zero_axes, tables):
    sc, _, osc = build(make)
    cam = bytes(adversarial.camera_for(sc, W, H, 3, zero_axes))
    segs = min(SEGS, depth)
    ref = oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=segs)
    assert ref["stats"]["primary_hits"] > 0
    devcheck.devcheck_set_fast(tables)
    try:
        out, dep, tr, _, rays = run_devcheck(devcheck, osc, cam, W, H, depth, segs, 1)
        counts = fast_counts(devcheck)
    finally:
        devcheck.devcheck_set_fast(0)
    assert rays == ref["stats"]["rays"] == counts["rays"]
    for s in range(segs):
        a, b = tr[s], ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"])
        live = b["hit"] != 0xFFFFFFFF
        for f in HIT_FIELDS:
            assert np.array_equal(a[f][live].view(np.uint32), b[f][live].view(np.uint32)), f"segment {s} field {f}"
    assert np.array_equal(out, ref["rgba8"])
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))
    frac = counts["retraced"] / counts["rays"]
    print(f"{name}: {counts}, {100 * frac:.1f} % re-traced, primary hits {ref['stats']['primary_hits']}")
    if name == "coplanar_duplicates":   # every hit is a tie
        hits = sum(int((ref["trace"][s]["hit"] == 1).sum()) for s in range(segs))
        assert counts["ties"] >= hits > 0
    if name == "small_prop_in_a_large_room":
        assert frac < 0.02, "a small prop inside the scene must not push its rays to the exact traversal"
    if name in ("coplanar_duplicates", "rays_in_the_plane_x0", "rays_along_minus_z", "far_tiny_instance"):
        assert frac > 0.05, "this case is meant to send more than 5 % of the rays to the exact traversal"


def test_records_equal_helper_sees_a_difference():
    a = np.zeros(4, dtype=oracle.TRACE_DTYPE)
    b = a.copy()
    assert records_equal(a, b)[0]
    b["t"][2] = 1.0
    assert not records_equal(a, b)[0]
