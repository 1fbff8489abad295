"""Material breadth: roughness / metallic textures, sRGB albedo layers and any number of materials per
instance, behind "#define GDPT_MATERIAL_EXT" (include/gdpt_wire.h).  The reference has none of this, so there is nothing of
its own to pin the extension to: oracle/pt_oracle.cpp defines it and the device functions are held to that, bit for bit.
With the extension off the very same scene description produces the reference's bytes and the reference's shading."""
import numpy as np
import pytest

from gdpathtracing_b200 import nodes, scenes
from material_ext_scene import material_ext_scene
from oracle import oracle
from test_device_functions_cpu import HIT_FIELDS, fast_counts, run_devcheck

W, H, DEPTH = 160, 120, 5


def build(ext, many_surfaces=True):
    sc = material_ext_scene(ext, many_surfaces)
    grp = scenes.populate(sc)
    grp.build()
    return sc, grp, oracle.Scene(grp.buffers(), grp.texture_layers())


def test_extension_words_and_material_table():
    _, grp, _ = build(True)
    b = grp.buffers()
    mats = np.frombuffer(b["materials"], np.uint32).reshape(-1, 16)
    assert mats[:, 11].max() > 0 and mats[:, 12].max() > 0 and (mats[:, 13] & 1).any(), "roughness / metallic layers and the sRGB flag"
    table = np.frombuffer(b["surface_materials"], np.uint32)
    n_inst = len(b["blas"]) // 176
    off = table[:n_inst + 1]
    assert off[0] == n_inst + 1 and off[-1] == len(table) and np.diff(off).tolist() == [3, 5, 6, 5]
    blas = np.frombuffer(b["blas"], np.uint32).reshape(-1, 44)
    for i in range(n_inst):  # the first three ids are also where the reference keeps them (bvh.h:71)
        assert table[off[i]:off[i] + 3].tolist() == blas[i, 41:44].tolist()
    _, grp1, _ = build(True, many_surfaces=False)
    _, grp0, _ = build(False, many_surfaces=False)
    b1, b0 = grp1.buffers(), grp0.buffers()
    assert "surface_materials" not in b0
    m0 = np.frombuffer(b0["materials"], np.uint32).reshape(-1, 16)
    assert not m0[:, 11:].any(), "extension off: the padding stays zero, as upstream leaves it"
    for k in ("triangles_geometry", "triangles_data", "bvh", "blas", "tlas"):
        assert b1[k] == b0[k], k


@pytest.mark.parametrize("fast", [0, 2], ids=["reference_order", "closest_hit_search"])
def test_device_functions_match_the_oracle_with_the_extension(devcheck, fast):
    sc, _, osc = build(True)
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, 2))
    ref = oracle.path_trace(osc, W, H, cam, max_depth=DEPTH, trace_segments=4, radiance=True)
    assert ref["stats"]["primary_hits"] > W * H // 2
    devcheck.devcheck_set_fast(fast)
    try:
        out, dep, tr, _, rays = run_devcheck(devcheck, osc, cam, W, H, DEPTH, 4, 1)
    finally:
        devcheck.devcheck_set_fast(0)
    assert rays == ref["stats"]["rays"]
    for s in range(4):
        a, b = tr[s], ref["trace"][s]
        assert np.array_equal(a["hit"], b["hit"])
        live = b["hit"] != 0xFFFFFFFF
        for f in HIT_FIELDS:
            assert np.array_equal(a[f][live].view(np.uint32), b[f][live].view(np.uint32)), f"segment {s} field {f}"
    assert np.array_equal(out, ref["rgba8"])
    assert np.array_equal(dep.view(np.uint32), ref["depth"].view(np.uint32))


def test_the_extension_changes_the_picture_and_its_absence_is_the_reference():
    sc, _, on = build(True, many_surfaces=False)
    _, _, off = build(False, many_surfaces=False)
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, 2))
    a = oracle.path_trace(on, W, H, cam, max_depth=DEPTH)
    b = oracle.path_trace(off, W, H, cam, max_depth=DEPTH)
    assert not np.array_equal(a["rgba8"], b["rgba8"]), "roughness / metallic maps and the sRGB decode must be visible"
    if oracle.ref_shader_available():  # extension off: the restatement is still the reference's shader text on this scene
        c = oracle.path_trace(off, W, H, cam, max_depth=DEPTH, impl="reference")
        assert np.array_equal(b["rgba8"], c["rgba8"]) and b["stats"]["rays"] == c["stats"]["rays"]


def test_without_the_define_the_extension_words_are_ignored():
    """A reader that was not asked for the extension (material_ext == 0) treats the words as the padding they are upstream."""
    sc, grp1, _ = build(True, many_surfaces=False)
    _, _, off = build(False, many_surfaces=False)
    bufs = grp1.buffers()
    del bufs["surface_materials"]
    deaf = oracle.Scene(bufs, grp1.texture_layers())
    assert deaf.c.material_ext == 0
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, 2))
    a = oracle.path_trace(deaf, W, H, cam, max_depth=DEPTH)
    b = oracle.path_trace(off, W, H, cam, max_depth=DEPTH)
    assert np.array_equal(a["rgba8"], b["rgba8"]) and a["stats"]["rays"] == b["stats"]["rays"]
