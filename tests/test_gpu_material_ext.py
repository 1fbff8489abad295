"""GPU tier of the material-breadth extension ("#define GDPT_MATERIAL_EXT", include/gdpt_wire.h): the rendering
kernels against oracle/pt_oracle.cpp, which defines the extension (the reference has none).  Frames, depth, ray counts and
hit records bit for bit, with five- and six-surface meshes, roughness / metallic maps and an sRGB albedo layer; and the same
scene with the extension off renders as before."""
import numpy as np
import pytest

from gdpathtracing_b200 import PathTracingCamera, scenes
from material_ext_scene import material_ext_scene
from oracle import oracle

pytestmark = pytest.mark.gpu
W, H, DEPTH = 240, 180, 6


def camera_for(sc, grp, variant, mode=PathTracingCamera.NONE):
    cam = PathTracingCamera()
    cam.fov = sc.fov
    cam.geometry_group = grp
    cam.denoising_mode = mode
    cam.set_window_size(W, H)
    cam.set_global_transform(sc.camera_transform12)
    cam.set_max_depth(DEPTH)
    cam.set_record_hits(4)
    cam.set_variant(variant)
    cam.init()
    return cam


@pytest.mark.parametrize("variant", [6, 3], ids=["pooled_paths", "reference_order"])
@pytest.mark.parametrize("ext,many", [(True, True), (True, False), (False, False)], ids=["ext_many_surfaces", "ext_three_surfaces", "ext_off"])
def test_material_extension_frames_equal_the_oracle(variant, ext, many):
    sc = material_ext_scene(ext, many)
    grp = scenes.populate(sc)
    cam = camera_for(sc, grp, variant)
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    assert osc.c.material_ext == (1 if ext else 0)
    for _ in range(2):
        frame = cam.render().copy()
        st = cam.stats()
        ref = oracle.path_trace(osc, W, H, bytes(cam.camera_block()), max_depth=DEPTH, trace_segments=4)
        assert st["rays"] == ref["stats"]["rays"] and st["primary_hits"] == ref["stats"]["primary_hits"] > W * H // 2
        for s in range(4):
            a, b = cam.read_trace(s), ref["trace"][s]
            assert np.array_equal(a["hit"], b["hit"])
            live = b["hit"] != 0xFFFFFFFF
            for f in ("triangle", "blas", "front", "t", "u", "v"):
                assert np.array_equal(a[f][live].view(np.uint32), b[f][live].view(np.uint32)), f"segment {s} field {f}"
        assert np.array_equal(frame, ref["rgba8"])
        assert np.array_equal(cam.read_image("depth").view(np.uint32), ref["depth"].view(np.uint32))


def test_material_extension_accumulates_like_any_frame():
    sc = material_ext_scene(True, True)
    grp = scenes.populate(sc)
    cam = camera_for(sc, grp, 6, mode=PathTracingCamera.PROGRESSIVE_RENDERING)
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    acc = np.zeros((H, W, 4), np.float32)
    for f in range(1, 4):
        got = cam.render().copy()
        ref = oracle.path_trace(osc, W, H, bytes(cam.camera_block()), max_depth=DEPTH)
        screen = ref["rgba8"].copy()
        oracle.progressive(screen, acc, f)
        assert np.array_equal(got, screen), f"frame {f}"
