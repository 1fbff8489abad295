"""The host BLAS/TLAS builder (gdpathtracing_b200/csrc/host/accel_build.cpp) emits the REFERENCE's
bytes: compared with the reference's own bvh.cpp compiled in place (oracle/_ref) when that
library is present, and with the committed SHA-256 fixtures (made from it) always."""
import hashlib
import json
import os

import numpy as np
import pytest

from accel_cases import CASES, product_buffers, reference_buffers
from conftest import REPO
from oracle import oracle

GOLDEN = json.load(open(os.path.join(REPO, "tests", "golden", "accel_hashes.json")))["cases"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_builder_matches_committed_reference_hashes(name):
    got, _ = product_buffers(CASES[name]())
    for key, arr in got.items():
        assert len(arr) == GOLDEN[name]["counts"][key], f"{name}/{key}: size differs"
        assert hashlib.sha256(arr.tobytes()).hexdigest() == GOLDEN[name][key], f"{name}/{key}: bytes differ from the reference"


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libgdpt_refbvh.so not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_builder_matches_reference_compiled_in_place(name):
    sc = CASES[name]()
    got, _ = product_buffers(sc)
    ref = reference_buffers(sc)
    for key in got:
        assert np.array_equal(got[key], ref[key]), f"{name}/{key}: product builder differs from reference bvh.cpp"


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libgdpt_refbvh.so not built")
def test_builder_matches_reference_on_100k_soup():
    from gdpathtracing_b200 import scenes
    sc = scenes.triangle_soup(100_000, seed=5)
    got, _ = product_buffers(sc)
    ref = reference_buffers(sc)
    for key in got:
        assert np.array_equal(got[key], ref[key])


def test_quirk_q1_every_blas_box_contains_the_origin_planes():
    """bvh.cpp:6-10 + vec.h:49: boxes start at (FLT_MAX,0,0)/(FLT_MIN,0,0), so y=0 and z=0 are always inside."""
    _, b = product_buffers(CASES["soup20k"]())
    n = np.frombuffer(b["bvh"], np.float32).reshape(-1, 12)
    assert (n[:, 1] <= 0).all() and (n[:, 2] <= 0).all() and (n[:, 5] >= 0).all() and (n[:, 6] >= 0).all()
    assert (n[:, 4] > 0).all()  # max.x >= FLT_MIN_POSITIVE


def test_tlas_layout_and_16_bit_links():
    """TLAS: node 0 = copy of the last merged node, leaves at 1..I in instance order, 2*I nodes (bvh.cpp:264-317)."""
    _, b = product_buffers(CASES["instanced27x800"]())
    t = np.frombuffer(b["tlas"], np.uint32).reshape(-1, 8)
    n_inst = len(b["blas"]) // 176
    assert len(t) == 2 * n_inst
    assert (t[1:n_inst + 1, 3] == 0).all() and (t[1:n_inst + 1, 7] == np.arange(n_inst)).all()
    assert np.array_equal(t[0], t[-1])
    internal = t[n_inst + 1:]
    assert ((internal[:, 3] & 0xFFFF) < len(t)).all() and ((internal[:, 3] >> 16) < len(t)).all()


def test_material_and_texture_conversion():
    """StandardMaterial3D -> GpuMaterial (geometry_group3d.cpp:271-292): default at index 0, de-duplication by
    identity, emission rgb + energy in .w, texture indices in first-use order, blank layer when no textures."""
    from gdpathtracing_b200 import scenes
    grp = scenes.populate(scenes.demo_scene())
    grp.build()
    m = np.frombuffer(grp.buffer("materials"), np.float32).reshape(-1, 16)
    mi = np.frombuffer(grp.buffer("materials"), np.int32).reshape(-1, 16)
    assert len(m) == 6
    assert np.allclose(m[0, :4], [1, 1, 1, 1]) and mi[0, 10] == -1 and m[0, 9] == 1.0   # default StandardMaterial3D
    assert mi[1, 10] == 0 and mi[4, 10] == 1 and mi[5, 10] == 2                          # grass, icon, checker
    assert np.allclose(m[2, :3], [1.0, 0.16, 0.16]) and np.allclose(m[:, 7], 1.0)        # energy multiplier default 1
    layers = grp.texture_layers()
    assert len(layers) == 3 and layers[0].shape == (1024, 1024, 4)
    blas = np.frombuffer(grp.buffer("blas"), np.uint32).reshape(-1, 44)
    assert blas[0, 41:44].tolist() == [1, 2, 3]      # room: grass, red, green
    assert blas[1, 41] == 0                          # Suzanne without override -> default
    assert blas[3, 41:44].tolist() == [0, 5, 5]      # Gobot: surface 0 has no override -> default (Appendix B)

    grp2 = scenes.populate(scenes.cornell32())
    grp2.build()
    assert len(grp2.texture_layers()) == 1 and not grp2.texture_layers()[0].any()
    m2 = np.frombuffer(grp2.buffer("materials"), np.float32).reshape(-1, 16)
    assert np.allclose(m2[1, 4:8], [0.832472, 0.8072, 0.719802, 10.0])


def test_builtin_default_material_when_none_set():
    from gdpathtracing_b200 import scenes
    sc = scenes.cornell32()
    sc.default_material = -1
    grp = scenes.populate(sc)
    grp.build()
    m = np.frombuffer(grp.buffer("materials"), np.float32).reshape(-1, 16)
    assert np.allclose(m[0, :3], 0.5) and m[0, 9] == 0.5 and m[0, 8] == 0.0   # geometry_group3d.cpp:239-247


def _big_degenerate(n=40_000):
    """Half of the triangles coincide (one centroid), the rest spread along x only: flat axes, one-sided partitions
    and the nth_element fallback (bvh.cpp:54-55,170-177) at sizes where the builder runs its multi-threaded top."""
    from gdpathtracing_b200 import scenes
    sc = scenes.SceneDesc("big_degenerate", camera_transform12=scenes.transform12(None, (0, 0, 6)), fov=60.0)
    sc.materials = [dict()]
    sc.default_material = 0
    rng = np.random.default_rng(11)
    x = np.where(np.arange(n) < n // 2, 0.0, rng.integers(0, 97, n) * 0.37).astype(np.float32)
    base = np.array([[-0.5, -0.5, 0.0], [0.5, -0.5, 0.0], [0.0, 0.5, 0.0]], np.float32)
    p = (base[None, :, :] + np.stack([x, np.zeros_like(x), np.zeros_like(x)], 1)[:, None, :]).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0, 0, -1], np.float32), (len(p), 1))
    sc.meshes = [[{"positions": p, "normals": nrm, "uvs": np.zeros((len(p), 2), np.float32),
                   "indices": np.arange(len(p), dtype=np.int32)}]]
    sc.instances = [dict(mesh=0)]
    return sc


def _buffers_with_threads(sc, threads):
    from gdpathtracing_b200 import scenes
    grp = scenes.populate(sc)
    grp.build_threads = threads
    assert grp.build_threads == threads
    grp.build()
    return {k: bytes(v) for k, v in grp.buffers().items()}


@pytest.mark.parametrize("make", [lambda: __import__("gdpathtracing_b200").scenes.triangle_soup(150_000, seed=9), _big_degenerate],
                         ids=["soup150k", "big_degenerate"])
def test_multithreaded_build_emits_the_single_thread_bytes(make):
    """the parallel build (thread-pooled top of the tree + independent subtrees) is byte-identical to the
    single-threaded one for every thread count."""
    sc = make()
    one = _buffers_with_threads(sc, 1)
    for threads in (2, 3, 8, 0):
        got = _buffers_with_threads(sc, threads)
        for key in one:
            assert got[key] == one[key], f"{key}: {threads} threads differ from 1 thread"


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libgdpt_refbvh.so not built")
def test_multithreaded_build_matches_reference_on_degenerate_input():
    sc = _big_degenerate()
    got, _ = product_buffers(sc)   # default: all hardware threads
    ref = reference_buffers(sc)
    for key in got:
        assert np.array_equal(got[key], ref[key]), f"{key}: differs from reference bvh.cpp"


def test_texture_layers_do_not_depend_on_the_thread_count():
    """The bilinear resize to the texture-array resolution (geometry_group3d.cpp:294-303) runs in bands of rows on the build
    threads; every pixel is computed from the source alone, so the layers are the single-thread bytes."""
    from gdpathtracing_b200 import scenes
    sc = scenes.demo_scene()
    layers = {}
    for threads in (1, 3, 8, 0):
        grp = scenes.populate(sc)
        grp.build_threads = threads
        grp.build()
        layers[threads] = [l.copy() for l in grp.texture_layers()]
    assert len(layers[1]) >= 1 and layers[1][0].shape[0] >= 64
    for threads in (3, 8, 0):
        assert len(layers[threads]) == len(layers[1])
        for a, b in zip(layers[threads], layers[1]):
            assert np.array_equal(a, b), f"{threads} threads"
