"""CPU tier: the restatement (oracle/pt_oracle.cpp) against the REFERENCE'S OWN SHADER TEXT.

oracle/_ref/libgdpt_refshader.so is main.glsl + brdfs.glsl + progressive_rendering.glsl +
temporal_reprojection.glsl of the reference, compiled as C++ (oracle/ref_shader_bridge.cpp,
oracle/glsl_shim/).  Everything the CUDA kernels are compared with elsewhere -- frames, depth, radiance
before quantisation, per-segment hit ids, node-visit order, work counters, single-ray hit records, K2
accumulation, K3 -- is compared here between the restatement and that text, bit for bit.  Where the library
is absent (a box without /root/reference and without the prebuilt file) the committed SHA-256 fixtures of
the reference's outputs (tests/golden/ref_shader_hashes.json, written by tools/make_ref_shader_golden.py)
take its place.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from gdpathtracing_b200 import nodes, scenes
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_shader_hashes.json")

# (name, scene, W, H, depth, frame_index): C1 at full size, C2 at 480x270 depth 8, the shader's
# own depth 5, an instanced scene and a soup
CASES = [
    ("cornell32_256_d4", lambda: scenes.cornell32(), 256, 256, 4, 1),
    ("demo_480x270_d8", lambda: scenes.demo_scene(), 480, 270, 8, 1),
    ("demo_320x180_d5_f7", lambda: scenes.demo_scene(), 320, 180, 5, 7),
    ("instanced27x800_d5", lambda: scenes.instanced_grid(3, 800, seed=3), 192, 108, 5, 3),
    ("soup20k_d2", lambda: scenes.triangle_soup(20000, seed=1), 160, 90, 2, 2),
]
SEGS = 4
VISITS = 48
# fields the reference's traversal log yields (t / u / v / front come from trace_rays, max_stack is not observable)
LOG_FIELDS = ("hit", "triangle", "blas", "node_pops", "box_tests", "tri_tests", "tlas_leaves", "visit_hash_lo", "visit_hash_hi")
pin = pytest.mark.skipif(not oracle.ref_shader_available(), reason="oracle/_ref/libgdpt_refshader.so not built")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build(make):
    sc = make()
    grp = scenes.populate(sc)
    grp.build()
    return sc, oracle.Scene(grp.buffers(), grp.texture_layers())


def render(osc, sc, W, H, depth, frame, impl):
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, frame))
    return oracle.path_trace(osc, W, H, cam, max_depth=depth, trace_segments=SEGS, visits_per_ray=VISITS, radiance=True, impl=impl)


def frame_digest(r):
    """What a frame is summarised by in the committed fixtures."""
    tr = r["trace"]
    live = tr["hit"] != 0xFFFFFFFF
    d = {"rgba8": sha(r["rgba8"]), "depth": sha(r["depth"]), "radiance": sha(r["radiance"]), "visits": sha(r["visits"]),
         "rays": int(r["stats"]["rays"]), "primary_hits": int(r["stats"]["primary_hits"])}
    for k in ("node_pops", "box_tests", "tri_tests", "tlas_leaves"):
        d[k] = int(r["stats"][k])
    for f in LOG_FIELDS:
        d["trace_" + f] = sha(np.where(live, tr[f], 0))
    return d


@pin
@pytest.mark.parametrize("name,make,W,H,depth,frame", CASES, ids=[c[0] for c in CASES])
def test_restatement_equals_reference_shader_text(name, make, W, H, depth, frame):
    sc, osc = build(make)
    a = render(osc, sc, W, H, depth, frame, "restatement")
    b = render(osc, sc, W, H, depth, frame, "reference")
    assert b["stats"]["primary_hits"] > 0
    assert np.array_equal(a["rgba8"], b["rgba8"]), "RGBA8 frame"
    assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32)), "depth bits"
    assert np.array_equal(a["radiance"].view(np.uint32), b["radiance"].view(np.uint32)), "radiance bits before quantisation"
    for k in ("rays", "primary_hits", "node_pops", "box_tests", "tri_tests", "tlas_leaves"):
        assert a["stats"][k] == b["stats"][k], k
    assert np.array_equal(a["trace"]["hit"], b["trace"]["hit"]), "which segments exist / hit"
    live = b["trace"]["hit"] != 0xFFFFFFFF
    for f in LOG_FIELDS:
        assert np.array_equal(a["trace"][f][live], b["trace"][f][live]), f"per-segment {f}"
    assert np.array_equal(a["visits"], b["visits"]), "node-visit order of the camera rays"


def probe_rays(sc, n, seed):
    """Rays that stress the traversal: random ones through the scene, rays from the camera position, axis-parallel
    rays (zero direction components: 1/0 = inf in rD, 0*inf = NaN in the slab test) and rays lying in the planes of
    the Cornell walls."""
    r = scenes.splitmix64_floats(seed, n * 6).reshape(n, 6)
    o = ((r[:, :3] * 2 - 1) * 6).astype(np.float32)
    d = (r[:, 3:] * 2 - 1).astype(np.float32)
    cam_pos = np.asarray(sc.camera_transform12[9:12], np.float32)
    o[: n // 4] = cam_pos
    k = n // 8
    d[n // 4: n // 4 + k, 0] = 0.0                     # one zero component
    d[n // 4 + k: n // 4 + 2 * k, 1:] = 0.0            # axis-parallel, +-x
    d[n // 4 + 2 * k: n // 4 + 3 * k, 0] = 0.0
    d[n // 4 + 2 * k: n // 4 + 3 * k, 2] = 0.0         # axis-parallel, +-y
    o[n // 4 + 3 * k: n // 4 + 4 * k, 1] = np.float32(-3.0)   # origin in the plane of the room's floor (y = -5 * 0.6)
    d[n // 4 + 3 * k: n // 4 + 4 * k, 1] = 0.0
    return o, d


@pin
@pytest.mark.parametrize("name,make", [(c[0], c[1]) for c in CASES[:1] + CASES[1:2] + CASES[3:]],
                         ids=[c[0] for c in CASES[:1] + CASES[1:2] + CASES[3:]])
def test_single_ray_hit_records_equal_reference_shader_text(name, make):
    """ray_trace_tlas of the shader text on 20 000 rays: hit, triangle, instance, front, t / u / v bits, counters and
    visit hash equal the restatement's."""
    sc, osc = build(make)
    o, d = probe_rays(sc, 20000, 11)
    a = oracle.trace_rays(osc, o, d, "restatement")
    b = oracle.trace_rays(osc, o, d, "reference")
    assert 0 < b["hit"].sum() < len(b)
    for f in a.dtype.names:
        if f == "max_stack":
            continue
        x, y = a[f], b[f]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        assert np.array_equal(x, y), f"{f}: {(x != y).sum()} rays differ"


@pin
def test_debug_steps_mode_equals_reference_shader_text():
    """main.glsl:4 `#define DEBUG_STEPS`: the heat map of triangle tests per camera ray."""
    sc, osc = build(lambda: scenes.demo_scene())
    cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, 160, 90, 1))
    a = oracle.path_trace(osc, 160, 90, cam, debug_steps=True, radiance=True)
    b = oracle.path_trace(osc, 160, 90, cam, debug_steps=True, radiance=True, impl="reference")
    assert np.array_equal(a["rgba8"], b["rgba8"]) and a["stats"]["rays"] == b["stats"]["rays"] == 160 * 90
    assert np.array_equal(a["radiance"].view(np.uint32), b["radiance"].view(np.uint32))


@pin
def test_rng_of_the_shader_text_reproduces_the_known_answers():
    kat = json.load(open(os.path.join(HERE, "golden", "rng_kat.json")))
    lib = oracle.ref_shader()
    for v in kat["vectors"]:
        seed = np.zeros(2, np.uint32)
        lib.refsh_prng_seed(v["pixel"][0], v["pixel"][1], v["frame"], seed.ctypes.data)
        assert [int(x) for x in seed] == [int(x, 16) for x in v["seed"]]
        out = np.zeros(2, np.float32)
        lib.refsh_pcg2d(seed.ctypes.data, out.ctypes.data)
        assert [int(x) for x in seed] == [int(x, 16) for x in v["state1"]]
        assert [int(x) for x in out.view(np.uint32)] == [int(x, 16) for x in v["r_bits"]]
        lib.refsh_pcg2d(seed.ctypes.data, out.ctypes.data)
        assert [int(x) for x in seed] == [int(x, 16) for x in v["state2"]]


def progressive_run(impl, frames=6):
    """K1 + K2 for a fixed camera, frame_count 1, 2, ... (progressive_rendering.cpp:47-66 on a still camera that is
    not at the identity pose)."""
    sc, osc = build(lambda: scenes.cornell32())
    W = H = 96
    accum = np.zeros((H, W, 4), np.float32)
    out = []
    for f in range(1, frames + 1):
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, f))
        screen = oracle.path_trace(osc, W, H, cam, max_depth=5, impl=impl)["rgba8"].copy()
        oracle.progressive(screen, accum, f, impl=impl)
        out.append((screen.copy(), accum.copy()))
    return out


@pin
def test_progressive_accumulation_equals_reference_shader_text():
    for (sa, aa), (sb, ab) in zip(progressive_run("restatement"), progressive_run("reference")):
        assert np.array_equal(sa, sb), "tone-mapped frame"
        assert np.array_equal(aa.view(np.uint32), ab.view(np.uint32)), "accumulation buffer bits"


def temporal_run(impl):
    rng = np.random.default_rng(5)
    W, H = 64, 48
    outs = []
    fb1 = rng.random((H, W, 4), np.float32)
    fb2 = rng.random((H, W, 4), np.float32)
    for frame_count in (0, 1, 2, 3):
        screen = rng.integers(0, 256, (H, W, 4), np.uint8)
        depth = (rng.random((H, W), np.float32) * np.float32(0.2) + np.float32(0.4)).astype(np.float32)
        par = np.zeros(22, np.float32)
        m = np.eye(4, dtype=np.float32) + (rng.random((4, 4), np.float32) - np.float32(0.5)) * np.float32(0.05)
        m[3] = (0, 0, 0, 1)
        par[:16] = m.T.reshape(16)
        pv = par.view(np.uint32)
        pv[16], pv[17], pv[18] = W, H, frame_count
        par[19:22] = (0.9, 0.01, 1000.0)
        oracle.temporal(par.tobytes(), screen, depth, fb1, fb2, impl=impl)
        outs.append((screen.copy(), fb1.copy(), fb2.copy()))
    return outs


@pin
def test_temporal_reprojection_equals_reference_shader_text():
    for a, b in zip(temporal_run("restatement"), temporal_run("reference")):
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
        assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


@pytest.mark.parametrize("name,make,W,H,depth,frame", CASES, ids=[c[0] for c in CASES])
def test_restatement_reproduces_the_committed_reference_shader_fixtures(name, make, W, H, depth, frame):
    """Holds where the reference library cannot be built: digests of the REFERENCE shader's outputs, committed."""
    golden = json.load(open(GOLDEN))["frames"][name]
    sc, osc = build(make)
    got = frame_digest(render(osc, sc, W, H, depth, frame, "restatement"))
    for k, v in golden.items():
        assert got[k] == v, k
