"""CPU tier of K3 (temporal reprojection, temporal_reprojection.glsl:31-71 + temporal_reprojection.cpp:53-69):
the kernel's per-pixel function (pt_post.cuh, compiled for the host by tests/devcheck) against the oracle bit for
bit, the host parameter update against the oracle's restatement of the godot-cpp matrix arithmetic, and the
size-independent properties of the filter."""
import ctypes

import numpy as np
import pytest

from conftest import ptr
from gdpathtracing_b200 import _lib, nodes, scenes
from oracle import oracle

IDENTITY16 = np.eye(4, dtype=np.float32).reshape(16)


def params_block(delta16, w, h, frame_count):
    p = _lib.TemporalParams()
    for i, v in enumerate(np.asarray(delta16, np.float32).reshape(16)):
        p.delta_matrix[i] = float(v)
    p.width, p.height, p.frame_count = w, h, frame_count
    p.blend_factor, p.near_plane, p.far_plane = 0.75, 0.01, 1000.0
    return p


def random_images(rng, w, h):
    screen = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    screen[..., 3] = 255
    depth = rng.uniform(0.9, 1.0, (h, w)).astype(np.float32)
    fb1 = rng.uniform(0.0, 4.0, (h, w, 4)).astype(np.float32)
    fb2 = rng.uniform(0.0, 4.0, (h, w, 4)).astype(np.float32)
    return screen, depth, fb1, fb2


def delta_cases(rng):
    shift = IDENTITY16.copy(); shift[12] = 0.031; shift[13] = -0.017          # a pan of a few pixels
    zoom = IDENTITY16.copy(); zoom[0] = 1.07; zoom[5] = 0.93; zoom[14] = 0.04  # scale + depth offset (fails the 0.1 test partly)
    far = IDENTITY16.copy(); far[12] = 5.0                                     # everything leaves the image
    wzero = IDENTITY16.copy(); wzero[15] = 0.0                                 # division by w == 0 -> inf / NaN positions
    nan = IDENTITY16.copy(); nan[0] = np.nan
    huge = IDENTITY16.copy(); huge[0] = 3e38; huge[5] = -3e38                  # float -> int conversion out of range
    rnd = rng.normal(0, 1, 16).astype(np.float32)
    return {"identity": IDENTITY16, "pan": shift, "zoom": zoom, "off_screen": far, "w_zero": wzero, "nan": nan,
            "huge": huge, "random": rnd}


@pytest.mark.parametrize("frame_count", [2, 3, 0])
@pytest.mark.parametrize("size", [(64, 48), (33, 7), (1, 1)])
def test_temporal_pixel_function_matches_oracle(devcheck, size, frame_count):
    w, h = size
    rng = np.random.default_rng(1234 + w * 7 + frame_count)
    for name, delta in delta_cases(rng).items():
        screen, depth, fb1, fb2 = random_images(rng, w, h)
        p = params_block(delta, w, h, frame_count)
        a = [screen.copy(), fb1.copy(), fb2.copy()]
        b = [screen.copy(), fb1.copy(), fb2.copy()]
        oracle.temporal(bytes(p), a[0], depth, a[1], a[2])
        devcheck.devcheck_temporal(ctypes.byref(p), ptr(b[0]), ptr(depth), ptr(b[1]), ptr(b[2]))
        for x, y, what in zip(a, b, ("screen", "frameBuffer1", "frameBuffer2")):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), f"{name}: {what} differs"


def test_ping_pong_roles_follow_frame_count_parity():
    """Even frameCount reads frameBuffer1 and writes frameBuffer2, odd the other way round (temporal_reprojection.glsl:46,62,66)."""
    rng = np.random.default_rng(5)
    w, h = 16, 8
    screen, depth, fb1, fb2 = random_images(rng, w, h)
    for fc, read, written in ((2, "fb1", "fb2"), (3, "fb2", "fb1")):
        s, bufs = screen.copy(), {"fb1": fb1.copy(), "fb2": fb2.copy()}
        oracle.temporal(bytes(params_block(IDENTITY16, w, h, fc)), s, depth, bufs["fb1"], bufs["fb2"])
        before = {"fb1": fb1, "fb2": fb2}
        assert np.array_equal(bufs[read], before[read]), "the history buffer must stay untouched"
        cur = screen[..., :3].astype(np.float32) / np.float32(255.0)
        want = cur * np.float32(0.25) + before[read][..., :3] * np.float32(0.75)
        assert np.array_equal(bufs[written][..., :3], want) and np.all(bufs[written][..., 3] == 1.0)


def test_static_camera_converges_to_the_input_colour():
    """With an identity delta and a constant input the history is h_n = 0.25 c + 0.75 h_(n-1): a geometric approach to c."""
    w, h = 8, 4
    screen0 = np.zeros((h, w, 4), np.uint8); screen0[..., 0] = 200; screen0[..., 1] = 50; screen0[..., 3] = 255
    depth = np.full((h, w), 0.95, np.float32)
    fb1, fb2 = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    c = screen0[0, 0, :3].astype(np.float32) / np.float32(255.0)
    for n, fc in enumerate(range(2, 42), start=1):
        s = screen0.copy()
        oracle.temporal(bytes(params_block(IDENTITY16, w, h, fc)), s, depth, fb1, fb2)
        latest = fb2 if fc % 2 == 0 else fb1
        assert np.allclose(latest[0, 0, :3], c * (1.0 - 0.75 ** n), rtol=1e-5, atol=1e-6)
    assert np.allclose(latest[..., :3], c, atol=1e-4)


def test_depth_mismatch_and_off_screen_fall_back_to_the_current_colour():
    w, h = 8, 8
    rng = np.random.default_rng(9)
    screen, depth, fb1, fb2 = random_images(rng, w, h)
    cur = screen[..., :3].astype(np.float32) / np.float32(255.0)
    aces = lambda x: (x * (np.float32(2.51) * x + np.float32(0.03))) / (x * (np.float32(2.43) * x + np.float32(0.59)) + np.float32(0.14))
    for delta in (delta_cases(rng)["off_screen"], np.where(np.arange(16) == 14, np.float32(0.5), IDENTITY16).astype(np.float32)):
        s, a, b = screen.copy(), fb1.copy(), fb2.copy()
        oracle.temporal(bytes(params_block(delta, w, h, 2)), s, depth, a, b)
        assert np.array_equal(b[..., :3], cur * np.float32(0.25) + cur * np.float32(0.75))  # mix(c, c, 0.75), not necessarily == c
        assert np.array_equal(s[..., :3], np.rint(np.clip(aces(b[..., :3]), 0, 1) * np.float32(255.0)).astype(np.uint8))


def test_host_parameter_update_matches_the_oracle_restatement():
    """TemporalReprojection::render's delta matrix (host C++ twin) == the oracle's composition of the same godot-cpp operations."""
    sc = scenes.demo_scene()
    W, H = 320, 180
    prev = IDENTITY16.copy()  # a default-constructed Projection (temporal_reprojection.h:52)
    rng = np.random.default_rng(3)
    t = np.asarray(sc.camera_transform12, np.float32).copy()
    for step in range(6):
        vp, delta = nodes.make_temporal_delta(prev, t, sc.fov, W, H)
        cam = nodes.make_camera_block(t, sc.fov, W, H, 1)
        assert np.array_equal(vp, np.array(cam.vp[:], np.float32)), "same view-projection as the camera block (render_parameters.h:27-33)"
        want = oracle.temporal_delta(prev, vp)
        assert np.array_equal(delta.view(np.uint32), want.view(np.uint32)), f"step {step}"
        assert delta[3] == 0 and delta[7] == 0 and delta[11] == 0 and delta[15] == 1  # the projective row is dropped (:58)
        if step >= 2 and step % 2 == 0:
            pass  # camera rests: the next delta must be close to identity
        else:
            t[9:12] += rng.normal(0, 0.2, 3).astype(np.float32)
        prev = vp
    # resting camera: delta ~ identity (not exactly: fl(vp * vp^-1))
    vp, delta = nodes.make_temporal_delta(prev, t, sc.fov, W, H)
    vp2, delta2 = nodes.make_temporal_delta(vp, t, sc.fov, W, H)
    assert np.allclose(delta2, IDENTITY16, atol=2e-3)
