"""GPU tier of K3 (TEMPORAL_REPROJECTION, path_tracing_camera.cpp:215-221): k_temporal through the C-ABI -- fused
frame call, the reference's dispatch-by-dispatch sequence and the pipelined call -- against oracle K1 + oracle K3,
bit for bit: presented frame, both ping-pong history buffers, and the 88 B Params block the host computed."""
import numpy as np
import pytest

from gdpathtracing_b200 import PathTracingCamera, _lib, scenes
from oracle import oracle
from test_gpu_parity import make_camera, oracle_scene

pytestmark = pytest.mark.gpu

IDENTITY16 = np.eye(4, dtype=np.float32).reshape(16)


def camera_path(sc, frames):
    """Rest for three frames, then dolly/pan a little every frame (forces real reprojection and depth rejections)."""
    t = np.asarray(sc.camera_transform12, np.float32).copy()
    out = []
    for f in range(frames):
        if f >= 3:
            t = t.copy()
            t[9] += np.float32(0.11); t[10] -= np.float32(0.05); t[11] -= np.float32(0.21)
        out.append(t)
    return out


def oracle_sequence(sc, grp, W, H, depth, poses, cams):
    """Frame by frame: oracle K1 with the camera block the product used, the oracle's own delta matrix, oracle K3."""
    osc = oracle_scene(grp)
    fb1, fb2 = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
    prev_vp = IDENTITY16.copy()
    frame_count = 1
    for f, (pose, cam_block, params) in enumerate(zip(poses, cams["block"], cams["params"])):
        ref = oracle.path_trace(osc, W, H, cam_block, max_depth=depth)
        vp = np.frombuffer(cam_block, np.float32, 16).copy()
        delta = oracle.temporal_delta(prev_vp, vp)
        prev_vp = vp
        frame_count += 1
        want = _lib.TemporalParams()
        for i in range(16):
            want.delta_matrix[i] = float(delta[i])
        want.width, want.height, want.frame_count = W, H, frame_count
        want.blend_factor, want.near_plane, want.far_plane = 0.75, 0.01, 1000.0
        assert bytes(want) == params, f"frame {f}: host Params block differs from the oracle's"
        screen = ref["rgba8"].copy()
        oracle.temporal(params, screen, ref["depth"], fb1, fb2)
        yield f, screen, fb1, fb2


@pytest.mark.parametrize("how", ["fused", "split", "pipelined"])
def test_temporal_reprojection_frames(how):
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    W, H, depth, frames = 256, 256, 4, 7
    poses = camera_path(sc, frames)
    cam = make_camera(sc, grp, W, H, depth, mode=PathTracingCamera.TEMPORAL_REPROJECTION, fused=(how != "split"))
    got, record = [], {"block": [], "params": []}
    in_flight = 0
    for pose in poses:
        cam.set_global_transform(pose)
        if how == "pipelined":
            cam.render_begin()
            in_flight += 1
            record["block"].append(bytes(cam.camera_block())); record["params"].append(bytes(cam.temporal_params()))
            if in_flight == 2:
                got.append(cam.render_wait()[0].copy()); in_flight -= 1
        else:
            got.append(cam.render().copy())
            record["block"].append(bytes(cam.camera_block())); record["params"].append(bytes(cam.temporal_params()))
    while in_flight:
        got.append(cam.render_wait()[0].copy()); in_flight -= 1
    moved = 0
    for f, screen, fb1, fb2 in oracle_sequence(sc, grp, W, H, depth, poses, record):
        assert np.array_equal(got[f], screen), f"{how}: frame {f} differs in {(got[f] != screen).any(axis=2).sum()} pixels"
        if f >= 3:
            moved += int((got[f] != got[2]).any(axis=2).sum())
    assert moved > 0, "the moving part of the sequence must change the picture"
    assert np.array_equal(cam.read_image("history1").view(np.uint32), fb1.view(np.uint32))
    assert np.array_equal(cam.read_image("history2").view(np.uint32), fb2.view(np.uint32))


def test_temporal_on_the_demo_scene_and_mode_switch():
    """Demo scene at 480x270, depth 8; then the same camera switches to progressive rendering and back (each post
    process keeps its own state, path_tracing_camera.cpp:206-226)."""
    sc = scenes.demo_scene()
    grp = scenes.populate(sc)
    W, H, depth = 480, 270, 8
    poses = camera_path(sc, 5)
    cam = make_camera(sc, grp, W, H, depth, mode=PathTracingCamera.TEMPORAL_REPROJECTION)
    got, record = [], {"block": [], "params": []}
    for pose in poses:
        cam.set_global_transform(pose)
        got.append(cam.render().copy())
        record["block"].append(bytes(cam.camera_block())); record["params"].append(bytes(cam.temporal_params()))
    for f, screen, fb1, fb2 in oracle_sequence(sc, grp, W, H, depth, poses, record):
        assert np.array_equal(got[f], screen), f"frame {f}"
    cam.denoising_mode = PathTracingCamera.PROGRESSIVE_RENDERING
    a = cam.render().copy()
    osc = oracle_scene(grp)
    ref = oracle.path_trace(osc, W, H, bytes(cam.camera_block()), max_depth=depth)
    screen, acc = ref["rgba8"].copy(), np.zeros((H, W, 4), np.float32)
    oracle.progressive(screen, acc, cam.last_frame_count())
    assert cam.last_frame_count() == 1 and np.array_equal(a, screen)
    cam.denoising_mode = PathTracingCamera.TEMPORAL_REPROJECTION
    cam.render()
    assert cam.temporal_params().frame_count == 7  # 1 + six temporal dispatches


def test_temporal_refuses_a_row_sharded_frame():
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    cam = make_camera(sc, grp, 64, 64, 2, mode=PathTracingCamera.TEMPORAL_REPROJECTION, shard=(0, 2, 8))
    with pytest.raises(_lib.GdptError):
        cam.render_begin()
