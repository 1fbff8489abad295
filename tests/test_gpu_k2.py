"""GPU tier: K2's quotients (k_progressive, pt_kernels.cu) on crafted inputs.  The kernel runs the six divisions of a pixel
(acc / frame_count, the ACES fraction) as its own reciprocal + FMA sequence when every operand is +0 or inside a window of
exponents, and as plain IEEE divisions otherwise; both must give the oracle's bytes (progressive_rendering.glsl:28-46:
each `/` one correctly rounded binary32 division).  Inputs: every byte value, accumulated sums over 60 binades on both
sides of the window's edges, +0, tiny, huge, negative and non-finite values, and frame counts from 1 to 2^32 - 1."""
import numpy as np
import pytest

from gdpathtracing_b200 import PathTracingCamera, multigpu, scenes
from oracle import oracle

pytestmark = pytest.mark.gpu

FRAME_COUNTS = [1, 2, 3, 5, 7, 255, 256, 257, 1000, 65535, 2 ** 24 - 1, 2 ** 24, 2 ** 24 + 2, 2 ** 31, 2 ** 32 - 1]


def crafted_accumulation(rng, H, W):
    """Non-negative sums of k / 255 (what K2 itself produces), values across 2^-60 .. 2^60, exact window edges and their
    neighbours, +0 / -0, denormals, negative numbers."""
    acc = np.zeros((H, W, 4), np.float32)
    n = H * W * 3
    v = np.empty(n, np.float32)
    kinds = rng.integers(0, 8, n)
    sums = (rng.integers(0, 256, n).astype(np.float32) / np.float32(255.0)) * rng.integers(0, 4000, n).astype(np.float32)
    wide = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-60, 61, n)).astype(np.float32)
    edges = np.array([2.0 ** -40, 2.0 ** 40, 1.0, 0.0], np.float32).view(np.uint32)
    near = (edges[rng.integers(0, 4, n)].astype(np.int64) + rng.integers(-3, 4, n)).clip(0, 0x7F7FFFFF).astype(np.uint32).view(np.float32)
    v[:] = sums
    v[kinds == 1] = wide[kinds == 1]
    v[kinds == 2] = near[kinds == 2]
    v[kinds == 3] = 0.0
    tiny = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-149, -100, n)).astype(np.float32)
    v[kinds == 4] = tiny[kinds == 4]
    v[kinds == 5] = -wide[kinds == 5]
    acc[..., :3] = v.reshape(H, W, 3)
    acc[..., 3] = 1.0
    return acc


@pytest.mark.parametrize("finite", [True, False], ids=["finite", "with_inf_and_nan"])
def test_k2_quotients_on_crafted_inputs(finite):
    import torch
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    cam = PathTracingCamera()
    cam.fov = sc.fov; cam.geometry_group = grp; cam.denoising_mode = PathTracingCamera.NONE
    cam.set_window_size(64, 64); cam.set_global_transform(sc.camera_transform12); cam.set_max_depth(2)
    cam.init()
    k2 = multigpu.cuda_k2(cam)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(cam.stream())
    rng = np.random.default_rng(20 if finite else 21)
    H, W = 384, 500  # not a multiple of the 128-pixel pieces the blocks take
    for fc in FRAME_COUNTS:
        raw = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        raw[0, :256, 0] = np.arange(256); raw[1, :256, 1] = np.arange(256); raw[2, :256, 2] = np.arange(256)
        raw[3:6] = 0
        acc = crafted_accumulation(rng, H, W)
        if not finite:
            bad = rng.integers(0, 40, (H, W, 3))
            acc[..., :3][bad == 0] = np.inf
            acc[..., :3][bad == 1] = -np.inf
            acc[..., :3][bad == 2] = np.nan
        want_screen, want_acc = raw.copy(), acc.copy()
        oracle.progressive(want_screen, want_acc, fc)
        raw_t, acc_t = torch.from_numpy(raw).to(dev), torch.from_numpy(acc).to(dev)
        screen_t = torch.empty_like(raw_t)
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            k2(raw_t, screen_t, acc_t, fc)
        torch.cuda.synchronize()
        got_screen, got_acc = screen_t.cpu().numpy(), acc_t.cpu().numpy()
        nan = np.isnan(want_acc)
        assert np.array_equal(np.isnan(got_acc), nan), f"frame_count {fc}"
        assert np.array_equal(got_acc.view(np.uint32)[~nan], want_acc.view(np.uint32)[~nan]), f"frame_count {fc}: accumulation"
        differ = (got_screen != want_screen).any(axis=-1)
        assert not differ.any(), (f"frame_count {fc}: {int(differ.sum())} pixels differ, first at {np.argwhere(differ)[0]}: "
                                  f"acc {acc[tuple(np.argwhere(differ)[0])]} raw {raw[tuple(np.argwhere(differ)[0])]} "
                                  f"got {got_screen[tuple(np.argwhere(differ)[0])]} want {want_screen[tuple(np.argwhere(differ)[0])]}")
