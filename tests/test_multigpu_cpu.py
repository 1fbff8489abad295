"""N > 1 host logic on CPU: two gloo ranks (127.0.0.1) run the same partition / gather / reduce code the
NCCL path runs (gdpathtracing_b200/multigpu.py).  The per-rank pixels come from the oracle (test
infrastructure) because there is no GPU here; what is under test is the partition arithmetic: the row-band
rule the kernels use, ragged ownership (bit-exact reassembly), and that sample-index accumulation reduces to
the sequential accumulation within float re-association error (stated in the test)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
W, H, DEPTH, FRAMES = 64, 44, 3, 4  # 44 rows with band 8 over 2 ranks: ragged (24 vs 20 rows)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _frames():
    from gdpathtracing_b200 import nodes, scenes
    from oracle import oracle
    sc = scenes.cornell32()
    grp = scenes.populate(sc)
    grp.build()
    osc = oracle.Scene(grp.buffers(), grp.texture_layers())
    out = []
    for f in range(1, FRAMES + 1):
        cam = bytes(nodes.make_camera_block(sc.camera_transform12, sc.fov, W, H, f))
        out.append(oracle.path_trace(osc, W, H, cam, max_depth=DEPTH)["rgba8"].copy())
    return out


def _worker(rank, world, port, result_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from gdpathtracing_b200 import multigpu
    from oracle import oracle
    r, _, w = multigpu.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    frames = _frames()

    # ---- row bands: every rank holds only the rows it owns of frame 1
    band = 8
    mine = multigpu.band_rows(H, band, rank, world)
    part = np.zeros_like(frames[0])
    part[mine] = frames[0][mine]
    whole = multigpu.gather_row_bands(torch.from_numpy(part), band, rank, world).numpy()
    assert np.array_equal(whole, frames[0]), "row-band gather does not reassemble the frame"

    # ---- sample index: rank r accumulates frames r, r+world, ...; one sum-reduce; tone-map on rank 0
    acc = np.zeros((H, W, 4), np.float32)
    n_mine = 0
    for i in range(rank, FRAMES, world):
        screen = frames[i].copy()
        n_mine += 1
        oracle.progressive(screen, acc, n_mine)  # acc += rgba8(screen) (frame_count 1 resets)
    t = torch.from_numpy(acc)
    multigpu.reduce_accumulations(t, dst=0)
    if rank == 0:
        seq = np.zeros((H, W, 4), np.float32)
        for i in range(FRAMES):
            oracle.progressive(frames[i].copy(), seq, i + 1)
        got = t.numpy()
        # Not bit-equal by construction: the addends are fl(k/255), so float addition in a different order rounds
        # differently (a few ulp).  Stated tolerance: 4 ulp of the sum (2^-21 relative).
        assert np.allclose(got[..., :3], seq[..., :3], rtol=2.0 ** -21, atol=0.0), \
            "sum of per-rank accumulations is not the sequential accumulation within 4 ulp"
        assert not np.array_equal(got[..., :3], np.zeros_like(got[..., :3]))
        open(os.path.join(result_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_band_rule_partitions_every_row_exactly_once():
    from gdpathtracing_b200 import multigpu
    for height, band, parts in [(1080, 8, 8), (2160, 32, 4), (44, 8, 2), (7, 4, 3)]:
        rows = np.concatenate([multigpu.band_rows(height, band, p, parts) for p in range(parts)])
        assert sorted(rows.tolist()) == list(range(height))


@pytest.mark.timeout(300)
def test_two_gloo_ranks_gather_bands_and_reduce_samples(tmp_path, built):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
