"""N > 1 host logic on CPU: two gloo ranks (127.0.0.1) run the same partition / gather / reduce code the
NCCL path runs (gdpathtracing_b200/multigpu.py).  The per-rank pixels come from the oracle (test
infrastructure) because there is no GPU here; what is under test is the partition arithmetic: the row-band
rule the kernels use, ragged ownership (bit-exact reassembly), and that the sample-index accumulation (row blocks
swapped between the ranks, K2 in frame order) is bit-identical to the sequential accumulation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
W, H, DEPTH, FRAMES = 64, 45, 3, 8  # 45 rows: ragged row bands (band 8 over 2 ranks: 24 vs 21 rows) and row blocks (22 vs 23)


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

    # ---- sample index: rank r renders frames r+1, r+1+world, ...; the ranks swap row blocks of their frames and every
    # rank accumulates its rows in frame order (SampleIndexAccumulator): the float additions of ONE GPU accumulating
    # frames 1, 2, 3, ... -- bit-identical accumulation buffer and presented frame, no tolerance
    def k2(raw, screen, accum, frame_count):
        s_np, a_np = raw.numpy().copy(), accum.numpy()
        oracle.progressive(s_np, a_np, frame_count)  # in place on the tensors' memory
        screen.copy_(torch.from_numpy(s_np))

    acc = multigpu.SampleIndexAccumulator(H, W, rank, world, torch.device("cpu"), k2)
    mine_frames = torch.from_numpy(np.stack([frames[i] for i in range(rank, FRAMES, world)]))
    half = mine_frames.shape[0] // 2
    acc.add(mine_frames[:half])          # two batches: the frame counter carries over
    first_presented = acc.present().numpy().copy()
    acc.add(mine_frames[half:])
    presented = acc.present().numpy()
    seq = np.zeros((H, W, 4), np.float32)
    seq_screens = []
    for i in range(FRAMES):
        screen = frames[i].copy()
        oracle.progressive(screen, seq, i + 1)
        seq_screens.append(screen)
    b, e = acc.blocks[rank]
    assert np.array_equal(acc.accum.numpy().view(np.uint32), seq[b:e].view(np.uint32)), "accumulation rows differ from the sequential accumulation"
    assert np.array_equal(presented, seq_screens[FRAMES - 1]), "presented frame differs from the sequential one"
    assert np.array_equal(first_presented, seq_screens[half * world - 1])
    assert acc.frames_done == FRAMES and (world == 1 or acc.bytes_exchanged > 0)
    if rank == 0:
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
