"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The path shards with NO collective inside a frame: pixels are independent and the seed depends
only on (x, y, frame_index) (main.glsl:409), the scene is replicated.  Two partitions:

  sample-index  rank r renders whole frames with frame_index = step * world + r + 1 (K1 only).  The
                accumulation of the reference is a chain of float additions in FRAME order
                (progressive_rendering.glsl:33-38), so summing per-rank accumulations would differ from it by
                rounding.  Instead every rank takes a block of rows: the ranks swap the row blocks of their
                frames (all-to-all over NVLink) and each rank runs K2 over its rows in frame order --
                bit-identical to one GPU accumulating the same frames (SampleIndexAccumulator).
  row bands     every rank renders the rows y with (y // band) % world == rank of the SAME frame
                (bit-identical to the single-GPU frame); a presented image needs one all-gather
                of the RGBA8 bands.

PyTorch is used here for what it is good at -- wrapping device memory, NCCL process groups --
not for any rendering.
"""
import numpy as np
import torch
import torch.distributed as dist


class _DevicePointer:
    """Expose a raw device allocation to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def as_tensor(ptr, shape, dtype, device):
    typestr = {torch.uint8: "|u1", torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevicePointer(ptr, shape, typestr), device=device)


def band_rows(height, band, part, parts):
    """Row indices owned by `part` (same rule as the kernels: (y // band) % parts == part)."""
    y = np.arange(height)
    return y[((y // band) % parts) == part]


def init_process_group(backend=None):
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def gather_row_bands(frame, band, rank, world):
    """frame: (H, W, 4) uint8 tensor holding this rank's rows; returns the assembled frame on every rank.
    Ragged ownership (H not a multiple of band * world) is handled by padding each contribution to
    the largest row count."""
    if world == 1:
        return frame
    H = frame.shape[0]
    rows = [torch.as_tensor(band_rows(H, band, r, world), device=frame.device) for r in range(world)]
    most = max(len(r) for r in rows)
    mine = torch.zeros((most,) + tuple(frame.shape[1:]), dtype=frame.dtype, device=frame.device)
    mine[:len(rows[rank])] = frame[rows[rank]]
    flat = torch.empty((world * most,) + tuple(mine.shape[1:]), dtype=frame.dtype, device=frame.device)
    dist.all_gather_into_tensor(flat, mine)  # concatenated along dim 0: the form NCCL and gloo both accept
    gathered = flat.view((world,) + tuple(mine.shape))
    out = torch.empty_like(frame)
    for r in range(world):
        out[rows[r]] = gathered[r, :len(rows[r])]
    return out


class PeerFrame:
    """Row-band frame whose presentation needs no gather: every rank's accumulate/tone-map kernel also writes its rows
    into the other ranks' output images over NVLink (CUDA IPC peer memory), so after `barrier()` each rank's image is
    the whole frame.  One process per GPU; handles travel through torch.distributed."""

    def __init__(self, cam, rank, world):
        self.cam, self.rank, self.world, self.opened = cam, rank, world, []
        if world == 1:
            return
        handles = [None] * world
        dist.all_gather_object(handles, cam.export_output_handle())
        self.opened = [cam.open_peer_image(h) for r, h in enumerate(handles) if r != rank]
        cam.set_peer_screens(self.opened)
        self._token = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        self._stream = torch.cuda.ExternalStream(cam.stream())

    def barrier(self):
        """All ranks' kernels of the frame enqueued so far have completed once this returns on the device side:
        a one-element NCCL all-reduce on the backend's own stream.  After it every rank's image holds frame N."""
        if self.world > 1:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self._token)

    def frame_consumed(self):
        """Second half of the handshake: call once this rank's reads of frame N (read-back, copy, display) are enqueued
        on the backend's stream or finished, BEFORE the next frame is launched.  A peer's K2 of frame N+1 writes into
        this rank's image; without this collective a fast peer could overwrite rows this rank is still reading
        (write-after-read).  Same one-element all-reduce: nobody passes it until everybody has consumed frame N."""
        self.barrier()

    def close(self):
        if self.opened:
            self.cam.set_peer_screens([])
            self.cam.synchronize()
            dist.barrier()
            for p in self.opened:
                self.cam.close_peer_image(p)
            self.opened = []


def row_blocks(height, parts):
    """[begin, end) of the contiguous block of rows every part accumulates (sample-index partition)."""
    edges = [(height * p) // parts for p in range(parts + 1)]
    return [(edges[p], edges[p + 1]) for p in range(parts)]


class SampleIndexAccumulator:
    """Progressive accumulation of the sample-index partition, bit-identical to the sequential one.

    Rank r renders frames r+1, r+1+world, ... (K1 only).  `add(frames)` takes this rank's next `n` frames
    ((n, H, W, 4) uint8, global indices first+r, first+r+world, ...; every rank passes the same n), sends rows
    [b_d, e_d) of each to rank d and receives its own rows of everybody's frames, then runs K2
    (progressive_rendering.glsl:28-46) over its row block once per frame IN FRAME ORDER, with frame_count = the
    frame's global index.  The float additions are those of a single GPU accumulating frames 1, 2, 3, ... -- only
    spread over the ranks by rows.  `present()` all-gathers the tone-mapped row blocks into the whole frame.

    `k2(raw, screen, accum, frame_count)` runs one K2 over a (rows, W) block: the CUDA backend's
    gdpt_progressive_accumulate on the GPU (cuda_k2), the oracle in the CPU tests."""

    def __init__(self, height, width, rank, world, device, k2):
        self.H, self.W, self.rank, self.world, self.device, self.k2 = height, width, rank, world, device, k2
        self.blocks = row_blocks(height, world)
        b, e = self.blocks[rank]
        self.accum = torch.zeros((e - b, width, 4), dtype=torch.float32, device=device)
        self.screen = torch.zeros((e - b, width, 4), dtype=torch.uint8, device=device)
        self.frames_done = 0
        self.bytes_exchanged = 0

    def add(self, frames):
        n = int(frames.shape[0])
        b, e = self.blocks[self.rank]
        if self.world == 1:
            mine = [frames]
        else:
            mine = [frames[:, b:e] if s == self.rank else torch.empty((n, e - b, self.W, 4), dtype=torch.uint8, device=self.device)
                    for s in range(self.world)]
            ops = []
            for peer in range(self.world):
                if peer == self.rank:
                    continue
                pb, pe = self.blocks[peer]
                send = frames[:, pb:pe].contiguous()
                ops.append(dist.P2POp(dist.isend, send, peer))
                ops.append(dist.P2POp(dist.irecv, mine[peer], peer))
                self.bytes_exchanged += send.numel()
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            mine[self.rank] = frames[:, b:e].contiguous()
        for j in range(n):          # frame order: global index first + j * world + s
            for s in range(self.world):
                self.frames_done += 1
                self.k2(mine[s][j], self.screen, self.accum, self.frames_done)
        return self.frames_done

    def present(self):
        """The tone-mapped frame after the frames added so far, assembled on every rank."""
        if self.world == 1:
            return self.screen
        most = max(e - b for b, e in self.blocks)
        pad = torch.zeros((most, self.W, 4), dtype=torch.uint8, device=self.device)
        pad[: self.screen.shape[0]] = self.screen
        flat = torch.empty((self.world * most, self.W, 4), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(flat, pad)
        out = torch.empty((self.H, self.W, 4), dtype=torch.uint8, device=self.device)
        for r, (b, e) in enumerate(self.blocks):
            out[b:e] = flat[r * most: r * most + (e - b)]
        return out


def cuda_k2(cam):
    """K2 of the CUDA backend on torch tensors of `cam`'s device (SampleIndexAccumulator's k2 on the GPU)."""
    from ._lib import check, cuda

    def k2(raw, screen, accum, frame_count):
        assert raw.is_contiguous() and screen.is_contiguous() and accum.is_contiguous()
        h, w = int(raw.shape[0]), int(raw.shape[1])
        check(cuda.gdpt_progressive_accumulate(cam.device, raw.data_ptr(), screen.data_ptr(), accum.data_ptr(), w, h, int(frame_count)),
              cam.device, "progressive_accumulate")
    return k2


def reduce_accumulations(accum, dst=0):
    """Sum the per-rank RGBA32F accumulation buffers onto `dst` (sample-index partition)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
