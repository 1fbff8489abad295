"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The path shards with NO collective inside a frame: pixels are independent and the seed depends
only on (x, y, frame_index) (main.glsl:409), the scene is replicated.  Two partitions:

  sample-index  rank r renders whole frames with frame_index = step * world + r + 1 and keeps its
                own accumulation; a presented image needs one sum-reduce of the accumulations.
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
        a one-element NCCL all-reduce on the backend's own stream (frames are separated by it)."""
        if self.world > 1:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self._token)

    def close(self):
        if self.opened:
            self.cam.set_peer_screens([])
            self.cam.synchronize()
            dist.barrier()
            for p in self.opened:
                self.cam.close_peer_image(p)
            self.opened = []


def reduce_accumulations(accum, dst=0):
    """Sum the per-rank RGBA32F accumulation buffers onto `dst` (sample-index partition)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
