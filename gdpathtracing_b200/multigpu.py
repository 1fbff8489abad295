"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The path shards with NO collective inside a frame: pixels are independent and the seed depends
only on (x, y, frame_index) (main.glsl:409), the scene is replicated.  Two partitions:

  sample-index  rank r renders whole frames with frame_index = step * world + r + 1 (K1 only).  The
                accumulation of the reference is a chain of float additions in FRAME order
                (progressive_rendering.glsl:33-38), so summing per-rank accumulations would differ from it by
                rounding.  Instead every rank takes a block of rows and runs K2 over its rows of EVERY rank's
                frames in frame order -- bit-identical to one GPU accumulating the same frames
                (SampleIndexAccumulator).  On GPUs the exchange is K2's own loads: the frames sit in device memory
                the other ranks map through CUDA IPC (PeerFrameStore, add_from_store); add() is the same with NCCL
                send/recv of the row blocks (the form the gloo CPU tests run).
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

    def add_from_store(self, store, n=None):
        """The same accumulation with the exchange fused into K2: every rank's frames sit in a PeerFrameStore, and this rank
        runs K2 over its row block of frame j of rank s by READING that block from rank s's memory over NVLink.  `k2` must
        accept a device address (cuda_k2_at).  Frame order and arithmetic are those of add()."""
        n = store.n if n is None else int(n)
        b, e = self.blocks[self.rank]
        store.barrier()  # every rank's frames are complete
        for j in range(n):
            for s in range(self.world):
                self.frames_done += 1
                self.k2(store.block_address(s, j, b), e - b, self.W, self.screen, self.accum, self.frames_done)
        store.barrier()  # every rank has read what it needs: the stores may be overwritten
        self.bytes_exchanged += n * (self.world - 1) * (e - b) * self.W * 4  # read from peers
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


class PeerFrameStore:
    """The frames a rank keeps for the sample-index accumulation, in device memory every other rank can READ over NVLink
    (CUDA IPC peer memory): the owner of a row block runs K2 straight from the source rank's copy of the frame, so the
    exchange step is the accumulate kernel's own loads -- no send / recv, no staging copies.  One process per GPU; handles
    travel through torch.distributed.  `frames` is this rank's (n, H, W, 4) uint8 tensor over the store."""

    def __init__(self, cam, n, height, width, rank, world, device):
        from ._lib import check, cuda
        self.cam, self.n, self.H, self.W, self.rank, self.world = cam, int(n), int(height), int(width), rank, world
        size = self.n * self.H * self.W * 4
        self.rid = cuda.gdpt_device_create_buffer(cam.device, size)
        if not self.rid:
            raise RuntimeError("gdpt_device_create_buffer failed: " + (cuda.gdpt_last_error(cam.device) or b"").decode())
        import ctypes
        p, s = ctypes.c_uint64(), ctypes.c_uint64()
        check(cuda.gdpt_rid_device_pointer(cam.device, self.rid, ctypes.byref(p), ctypes.byref(s)), cam.device, "rid_device_pointer")
        self.frames = as_tensor(p.value, (self.n, self.H, self.W, 4), torch.uint8, device)
        self.base = [0] * world
        self.base[rank] = p.value
        self.opened = []
        if world > 1:
            buf = ctypes.create_string_buffer(64)
            check(cuda.gdpt_rid_ipc_export(cam.device, self.rid, buf), cam.device, "rid_ipc_export")
            handles = [None] * world
            dist.all_gather_object(handles, buf.raw)
            for q, h in enumerate(handles):
                if q != rank:
                    self.base[q] = cam.open_peer_image(h)
                    self.opened.append(self.base[q])
        self._token = torch.zeros(1, dtype=torch.int32, device=device)
        self._stream = torch.cuda.ExternalStream(cam.stream())

    def block_address(self, source_rank, frame, first_row):
        return self.base[source_rank] + ((frame * self.H + first_row) * self.W) * 4

    def barrier(self):
        """Device-side: everything every rank enqueued on its backend stream so far has completed once this has
        (a one-element NCCL all-reduce on that stream).  Called once after the frames were written (they may now be read)
        and once after they were read (they may now be overwritten)."""
        if self.world > 1:
            with torch.cuda.stream(self._stream):
                dist.all_reduce(self._token)

    def close(self):
        from ._lib import cuda
        self.cam.synchronize()
        if self.world > 1:
            dist.barrier()
        for p in self.opened:
            self.cam.close_peer_image(p)
        self.opened = []
        self.frames = None
        cuda.gdpt_device_free_buffer(self.cam.device, self.rid)


def cuda_k2(cam):
    """K2 of the CUDA backend on torch tensors of `cam`'s device (SampleIndexAccumulator's k2 on the GPU)."""
    from ._lib import check, cuda

    def k2(raw, screen, accum, frame_count):
        assert raw.is_contiguous() and screen.is_contiguous() and accum.is_contiguous()
        h, w = int(raw.shape[0]), int(raw.shape[1])
        check(cuda.gdpt_progressive_accumulate(cam.device, raw.data_ptr(), screen.data_ptr(), accum.data_ptr(), w, h, int(frame_count)),
              cam.device, "progressive_accumulate")
    return k2


def cuda_k2_at(cam):
    """K2 of the CUDA backend with the raw frame given as a device address (possibly a peer's, PeerFrameStore)."""
    from ._lib import check, cuda

    def k2(raw_address, rows, width, screen, accum, frame_count):
        check(cuda.gdpt_progressive_accumulate(cam.device, int(raw_address), screen.data_ptr(), accum.data_ptr(), int(width), int(rows),
                                               int(frame_count)), cam.device, "progressive_accumulate")
    return k2


def reduce_accumulations(accum, dst=0):
    """Sum the per-rank RGBA32F accumulation buffers onto `dst` (sample-index partition)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
