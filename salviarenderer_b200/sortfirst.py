"""Sort-first multi-GPU frame assembly — host-side plumbing, one process per GPU.

Rank r of R owns the 64x64 screen tiles with (tx + 3*ty) % R == r (`slv_set_tile_shard`).  Every rank issues the same
command stream; geometry is replicated, the raster kernels only touch owned tiles.  Once per frame the owned tiles of
the RESOLVED colour surface have to reach rank 0.  Two transports:

* ``p2p`` (the CUDA product, default): rank 0 exports its resolved surface and its flag block as CUDA IPC handles; every
  other rank opens them and redirects `slv_resolve` into rank 0's memory, so the MSAA resolve and the exchange are ONE
  kernel storing over NVLink / NVSwitch — no staging buffer, no pack / unpack kernels, no collective.  Per frame the
  ranks' streams are ordered on the device by flags in rank 0's memory (`slv_peer_signal` / `slv_flags_wait`):
  rank r raises flags[r] = k+1 after its resolve of frame k; rank 0 waits for all of them ON ITS COPY STREAM
  (`slv_assembly_wait`: consumers of the assembled frame order behind it, rank 0's render stream goes straight on to the next
  frame in the other buffer), and raises flags[0] = k - nbuf + 1 there once the consumers of that buffer's previous frame are
  done, which the other ranks poll before they overwrite it with frame k.  No host synchronisation anywhere.
* ``gather`` (fallback, and what the CPU checkers use under gloo): `slv_pack_tiles` -> `torch.distributed.gather`
  (NCCL on the GPUs) -> `slv_unpack_tiles` on rank 0.

The staging buffers of the fallback are torch tensors: device tensors for the CUDA product (all work is enqueued on
the stream the library was given with `slv_set_stream`), CPU tensors for the checkers.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import abi


class HostFrame:
    """End-to-end sort-first assembly on the HOST: `nbuf` frames in POSIX shared memory, mapped by every rank's process and
    page-locked for its GPU (slv_host_register).  Rank r writes the tiles it owns of frame k into frames[k % nbuf] over its own
    PCIe link (slv_texture_export_tiles_async); the application reads the assembled frame through `view(k)` on any rank once
    every rank's export has landed (readback_wait on each rank + a barrier, or the ranks' own completion protocol).

    The owned tiles never travel between GPUs on this path - each rank exports what it rendered - so the device -> host
    bandwidth of a frame is N host links instead of rank 0's one."""

    def __init__(self, be: abi.Backend, nbytes: int, rank: int, nranks: int, nbuf: int = 2):
        import mmap
        import tempfile
        self.be, self.nbytes, self.rank, self.n, self.nbuf = be, nbytes, rank, nranks, nbuf
        page = mmap.PAGESIZE
        self.stride = (nbytes + page - 1) // page * page
        total = self.stride * nbuf
        name = [None]
        if rank == 0:
            shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            fd, path = tempfile.mkstemp(prefix="slv_hostframe_", dir=shm_dir)
            os.ftruncate(fd, total)
            name = [path]
        else:
            fd = -1
        if nranks > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self.map = mmap.mmap(fd, total)
        os.close(fd)
        import ctypes
        self.base = ctypes.addressof(ctypes.c_char.from_buffer(self.map))
        be.host_register(self.base, total)
        if nranks > 1:
            dist.barrier()
        if rank == 0:
            os.unlink(self.path)  # every rank holds its own mapping now

    def ptr(self, k: int) -> int:
        return self.base + (k % self.nbuf) * self.stride

    def view(self, k: int):
        import numpy as np
        off = (k % self.nbuf) * self.stride
        return np.frombuffer(self.map, dtype=np.uint8, count=self.nbytes, offset=off)

    def export(self, tex: abi.Texture, k: int):
        self.be.export_tiles_async(tex, self.ptr(k), self.nbytes)

    def close(self):
        self.be.readback_wait()
        self.be.host_unregister(self.base)
        # the mmap object is released with the last numpy view of it


class _DevicePtr:
    """Hands a raw device pointer to torch (zero-copy, __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class ShardedUpload:
    """Per-frame upload of replicated geometry on N ranks with every byte crossing a host link ONCE: rank r copies the r-th slice
    of the (concatenated) vertex + index data from pinned host memory to its GPU, an all-gather over NVLink (NCCL) assembles the
    whole on every GPU, and two device-to-device copies drop it into the library's buffers.  Everything runs on a side stream:
    slv_external_write_begin / _end order it after the previous frame's geometry pass (the buffers' reader) and before the next
    one, so the upload of frame k+1 overlaps the back half of frame k."""

    def __init__(self, be: abi.Backend, buffer_sets, host_arrays, rank: int, nranks: int):
        """`buffer_sets`: one list of buffer handles (matching `host_arrays`) per set; upload k writes set k % len(buffer_sets).
        With two sets the upload of frame k+1 does not wait for the geometry pass of frame k (which reads the other set)."""
        import numpy as np
        self.be, self.rank, self.n = be, rank, nranks
        self.side = torch.cuda.Stream()
        raw = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in host_arrays]
        self.sizes = [r.size for r in raw]
        total = sum(self.sizes)
        self.chunk = (total + nranks * 16 - 1) // (nranks * 16) * 16
        padded = np.zeros(self.chunk * nranks, np.uint8)
        padded[:total] = np.concatenate(raw)
        self.host_slice = torch.from_numpy(padded[rank * self.chunk:(rank + 1) * self.chunk].copy()).pin_memory()
        self.stage = torch.empty(self.chunk, dtype=torch.uint8, device="cuda")
        self.gathered = torch.empty(self.chunk * nranks, dtype=torch.uint8, device="cuda")
        self.sets = []
        for buffers in buffer_sets:
            views, off = [], 0
            for h, nbytes in zip(buffers, self.sizes):
                ptr, cap = be.buffer_device_ptr(h)
                assert cap >= nbytes
                views.append((torch.as_tensor(_DevicePtr(ptr, nbytes), device="cuda"), off, nbytes))
                off += nbytes
            self.sets.append(views)
        self.h2d_bytes_per_rank = self.chunk
        self.k = 0

    def upload(self) -> int:
        """Enqueues the upload into the next buffer set; returns the index of that set (the frame's draws must read it)."""
        which = self.k % len(self.sets)
        self.k += 1
        self.be.external_write_begin(self.side.cuda_stream, skip_latest=len(self.sets) - 1)
        with torch.cuda.stream(self.side):
            self.stage.copy_(self.host_slice, non_blocking=True)
            dist.all_gather_into_tensor(self.gathered, self.stage)
            for view, off, nbytes in self.sets[which]:
                view.copy_(self.gathered[off:off + nbytes], non_blocking=True)
        self.be.external_write_end(self.side.cuda_stream)
        return which


def tile_owner(tx: int, ty: int, nranks: int) -> int:
    """The rank that owns tile (tx, ty) — must equal `tile_owned` in csrc/slv_kernels.cuh."""
    return (tx + 3 * ty) % nranks


class FrameGather:
    """Per-frame assembly of the owned tiles of the resolved (single-sampled) frame on rank 0.

    Per frame:  begin_frame(); <clears and draws>; before_resolve(); <slv_resolve into target()>; gather()

    `surface` may be a list of textures: rank 0 then assembles frame k in surfaces[k % len] (the other ranks always resolve
    "into" surfaces[0], which the peer-memory transport redirects to rank 0's buffer of the frame).  With two buffers a rank
    may run one frame ahead of rank 0's consumer (e.g. an asynchronous readback) instead of in lockstep with it.
    """

    def __init__(self, be: abi.Backend, surface, rank: int, nranks: int, device: str | torch.device,
                 transport: str | None = None):
        if nranks < 1 or not (0 <= rank < nranks):
            raise ValueError("bad rank / nranks")
        surfaces = list(surface) if isinstance(surface, (list, tuple)) else [surface]
        if not surfaces or any(t.samples != 1 for t in surfaces):
            raise ValueError("sort-first gather works on the resolved (single-sampled) surface")
        surface = surfaces[0]
        self.surfaces = surfaces
        self.be, self.surface, self.rank, self.n, self.device = be, surface, rank, nranks, device
        # flag values only ever grow: a second FrameGather on the same device continues the numbering (all ranks run the
        # same frames, so they agree on it)
        self.frame = getattr(be, "_sortfirst_frame", 0)
        self.transport = "none" if nranks == 1 else "gather"
        self.root_surface = self.root_flags = None
        self.root_surfaces = []
        be.set_tile_shard(rank, nranks)
        want = transport or os.environ.get("SLV_SORTFIRST_TRANSPORT", "p2p")
        if nranks > 1 and want == "p2p" and be.name.startswith("cuda") and nranks <= 64:
            self._setup_p2p()
        if self.transport == "gather":
            self.sizes = [be.packed_tiles_bytes(surface, r, nranks) for r in range(nranks)]
            mx = max(self.sizes)
            self.stage = torch.empty(mx, dtype=torch.uint8, device=device)
            self.gather_list = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(nranks)] if rank == 0 else None

    # ---- p2p transport ----
    def _setup_p2p(self):
        be = self.be
        payload = [None]
        if self.rank == 0:
            try:
                payload = [([be.peer_export_texture(t) for t in self.surfaces], be.peer_export_flags())]
            except abi.SlvError:
                payload = [None]
        dist.broadcast_object_list(payload, src=0)
        ok = 1 if payload[0] is not None else 0
        if ok and self.rank != 0:
            try:
                self.root_surfaces = [be.peer_open(h) for h in payload[0][0]]
                self.root_surface = self.root_surfaces[0]
                self.root_flags = be.peer_open(payload[0][1])
                be.resolve_target_peer(self.surface, self.root_surface)
            except abi.SlvError:
                ok = 0
        # control traffic rides on whatever backend the job uses (gloo in the tests: CPU tensors)
        t = torch.tensor([ok], dtype=torch.int32, device="cpu" if dist.get_backend() == "gloo" else self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 1:
            self.transport = "p2p"
        else:  # some rank cannot map rank 0's memory (no peer access between the devices): everybody falls back
            self._close_p2p()

    def _close_p2p(self):
        be = self.be
        if self.root_surfaces:
            be.resolve_target_peer(self.surface, None)
            for p in self.root_surfaces:
                be.peer_close(p)
        if self.root_flags:
            be.peer_close(self.root_flags)
        self.root_surface = self.root_flags = None
        self.root_surfaces = []
        self.root_surfaces = []

    def close(self):
        if self.transport == "p2p":
            self.be.flush()
            dist.barrier()
            self._close_p2p()
            self.transport = "gather"

    @property
    def bytes_into_rank0(self) -> int:
        return sum(self.be.packed_tiles_bytes(self.surface, r, self.n) for r in range(1, self.n))

    # ---- per frame ----
    @property
    def nbuf(self) -> int:
        return len(self.surfaces)

    def target(self) -> abi.Texture:
        """The texture this rank resolves the current frame into (rank 0: the frame's buffer; others: their local stand-in)."""
        return self.surfaces[self.frame % self.nbuf] if self.rank == 0 else self.surface

    def begin_frame(self):
        """Rank 0: frame k reuses the buffer of frame k - nbuf; everything rank 0 enqueued that reads that buffer (including an
        asynchronous readback) is ordered ahead, then the other ranks are told they may overwrite it.  Other ranks (peer
        memory): point this frame's resolve at rank 0's buffer."""
        if self.transport != "p2p":
            return
        k = self.frame
        if self.rank == 0:
            if k >= self.nbuf:
                # behind the assembly wait and every copy-stream consumer (readback) of this buffer's previous frame - and of nothing
                # else; rank 0's render stream is never stalled by the protocol (its own next write of the buffer waits for the
                # texture's event)
                self.be.peer_signal_after_consumers(self.surfaces[k % self.nbuf], None, 0, k - self.nbuf + 1)
        elif self.nbuf > 1:
            self.be.resolve_target_peer(self.surface, self.root_surfaces[k % self.nbuf])

    def before_resolve(self):
        """Other ranks: do not store frame k into rank 0's buffer before rank 0 released frame k - nbuf."""
        if self.transport == "p2p" and self.rank != 0 and self.frame >= self.nbuf:
            self.be.flags_wait(self.root_flags, 0, 1, self.frame - self.nbuf + 1)

    def gather(self):
        """Call after the frame's resolve.  On return rank 0's target() holds the whole frame (stream-ordered)."""
        if self.transport == "none":
            self.frame += 1
            return
        if self.transport == "p2p":
            if self.rank == 0:
                self.be.assembly_wait(self.target(), None, 1, self.n - 1, self.frame + 1)
            else:
                self.be.peer_signal(self.root_flags, self.rank, self.frame + 1)
            self.frame += 1
            self.be._sortfirst_frame = self.frame
            return
        tgt = self.target()
        self.be.pack_tiles(tgt, self.rank, self.n, self.stage.data_ptr())
        dist.gather(self.stage, self.gather_list, dst=0)
        if self.rank == 0:
            for r in range(1, self.n):
                self.be.unpack_tiles(tgt, r, self.n, self.gather_list[r].data_ptr())
        self.frame += 1
