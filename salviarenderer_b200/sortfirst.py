"""Sort-first multi-GPU frame assembly — host-side plumbing, one process per GPU.

Rank r of R owns the 64x64 screen tiles with (tx + 3*ty) % R == r (`slv_set_tile_shard`).  Every rank issues the same
command stream; geometry is replicated, the raster kernels only touch owned tiles.  Once per frame the owned tiles of
the RESOLVED colour surface have to reach rank 0.  Two transports:

* ``p2p`` (the CUDA product, default): rank 0 exports its resolved surface and its flag block as CUDA IPC handles; every
  other rank opens them and redirects `slv_resolve` into rank 0's memory, so the MSAA resolve and the exchange are ONE
  kernel storing over NVLink / NVSwitch — no staging buffer, no pack / unpack kernels, no collective.  Per frame the
  ranks' streams are ordered on the device by flags in rank 0's memory (`slv_peer_signal` / `slv_flags_wait`):
  rank r raises flags[r] = k+1 after its resolve of frame k; rank 0's stream waits for all of them, and raises
  flags[0] = k once everything it enqueued that reads frame k-1.. is behind it, which the other ranks poll before they
  overwrite the surface with frame k.  No host synchronisation anywhere.
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


def tile_owner(tx: int, ty: int, nranks: int) -> int:
    """The rank that owns tile (tx, ty) — must equal `tile_owned` in csrc/slv_kernels.cuh."""
    return (tx + 3 * ty) % nranks


class FrameGather:
    """Per-frame assembly of the owned tiles of `surface` (single-sampled) on rank 0.

    Per frame:  begin_frame(); <clears and draws>; before_resolve(); <slv_resolve into `surface`>; gather()
    """

    def __init__(self, be: abi.Backend, surface: abi.Texture, rank: int, nranks: int, device: str | torch.device,
                 transport: str | None = None):
        if nranks < 1 or not (0 <= rank < nranks):
            raise ValueError("bad rank / nranks")
        if surface.samples != 1:
            raise ValueError("sort-first gather works on the resolved (single-sampled) surface")
        self.be, self.surface, self.rank, self.n, self.device = be, surface, rank, nranks, device
        # flag values only ever grow: a second FrameGather on the same device continues the numbering (all ranks run the
        # same frames, so they agree on it)
        self.frame = getattr(be, "_sortfirst_frame", 0)
        self.transport = "none" if nranks == 1 else "gather"
        self.root_surface = self.root_flags = None
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
                payload = [(be.peer_export_texture(self.surface), be.peer_export_flags())]
            except abi.SlvError:
                payload = [None]
        dist.broadcast_object_list(payload, src=0)
        ok = 1 if payload[0] is not None else 0
        if ok and self.rank != 0:
            try:
                self.root_surface = be.peer_open(payload[0][0])
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
        if self.root_surface:
            be.resolve_target_peer(self.surface, None)
            be.peer_close(self.root_surface)
        if self.root_flags:
            be.peer_close(self.root_flags)
        self.root_surface = self.root_flags = None

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
    def begin_frame(self):
        """Rank 0: everything enqueued so far that reads the assembled surface is ahead in the stream, so the other
        ranks may overwrite it with frame `self.frame`."""
        if self.transport == "p2p" and self.rank == 0 and self.frame > 0:
            self.be.peer_signal(None, 0, self.frame)

    def before_resolve(self):
        """Other ranks: do not store frame k into rank 0's surface before rank 0 released frame k-1."""
        if self.transport == "p2p" and self.rank != 0 and self.frame > 0:
            self.be.flags_wait(self.root_flags, 0, 1, self.frame)

    def gather(self):
        """Call after the frame's resolve.  On return rank 0's `surface` holds the whole frame (stream-ordered)."""
        if self.transport == "none":
            return
        if self.transport == "p2p":
            if self.rank == 0:
                self.be.flags_wait(None, 1, self.n - 1, self.frame + 1)
            else:
                self.be.peer_signal(self.root_flags, self.rank, self.frame + 1)
            self.frame += 1
            self.be._sortfirst_frame = self.frame
            return
        self.be.pack_tiles(self.surface, self.rank, self.n, self.stage.data_ptr())
        dist.gather(self.stage, self.gather_list, dst=0)
        if self.rank == 0:
            for r in range(1, self.n):
                self.be.unpack_tiles(self.surface, r, self.n, self.gather_list[r].data_ptr())
        self.frame += 1
