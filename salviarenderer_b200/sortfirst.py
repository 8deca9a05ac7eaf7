"""Sort-first multi-GPU frame assembly — host-side plumbing, one process per GPU.

Rank r of R owns the 64x64 screen tiles with (tx + 3*ty) % R == r (`slv_set_tile_shard`).  Every rank issues the same
command stream; geometry is replicated, the raster kernel only touches owned tiles.  Once per frame the owned tiles
of the RESOLVED colour surface are packed into a dense staging buffer (`slv_pack_tiles`), gathered to rank 0 with
`torch.distributed.gather` (NCCL over NVLink on the GPUs; gloo in the CPU tests, where the checker backends stage in
host memory) and scattered back into rank 0's linear surface (`slv_unpack_tiles`).  There is no other collective on
the path.

The staging buffers are torch tensors: device tensors for the CUDA product (all work is enqueued on the stream the
library was given with `slv_set_stream`, so the pack kernel, the NCCL gather and the unpack kernels order on the
device without host synchronisation), CPU tensors for the checkers.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import abi


def tile_owner(tx: int, ty: int, nranks: int) -> int:
    """The rank that owns tile (tx, ty) — must equal `tile_owned` in csrc/slv_kernels.cuh."""
    return (tx + 3 * ty) % nranks


class FrameGather:
    """Per-frame gather of the owned tiles of `surface` (single-sampled) to rank 0."""

    def __init__(self, be: abi.Backend, surface: abi.Texture, rank: int, nranks: int, device: str | torch.device):
        if nranks < 1 or not (0 <= rank < nranks):
            raise ValueError("bad rank / nranks")
        if surface.samples != 1:
            raise ValueError("sort-first gather works on the resolved (single-sampled) surface")
        self.be, self.surface, self.rank, self.n = be, surface, rank, nranks
        self.sizes = [be.packed_tiles_bytes(surface, r, nranks) for r in range(nranks)]
        mx = max(self.sizes)
        self.stage = torch.empty(mx, dtype=torch.uint8, device=device)
        self.gather_list = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(nranks)] if rank == 0 else None
        be.set_tile_shard(rank, nranks)

    @property
    def bytes_into_rank0(self) -> int:
        return sum(self.sizes[1:])

    def gather(self):
        """Call after the frame's resolve.  On return rank 0's `surface` holds the whole frame (stream-ordered)."""
        if self.n == 1:
            return
        self.be.pack_tiles(self.surface, self.rank, self.n, self.stage.data_ptr())
        dist.gather(self.stage, self.gather_list, dst=0)
        if self.rank == 0:
            for r in range(1, self.n):
                self.be.unpack_tiles(self.surface, r, self.n, self.gather_list[r].data_ptr())
