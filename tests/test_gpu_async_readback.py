"""GPU suite: slv_texture_readback_async — frames read back on the copy stream while the next frame renders into the other
resolve target; every buffer must hold exactly the frame it was issued for (the write-after-read ordering on the device)."""
import numpy as np
import pytest
import torch

from salviarenderer_b200 import scenes as S

pytestmark = pytest.mark.gpu


def test_async_readback_pipelined_frames(cuda):
    w, h = 480, 272
    sc = S.SponzaLike(w, h, 4, tex_size=64)
    sc.setup(cuda)
    targets = [sc.t.resolved, cuda.create_texture(w, h, 1, sc.t.resolved.fmt)]
    frames = [0, 3, 5, 7, 2, 6, 1]
    nbytes = w * h * 4
    bufs = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in frames]
    # geometry is re-uploaded every frame as well (uploads ride the front stream when pipelining)
    vb, ib = sc.mesh.upload(cuda)
    vb_np = np.ascontiguousarray(sc.mesh.streams[0])
    ib_np = np.ascontiguousarray(sc.mesh.indices)
    for i, f in enumerate(frames):
        cuda.upload_from_ptr(vb[0], vb_np.ctypes.data, vb_np.nbytes)
        cuda.upload_from_ptr(ib, ib_np.ctypes.data, ib_np.nbytes)
        sc.t.resolved = targets[i % 2]
        sc.render(cuda, f)
        cuda.read_texture_into_async(targets[i % 2], bufs[i].data_ptr(), nbytes)
    cuda.readback_wait()
    sc.t.resolved = targets[0]
    for i, f in enumerate(frames):
        sc.render(cuda, f)
        want = cuda.read_texture(targets[0]).reshape(-1)
        assert np.array_equal(bufs[i].numpy(), want), f"frame {f} (slot {i})"
