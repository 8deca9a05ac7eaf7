"""GPU suite: sort-first frame assembly over peer memory (salviarenderer_b200/sortfirst.py, transport "p2p").

Two processes (one per GPU when the box has two, else both on cuda:0 — CUDA IPC works either way) render the same
Sponza-like frames with interleaved tile ownership; rank 1 resolves its tiles straight into rank 0's surface and the
ranks' streams are ordered by device-side flags only.  Rank 0's assembled frames must equal an unsharded render of the
same frames on the same device, bit for bit, for several frames in a row (so the release / wait handshake is used)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu

FRAMES = (0, 3, 5, 7, 2)
W, H, S = 576, 320, 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path, transport, nbuf=1):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import salviarenderer_b200 as pkg
    from salviarenderer_b200 import scenes, sortfirst
    ordinal = rank % torch.cuda.device_count()
    be = pkg.load(ordinal)
    # as bench.py does: the library, torch and the collective backend all order on one stream
    torch.cuda.set_device(ordinal)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be.set_stream(stream.cuda_stream)
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(be)
    targets = [sc.t.resolved] + [be.create_texture(W, H, 1, sc.t.resolved.fmt) for _ in range(nbuf - 1)]
    fg = sortfirst.FrameGather(be, targets if nbuf > 1 else targets[0], rank, world, f"cuda:{ordinal}", transport=transport)
    frames = []
    pinned = [torch.zeros(W * H * 4, dtype=torch.uint8).pin_memory() for _ in FRAMES]
    for i, f in enumerate(FRAMES):
        fg.begin_frame()
        sc.t.resolved = fg.target()
        tgt = fg.target()
        sc.render(be, f, before_resolve=fg.before_resolve)
        fg.gather()
        if rank == 0:
            if nbuf > 1:  # asynchronous: frame f is copied out while the ranks already render the next one into the other buffer
                be.read_texture_into_async(tgt, pinned[i].data_ptr(), W * H * 4)
            else:
                frames.append(be.read_texture(tgt).copy())  # synchronous: the app owns frame f now
    be.flush()
    if rank == 0:
        if nbuf > 1:
            frames = [p.numpy().reshape(H, W, 1, 4).copy() for p in pinned]
        np.save(out_path, np.stack(frames))
        open(out_path + ".transport", "w").write(fg.transport)
    dist.barrier()
    fg.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("transport,nbuf", [("p2p", 1), ("gather", 1), ("p2p", 2)])
def test_world2_assembled_frames_equal_unsharded(cuda, tmp_path, transport, nbuf):
    from salviarenderer_b200 import scenes
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker, args=(2, _free_port(), out, transport, nbuf), nprocs=2, join=True)
    got = np.load(out)
    assert open(out + ".transport").read() == transport
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(cuda)
    cuda.set_tile_shard(0, 1)
    for i, f in enumerate(FRAMES):
        sc.render(cuda, f)
        want = cuda.read_texture(sc.t.resolved)
        assert np.array_equal(got[i].reshape(want.shape), want), f"frame {f} ({transport}, {nbuf} buffers)"


def _worker_hostframe(rank, world, port, out_path):
    """End-to-end assembly on the host: every rank exports the tiles it owns into one shared pinned host frame."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import salviarenderer_b200 as pkg
    from salviarenderer_b200 import scenes, sortfirst
    ordinal = rank % torch.cuda.device_count()
    be = pkg.load(ordinal)
    torch.cuda.set_device(ordinal)
    be.set_tile_shard(rank, world)
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(be)
    local = [sc.t.resolved, be.create_texture(W, H, 1, sc.t.resolved.fmt)]
    hf = sortfirst.HostFrame(be, W * H * 4, rank, world, nbuf=2)
    frames = []
    for i, f in enumerate(FRAMES):
        sc.t.resolved = local[i % 2]
        sc.render(be, f)
        hf.export(local[i % 2], i)
        if i % 2 == 1 or i == len(FRAMES) - 1:  # every second frame: wait for both buffers, as an application consuming them would
            be.readback_wait()
            dist.barrier()  # every rank's tiles of the frames in flight have landed
            if rank == 0:
                for j in range(max(0, i - 1 if i % 2 == 1 else i), i + 1):
                    frames.append(hf.view(j).reshape(H, W, 1, 4).copy())
            dist.barrier()  # nobody overwrites a buffer before rank 0 has read it
    if rank == 0:
        np.save(out_path, np.stack(frames))
    hf.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_host_frame_export_equals_unsharded(cuda, tmp_path):
    """Multi-GPU end to end: the tiles each rank exports over its own host link (slv_texture_export_tiles_async into POSIX shared
    memory registered by both processes) assemble, on the host, the frames an unsharded render produces."""
    from salviarenderer_b200 import scenes
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker_hostframe, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(cuda)
    cuda.set_tile_shard(0, 1)
    assert len(got) == len(FRAMES)
    for i, f in enumerate(FRAMES):
        sc.render(cuda, f)
        want = cuda.read_texture(sc.t.resolved)
        assert np.array_equal(got[i].reshape(want.shape), want), f"frame {f}"


def _worker_sharded_upload(rank, world, port, out_path):
    """bench.py's e2e path at N > 1: sharded upload (each rank a slice + NCCL all-gather) + host-frame export."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import salviarenderer_b200 as pkg
    from salviarenderer_b200 import scenes, sortfirst
    be = pkg.load(rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    be.set_stream(stream.cuda_stream)
    be.set_tile_shard(rank, world)
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(be)
    vb_h, ib_h = sc.mesh.upload(be)[0][0], sc.mesh.upload(be)[1]
    # scramble the resident geometry: only a correct sharded upload restores it
    zeros = np.zeros_like(np.ascontiguousarray(sc.mesh.streams[0], dtype=np.float32))
    be.upload_from_ptr(vb_h, zeros.ctypes.data, zeros.nbytes)
    be.flush()
    su = sortfirst.ShardedUpload(be, [[vb_h, ib_h]], [sc.mesh.streams[0], sc.mesh.indices], rank, world)
    local = [sc.t.resolved, be.create_texture(W, H, 1, sc.t.resolved.fmt)]
    hf = sortfirst.HostFrame(be, W * H * 4, rank, world, nbuf=2)
    frames = []
    for i, f in enumerate(FRAMES):
        su.upload()
        sc.t.resolved = local[i % 2]
        sc.render(be, f)
        hf.export(local[i % 2], i)
        be.readback_wait()
        dist.barrier()
        if rank == 0:
            frames.append(hf.view(i).reshape(H, W, 1, 4).copy())
        dist.barrier()
    if rank == 0:
        np.save(out_path, np.stack(frames))
    hf.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_sharded_upload_and_host_frame(cuda, tmp_path):
    """Needs two GPUs (NCCL): every rank uploads half of the vertex + index data, an all-gather over NVLink assembles the library's
    buffers on both, the frames exported into the shared host frame equal the unsharded render."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL all-gather)")
    from salviarenderer_b200 import scenes
    out = str(tmp_path / "frames.npy")
    mp.spawn(_worker_sharded_upload, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.SponzaLike(W, H, S, tex_size=64)
    sc.setup(cuda)
    cuda.set_tile_shard(0, 1)
    for i, f in enumerate(FRAMES):
        sc.render(cuda, f)
        want = cuda.read_texture(sc.t.resolved)
        assert np.array_equal(got[i].reshape(want.shape), want), f"frame {f}"
