"""CPU suite: the N>1 host path (sort-first tile sharding + per-frame gather, salviarenderer_b200/sortfirst.py) with
world_size 2 over gloo.  The compute behind the C ABI is the CPU checker here (there is no GPU in this container);
the host logic — shard assignment, pack / gather / unpack, rank-0 assembly — is the code bench.py runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ORACLE_LIB, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from salviarenderer_b200 import abi, scenes, sortfirst
    be = abi.Backend(ORACLE_LIB)
    sc = scenes.SponzaLike(320, 192, 4, tex_size=64)
    sc.setup(be)
    resolved = sc.t.resolved
    fg = sortfirst.FrameGather(be, resolved, rank, world, "cpu")
    sc.render(be, 1)
    be.flush()
    own = be.read_texture(resolved).copy()
    fg.gather()
    be.flush()
    stats = be.query_get()
    ps = torch.tensor([float(stats["ps_invocations"])], dtype=torch.float64)
    dist.all_reduce(ps)
    if rank == 0:
        np.save(out_path, be.read_texture(resolved))
        np.save(out_path + ".ps.npy", ps.numpy())
    np.save(out_path + f".own{rank}.npy", own)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_sort_first_gather_equals_single_rank(oracle, tmp_path):
    from salviarenderer_b200 import scenes, sortfirst
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.SponzaLike(320, 192, 4, tex_size=64)
    sc.setup(oracle)
    oracle.set_tile_shard(0, 1)
    oracle.query_begin()
    sc.render(oracle, 1)
    oracle.flush()
    want = oracle.read_texture(sc.t.resolved)
    assert np.array_equal(got, want)
    assert float(np.load(out + ".ps.npy")[0]) == float(oracle.query_get()["ps_invocations"])
    # before the gather each rank held exactly its own tiles (the others still hold the clear colour)
    h, w = want.shape[:2]
    ty, tx = np.mgrid[0:h, 0:w]
    owner = (tx // 64 + 3 * (ty // 64)) % 2
    assert sortfirst.tile_owner(2, 1, 2) == (2 + 3) % 2
    for r in range(2):
        own = np.load(out + f".own{r}.npy")
        assert np.array_equal(own[owner == r], want[owner == r])
        assert not np.array_equal(own[owner != r], want[owner != r])


def test_frame_gather_rejects_bad_arguments(oracle):
    from salviarenderer_b200 import abi, sortfirst
    t = oracle.create_texture(128, 128, 4, abi.PF_RGBA8)
    with pytest.raises(ValueError):
        sortfirst.FrameGather(oracle, t, 0, 2, "cpu")  # multi-sampled surface
    t1 = oracle.create_texture(128, 128, 1, abi.PF_RGBA8)
    with pytest.raises(ValueError):
        sortfirst.FrameGather(oracle, t1, 2, 2, "cpu")
    sizes = [oracle.packed_tiles_bytes(t1, r, 3) for r in range(3)]
    assert sum(sizes) == 4 * 64 * 64 * 4  # 2x2 tiles of 64x64 rgba8, every tile owned by exactly one rank
    oracle.set_tile_shard(0, 1)


def _worker_hostframe(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from salviarenderer_b200 import abi, scenes, sortfirst
    be = abi.Backend(ORACLE_LIB)
    be.set_tile_shard(rank, world)
    sc = scenes.SponzaLike(320, 192, 4, tex_size=64)
    sc.setup(be)
    hf = sortfirst.HostFrame(be, 320 * 192 * 4, rank, world, nbuf=2)
    for k, f in enumerate((1, 4, 6)):
        sc.render(be, f)
        hf.export(sc.t.resolved, k)
        be.readback_wait()
        dist.barrier()
        if rank == 0:
            np.save(out_path + f".{k}.npy", hf.view(k).reshape(192, 320, 1, 4).copy())
        dist.barrier()
    hf.close()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_host_frame_assembly_equals_single_rank(oracle, tmp_path):
    """The end-to-end transport of bench.py at N > 1 (sortfirst.HostFrame): every rank exports the tiles it owns into one host
    frame in POSIX shared memory; three frames through the two alternating buffers equal the unsharded render."""
    from salviarenderer_b200 import scenes
    out = str(tmp_path / "hostframe")
    mp.spawn(_worker_hostframe, args=(2, _free_port(), out), nprocs=2, join=True)
    sc = scenes.SponzaLike(320, 192, 4, tex_size=64)
    sc.setup(oracle)
    oracle.set_tile_shard(0, 1)
    for k, f in enumerate((1, 4, 6)):
        sc.render(oracle, f)
        oracle.flush()
        assert np.array_equal(np.load(out + f".{k}.npy"), oracle.read_texture(sc.t.resolved)), f"frame {f}"
