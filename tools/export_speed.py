"""Developer tool: throughput of slv_texture_export_tiles_async (zero-copy stores of the owned tiles into a registered host frame)
against the DMA readback of the whole surface, for 1/2/8-way tile ownership."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg
from salviarenderer_b200 import abi as A, sortfirst
import torch
be = pkg.load(0)
W, H = 3840, 2160
tex = be.create_texture(W, H, 1, A.PF_BGRA8)
nbytes = W * H * 4
hf = sortfirst.HostFrame(be, nbytes, 0, 1, nbuf=1)
pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
for n in (1, 2, 4, 8):
    be.set_tile_shard(0, n)
    own = be.packed_tiles_bytes(tex, 0, n)
    for _ in range(3):
        hf.export(tex, 0)
    be.readback_wait()
    t0 = time.perf_counter()
    for _ in range(20):
        hf.export(tex, 0)
    be.readback_wait()
    dt = (time.perf_counter() - t0) / 20
    print(f"export 1/{n} of the frame ({own / 1e6:.1f} MB): {dt * 1e3:.3f} ms = {own / dt / 1e9:.1f} GB/s", flush=True)
be.set_tile_shard(0, 1)
t0 = time.perf_counter()
for _ in range(20):
    be.read_texture_into_async(tex, pinned.data_ptr(), nbytes)
be.readback_wait()
dt = (time.perf_counter() - t0) / 20
print(f"DMA readback of the whole frame ({nbytes / 1e6:.1f} MB): {dt * 1e3:.3f} ms = {nbytes / dt / 1e9:.1f} GB/s")
