"""Developer tool: host cost of one frame's command stream.  The Sponza-like scene's 27 API calls per frame (2 clears, 24 draws,
1 resolve) are issued for a scene whose GPU work is negligible (24 draws of the same mesh clipped away / a 64x64 target), so the
frame rate is bound by the host: Python + ctypes + the library's validation / batching + the CUDA launches of one batch flush.
    python tools/host_overhead.py [--frames 2000]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2000)
a = ap.parse_args()
be = pkg.load(0)
sc = S.SponzaLike(64, 64, 4, tex_size=16)
# keep one triangle per draw: the command stream is the same, the GPU has nothing to do
sc.groups = [(m, start, 1) for m, start, _ in sc.groups]
sc.setup(be)
for f in range(20):
    sc.render(be, f % sc.n_frames)
be.flush()
for rep in range(3):
    t0 = time.perf_counter()
    for f in range(a.frames):
        sc.render(be, f % sc.n_frames)
    t1 = time.perf_counter()
    be.flush()
    t2 = time.perf_counter()
    print(f"host-bound frame loop: {(t1 - t0) / a.frames * 1e6:.1f} us/frame enqueue, {(t2 - t0) / a.frames * 1e6:.1f} us/frame with the final sync "
          f"({a.frames} frames, 27 API calls + one batch flush each)", flush=True)
