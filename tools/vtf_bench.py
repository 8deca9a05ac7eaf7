"""Developer tool: the VertexTextureFetch scene at a size where the vertex stage matters (a 2B x 2B-quad plane displaced by a height
map in the vertex shader): GPU time per frame and per stage, with the post-transform vertex cache (default) or without
(SLV_VERTEX_CACHE=0).     python tools/vtf_bench.py [--block 512 --width 1920 --height 1080 --frames 50]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--block", type=int, default=256)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--frames", type=int, default=50)
a = ap.parse_args()
import numpy as np  # noqa: E402
sc = S.TerrainVTF(a.width, a.height, 1, block=32, tex_size=256)
b = a.block  # the sample's grid, larger and with 32-bit indices
sc.block = b
sc.plane = S.create_planar((-16.0, 0.0, -16.0), (0.5 * 32 / b, 0, 0), (0, 0, 0.5 * 32 / b), b * 2, b * 2, False, index_dtype=np.uint32)
sc.plane.elements = [(0, S._V4, 0, 0, 1.0), (1, S._V4, 2, 0, 0.0)]
be = pkg.load(0)
sc.setup(be)
for f in range(3):
    sc.render(be, f)
be.flush()
be.query_begin()
be.event_record(0)
for f in range(a.frames):
    sc.render(be, f % sc.n_frames)
be.event_record(1)
ms = be.event_elapsed_ms(0, 1) / a.frames
st = be.query_get()
print(f"SLV_VERTEX_CACHE={os.environ.get('SLV_VERTEX_CACHE', '(default)')}: {sc.plane.prim_count} triangles, {len(sc.plane.streams[0])} vertices: "
      f"{ms:.3f} ms/frame, vs_invocations/frame {st['vs_invocations'] // a.frames}", flush=True)
be.profile_enable(True)
be.query_begin()
sc.render(be, 0)
be.flush()
print("  stages (ms):", {k: round(v, 3) for k, v in be.profile_stages().items()})
