"""Developer tool: BASELINE configs[4] at full size — the 10,000,000-triangle height field at 7680x4320, two passes per frame
(depth-only shadow pass + colour pass): GPU time per frame and the pipeline counters.

    python tools/stress_10m.py [--nx 2500 --nz 2000 --width 7680 --height 4320 --frames 6]
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
ap.add_argument("--nx", type=int, default=2500)
ap.add_argument("--nz", type=int, default=2000)
ap.add_argument("--width", type=int, default=7680)
ap.add_argument("--height", type=int, default=4320)
ap.add_argument("--samples", type=int, default=1)
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--shadowed", action="store_true", help="colour pass = the StandardShadowMap sample's shaders, sampling pass 1's depth")
a = ap.parse_args()
t0 = time.time()
sc = S.HeightFieldTwoPass(a.width, a.height, a.samples, nx=a.nx, nz=a.nz, shadowed=a.shadowed)
print(f"mesh: {sc.mesh.prim_count} triangles, {len(sc.mesh.streams[0])} vertices, built in {time.time() - t0:.1f} s", flush=True)
be = pkg.load(0)
sc.setup(be)
for f in range(2):
    sc.render(be, f)
be.flush()
be.query_begin()
be.event_record(0)
for f in range(a.frames):
    sc.render(be, f % sc.n_frames)
be.event_record(1)
ms = be.event_elapsed_ms(0, 1) / a.frames
st = be.query_get()
tris = 2 * sc.mesh.prim_count  # two passes
print(f"{a.width}x{a.height}x{a.samples}{' shadowed (SSM colour pass)' if a.shadowed else ''}: {ms:.3f} ms/frame ({1e3 / ms:.1f} frames/s), {tris / ms / 1e6:.2f} G triangles/s in, "
      f"cprimitives/frame {st['cprimitives'] // a.frames}, ps_invocations/frame {st['ps_invocations'] // a.frames}", flush=True)
be.profile_enable(True)
be.query_begin()
sc.render(be, 0)
be.flush()
print("stages (ms, both passes):", {k: round(v, 3) for k, v in be.profile_stages().items()})
