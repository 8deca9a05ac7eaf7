"""Developer tool: per-stage GPU time of SponzaLike variants (decomposes where k_raster's time goes)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg
from salviarenderer_b200 import abi as A, scenes as S

be = pkg.load(0)
variants = {
    "4K x4 sponza ps (trilinear)": dict(w=3840, h=2160, samples=4),
    "4K x4 sponza ps untextured": dict(w=3840, h=2160, samples=4, textured=False),
    "4K x4 attr0 colour ps": dict(w=3840, h=2160, samples=4, ps_program=A.PS_ATTR0_COLOR),
    "4K x1 sponza ps": dict(w=3840, h=2160, samples=1),
    "1080p x4 sponza ps": dict(w=1920, h=1080, samples=4),
    "4K x4 sponza ps tex256": dict(w=3840, h=2160, samples=4, tex_size=256),
}
for name, kw in variants.items():
    sc = S.SponzaLike(**kw)
    sc.setup(be)
    for f in range(3):
        sc.render(be, f)
    be.flush()
    K = 8
    be.event_record(0)
    for f in range(K):
        sc.render(be, f)
    be.event_record(1)
    total = be.event_elapsed_ms(0, 1) / K
    be.profile_enable(True)
    be.query_begin()
    for f in range(K):
        sc.render(be, f)
    be.flush()
    pr = be.profile_get(); st = be.query_get(); tr = be.traffic()
    be.profile_enable(False)
    print(f"{name:34s} frame {total:7.3f} ms | geom {pr['clipping']/1e6/K:6.3f} bin+sort {pr['tri_dispatch']/1e6/K:6.3f} raster {pr['ras']/1e6/K:7.3f} ms | "
          f"ps_px/frame {st['ps_invocations']/K/1e6:6.2f}M cprims {st['cprimitives']/K:8.0f} z_tested {tr['z_tested']/K/1e6:6.1f}M c_written {tr['c_written']/K/1e6:6.1f}M", flush=True)
    for t in sc.textures: be.release(t.handle)
    for t in (sc.t.color, sc.t.ds, sc.t.resolved):
        if t is not None: be.release(t.handle)
