"""Developer tool: per-frame GPU time and counters of the Sponza-like scene (frames 0..7)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg
from salviarenderer_b200 import abi as A, scenes as S
be = A.Backend(os.environ['SLV_LIB']) if os.environ.get('SLV_LIB') else pkg.load(0)
sc = S.SponzaLike(3840, 2160, 4)
sc.setup(be)
if os.environ.get('SHARD'):
    r, n = map(int, os.environ['SHARD'].split(','))
    be.set_tile_shard(r, n)
    print('tile shard', r, 'of', n)
import time
for f in range(8):
    sc.render(be, f)
be.flush()
NF = int(os.environ.get('FRAMES', '40'))
for rep in range(int(os.environ.get('REPS', '1'))):
    t0 = time.perf_counter()
    for f in range(NF):
        sc.render(be, f % 8)
    t1 = time.perf_counter()
    be.flush()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0) / NF:.3f} ms/frame, with the final sync {1e3 * (t2 - t0) / NF:.3f} ms/frame ({NF} frames)", flush=True)
for f in range(0 if os.environ.get('NO_STAGES') else 8):
    be.profile_enable(True)
    be.query_begin()
    sc.render(be, f)
    be.flush()
    sg = be.profile_stages(); pr = be.profile_get(); st = be.query_get(); tr = be.traffic()
    be.profile_enable(False)
    be.event_record(0)
    for _ in range(3):
        sc.render(be, f)
    be.event_record(1)
    ms = be.event_elapsed_ms(0, 1) / 3
    print(f"frame {f}: {ms:6.3f} ms | geom {pr['clipping']/1e6:5.3f} bin {pr['tri_dispatch']/1e6:5.3f} raster {pr['ras']/1e6:6.3f} (cover {sg['raster_or_cover']:5.3f} shade {sg['shade']:5.3f} sort {sg['sort']:5.3f} rbin {sg['region_bin']:5.3f}) | ps {st['ps_invocations']/1e6:6.2f}M cprims {st['cprimitives']:7d} ztest {tr['z_tested']/1e6:6.1f}M cwr {tr['c_written']/1e6:6.1f}M | scanned {tr['list_entries_scanned']/1e6:6.2f}M surv {tr['region_survivors']/1e6:6.2f}M pairs {tr['warp_pairs']/1e6:6.2f}M quads {tr['quads_shaded']/1e6:6.2f}M ps_exec {tr['ps_executed']/1e6:6.2f}M", flush=True)
