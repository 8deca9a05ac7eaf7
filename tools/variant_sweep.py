"""Developer tool: frame time of the headline workload (built-in twins of the SASL pair, 16x AF) and of the trilinear built-in
workload for alternative builds of the library (build_variants/lib_*.so: other occupancy targets / group sizes of k_cover /
k_shade) and resident-CTA settings.   python tools/variant_sweep.py"""
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from salviarenderer_b200 import abi as A, scenes as S  # noqa: E402

libs = [os.path.join(ROOT, "salviarenderer_b200", "csrc", "libsalvia_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "build_variants", "lib_*.so")))
for lib in libs:
    name = os.path.basename(lib)
    ctas = ["", "8"]
    if "10" in name:
        ctas = ["8", "9", "10"]
    if "6.so" in name:
        ctas = ["6"]
    for bc in ctas:
        if bc:
            os.environ["SLV_BACK_CTAS"] = bc
        try:
            be = A.Backend(lib)
        finally:
            os.environ.pop("SLV_BACK_CTAS", None)
        out = []
        for label, kw in (("AF16 twins", dict(max_aniso=16, ps_program=A.PS_SPONZA_GRAD)), ("trilinear", dict())):
            sc = S.SponzaLike(3840, 2160, 4, **kw)
            sc.setup(be)
            for f in range(8):
                sc.render(be, f)
            be.flush()
            best = 1e9
            for rep in range(3):
                be.event_record(0)
                for f in range(80):
                    sc.render(be, f % 8)
                be.event_record(1)
                best = min(best, be.event_elapsed_ms(0, 1) / 80)
            be.profile_enable(True)
            be.query_begin()
            for f in range(8):
                sc.render(be, f)
            be.flush()
            sg = be.profile_stages()
            be.profile_enable(False)
            out.append(f"{label}: {best:.4f} ms (geom {sg['geometry'] / 8:.3f} cover {sg['raster_or_cover'] / 8:.3f} shade {sg['shade'] / 8:.3f})")
        if "GEOM" in name or name == "libsalvia_b200.so":  # the geometry-bound stress mesh: 2 x 1.6 M triangles at 3840x2160
            sc = S.HeightFieldTwoPass(3840, 2160, 1, nx=1000, nz=800)
            sc.setup(be)
            for f in range(2):
                sc.render(be, f)
            be.flush()
            be.profile_enable(True)
            be.query_begin()
            for f in range(4):
                sc.render(be, f % sc.n_frames)
            be.flush()
            sg = be.profile_stages()
            be.profile_enable(False)
            out.append("heightfield 2x1.6M: " + " ".join(f"{k}={v / 4:.3f}" for k, v in sg.items()))
        print(f"{name:22s} BACK_CTAS={bc or 'default':7s} " + " | ".join(out), flush=True)
        be.close()
