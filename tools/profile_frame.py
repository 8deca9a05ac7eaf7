"""Developer tool: renders warm-up frames, then ONE frame of the bench workload between cudaProfilerStart/Stop, so that
`ncu --profile-from-start off ...` captures exactly the kernels of one steady-state frame.

    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/prof -f \
        python tools/profile_frame.py [--frame 3] [--width 3840 --height 2160 --samples 4]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frame", type=int, default=3)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--samples", type=int, default=4)
ap.add_argument("--tex-size", type=int, default=1024)
ap.add_argument("--aniso", type=int, default=0)
ap.add_argument("--shaders", default="builtin", choices=["builtin", "sasl", "twins"],
                help="builtin: SLV_VS_SPONZA + SLV_PS_SPONZA; sasl: bench.py's SASL pair compiled at run time; twins: SLV_PS_SPONZA_GRAD")
ap.add_argument("--scene", default="sponza", help="sponza | c1 | c2 | c3a | c3b | c5 (bench.small_scene: the other BASELINE.json configs at full size)")
ap.add_argument("--shard", default=None, help="r,n: render only the tiles sort-first rank r of n owns (what one GPU of an n-GPU run executes)")
a = ap.parse_args()
be = pkg.load(0)
if a.shard:
    r_, n_ = map(int, a.shard.split(","))
    be.set_tile_shard(r_, n_)
from salviarenderer_b200 import abi as A  # noqa: E402
if a.scene != "sponza":
    import bench  # noqa: E402
    sc, what = bench.small_scene(a.scene, S)
    print(what)
elif a.shaders == "twins":
    sc = S.SponzaLike(a.width, a.height, a.samples, tex_size=a.tex_size, max_aniso=a.aniso, ps_program=A.PS_SPONZA_GRAD)
else:
    sc = S.SponzaLike(a.width, a.height, a.samples, tex_size=a.tex_size, max_aniso=a.aniso)
if a.scene == "sponza" and a.shaders == "sasl":
    import bench  # noqa: E402
    bench.install_sasl_shaders(sc, be, A)
sc.setup(be)
for f in range(4):
    sc.render(be, f % sc.n_frames)
be.flush()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
sc.render(be, a.frame % sc.n_frames)
be.flush()
rt.cudaProfilerStop()
print("profiled frame", a.frame, be.query_get())
