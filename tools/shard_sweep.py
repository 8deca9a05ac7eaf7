"""Developer tool: what ONE rank of an N-way sort-first run costs per frame, on a single GPU (slv_set_tile_shard(r, N) without
peers), across the library's scheduling knobs.  The multi-GPU frame rate is bounded by the slowest rank, so this is the cheap
way to tune the per-rank pipeline (front half replicated, back half an N-th of the frame).

    python tools/shard_sweep.py [--shards 0,8 3,8] [--frames 200] [--settings "SLV_BACK_CTAS=4" "SLV_PERSISTENT=0" ...]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import abi as A, scenes as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shards", nargs="*", default=["0,8", "3,8"])
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--settings", nargs="*", default=[""])
ap.add_argument("--aniso", type=int, default=16)
ap.add_argument("--stages", action="store_true")
a = ap.parse_args()

for setting in a.settings:
    env = dict(kv.split("=") for kv in setting.split() if "=" in kv)
    for k, v in env.items():
        os.environ[k] = v
    try:
        be = pkg.load(0)
    finally:
        for k in env:
            del os.environ[k]
    sc = S.SponzaLike(3840, 2160, 4, max_aniso=a.aniso, ps_program=A.PS_SPONZA_GRAD if a.aniso > 1 else A.PS_SPONZA)
    sc.setup(be)
    for shard in a.shards:
        r, n = map(int, shard.split(","))
        be.set_tile_shard(r, n)
        for f in range(8):
            sc.render(be, f)
        be.flush()
        res = []
        for rep in range(3):
            be.event_record(0)
            for f in range(a.frames):
                sc.render(be, f % 8)
            be.event_record(1)
            res.append(be.event_elapsed_ms(0, 1) / a.frames)
        line = f"[{setting or 'default':40s}] shard {shard:5s}: {min(res):.4f} ms/frame (3 reps: {', '.join(f'{x:.4f}' for x in res)})"
        if a.stages:
            be.profile_enable(True)
            be.query_begin()
            for f in range(8):
                sc.render(be, f)
            be.flush()
            sg = be.profile_stages()
            be.profile_enable(False)
            line += " | stages alone: " + " ".join(f"{k}={v / 8:.3f}" for k, v in sg.items())
        print(line, flush=True)
    be.set_tile_shard(0, 1)
    be.close()
