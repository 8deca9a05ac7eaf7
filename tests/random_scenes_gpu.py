"""Random textured scenes (tests/fuzz.py scene_from_seed), the CUDA product against the oracle: every buffer and counter.
No torch, no pytest: starts in a second, stops at the time limit and reports how far it got.
    python tests/random_scenes_gpu.py [--viewports] [first_seed] [last_seed] [seconds]  ->  gpurun_out/random_scenes_gpu.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repository root
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import cases  # noqa: E402
import fuzz  # noqa: E402
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import abi  # noqa: E402

viewports = "--viewports" in sys.argv  # the random-viewport soups (fuzz.viewport_soup_from_seed) instead of the textured scenes
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
first, last, limit = (int(a) for a in (argv[:3] + ["0", "400", "60"][len(argv):]))
t0 = time.time()
cuda, oracle = pkg.load(0), abi.Backend(os.path.join(ROOT, "oracle", "libsalvia_oracle.so"))
out = {"first": first, "equal": 0, "bad": [], "last_seed_run": None}
for seed in range(first, last):
    if time.time() - t0 > limit:
        break
    if viewports:
        (kw, a), (_, b), f = fuzz.viewport_soup_from_seed(seed), fuzz.viewport_soup_from_seed(seed), 0
        what = str(kw["viewport"])
    else:
        a, f, what = fuzz.scene_from_seed(seed)
        b, _, _ = fuzz.scene_from_seed(seed)
    a.setup(cuda)
    b.setup(oracle)
    msgs = cases.compare_frames(a.run(cuda, f), b.run(oracle, f), color_tol=fuzz.scene_tolerance(a))
    out["last_seed_run"] = seed
    if msgs:
        out["bad"].append({"seed": seed, "what": what, "scene": type(a).__name__, "frame": f, "msgs": [str(m) for m in msgs[:3]]})
    else:
        out["equal"] += 1
out["seconds"] = round(time.time() - t0, 1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "random_scenes_gpu.json"), "w"), indent=1)
print(json.dumps(out))
