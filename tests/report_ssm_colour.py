"""Test infrastructure (report, not a test): how close the CUDA colour of the StandardShadowMap cases (expf / logf / pow in the pixel shader) is to the
CPU oracle's — number of differing samples and the largest difference in LSB.  Test infrastructure (uses oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import salviarenderer_b200 as pkg  # noqa: E402
from salviarenderer_b200 import abi as A, scenes as S  # noqa: E402

cuda = pkg.load(0)
ora = A.Backend(os.path.join(ROOT, "oracle", "libsalvia_oracle.so"))
for w, h, s, frames in ((640, 360, 1, range(8)), (1920, 1080, 4, (0, 4))):
    a, b = S.StandardShadowMap(w, h, s), S.StandardShadowMap(w, h, s)
    a.setup(cuda)
    b.setup(ora)
    for f in frames:
        ra, rb = a.run(cuda, f), b.run(ora, f)
        d = np.abs(ra.color.astype(np.int32) - rb.color.astype(np.int32))
        print(f"{w}x{h}x{s} frame {f}: colour samples differing {int((d > 0).any(-1).sum())} of {d[..., 0].size}, max |d| {int(d.max())} LSB; "
              f"depth equal {np.array_equal(ra.depth.view(np.uint32), rb.depth.view(np.uint32))}, shadow map equal "
              f"{np.array_equal(ra.count.view(np.uint32), rb.count.view(np.uint32))}, counters equal {all(ra.stats[k] == rb.stats[k] for k in ('cprimitives', 'ps_invocations'))}",
              flush=True)
