"""Generates tests/golden/golden_fullsize.json: fingerprints of FULL-SIZE BASELINE.json configurations rendered by the UNMODIFIED
reference (oracle/_ref/libsalvia_ref.so) - the cases that are too slow for the CPU suite and for make_golden.py's double runs.

    python tests/golden/make_golden_fullsize.py          (build container: needs /root/reference for oracle/_ref)

configs[4]: the synthetic 10,000,000-triangle height field at 7680x4320, two passes (depth-only shadow pass + colour pass),
frame 0 - about a minute of reference time on 8 cores.  The GPU suite renders the same frame with the product and compares the
sha-256 of every buffer and the gated counters (tests/test_gpu_parity.py::test_full_size_configs4_equals_reference_fixture)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from salviarenderer_b200 import abi, scenes as S  # noqa: E402
import cases  # noqa: E402

FULL = {
    "c5_10m_tris_7680x4320": (lambda: S.HeightFieldTwoPass(7680, 4320, 1, nx=2500, nz=2000), (0,)),
}


def main():
    ref = abi.Backend(os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so"))
    assert ref.name == "reference"
    out = {}
    for name, (mk, frames) in FULL.items():
        sc = mk()
        sc.setup(ref)
        out[name] = {}
        for f in frames:
            t0 = time.time()
            out[name][str(f)] = cases.summarize(sc.run(ref, f))
            print(name, f, f"{time.time() - t0:.1f} s", out[name][str(f)]["stats"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden_fullsize.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden_fullsize.py", "source": "oracle/_ref (unmodified reference)", "cases": out},
                  fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
