"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/libsalvia_ref.so).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The fixture stores, per case and frame, a sha256 prefix of every output buffer (colour, depth bits, stencil,
resolved colour, per-sample coverage counter) and the gated pipeline counters — an exact fingerprint.
The reference is rendered twice per frame; cases that are not deterministic run-to-run are refused.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from salviarenderer_b200 import abi  # noqa: E402
import cases  # noqa: E402


def main():
    ref = abi.Backend(os.path.join(ROOT, "oracle", "_ref", "libsalvia_ref.so"))
    assert ref.name == "reference"
    out = {}
    for name, (mk, frames) in cases.CASES.items():
        sc = mk()
        sc.setup(ref)
        out[name] = {}
        for f in frames:
            a = cases.summarize(sc.run(ref, f))
            b = cases.summarize(sc.run(ref, f))
            assert a == b, f"reference not deterministic on {name} frame {f}"
            out[name][str(f)] = a
        print(name, "ok", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as fh:
        json.dump({"generator": "tests/golden/make_golden.py", "source": "oracle/_ref (unmodified reference)",
                   "cases": out}, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
