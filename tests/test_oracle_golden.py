"""CPU suite: the oracle restatement against the frozen fixtures generated from the unmodified reference
(tests/golden/golden.json, made by tests/golden/make_golden.py) — the parity pin of the oracle."""
import json
import os

import pytest

import cases
from conftest import ROOT

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]


@pytest.mark.parametrize("name", cases.CPU_CASES)
def test_oracle_matches_reference_golden(oracle, name):
    mk, frames = cases.CASES[name]
    sc = mk()
    sc.setup(oracle)
    for f in frames:
        got = cases.summarize(sc.run(oracle, f))
        assert got == GOLDEN[name][str(f)], f"{name} frame {f}"


def test_golden_covers_every_case():
    assert sorted(GOLDEN) == sorted(cases.CASES)
