"""GPU suite: the same seeded random soups as tests/test_fuzz_oracle_vs_reference.py (which pins the oracle to the live reference
on them), the CUDA product against the oracle: every buffer and the six counters bit for bit."""
import pytest

import cases
import fuzz

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("block", range(6))
def test_cuda_equals_oracle_on_random_soups(cuda, oracle, block):
    for seed in range(block * 20, block * 20 + 20):
        kw, a = fuzz.soup_from_seed(seed)
        _, b = fuzz.soup_from_seed(seed)
        a.setup(cuda)
        b.setup(oracle)
        msgs = cases.compare_frames(a.run(cuda, 0), b.run(oracle, 0))
        assert not msgs, f"seed {seed}: {msgs} {kw}"
