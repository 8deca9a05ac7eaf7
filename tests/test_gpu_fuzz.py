"""GPU suite: the same seeded random soups as tests/test_fuzz_oracle_vs_reference.py (which pins the oracle to the live reference
on them), the CUDA product against the oracle: every buffer and the six counters bit for bit - and the same for the seeded random
instances of the textured scenes (fuzz.scene_from_seed, 400 seeds; profiles/r02_random_scenes_gpu.json)."""
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


@pytest.mark.parametrize("block", range(8))
def test_cuda_equals_oracle_on_random_textured_scenes(cuda, oracle, block):
    for seed in range(block * 50, block * 50 + 50):
        a, frame, what = fuzz.scene_from_seed(seed)
        b, _, _ = fuzz.scene_from_seed(seed)
        a.setup(cuda)
        b.setup(oracle)
        msgs = cases.compare_frames(a.run(cuda, frame), b.run(oracle, frame), color_tol=fuzz.scene_tolerance(a))
        assert not msgs, f"seed {seed} ({what}, {type(a).__name__}, frame {frame}): {msgs}"
