"""CPU suite (build container only): seeded random points of the TriangleSoup parameter space (tests/fuzz.py: target size,
MSAA, cull, depth / two-sided stencil state, blend shader and colour format, strips, index width, base vertex, split and
non-indexed draws, attribute modifiers), the oracle restatement against the LIVE unmodified reference, every buffer and counter.
500 seeds were run when this was written (all equal); the suite keeps the first 120.  The second test does the same for random
instances of the TEXTURED scenes (fuzz.scene_from_seed: texture + blending with every mip filter and both derivative conventions,
the Sponza-like atrium trilinear / anisotropic, the AnisotropicFilter sample, shadow map, vertex texture fetch, two-pass height
field; random size, sample count and frame): 400 seeds, all equal."""
import pytest

import cases
import fuzz


@pytest.mark.parametrize("block", range(6))
def test_oracle_equals_live_reference_on_random_soups(oracle, reference, block):
    for seed in range(block * 20, block * 20 + 20):
        kw, a = fuzz.soup_from_seed(seed)
        _, b = fuzz.soup_from_seed(seed)
        a.setup(oracle)
        b.setup(reference)
        msgs = cases.compare_frames(a.run(oracle, 0), b.run(reference, 0))
        assert not msgs, f"seed {seed}: {msgs} {kw}"


@pytest.mark.parametrize("block", range(8))
def test_oracle_equals_live_reference_on_random_textured_scenes(oracle, reference, block):
    for seed in range(block * 50, block * 50 + 50):
        a, frame, what = fuzz.scene_from_seed(seed)
        b, _, _ = fuzz.scene_from_seed(seed)
        a.setup(oracle)
        b.setup(reference)
        msgs = cases.compare_frames(a.run(oracle, frame), b.run(reference, frame), color_tol=fuzz.scene_tolerance(a))
        assert not msgs, f"seed {seed} ({what}, {type(a).__name__}, frame {frame}): {msgs}"


@pytest.mark.parametrize("block", range(4))
def test_oracle_equals_live_reference_on_random_viewports(oracle, reference, block):
    """Sub-rectangle, fractional and oversized viewports and depth ranges other than 0..1 (fuzz.viewport_soup_from_seed): 500
    seeds equal when this was written, the suite keeps 120."""
    for seed in range(block * 30, block * 30 + 30):
        kw, a = fuzz.viewport_soup_from_seed(seed)
        _, b = fuzz.viewport_soup_from_seed(seed)
        a.setup(oracle)
        b.setup(reference)
        msgs = cases.compare_frames(a.run(oracle, 0), b.run(reference, 0))
        assert not msgs, f"seed {seed}: {msgs} {kw}"
