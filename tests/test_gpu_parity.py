"""GPU suite (-m gpu): the CUDA product, through the C ABI, against (1) the frozen reference fixtures,
(2) the CPU oracle buffer by buffer, (3) the sampler matrix, and (4) size-independent properties at the
full BASELINE.json sizes.  Integer decisions (coverage, depth, stencil, counters) are gated bit-exact; colour is
ALSO bit-exact in practice (the north_star tolerance is <= 1 LSB on < 0.01 % of pixels)."""
import json
import os

import numpy as np
import pytest

import cases
from conftest import ROOT
from salviarenderer_b200 import abi as A, scenes as S
from test_oracle_vs_reference import run_probe_matrix

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))["cases"]
COLOR_TOL_LSB = 1  # north_star: colour may differ by at most 1 LSB per 8-bit channel on < 0.01 % of pixels


def test_native_library_is_the_one_running(cuda):
    assert cuda.name == "cuda-sm100a"
    assert cuda.lib_path.endswith("salviarenderer_b200/csrc/libsalvia_b200.so")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_matches_reference_golden(cuda, name):
    mk, frames = cases.CASES[name]
    sc = mk()
    sc.setup(cuda)
    for f in frames:
        got = cases.summarize(sc.run(cuda, f))
        want = GOLDEN[name][str(f)]
        for k in ("stats", "depth", "stencil", "count"):  # bit-exact class
            assert got.get(k) == want.get(k), f"{name} frame {f}: {k}"
        if got["color"] != want["color"] or got.get("resolved") != want.get("resolved"):
            pytest.fail(f"{name} frame {f}: colour hash differs from the reference fixture "
                        f"(see test_cuda_matches_oracle for the per-pixel tolerance check)")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_matches_oracle(cuda, oracle, name):
    mk, frames = cases.CASES[name]
    a, b = mk(), mk()
    a.setup(cuda)
    b.setup(oracle)
    for f in frames[:2]:
        msgs = cases.compare_frames(a.run(cuda, f), b.run(oracle, f), color_tol=COLOR_TOL_LSB)
        assert not msgs, f"{name} frame {f}: {msgs}"


def test_sampler_matrix_cuda_equals_oracle(cuda, oracle):
    bad = run_probe_matrix(cuda, oracle, fmts=(A.PF_RGBA8, A.PF_RGBA32F, A.PF_RG32F))
    assert not bad, bad[:10]


def test_empty_and_degenerate_draws(cuda, oracle):
    """prim_count = 0, fully culled, fully clipped and zero-area primitives leave the targets untouched."""
    for be in (cuda, oracle):
        sc = S.TriangleSoup(samples=4, n=64, seed=3)
        sc.setup(be)
        sc.mesh.prim_count = 0
        r0 = sc.run(be)
        assert r0.stats["ps_invocations"] == 0 and r0.count.max() == 0
    a, b = S.TriangleSoup(samples=1, n=50, seed=4, size=0.0), S.TriangleSoup(samples=1, n=50, seed=4, size=0.0)
    a.setup(cuda)
    b.setup(oracle)
    assert not cases.compare_frames(a.run(cuda), b.run(oracle))


def test_render_is_deterministic_and_idempotent_clear(cuda):
    sc = S.SponzaLike(1280, 720, 4, tex_size=256)
    sc.setup(cuda)
    r1, r2 = sc.run(cuda, 2), sc.run(cuda, 2)
    assert not cases.compare_frames(r1, r2)


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_tile_sharding_partitions_the_frame(cuda, nranks):
    """Sort-first split: the union of the ranks' owned tiles equals the unsharded frame, counters add up."""
    sc = S.SponzaLike(1920, 1080, 4, tex_size=128)
    sc.setup(cuda)
    full = sc.run(cuda, 1)
    acc_color = np.zeros_like(full.color)
    acc_depth = np.zeros_like(full.depth)
    ps = 0
    ty, tx = np.mgrid[0:full.color.shape[0], 0:full.color.shape[1]]
    owner = ((tx // 64) + 3 * (ty // 64)) % nranks
    try:
        for r in range(nranks):
            cuda.set_tile_shard(r, nranks)
            part = sc.run(cuda, 1)
            m = owner == r
            acc_color[m] = part.color[m]
            acc_depth[m] = part.depth[m]
            ps += part.stats["ps_invocations"]
            # nothing outside the owned tiles was touched (still the clear colour / depth)
            assert np.all(part.depth[~m] == 1.0)
    finally:
        cuda.set_tile_shard(0, 1)
    assert np.array_equal(acc_color, full.color)
    assert np.array_equal(acc_depth.view(np.uint32), full.depth.view(np.uint32))
    assert ps == full.stats["ps_invocations"]


def test_full_size_sponza_4k_msaa4_properties(cuda, oracle):
    """BASELINE.json configs[3] at full size (3840x2160, 4x MSAA): (a) counters that are exact integer
    functions of the scene, (b) a 256x256 window re-rendered by the oracle through a viewport-independent check is
    too slow on CPU, so parity at this size is pinned by: resolve == box filter of the samples (recomputed in
    numpy with the reference's op order), depth never above the clear value, and determinism."""
    sc = S.SponzaLike(3840, 2160, 4, tex_size=512)
    sc.setup(cuda)
    r = sc.run(cuda, 3)
    assert r.stats["ia_primitives"] == 262249 and r.stats["ia_vertices"] == 786747 and r.stats["cinvocations"] == 262249
    assert r.stats["ps_invocations"] % 4 == 0 and r.stats["ps_invocations"] >= r.stats["backend_input_pixels"] > 0
    assert float(r.depth.max()) <= 1.0
    # resolve: sum of to_rgba32f(sample) in order, * (1/S), RNE (surface.cpp:123-140)
    inv255 = np.float32(1.0) / np.float32(255)
    s = r.color.astype(np.float32) * inv255
    acc = np.zeros(s.shape[:2] + (4,), np.float32)
    for k in range(4):
        acc = (acc + s[:, :, k, :]).astype(np.float32)
    acc = (acc * np.float32(0.25)).astype(np.float32)
    want = np.rint(np.clip(acc * np.float32(255), 0, 255)).astype(np.uint8)
    assert np.array_equal(want, r.resolved[:, :, 0, :])
    r2 = sc.run(cuda, 3)
    assert not cases.compare_frames(r, r2)


def test_full_size_texture_and_blending_1080p(cuda, oracle):
    """BASELINE.json configs[1] at full size against the oracle, buffer by buffer."""
    a, b = S.TextureAndBlending(1920, 1080), S.TextureAndBlending(1920, 1080)
    a.setup(cuda)
    b.setup(oracle)
    for f in (0, 4):
        assert not cases.compare_frames(a.run(cuda, f), b.run(oracle, f), color_tol=COLOR_TOL_LSB)


def test_full_size_antialiasing_1080p_msaa4(cuda, oracle):
    """BASELINE.json configs[2] at full size (1920x1080, 4x MSAA + resolve) against the oracle."""
    a, b = S.ColorizedTriangle(1920, 1080, 4), S.ColorizedTriangle(1920, 1080, 4)
    a.setup(cuda)
    b.setup(oracle)
    assert not cases.compare_frames(a.run(cuda, 2), b.run(oracle, 2), color_tol=COLOR_TOL_LSB)
