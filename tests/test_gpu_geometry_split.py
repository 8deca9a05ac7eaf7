"""GPU suite: two alternative forms of the front half, both off by default (they did not pay: profiles/
r02_front_half_experiments.txt) and both required to be bit-identical to the default on every parity case:
* the two-kernel geometry (SLV_GEOMETRY_SPLIT=1: k_geometry_cull runs the position pass of every primitive at full occupancy,
  survivors -> k_geometry) - clipping-heavy soups (clipped primitives survive the first kernel and are counted by the second),
  culling, strips, base vertices, vertex texture fetch, SASL vertex shaders, a sharded device;
* the big-triangle queue (SLV_BIG_TILES=1: triangles spanning many tiles are counted by a warp of k_big_tiles)."""
import os

import numpy as np
import pytest

import cases
from salviarenderer_b200 import abi as A, scenes as S

pytestmark = pytest.mark.gpu


def _device(mode, var="SLV_GEOMETRY_SPLIT"):
    import salviarenderer_b200 as pkg
    os.environ[var] = str(mode)
    try:
        return pkg.load(0)
    finally:
        del os.environ[var]


@pytest.fixture(scope="module")
def cuda_split(built):
    return _device(1)


@pytest.fixture(scope="module")
def cuda_fused(built):
    return _device(0)


@pytest.fixture(scope="module")
def cuda_big_tiles(built):
    return _device(1, "SLV_BIG_TILES")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_big_triangle_queue_equals_default(cuda_big_tiles, cuda_fused, name):
    mk, frames = cases.CASES[name]
    a, b = mk(), mk()
    a.setup(cuda_big_tiles)
    b.setup(cuda_fused)
    for f in frames[:1]:
        msgs = cases.compare_frames(a.run(cuda_big_tiles, f), b.run(cuda_fused, f))
        assert not msgs, f"{name} frame {f}: {msgs}"


def test_big_triangle_queue_with_many_fullscreen_triangles(cuda_big_tiles, cuda_fused):
    mk = lambda: S.TriangleSoup(w=1920, h=1080, samples=1, n=200, size=4.0, seed=3, bs=A.BS_REPLACE)  # noqa: E731
    a, b = mk(), mk()
    a.setup(cuda_big_tiles)
    b.setup(cuda_fused)
    assert not cases.compare_frames(a.run(cuda_big_tiles, 0), b.run(cuda_fused, 0))


@pytest.mark.parametrize("name", list(cases.CASES))
def test_split_equals_fused(cuda_split, cuda_fused, name):
    mk, frames = cases.CASES[name]
    a, b = mk(), mk()
    a.setup(cuda_split)
    b.setup(cuda_fused)
    for f in frames[:2]:
        msgs = cases.compare_frames(a.run(cuda_split, f), b.run(cuda_fused, f))
        assert not msgs, f"{name} frame {f}: {msgs}"


def test_split_with_sasl_vertex_shader_and_sharding(cuda_split, cuda_fused):
    import bench
    for be in (cuda_split, cuda_fused):
        be.set_tile_shard(3, 8)
    try:
        a, b = S.SponzaLike(960, 540, 4, tex_size=64), S.SponzaLike(960, 540, 4, tex_size=64)
        bench.install_sasl_shaders(a, cuda_split, A)
        bench.install_sasl_shaders(b, cuda_fused, A)
        a.setup(cuda_split)
        b.setup(cuda_fused)
        for f in (0, 5):
            ra, rb = a.run(cuda_split, f), b.run(cuda_fused, f)
            assert not cases.compare_frames(ra, rb)
            assert ra.stats["cprimitives"] > 0 and 0 < ra.stats["ps_invocations"]
    finally:
        for be in (cuda_split, cuda_fused):
            be.set_tile_shard(0, 1)
