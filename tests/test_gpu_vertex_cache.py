"""GPU suite: the post-transform vertex cache (reference salvia/src/core/default_vertex_cache.cpp:128-197).
k_vertex_mark flags the vertices the queued indexed draws reference, k_vertex_shade runs the vertex shader once per flagged
vertex, k_geometry gathers clip-space positions and attributes.  The vertex shader is a pure function of the vertex, so the
cached pipeline must produce the same bits as the per-corner recompute (SLV_VERTEX_CACHE=0) on every scene, including
base_vertex != 0, 16-bit indices, strips, clipping-heavy soups, vertex texture fetch and SASL vertex shaders."""
import os

import numpy as np
import pytest

import cases
from salviarenderer_b200 import abi as A, scenes as S

pytestmark = pytest.mark.gpu


def _device(mode):
    import salviarenderer_b200 as pkg
    os.environ["SLV_VERTEX_CACHE"] = str(mode)
    try:
        return pkg.load(0)
    finally:
        del os.environ["SLV_VERTEX_CACHE"]


@pytest.fixture(scope="module")
def cuda_vc_off(built):
    return _device(0)


@pytest.fixture(scope="module")
def cuda_vc_all(built):
    """Every indexed draw is cached, however small (the default only caches groups with >= 768 index references)."""
    return _device(2)


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cached_equals_recompute(cuda_vc_all, cuda_vc_off, name):
    mk, frames = cases.CASES[name]
    a, b = mk(), mk()
    a.setup(cuda_vc_all)
    b.setup(cuda_vc_off)
    for f in frames[:2]:
        ra, rb = a.run(cuda_vc_all, f), b.run(cuda_vc_off, f)
        msgs = cases.compare_frames(ra, rb)
        assert not msgs, f"{name} frame {f}: {msgs}"


def test_vs_invocations_counts_unique_vertices(cuda_vc_all, cuda_vc_off):
    cuda = cuda_vc_all
    """Cached: the shader runs once per referenced vertex of a vertex-state group (the 24 material draws of the Sponza-like
    mesh share one state, so the whole mesh is transformed once); recompute: three runs per primitive."""
    a, b = S.SponzaLike(640, 360, 4, tex_size=32), S.SponzaLike(640, 360, 4, tex_size=32)
    a.setup(cuda)
    b.setup(cuda_vc_off)
    ra, rb = a.run(cuda, 1), b.run(cuda_vc_off, 1)
    assert not cases.compare_frames(ra, rb)
    n_unique = len(np.unique(a.mesh.indices))
    assert ra.stats["vs_invocations"] == n_unique, (ra.stats["vs_invocations"], n_unique)
    assert rb.stats["vs_invocations"] == 3 * ra.stats["ia_primitives"]
    assert ra.stats["ia_vertices"] == rb.stats["ia_vertices"] == 3 * ra.stats["ia_primitives"]


def test_two_passes_are_two_groups(cuda_vc_all):
    cuda = cuda_vc_all
    """The two passes of the shadow-map scene bind different matrices: two vertex states, the mesh is transformed twice."""
    sc = S.HeightFieldTwoPass(320, 180, 1, nx=60, nz=48)
    sc.setup(cuda)
    r = sc.run(cuda, 0)
    assert r.stats["vs_invocations"] == 2 * 61 * 49


@pytest.mark.parametrize("base", [64, -37])
def test_base_vertex(cuda_vc_all, cuda_vc_off, base):
    """The cache is indexed by the FINAL vertex index (stored index + base_vertex, index_fetcher.cpp:26-115)."""
    mk = lambda: S.TriangleSoup(w=200, h=160, n=400, seed=11, base_vertex=base)  # noqa: E731
    a, b = mk(), mk()
    a.setup(cuda_vc_all)
    b.setup(cuda_vc_off)
    assert not cases.compare_frames(a.run(cuda_vc_all, 0), b.run(cuda_vc_off, 0))


def test_vertex_buffer_update_between_frames(cuda_vc_all, cuda_vc_off):
    cuda = cuda_vc_all
    """The cache is rebuilt per batch: re-uploading the vertex buffer between frames changes the next frame."""
    a, b = S.SponzaLike(480, 270, 1, tex_size=32), S.SponzaLike(480, 270, 1, tex_size=32)
    a.setup(cuda)
    b.setup(cuda_vc_off)
    r0 = a.run(cuda, 0)
    for sc, be in ((a, cuda), (b, cuda_vc_off)):
        vb = sc.mesh.streams[0].copy()
        vb[:, 1] *= 0.5  # squash the mesh
        vb = np.ascontiguousarray(vb, dtype=np.float32)
        be.upload_from_ptr(sc.mesh.upload(be)[0][0], vb.ctypes.data, vb.nbytes)
    ra, rb = a.run(cuda, 0), b.run(cuda_vc_off, 0)
    assert not cases.compare_frames(ra, rb)
    assert not np.array_equal(r0.depth, ra.depth)


def test_default_policy_caches_texture_sampling_vertex_shaders(cuda, cuda_vc_off):
    """Default device: the cache is used where it pays - vertex texture fetch (one sampler run per vertex instead of up to six
    per primitive) - and arithmetic-only programs keep the per-corner recompute, which is faster for them."""
    sc = S.TerrainVTF(640, 360, 1)
    sc.setup(cuda)
    r = sc.run(cuda, 0)
    assert r.stats["vs_invocations"] == len(np.unique(sc.plane.indices)) < 3 * r.stats["ia_primitives"]
    sp = S.SponzaLike(320, 180, 1, tex_size=16)
    sp.setup(cuda)
    r = sp.run(cuda, 0)
    assert r.stats["vs_invocations"] == 3 * r.stats["ia_primitives"]
