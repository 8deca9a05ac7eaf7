"""The sampler on textures that are NOT a power of two and on textures larger than 1024 - the sizes on which wrap addressing
goes through the reference's float modulo (sampler.cpp:42-44, 88-92) instead of an exact integer mask, and on which
make_mip_surface's unchecked 2x+1 / 2y+1 reads matter (surface.cpp:71-79).

CPU (build container): the oracle against the LIVE unmodified reference.  GPU (-m gpu): the CUDA product against the oracle,
and k_mipgen against the numpy restatement of the defined chain.  Textures: the reference's own font/font_enu.png (400x400),
a non-square 100x60 one (odd widths AND heights down the chain) and a 2048x2048 one (> 1024: float modulo on a power of two).
Levels whose texels depend on reads past the allocation (undefined upstream) are overwritten with the defined chain on every
backend (scenes.make_texture(defined_mips=True)); all other levels must agree byte for byte as generated."""
import numpy as np
import pytest

from salviarenderer_b200 import abi as A, scenes as S
from test_oracle_vs_reference import probe_inputs, sampler_configs


def npot_textures():
    yield "font_enu_400x400", S.asset_texture("font_enu.png")
    yield "noise_100x60", np.ascontiguousarray(S.noise_texture(128, 11)[:60, :100])
    yield "noise_2048x2048", np.ascontiguousarray(np.tile(S.noise_texture(512, 12), (4, 4, 1)) ^ np.arange(2048, dtype=np.uint8)[None, :, None])


def first_undefined_level(chain):
    return next((l + 1 for l, c in enumerate(chain[:-1]) if (c.shape[0] & 1) or (c.shape[1] & 1)), len(chain))


def run_npot_matrix(bea, beb, n_probes=600):
    coords, ddx, ddy, lod = probe_inputs(n_probes, seed=9)
    bad = []
    for name, img in npot_textures():
        chain = S.mip_chain_rgba8(img)
        ta, tb = S.make_texture(bea, img, defined_mips=True), S.make_texture(beb, img, defined_mips=True)
        assert bea.level_count(ta) == beb.level_count(tb) == len(chain), name
        for l in range(len(chain)):
            assert np.array_equal(bea.read_texture(ta, l), beb.read_texture(tb, l)), f"{name}: mip level {l}"
        big = img.shape[0] > 1024
        for cfg in sampler_configs():
            minf, magf, mipf, q, au, av, an = cfg
            if big and (q != 1 or (au, av) not in ((0, 0), (1, 1), (0, 2))):
                continue  # the 2048^2 texture: the wrap / mirror paths at one LOD quality (the CPU reference is slow on it)
            d = A.sampler_desc(minf, magf, mipf, q, au, av, an, border=(0.25, 0.5, 0.75, 1.0))
            sa, sb = bea.create_sampler(d, ta), beb.create_sampler(d, tb)
            x, y = bea.sampler_probe(sa, coords, ddx, ddy), beb.sampler_probe(sb, coords, ddx, ddy)
            if not np.array_equal(x.view(np.uint32), y.view(np.uint32)):
                bad.append((name, "grad") + cfg)
            if mipf != 2:
                x, y = bea.sampler_probe(sa, coords, lod=lod), beb.sampler_probe(sb, coords, lod=lod)
                if not np.array_equal(x.view(np.uint32), y.view(np.uint32)):
                    bad.append((name, "lod") + cfg)
    return bad


def test_numpy_chain_equals_oracle_mipgen(oracle):
    for name, img in npot_textures():
        chain = S.mip_chain_rgba8(img)
        t = S.make_texture(oracle, img)
        for l, c in enumerate(chain):
            assert np.array_equal(oracle.read_texture(t, l)[:, :, 0, :], c), f"{name}: level {l}"


def test_reference_mipgen_equals_chain_where_defined(oracle, reference):
    """The unmodified reference's gen_mipmap equals the defined chain on every level above the first out-of-allocation read,
    and below it everywhere except the texels that descend from such a read (the last row / the last texel)."""
    for name, img in npot_textures():
        chain = S.mip_chain_rgba8(img)
        t = S.make_texture(reference, img)
        first = first_undefined_level(chain)
        for l in range(min(first, len(chain))):
            assert np.array_equal(reference.read_texture(t, l)[:, :, 0, :], chain[l]), f"{name}: level {l}"
        if first < len(chain):
            got, want = reference.read_texture(t, first)[:, :, 0, :], chain[first]
            assert np.array_equal(got[:-1, :-1], want[:-1, :-1]), f"{name}: level {first} away from the last row / column"


def test_npot_sampler_matrix_oracle_equals_reference(oracle, reference):
    bad = run_npot_matrix(oracle, reference)
    assert not bad, bad[:10]


@pytest.mark.gpu
def test_numpy_chain_equals_cuda_mipgen(cuda):
    for name, img in npot_textures():
        chain = S.mip_chain_rgba8(img)
        t = S.make_texture(cuda, img)
        for l, c in enumerate(chain):
            assert np.array_equal(cuda.read_texture(t, l)[:, :, 0, :], c), f"{name}: level {l}"


@pytest.mark.gpu
def test_npot_sampler_matrix_cuda_equals_oracle(cuda, oracle):
    bad = run_npot_matrix(cuda, oracle, n_probes=1500)
    assert not bad, bad[:10]
