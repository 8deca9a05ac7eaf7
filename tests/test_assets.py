"""CPU suite: asset / report formats (salviarenderer_b200/assets.py) restated from the reference's
ext sources (which need FreeImage and cannot be built here): OBJ + MTL layout and de-duplication, texture row order and
default alpha, PNG dump orientation, the *_Profiling.json schema — and an end-to-end check that a scene written to OBJ and
loaded back renders the same image through the C ABI."""
import json
import os

import numpy as np

from salviarenderer_b200 import abi as A, assets, scenes as S

OBJ = """# two materials, shared corners
mtllib scene.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 0.25
vn 0 0 1
usemtl red
f 1/1/1 2/2/1 3/3/1
usemtl blue
f 1/1/1 3/3/1 4//1
usemtl red
f 2/2/1 3/3/1 4 5 6
"""
MTL = """newmtl red
Ka 0.1 0.2 0.3
Kd 1 0 0
Ns 12
illum 2
map_Kd tex\\red.png
newmtl unused
Kd 1 1 1
newmtl blue
Kd 0 0 1
d 0.5
"""


def test_obj_loader_layout_dedupe_and_materials(tmp_path):
    (tmp_path / "scene.obj").write_text(OBJ)
    (tmp_path / "scene.mtl").write_text(MTL)
    m = assets.load_obj(str(tmp_path / "scene.obj"), flip_tex_v=True)
    assert m.vertices.dtype == np.float32 and m.vertices.shape[1] == 12 and m.vertices.strides[0] == 48
    # corners (1/1/1), (2/2/1), (3/3/1) are reused; (4//1) and (4) are distinct vertices (different index triples)
    assert m.indices.tolist() == [0, 1, 2, 0, 2, 3, 1, 2, 4]
    assert m.attrs.tolist() == [1, 2, 1]
    assert [x.name for x in m.materials] == ["", "red", "blue"]  # default material first, then first-use order
    assert np.array_equal(m.vertices[0], np.array([0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0], np.float32))  # v flipped: 1 - 0
    assert np.array_equal(m.vertices[2, 4:8], np.array([1, 0.75, 0, 0], np.float32))
    assert np.array_equal(m.vertices[3], np.array([0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0], np.float32))  # no texcoord: zeros
    assert np.array_equal(m.vertices[4, 8:], np.zeros(4, np.float32))                                  # no normal: zeros
    red, blue = m.materials[1], m.materials[2]
    assert red.diffuse == (1.0, 0.0, 0.0, 0.0) and red.shininess == 12 and red.is_specular and red.tex_name == "tex/red.png"
    assert abs(red.ambient[1] - np.float32(0.2)) < 1e-12 and blue.alpha == 0.5 and not blue.is_specular
    groups = m.material_groups()
    assert [g[0] for g in groups] == [1, 2]
    assert groups[0][1].tolist() == [0, 1, 2, 1, 2, 4] and groups[1][1].tolist() == [0, 2, 3]


def test_texture_rows_bottom_up_and_png_round_trip(tmp_path):
    from PIL import Image
    rgb = np.arange(2 * 3 * 3, dtype=np.uint8).reshape(2, 3, 3)
    Image.fromarray(rgb, "RGB").save(tmp_path / "t.png")
    t = assets.load_texture_rgba8(str(tmp_path / "t.png"))
    assert t.shape == (2, 3, 4)
    assert np.array_equal(t[0, :, :3], rgb[1]) and np.array_equal(t[1, :, :3], rgb[0])  # bottom row first
    assert (t[..., 3] == 0).all()                                                       # RGB: default alpha 0
    surf = np.random.default_rng(2).integers(0, 256, (5, 7, 4), dtype=np.uint8)
    assets.save_surface_png(str(tmp_path / "s.png"), surf, "bgra8")
    back = assets.load_texture_rgba8(str(tmp_path / "s.png"))
    assert np.array_equal(back, surf[..., [2, 1, 0, 3]])  # same rows (row 0 = bottom of the file), channels as RGBA
    top_row_of_file = np.asarray(Image.open(tmp_path / "s.png").convert("RGBA"))[0]
    assert np.array_equal(top_row_of_file, surf[-1][:, [2, 1, 0, 3]])


def test_profiling_json_schema(tmp_path):
    frames = [{"cinvocations": 10, "cprimitives": 4, "ia_primitives": 10, "ia_vertices": 30, "vs_invocations": 14, "ps_invocations": 100,
               "backend_input_pixels": 90, "ras": 5000, "clipping": 70},
              {"cinvocations": 20, "cprimitives": 9, "ia_primitives": 20, "ia_vertices": 60, "vs_invocations": 30, "ps_invocations": 301,
               "backend_input_pixels": 250, "ras": 7001, "clipping": 90}]
    p = assets.save_profiling_json("Sponza", "nvcc 12.9 / sm_100a", frames, str(tmp_path))
    assert os.path.basename(p) == "Sponza_Profiling.json"
    d = json.load(open(p))
    assert d["frames"] == "2" and d["compiler"].startswith("nvcc")
    assert d["async"]["pipeline_stat"]["ps_invocations"] == {"min": "100", "max": "301", "total": "401", "avg": "200"}
    assert d["async"]["internal_stat"]["backend_input_pixels"]["total"] == "340"
    assert set(d["async"]["pipeline_prof"]) == set(assets.PIPELINE_PROF_KEYS) and d["async"]["pipeline_prof"]["vp_trans"]["max"] == "0"


def test_scene_through_obj_renders_identically(oracle, tmp_path):
    """The Sponza-like mesh written as OBJ + MTL, loaded back and drawn per material group gives the same frame."""
    sc = S.SponzaLike(256, 144, 1, tex_size=32)
    vb, ib = sc.mesh.streams[0], sc.mesh.indices.astype(np.int64)
    lines = ["mtllib s.mtl"]
    vf = vb.astype(np.float64).tolist()  # repr of a double holding a float32 value round-trips exactly
    for v in vf:
        lines.append(f"v {v[0]!r} {v[1]!r} {v[2]!r}")
    for v in vf:
        lines.append(f"vt {v[4]!r} {v[5]!r}")
    for v in vf:
        lines.append(f"vn {v[8]!r} {v[9]!r} {v[10]!r}")
    for m, start, count in sc.groups:
        lines.append(f"usemtl m{m}")
        for t in ib[start * 3:(start + count) * 3].reshape(-1, 3) + 1:
            lines.append("f " + " ".join(f"{i}/{i}/{i}" for i in t))
    (tmp_path / "s.obj").write_text("\n".join(lines) + "\n")
    (tmp_path / "s.mtl").write_text("".join(f"newmtl m{m}\nKd 1 1 1\n" for m, _, _ in sc.groups))
    obj = assets.load_obj(str(tmp_path / "s.obj"))
    assert len(obj.indices) == len(ib) and [x.name for x in obj.materials[1:]] == [f"m{m}" for m, _, _ in sc.groups]

    sc.setup(oracle)
    want = sc.run(oracle, 2)
    # same scene, geometry from the OBJ: one shared 48-byte vertex buffer, one index buffer + draw per material
    elements = [(0, S._V4, 0, 0, 1.0), (1, S._V4, 0, 16, 0.0), (2, S._V4, 0, 32, 0.0)]
    t = sc.t
    oracle.query_begin()
    oracle.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
    oracle.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
    wvp, light, eye = sc.frame_uniforms(2)
    for (mi, idx), (m, _, _) in zip(obj.material_groups(), sc.groups):
        mesh = S.Mesh([obj.vertices], elements, idx, len(idx) // 3)
        d = S.base_desc(t, sc.w, sc.h, cull=A.CULL_BACK)
        mesh.fill_desc(oracle, d)
        d.vs = A.shader_binding(A.VS_SPONZA, S.pack_vs_sponza(wvp, light, eye))
        d.ps = A.shader_binding(A.PS_SPONZA, S.pack_ps_sponza(True), [sc.samplers[m]])
        d.bs = A.shader_binding(A.BS_REPLACE)
        oracle.draw(d)
    got = S.read_frame(oracle, t, oracle.query_get())
    assert np.array_equal(got.color, want.color) and np.array_equal(got.depth, want.depth)
    assert got.stats["cprimitives"] == want.stats["cprimitives"] and got.stats["ps_invocations"] == want.stats["ps_invocations"]
