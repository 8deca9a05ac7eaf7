"""CPU suite: asset / report formats (salviarenderer_b200/assets.py).  The OBJ + MTL loader is PINNED to the
reference's own loader: oracle/ref_obj_dump.cpp compiles salvia/src/ext/resource/mesh/{mesh_io_obj,mesh_impl,material}.cpp in
place (make -C oracle obj-dump) and prints a fingerprint of every mesh it builds; the tests compare assets.load_obj with it live
where the binary exists (this container) and with the committed copy of its output (tests/golden/obj_loader.json) everywhere.
The texture loader upstream needs FreeImage and cannot be built here: texture row order and default alpha, the PNG dump
orientation and the *_Profiling.json schema are restated - plus an end-to-end check that a scene written to OBJ and loaded back
renders the same image through the C ABI."""
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
    # corners (1/1/1), (2/2/1), (3/3/1) are reused; `4//1` and the bare `4` that follows BOTH key as (4, 3, 1): the texcoord /
    # normal indices of the key are stale from the corner before (upstream quirk) - one vertex, with zero uv
    assert m.indices.tolist() == [0, 1, 2, 0, 2, 3, 1, 2, 3]
    assert m.attrs.tolist() == [1, 2, 1]
    assert [x.name for x in m.materials] == ["default", "red", "blue"]  # default material first, then first-use order
    assert np.array_equal(m.vertices[0], np.array([0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 1, 0], np.float32))  # v flipped: 1 - 0
    assert np.array_equal(m.vertices[2, 4:8], np.array([1, 0.75, 0, 0], np.float32))
    assert np.array_equal(m.vertices[3], np.array([0, 1, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0], np.float32))  # no texcoord: zeros
    red, blue = m.materials[1], m.materials[2]
    # `newmtl unused` names a material the OBJ never used: upstream keeps `red` selected, so unused's Kd lands on red
    assert red.diffuse == (1.0, 1.0, 1.0, 0.0) and red.shininess == 12 and red.is_specular and red.tex_name == "tex/red.png"
    assert abs(red.ambient[1] - np.float32(0.2)) < 1e-12 and blue.alpha == 0.5 and blue.is_specular and blue.shininess == 2
    assert blue.ambient == m.materials[0].ambient and blue.ambient[3] == 1.0  # constructor defaults (material.cpp)
    groups = m.material_groups()
    assert [g[0] for g in groups] == [1, 2]
    assert groups[0][1].tolist() == [0, 1, 2, 1, 2, 3] and groups[1][1].tolist() == [0, 2, 3]


def _fnv(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).view(np.uint8).reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f"{h:016x}"


def _g9(x) -> str:
    return "%.9g" % float(np.float32(x))


def fingerprint(path: str, flip: bool) -> list:
    """The lines oracle/ref_obj_dump.cpp prints for the reference's loader, computed from assets.load_obj."""
    m = assets.load_obj(path, flip_tex_v=flip)
    groups = m.material_groups()
    out = [f"meshes {len(groups)}"]
    for i, (mi, idx) in enumerate(groups):
        mt = m.materials[mi]
        vec = lambda v: " ".join(_g9(c) for c in v)  # noqa: E731
        out.append(f"mesh {i} prims {len(idx) // 3} vb_bytes {m.vertices.nbytes} vb {_fnv(m.vertices)} ib_bytes {idx.nbytes} ib {_fnv(idx)} "
                   f"name {mt.name} ambient {vec(mt.ambient)} diffuse {vec(mt.diffuse)} specular {vec(mt.specular)} shininess {mt.shininess} "
                   f"tex_name {mt.tex_name}")
    for mt in m.materials:  # load order of the MTL file == the order load_texture was asked (one map_Kd per material here)
        pass
    return out


def _write_cases(tmp_path):
    """The OBJ / MTL files the loader is pinned on: the quirk file above and a 2,000-triangle mesh with shared corners,
    three materials interleaved, partial corner specifications and a material that is defined but unused."""
    (tmp_path / "scene.obj").write_text(OBJ)
    (tmp_path / "scene.mtl").write_text(MTL)
    rng = np.random.default_rng(31)
    n = 40
    lines = ["# grid", "mtllib grid.mtl"]
    for j in range(n + 1):
        for i in range(n + 1):
            lines.append(f"v {i * 0.25:.6f} {rng.uniform(-1, 1):.6f} {j * 0.25:.6f}")
    for j in range(n + 1):
        for i in range(n + 1):
            lines.append(f"vt {i / n:.6f} {j / n:.6f}")
    for k in range(7):
        lines.append(f"vn {rng.uniform(-1, 1):.6f} {rng.uniform(0, 1):.6f} {rng.uniform(-1, 1):.6f}")
    for j in range(n):
        for i in range(n):
            a, b, c, d = j * (n + 1) + i + 1, j * (n + 1) + i + 2, (j + 1) * (n + 1) + i + 1, (j + 1) * (n + 1) + i + 2
            lines.append(f"usemtl {['stone', 'grass', 'water'][(i // 5 + j // 7) % 3]}")
            kind = (i + 3 * j) % 4
            nn = (i * j) % 7 + 1
            if kind == 0:
                lines += [f"f {a}/{a}/{nn} {c}/{c}/{nn} {b}/{b}/{nn}", f"f {b}/{b}/{nn} {c}/{c}/{nn} {d}/{d}/{nn}"]
            elif kind == 1:
                lines += [f"f {a}//{nn} {c}//{nn} {b}//{nn}", f"f {b}/{b} {c}/{c} {d}/{d}"]
            elif kind == 2:
                lines += [f"f {a} {c} {b}", f"f {b}/{b}/{nn} {c} {d}//{nn} {a}"]
            else:
                lines += [f"f {a}/{a}/{nn} {c}/{c}/{nn} {b}/{b}/{nn}", f"f {b} {c}/{c}/{nn} {d}"]
    (tmp_path / "grid.obj").write_text("\n".join(lines) + "\n")
    (tmp_path / "grid.mtl").write_text("# materials\nnewmtl grass\nKa 0.05 0.2 0.05\nKd 0.1 0.8 0.2\nKs 0 0 0\nNs 3\nillum 1\nmap_Kd tex\\grass.png\n"
                                       "newmtl lava\nKd 1 0.3 0\nd 0.25\nnewmtl stone\nKd 0.5 0.5 0.55\nNs 40\nillum 2\n"
                                       "newmtl water\nKd 0 0.2 0.9\nTr 0.5\nmap_Kd water.png\n")
    return [("scene.obj", True), ("scene.obj", False), ("grid.obj", True), ("grid.obj", False)]


GOLDEN_OBJ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obj_loader.json")
REF_OBJ_DUMP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_obj_dump")


def test_obj_loader_equals_the_reference_loader(tmp_path):
    """assets.load_obj against salvia_ext's create_mesh_from_obj: per mesh the primitive count, the bytes of the shared vertex
    buffer and of the index buffer, and every material field.  Live against oracle/_ref/ref_obj_dump where it is built, and
    against the committed copy of its output."""
    import subprocess
    golden = json.load(open(GOLDEN_OBJ))
    live = os.path.exists(REF_OBJ_DUMP)
    for name, flip in _write_cases(tmp_path):
        mine = fingerprint(str(tmp_path / name), flip)
        key = f"{name} flip={int(flip)}"
        assert mine == golden["cases"][key], key
        if live:
            out = subprocess.run([REF_OBJ_DUMP, str(tmp_path / name), "1" if flip else "0"], capture_output=True, text=True, timeout=120)
            assert out.returncode == 0, out.stderr
            ref_lines = [ln for ln in out.stdout.splitlines() if not ln.startswith("texture ")]
            assert mine == ref_lines, key
            asked = [os.path.basename(ln.split(" ", 1)[1]) for ln in out.stdout.splitlines() if ln.startswith("texture ")]
            assert asked == golden["textures"][key]


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
