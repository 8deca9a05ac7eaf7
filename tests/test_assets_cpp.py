"""CPU suite: the C++ asset / report formats (salviarenderer_b200/host/salvia_b200_assets.hpp) against the Python twin
(salviarenderer_b200/assets.py) and, for the OBJ + MTL loader, against the fingerprints of the reference's own loader
(tests/golden/obj_loader.json - see tests/test_assets.py for how they were made)."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from salviarenderer_b200 import assets
from test_assets import GOLDEN_OBJ, _fnv, _g9, _write_cases, fingerprint


@pytest.fixture(scope="module")
def cli(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("assets_cli") / "assets_cli")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "salviarenderer_b200", "host"), "-o", exe,
                        os.path.join(ROOT, "tests", "cpp", "assets_cli.cpp"), "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def run(cli, *args):
    r = subprocess.run([cli, *map(str, args)], capture_output=True, text=True, timeout=120)
    return r.returncode, r.stdout.splitlines()


def test_obj_loader_equals_python_and_the_reference_fingerprints(cli, tmp_path):
    golden = json.load(open(GOLDEN_OBJ))
    for name, flip in _write_cases(tmp_path):
        path = str(tmp_path / name)
        rc, lines = run(cli, "obj", path, int(flip))
        assert rc == 0
        mesh_lines = [ln for ln in lines if ln.startswith(("meshes ", "mesh "))]
        assert mesh_lines == golden["cases"][f"{name} flip={int(flip)}"] == fingerprint(path, flip)
        m = assets.load_obj(path, flip_tex_v=flip)
        vec = lambda v: " ".join(_g9(c) for c in v)  # noqa: E731
        want = [f"material {mt.name} ambient {vec(mt.ambient)} diffuse {vec(mt.diffuse)} specular {vec(mt.specular)} alpha {_g9(mt.alpha)} "
                f"shininess {mt.shininess} is_specular {int(mt.is_specular)} tex_name {mt.tex_name} tex_path {mt.tex_path}" for mt in m.materials]
        assert [ln for ln in lines if ln.startswith("material ")] == want
        assert lines[-1] == f"indices {_fnv(m.indices)} attrs {_fnv(m.attrs)}"
    assert run(cli, "obj", str(tmp_path / "missing.obj"), 0)[0] == 1


def test_png_reader_equals_python(cli, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(7)
    w, h = 13, 9
    rgba = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    files = {}
    Image.fromarray(rgba, "RGBA").save(tmp_path / "rgba.png")
    Image.fromarray(rgba[..., :3], "RGB").save(tmp_path / "rgb.png")
    Image.fromarray(rgba[..., 0], "L").save(tmp_path / "l.png")
    Image.fromarray(rgba[..., :2], "LA").save(tmp_path / "la.png")
    pal = Image.fromarray(rgba[..., :3], "RGB").quantize(16)
    pal.save(tmp_path / "p.png")                                      # 4-bit palette
    Image.fromarray(rgba[..., :3], "RGB").quantize(200).save(tmp_path / "p8.png")   # 8-bit palette
    pal.save(tmp_path / "pt.png", transparency=bytes([0, 128, 255, 7]))             # palette + tRNS (shorter than the palette)
    Image.fromarray((rgba[..., 0] > 127).astype(np.uint8) * 255, "L").convert("1").save(tmp_path / "bit.png")  # 1-bit grey
    big = rng.integers(0, 256, (64, 200, 3), dtype=np.uint8)
    big[:, 1:] = (big[:, 1:] // 16) + big[:, :-1] // 2                # smooth rows: the encoder picks the Sub / Up / Average / Paeth filters
    Image.fromarray(big, "RGB").save(tmp_path / "filters.png", optimize=True)
    for name in ("rgba", "rgb", "l", "la", "p", "p8", "pt", "bit", "filters"):
        files[name] = str(tmp_path / f"{name}.png")
    for ref in ("/root/reference/resources/font/font_enu.png", "/root/reference/resources/texture_and_blending/chessboard.png"):
        if os.path.exists(ref):
            files[os.path.basename(ref)] = ref
    for name, path in files.items():
        t = assets.load_texture_rgba8(path)
        rc, lines = run(cli, "tex", path)
        assert rc == 0 and lines == [f"{t.shape[1]} {t.shape[0]} {_fnv(t)}"], name
    assert (assets.load_texture_rgba8(files["rgb"])[..., 3] == 0).all()                 # no alpha channel: alpha 0, in both
    (tmp_path / "junk.png").write_bytes(b"not a png")
    assert run(cli, "tex", str(tmp_path / "junk.png"))[0] == 2


def test_png_writer_round_trip(cli, tmp_path):
    from PIL import Image
    surf = np.random.default_rng(2).integers(0, 256, (5, 7, 4), dtype=np.uint8)
    (tmp_path / "raw.bin").write_bytes(surf.tobytes())
    for bgra in (0, 1):
        out = str(tmp_path / f"s{bgra}.png")
        assert run(cli, "png", 7, 5, bgra, str(tmp_path / "raw.bin"), out)[0] == 0
        want = surf[..., [2, 1, 0, 3]] if bgra else surf
        assert np.array_equal(assets.load_texture_rgba8(out), want)                                       # row 0 = bottom row of the file
        assert np.array_equal(np.asarray(Image.open(out).convert("RGBA"))[0], want[-1])                    # the file's top row is the surface's last
        ref = str(tmp_path / f"py{bgra}.png")
        assets.save_surface_png(ref, surf, "bgra8" if bgra else "rgba8")
        assert np.array_equal(np.asarray(Image.open(out)), np.asarray(Image.open(ref)))                    # same image as the Python writer
        rc, lines = run(cli, "tex", out)
        assert rc == 0 and lines == [f"7 5 {_fnv(want)}"]                                                  # and the C++ reader reads its own files


def test_profiling_json_is_byte_identical(cli, tmp_path):
    frames = [{"cinvocations": 10, "cprimitives": 4, "ia_primitives": 10, "ia_vertices": 30, "vs_invocations": 14, "ps_invocations": 100,
               "backend_input_pixels": 90, "ras": 5000, "clipping": 70},
              {"cinvocations": 20, "cprimitives": 9, "ia_primitives": 20, "ia_vertices": 60, "vs_invocations": 30, "ps_invocations": 301,
               "backend_input_pixels": 250, "ras": 7001, "clipping": 90}]
    for k, fr in enumerate((frames, [])):
        (tmp_path / "frames.txt").write_text("".join(" ".join(f"{a}={b}" for a, b in f.items()) + "\n" for f in fr))
        (tmp_path / f"py{k}").mkdir()
        (tmp_path / f"cpp{k}").mkdir()
        compiler = 'nvcc 12.9 / sm_100a "quoted"'
        p = assets.save_profiling_json("Sponza", compiler, fr, str(tmp_path / f"py{k}"))
        rc, lines = run(cli, "prof", compiler, str(tmp_path / "frames.txt"), str(tmp_path / f"cpp{k}"), "Sponza")
        assert rc == 0 and os.path.basename(lines[0]) == "Sponza_Profiling.json"
        assert open(lines[0]).read() == open(p).read()


def test_cpp_sample_application_renders_the_obj_scene(oracle, tmp_path):
    """tests/cpp/obj_viewer_test.cpp - a sample application written like samples/Sponza/Sponza.cpp, all in C++: OBJ + MTL and the
    map_Kd PNG textures through salvia_b200_assets.hpp, the Sponza shader twins through the host surface, the frame out as a PNG
    and the counters as ObjViewer_Profiling.json - against the same scene rendered through the Python path, on the oracle."""
    from conftest import ORACLE_LIB
    from salviarenderer_b200 import abi as A, scenes as S
    sc = S.SponzaLike(256, 144, 1, tex_size=32, color_fmt=A.PF_RGBA8)
    vb, ib = sc.mesh.streams[0], sc.mesh.indices.astype(np.int64)
    lines = ["mtllib s.mtl"]
    vf = vb.astype(np.float64).tolist()  # repr of a double holding a float32 value round-trips exactly
    lines += [f"v {v[0]!r} {v[1]!r} {v[2]!r}" for v in vf] + [f"vt {v[4]!r} {v[5]!r}" for v in vf] + [f"vn {v[8]!r} {v[9]!r} {v[10]!r}" for v in vf]
    for m, start, count in sc.groups:
        lines.append(f"usemtl m{m}")
        lines += ["f " + " ".join(f"{i}/{i}/{i}" for i in t) for t in ib[start * 3:(start + count) * 3].reshape(-1, 3) + 1]
    (tmp_path / "s.obj").write_text("\n".join(lines) + "\n")
    (tmp_path / "s.mtl").write_text("".join(f"newmtl m{m}\nKd 1 1 1\nmap_Kd tex_{m}.png\n" for m, _, _ in sc.groups))
    for m, _, _ in sc.groups:
        assets.save_surface_png(str(tmp_path / f"tex_{m}.png"), S.brick_texture(32, seed=100 + m))
    wvp, light, eye = sc.frame_uniforms(2)
    np.concatenate([np.asarray(wvp, np.float32).reshape(-1), np.asarray(light, np.float32), np.asarray(eye, np.float32)]).tofile(tmp_path / "uniforms.bin")

    exe = str(tmp_path / "obj_viewer_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "salviarenderer_b200", "host"),
                        os.path.join(ROOT, "tests", "cpp", "obj_viewer_test.cpp"), "-o", exe, "-ldl", "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    out = subprocess.run([exe, ORACLE_LIB, str(tmp_path / "s.obj"), str(tmp_path / "uniforms.bin"), "256", "144", str(tmp_path / "frame.png"), str(tmp_path)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr

    sc.setup(oracle)
    want = sc.run(oracle, 2)
    got = assets.load_texture_rgba8(str(tmp_path / "frame.png"))          # row 0 = bottom row of the file = surface row 0
    assert np.array_equal(got, want.color[:, :, 0, :])
    report = json.load(open(tmp_path / "ObjViewer_Profiling.json"))
    assert report["compiler"] == "oracle" and report["frames"] == "1"
    # (vs_invocations depends on how the indices are split over buffers - one per material here - through the vertex cache)
    for key in ("ia_primitives", "ia_vertices", "cinvocations", "cprimitives", "ps_invocations"):
        assert report["async"]["pipeline_stat"][key]["total"] == str(want.stats[key]), key
