"""Asset and report formats either side of the draw path: what the reference's samples load
and what they write, so that real assets (Sponza's OBJ + MTL + textures) drop in and outputs can be diffed against
upstream runs.

* OBJ + MTL -> the reference's mesh layout (salvia/src/ext/resource/mesh/mesh_io_obj.cpp:41-45, 161-288, 389-452):
  one shared vertex buffer of 48-byte vertices {pos.xyzw, uv.xyzw, normal.xyzw}, vertices de-duplicated by their
  (position, texcoord, normal) index triple in first-use order, one u32 index list per material in MTL/usemtl order.
* image -> rgba8 texels (salvia/src/ext/resource/texture/tex_io.cpp:28-75): rows bottom-up (FreeImage's scanline order,
  no flip), RGB images get alpha 0 (`default_alpha`, freeimage_utilities.h:46-58), RGBA keep theirs.
* surface -> PNG (tex_io.cpp:141-188): the surface's rows are written to FreeImage scanlines 0.., i.e. surface row 0
  is the BOTTOM row of the file; channels converted to bgra8 first.
* `<name>_Profiling.json` (salvia/src/utility/common/sample_app.cpp:448-567): boost::property_tree JSON — every leaf is
  a string; per counter {min, max, total, avg} over the frames.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np


def _f32(x):
    return float(np.float32(x))


@dataclass
class ObjMaterial:  # obj_material (material.h) with the constructor's defaults (salvia/src/ext/resource/mesh/material.cpp:4-13)
    name: str = "default"
    ambient: tuple = (_f32(0.2), _f32(0.2), _f32(0.2), 1.0)
    diffuse: tuple = (0.5, 0.5, 0.5, 1.0)
    specular: tuple = (_f32(0.7), _f32(0.7), _f32(0.7), 1.0)
    alpha: float = 1.0
    shininess: int = 2
    is_specular: bool = True
    tex_name: str = ""
    tex_path: str = ""


@dataclass
class ObjMesh:
    vertices: np.ndarray                  # (n, 12) float32: pos.xyzw, uv.xyzw, normal.xyzw  (48-byte stride)
    indices: np.ndarray                   # (3 * n_tris,) uint32 into `vertices`
    attrs: np.ndarray                     # (n_tris,) uint32 material index of each triangle
    materials: list = field(default_factory=list)

    def material_groups(self):
        """construct_meshes (mesh_io_obj.cpp:389-437): (material index, u32 index array) for every material that has
        triangles, in material order; all groups share the one vertex buffer."""
        out = []
        tri = self.indices.reshape(-1, 3)
        for m in range(len(self.materials)):
            sel = tri[self.attrs == m]
            if len(sel):
                out.append((m, np.ascontiguousarray(sel.reshape(-1), dtype=np.uint32)))
        return out


def _f(tok):
    try:
        return float(np.float32(tok))
    except ValueError:
        return 0.0


def load_mtl(path: str, materials: list):
    """load_material (mesh_io_obj.cpp:62-136): fills the materials `usemtl` created.  Upstream quirk, mirrored (pinned against
    the reference's own loader by tests/test_assets.py): a `newmtl` the OBJ never used does NOT deselect the current material,
    so its statements overwrite the previously selected one."""
    if not os.path.exists(path):
        return False
    cur = None
    base = os.path.dirname(path)
    for line in open(path, errors="replace"):
        tok = line.split()
        if not tok or tok[0].startswith("#"):
            continue
        cmd = tok[0]
        if cmd == "newmtl":
            hit = next((m for m in materials if m.name == (tok[1] if len(tok) > 1 else "")), None)
            if hit is not None:
                cur = hit
            continue
        if cur is None:
            continue
        if cmd in ("Ka", "Kd", "Ks") and len(tok) >= 4:
            v = (_f(tok[1]), _f(tok[2]), _f(tok[3]), 0.0)
            setattr(cur, {"Ka": "ambient", "Kd": "diffuse", "Ks": "specular"}[cmd], v)
        elif cmd in ("d", "Tr") and len(tok) >= 2:
            cur.alpha = _f(tok[1])
        elif cmd == "Ns" and len(tok) >= 2:
            cur.shininess = int(_f(tok[1]))
        elif cmd == "illum" and len(tok) >= 2:
            cur.is_specular = int(_f(tok[1])) == 2
        elif cmd == "map_Kd" and len(tok) >= 2:
            cur.tex_name = tok[1].replace("\\", "/")
            cur.tex_path = os.path.join(base, cur.tex_name)
    return True


def load_obj(path: str, flip_tex_v: bool = False) -> ObjMesh:
    """load_obj_mesh_c (mesh_io_obj.cpp:140-269).  Triangulated faces only (the reference reads exactly three corners of
    every `f` line and ignores the rest); material 0 is the default material ("default").  Upstream quirk, mirrored: the
    texcoord / normal indices of the de-duplication key are reset per FACE, not per corner, so a corner that omits them is keyed
    with the indices of the face's previous corner (its data are still zero) - pinned against the reference's own loader by tests/test_assets.py."""
    positions, uvs, normals = [], [], []
    verts, indices, attrs = [], [], []
    materials = [ObjMaterial()]
    subset, mtl_file = 0, ""
    seen: dict[tuple, int] = {}
    zero4 = (0.0, 0.0, 0.0, 0.0)
    for line in open(path, errors="replace"):
        tok = line.split()
        if not tok:
            continue
        cmd = tok[0]
        if cmd[0] == "#":
            continue
        if cmd == "v":
            x, y, z = (_f(t) for t in (tok[1:4] + ["0"] * 3)[:3])
            positions.append((x, y, z, 1.0))
        elif cmd == "vt":
            u, v = (_f(t) for t in (tok[1:3] + ["0"] * 2)[:2])
            uvs.append((u, float(np.float32(1.0) - np.float32(v)) if flip_tex_v else v, 0.0, 0.0))
        elif cmd == "vn":
            x, y, z = (_f(t) for t in (tok[1:4] + ["0"] * 3)[:3])
            normals.append((x, y, z, 0.0))
        elif cmd == "f":
            ti = ni = 0  # declared per face, outside the corner loop, upstream: a corner inherits the face's previous indices
            for corner in tok[1:4]:
                parts = corner.split("/")
                pi = int(parts[0])
                has_t = len(parts) > 1 and parts[1] != ""
                has_n = len(parts) > 2 and parts[2] != ""
                if has_t:
                    ti = int(parts[1])
                if has_n:
                    ni = int(parts[2])
                key = (pi, ti, ni)
                idx = seen.get(key)
                if idx is None:
                    idx = len(verts)
                    seen[key] = idx
                    verts.append(positions[pi - 1] + (uvs[ti - 1] if has_t else zero4) + (normals[ni - 1] if has_n else zero4))
                indices.append(idx)
            attrs.append(subset)
        elif cmd == "mtllib" and len(tok) > 1:
            mtl_file = tok[1]
        elif cmd == "usemtl" and len(tok) > 1:
            name = tok[1]
            for i, m in enumerate(materials):
                if m.name == name:
                    subset = i
                    break
            else:
                subset = len(materials)
                materials.append(ObjMaterial(name=name))
    if mtl_file:
        load_mtl(os.path.join(os.path.dirname(os.path.abspath(path)), mtl_file), materials)
    return ObjMesh(np.asarray(verts, dtype=np.float32).reshape(-1, 12), np.asarray(indices, dtype=np.uint32),
                   np.asarray(attrs, dtype=np.uint32), materials)


def load_texture_rgba8(path: str) -> np.ndarray:
    """load_texture(..., pixel_format_color_rgba8) (tex_io.cpp:28-95): (h, w, 4) uint8, row 0 = the file's bottom row."""
    from PIL import Image
    img = Image.open(path)
    if img.mode in ("RGBA", "LA", "PA") or (img.mode == "P" and "transparency" in img.info):
        a = np.asarray(img.convert("RGBA"), dtype=np.uint8)
    else:
        rgb = np.asarray(img.convert("RGB"), dtype=np.uint8)
        a = np.concatenate([rgb, np.zeros(rgb.shape[:2] + (1,), np.uint8)], axis=-1)  # default_alpha = 0
    return np.ascontiguousarray(a[::-1])


def save_surface_png(path: str, texels: np.ndarray, fmt: str = "rgba8"):
    """save_surface(..., pixel_format_color_bgra8) (tex_io.cpp:141-188).  `texels`: (h, w, 4) uint8 in the surface's own
    channel order (`fmt` 'rgba8' or 'bgra8'), row 0 first in memory = bottom row of the PNG."""
    from PIL import Image
    a = np.asarray(texels, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected (h, w, 4) uint8 texels")
    rgba = a[..., [2, 1, 0, 3]] if fmt == "bgra8" else a
    Image.fromarray(np.ascontiguousarray(rgba[::-1]), "RGBA").save(path, format="PNG")


PIPELINE_STAT_KEYS = ("cinvocations", "cprimitives", "ia_primitives", "ia_vertices", "vs_invocations", "ps_invocations")
PIPELINE_PROF_KEYS = ("gather_vtx", "vtx_proc", "clipping", "compact_clip", "vp_trans", "tri_dispatch", "ras")


def profiling_report(compiler: str, frames: list) -> dict:
    """sample_app::save_profiling_result (sample_app.cpp:468-567).  `frames`: one dict per frame with the keys of
    PIPELINE_STAT_KEYS, 'backend_input_pixels' and PIPELINE_PROF_KEYS (missing keys count as 0).  Leaves are strings, as
    boost::property_tree::write_json emits them."""
    def reduce(key):
        vals = [int(f.get(key, 0)) for f in frames]
        total = sum(vals)
        return {"min": str(min(vals) if vals else 0), "max": str(max(vals) if vals else 0), "total": str(total),
                "avg": str(total // len(vals) if vals else 0)}
    return {
        "compiler": compiler,
        "frames": str(len(frames)),
        "async": {
            "pipeline_stat": {k: reduce(k) for k in PIPELINE_STAT_KEYS},
            "internal_stat": {"backend_input_pixels": reduce("backend_input_pixels")},
            "pipeline_prof": {k: reduce(k) for k in PIPELINE_PROF_KEYS},
        },
    }


def save_profiling_json(benchmark_name: str, compiler: str, frames: list, directory: str = ".") -> str:
    path = os.path.join(directory, f"{benchmark_name}_Profiling.json")
    with open(path, "w") as f:
        json.dump(profiling_report(compiler, frames), f, indent=4)
    return path
