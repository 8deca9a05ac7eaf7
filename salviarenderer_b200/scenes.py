"""Frozen synthetic workloads of SURVEY.md §8d, expressed as slv_* command streams.

Every scene is a pure function of its parameters: float32 inputs are produced here once (numpy, fixed
seeds) and the SAME bytes are handed to whichever backend renders them — the CUDA product, the CPU
restatement or the unmodified reference — so results are comparable bit for bit.

Scene sources in the reference (what each restates):
  colorized_triangle : samples/ColorizedTriangle/ColorizedTriangle.cpp:108-205
  texture_and_blending: samples/TextureAndBlending/TextureAndBlending.cpp:196-347
  anisotropic_filter : samples/AnisotropicFilter/AnisotropicFilter.cpp
  sponza_like        : samples/Sponza/Sponza.cpp:143-278 (assets are Git-LFS pointers -> procedural atrium)
  grid_stress        : SURVEY §6 smoke probe (jittered height-field, one draw)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import struct
from dataclasses import dataclass, field

import numpy as np

from . import abi as A

f32 = np.float32


# ---- eflib-style float32 math (row-vector convention, eflib/src/math.cpp:142-154,418-570) ------------
def mat_identity():
    return np.eye(4, dtype=f32)


def mat_translate(x, y, z):
    m = np.eye(4, dtype=f32)
    m[3, :3] = (x, y, z)
    return m


def mat_scale(x, y, z):
    return np.diag(np.array([x, y, z, 1], dtype=f32)).astype(f32)


def mat_rotate_y(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = np.eye(4, dtype=f32)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def mat_mul(a, b):
    return (a.astype(f32) @ b.astype(f32)).astype(f32)


def _normalize(v):
    v = np.asarray(v, dtype=f32)
    return (v / f32(np.sqrt(np.sum(v * v, dtype=f32)))).astype(f32)


def mat_lookat(eye, target, up):
    eye, target, up = (np.asarray(v, dtype=f32) for v in (eye, target, up))
    z = _normalize(target - eye)
    x = _normalize(np.cross(up, z).astype(f32))
    y = np.cross(z, x).astype(f32)
    m = np.zeros((4, 4), dtype=f32)
    m[:3, 0], m[:3, 1], m[:3, 2] = x, y, z
    m[3, 0], m[3, 1], m[3, 2], m[3, 3] = -np.dot(x, eye), -np.dot(y, eye), -np.dot(z, eye), 1
    return m


def mat_perspective_fov(fovy, aspect, n, f):
    ys = f32(1.0 / math.tan(fovy / 2))
    xs = f32(ys / f32(aspect))
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0], m[1, 1] = xs, ys
    m[2, 2], m[2, 3] = f32(f / (f - n)), 1
    m[3, 2] = f32(-n * f / (f - n))
    return m


def mat_ortho(l, r, b, t, n, f):
    m = np.zeros((4, 4), dtype=f32)
    m[0, 0], m[1, 1], m[2, 2] = 2 / (r - l), 2 / (t - b), 1 / (f - n)
    m[3, 0], m[3, 1], m[3, 2], m[3, 3] = (l + r) / (l - r), (t + b) / (b - t), n / (n - f), 1
    return m.astype(f32)


# ---- geometry ----------------------------------------------------------------------------------------
@dataclass
class Mesh:
    """Host-side mesh: vertex streams (each an (n, k) float32 array), elements, index array."""
    streams: list            # list[np.ndarray float32 (nverts, comps)]
    elements: list           # list[(reg, fmt, slot, byte_offset, default_w)]
    indices: np.ndarray      # uint16 / uint32
    prim_count: int
    topology: int = A.TOPO_TRIANGLE_LIST
    handles: dict = field(default_factory=dict)

    def upload(self, be: A.Backend):
        key = id(be)
        if key not in self.handles:
            vb = [be.create_buffer(np.ascontiguousarray(s, dtype=f32)) for s in self.streams]
            ib = be.create_buffer(np.ascontiguousarray(self.indices)) if self.indices is not None else 0
            self.handles[key] = (vb, ib)
        return self.handles[key]

    def fill_desc(self, be: A.Backend, d: A.DrawDesc, start=0, prim_count=None, base_vertex=0):
        vb, ib = self.upload(be)
        d.n_streams = len(vb)
        for i, (h, s) in enumerate(zip(vb, self.streams)):
            d.streams[i].buffer, d.streams[i].stride, d.streams[i].offset = h, s.shape[1] * 4, 0
        d.n_elements = len(self.elements)
        for i, (reg, fmt, slot, off, dw) in enumerate(self.elements):
            e = d.elements[i]
            e.reg, e.format, e.slot, e.aligned_byte_offset, e.default_w = reg, fmt, slot, off, dw
        d.index_buffer = ib
        if self.indices is None:
            d.index_format = A.INDEX_NONE
        else:
            d.index_format = A.INDEX_R16_UINT if self.indices.dtype == np.uint16 else A.INDEX_R32_UINT
        d.topology = self.topology
        d.start, d.prim_count, d.base_vertex = start, self.prim_count if prim_count is None else prim_count, base_vertex


_V4 = A.FMT_R32G32B32A32_FLOAT


def create_planar(start, xdir, ydir, rx, ry, positive_normal=False, index_dtype=np.uint16) -> Mesh:
    """salvia/src/ext/resource/mesh/mesh_io.cpp:182-265: 3 vec4 streams (pos w=1, normal, uv)."""
    start, xdir, ydir = (np.asarray(v, dtype=f32) for v in (start, xdir, ydir))
    n = _normalize(np.cross(xdir, ydir).astype(f32))
    if not positive_normal:
        n = -n
    pos, nor, uv = [], [], []
    line = np.array([*start, 1], dtype=f32)
    x4, y4 = np.array([*xdir, 0], dtype=f32), np.array([*ydir, 0], dtype=f32)
    for i in range(rx + 1):
        p = line.copy()
        for j in range(ry + 1):
            pos.append(p.copy())
            nor.append(np.array([*n, 0], dtype=f32))
            uv.append(np.array([f32(i) / f32(rx), f32(j) / f32(ry), 0, 0], dtype=f32))
            p = (p + y4).astype(f32)
        line = (line + x4).astype(f32)
    idx = []
    for i in range(rx):
        for j in range(ry):
            q0 = i * (ry + 1) + j
            q2 = q0 + ry + 2
            idx += [q0, q0 + 1, q2, q2, q2 - 1, q0]
    return Mesh([np.array(pos, f32), np.array(nor, f32), np.array(uv, f32)],
                [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0), (2, _V4, 2, 0, 0.0)],
                np.array(idx, dtype=index_dtype), rx * ry * 2)


def create_box() -> Mesh:
    """mesh_io.cpp:35-180: unit box, 24 verts / 12 tris, streams pos/normal/uv as vec4, u16 indices."""
    faces = [  # (normal, 4 corner positions)
        ((1, 0, 0), [(1, 0, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1)]),
        ((-1, 0, 0), [(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0)]),
        ((0, 1, 0), [(0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 1, 0)]),
        ((0, -1, 0), [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1)]),
        ((0, 0, 1), [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]),
        ((0, 0, -1), [(0, 0, 0), (0, 1, 0), (1, 1, 0), (1, 0, 0)]),
    ]
    uvs4 = [(0, 0), (1, 0), (1, 1), (0, 1)]
    pos, nor, uv, idx = [], [], [], []
    for f, (n, corners) in enumerate(faces):
        for k, c in enumerate(corners):
            pos.append((*c, 1))
            nor.append((*n, 0))
            uv.append((*uvs4[k], 0, 0))
        b = f * 4
        idx += [b, b + 1, b + 2, b + 2, b + 3, b]
    return Mesh([np.array(pos, f32), np.array(nor, f32), np.array(uv, f32)],
                [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0), (2, _V4, 2, 0, 0.0)],
                np.array(idx, dtype=np.uint16), 12)


# ---- uniform packing -----------------------------------------------------------------------------------
def u_mvp_passthrough(wvp, src):
    src = list(src) + [0] * (5 - len(src))
    return wvp.astype(f32).tobytes() + struct.pack("<I5I", len([s for s in src[:5]]) if False else 0, *src[:5])


def pack_vs_mvp_passthrough(wvp, src):
    n = len(src)
    src = list(src) + [0] * (5 - n)
    return np.asarray(wvp, f32).tobytes() + struct.pack("<6I", n, *src)


def pack_vs_plane_xz(wvp):
    return np.asarray(wvp, f32).tobytes()


def pack_vs_lights3(wvp, lights):
    return np.asarray(wvp, f32).tobytes() + np.asarray(lights, f32).reshape(3, 4).tobytes()


def pack_vs_sponza(wvp, light, eye):
    return np.asarray(wvp, f32).tobytes() + np.asarray(light, f32).tobytes() + np.asarray(eye, f32).tobytes()


def pack_ps_tex_alpha(reg, alpha, sasl_derivatives=False):
    return struct.pack("<IfI", reg, alpha, int(sasl_derivatives))


def pack_ps_sponza(has_sampler):
    return struct.pack("<I", int(has_sampler))


def pack_ps_sponza_grad(has_sampler, sasl_derivatives=True):
    return struct.pack("<II", int(has_sampler), int(sasl_derivatives))


def pack_vs_ssm_draw(camera_wvp, light_wvp, light_pos, camera_pos):
    return (np.asarray(camera_wvp, f32).tobytes() + np.asarray(light_wvp, f32).tobytes() + np.asarray(light_pos, f32).tobytes() +
            np.asarray(camera_pos, f32).tobytes())


def pack_ps_ssm_draw(ambient, diffuse, specular, shininess, has_tex, has_depth):
    return struct.pack("<12fiII", *ambient, *diffuse, *specular, int(shininess), int(has_tex), int(has_depth))


# ---- frame targets ---------------------------------------------------------------------------------------
@dataclass
class Targets:
    color: A.Texture
    ds: A.Texture
    resolved: A.Texture | None
    count: A.Texture | None = None   # rgba32f coverage counter (MRT 1), optional


def create_targets(be: A.Backend, w, h, samples, color_fmt, with_count=False) -> Targets:
    color = be.create_texture(w, h, samples, color_fmt)
    ds = be.create_texture(w, h, samples, A.PF_RG32F)
    resolved = be.create_texture(w, h, 1, color_fmt) if samples > 1 else None
    count = be.create_texture(w, h, samples, A.PF_RGBA32F) if with_count else None
    return Targets(color, ds, resolved, count)


def base_desc(t: Targets, w, h, cull=A.CULL_BACK, ds=None) -> A.DrawDesc:
    d = A.DrawDesc()
    d.raster.cull_mode, d.raster.front_ccw = cull, 0
    d.ds = ds if ds is not None else A.depth_stencil_desc()
    d.stencil_ref = 0
    d.viewport.x, d.viewport.y, d.viewport.w, d.viewport.h = 0, 0, w, h
    d.viewport.minz, d.viewport.maxz = 0.0, 1.0
    if t.count is not None:
        d.n_color_targets = 2
        d.color_targets[0], d.color_targets[1] = t.color.handle, t.count.handle
    else:
        d.n_color_targets = 1
        d.color_targets[0] = t.color.handle
    d.ds_target = t.ds.handle
    return d


@dataclass
class FrameResult:
    color: np.ndarray                 # uint8 [h, w, S, 4]
    depth: np.ndarray                 # float32 [h, w, S]
    stencil: np.ndarray               # uint32 [h, w, S]
    resolved: np.ndarray | None       # uint8 [h, w, 1, 4]
    count: np.ndarray | None          # float32 [h, w, S] coverage counter
    stats: dict


def read_frame(be: A.Backend, t: Targets, stats=None) -> FrameResult:
    color = be.read_texture(t.color)
    ds = be.read_texture(t.ds)
    dsf = ds.view(np.float32).reshape(ds.shape[0], ds.shape[1], ds.shape[2], 2)
    depth = dsf[..., 0].copy()
    stencil = dsf[..., 1].copy().view(np.uint32)
    resolved = be.read_texture(t.resolved) if t.resolved is not None else None
    count = None
    if t.count is not None:
        c = be.read_texture(t.count)
        count = c.view(np.float32).reshape(c.shape[0], c.shape[1], c.shape[2], 4)[..., 0].copy()
    return FrameResult(color, depth, stencil, resolved, count, stats or {})


# ===========================================================================================================
# C1 / C3a: ColorizedTriangle (and AntiAliasing = same scene with 4x MSAA + resolve)
# ===========================================================================================================
class ColorizedTriangle:
    """samples/ColorizedTriangle/ColorizedTriangle.cpp:108-205; test-mode camera: angle -= 0.55f per frame
    accumulated in float32 BEFORE use (:166-175)."""

    def __init__(self, w=800, h=600, samples=1, with_count=False):
        self.w, self.h, self.samples, self.with_count = w, h, samples, with_count
        self.mesh = create_planar((-3.0, -1.0, -3.0), (6, 0, 0), (0, 0, 6), 1, 1, False)
        self.n_frames = 5

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_RGBA8, self.with_count)
        self.mesh.upload(be)

    def frame_uniforms(self, frame):
        ang = f32(0.0)
        for _ in range(frame + 1):
            ang = f32(ang - f32(0.55))
        ang = float(ang)
        camera = (math.cos(ang) * 2.3, 2.5, math.sin(ang) * 2.3)
        view = mat_lookat(camera, (0, 0, 0), (0, 1, 0))
        proj = mat_perspective_fov(math.pi / 2, f32(self.w) / f32(self.h), 0.1, 100.0)
        world = mat_translate(-0.5, 0, -0.5)
        wvp = mat_mul(world, mat_mul(view, proj))
        lights = [
            (math.sin(-ang * 1.5) * 2.2, 0.15, math.cos(ang * 0.9) * 1.8, 0.0),
            (math.sin(ang * 0.7) * 1.9, 0.15, math.cos(-ang * 0.4) * 2.5, 0.0),
            (math.sin(ang * 2.6) * 2.3, 0.15, math.cos(ang * 0.6) * 1.7, 0.0),
        ]
        return wvp, lights

    def render(self, be: A.Backend, frame: int):
        t = self.t
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        if t.count is not None:
            be.clear_color(t.count, (0, 0, 0, 0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        wvp, lights = self.frame_uniforms(frame)
        d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
        self.mesh.fill_desc(be, d)
        d.vs = A.shader_binding(A.VS_LIGHTS3, pack_vs_lights3(wvp, lights))
        d.ps = A.shader_binding(A.PS_LIGHTS3)
        d.bs = A.shader_binding(A.BS_REPLACE_AND_COUNT if t.count is not None else A.BS_REPLACE)
        be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


# ---- textures ----------------------------------------------------------------------------------------------
ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "assets")


def asset_texture(name: str) -> np.ndarray:
    """rgba8 texels of one of the reference's own image assets (tests/golden/assets/, copied from /root/reference/resources by
    tests/golden/fetch_assets.py), loaded the way load_texture does (tex_io.cpp:28-95: bottom-up rows, RGB gets alpha 0)."""
    from .assets import load_texture_rgba8
    return load_texture_rgba8(os.path.join(ASSET_DIR, name))


def chessboard_texture(n=32, cell=4) -> np.ndarray:
    """Procedural stand-in for resources/texture_and_blending/chessboard.png (32x32 RGBA); the scenes use the real file
    (asset_texture) and fall back to this only when the fixture directory is missing."""
    y, x = np.mgrid[0:n, 0:n]
    on = ((x // cell + y // cell) & 1).astype(np.uint8)
    img = np.empty((n, n, 4), dtype=np.uint8)
    img[..., 0] = 40 + on * 200
    img[..., 1] = 40 + on * 190
    img[..., 2] = 50 + on * 170
    img[..., 3] = 255
    return img


def noise_texture(size=512, seed=20240607) -> np.ndarray:
    """Seeded value-noise RGBA8 (stand-in for Dirt.jpg): uniform bytes + 3x3 box blur."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(size, size, 4), dtype=np.uint8).astype(np.uint32)
    acc = np.zeros_like(raw)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            acc += np.roll(np.roll(raw, dy, 0), dx, 1)
    img = (acc // 9).astype(np.uint8)
    img[..., 3] = 255
    return img


def brick_texture(size=1024, seed=1) -> np.ndarray:
    """Seeded brick/noise RGBA8 used by the Sponza-like atrium materials."""
    rng = np.random.default_rng(seed)
    base = rng.integers(60, 200, size=3)
    y, x = np.mgrid[0:size, 0:size]
    bh, bw = size // 16, size // 8
    row = y // bh
    xs = x + (row & 1) * (bw // 2)
    mortar = ((y % bh) < max(2, bh // 10)) | ((xs % bw) < max(2, bw // 16))
    grain = rng.integers(0, 48, size=(size, size), dtype=np.int32)
    brick_id = (row * 131 + xs // bw * 71) % 29
    img = np.empty((size, size, 4), dtype=np.uint8)
    for c in range(3):
        v = base[c] + (brick_id - 14) * 2 + grain - 24
        v = np.where(mortar, 200 + grain // 4, v)
        img[..., c] = np.clip(v, 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return img


def mip_chain_rgba8(img: np.ndarray) -> list:
    """texture_2d::gen_mipmap(filter_linear) for rgba8 texels, restated in numpy (surface.cpp:53-92, texture.h:26-35): level
    sizes (w + 1) / 2, texel = ((c0 + c1) + c2) + c3) * 0.25 over to_rgba32f texels, RNE back to unorm8.  The reference reads
    texels 2x + 1 / 2y + 1 without a bounds check: for an odd parent width the read wraps into the next
    row (linear addressing, mirrored here), for an odd parent height the last row reads PAST THE ALLOCATION - undefined upstream
    (heap garbage), defined as zeros here, in the product (k_mipgen) and in the oracle."""
    levels = [np.ascontiguousarray(img, dtype=np.uint8)]
    m = max(img.shape[0], img.shape[1])
    limit = 0
    while m > 0:
        m >>= 1
        limit += 1
    inv255 = f32(1.0) / f32(255.0)
    for _ in range(limit - 1):
        src = levels[-1]
        h, w = src.shape[:2]
        mh, mw = (h + 1) // 2, (w + 1) // 2
        flat = np.concatenate([src.reshape(-1, 4).astype(f32) * inv255, np.zeros((2 * w + 4, 4), f32)])
        yy, xx = np.mgrid[0:mh, 0:mw]

        def rd(dx, dy):
            return flat[(2 * yy + dy) * w + 2 * xx + dx]

        o = (((rd(0, 0) + rd(1, 0)) + rd(0, 1)) + rd(1, 1)) * f32(0.25)
        o = np.clip(o * f32(255.0), f32(0.0), f32(255.0))
        levels.append(np.rint(o).astype(np.uint8))
    return levels


def make_texture(be: A.Backend, img: np.ndarray, fmt=A.PF_RGBA8, mips=True, defined_mips=False) -> A.Texture:
    """`defined_mips`: for textures whose chain has odd-sized parents (not a power of two), the levels from the first
    undefined one on are overwritten with mip_chain_rgba8's - a no-op for the product and the oracle (their gen_mipmap IS that
    chain, tests/test_sampler_npot.py), and what makes the unmodified reference sample defined texels instead of whatever its
    out-of-bounds reads returned in this process."""
    h, w = img.shape[:2]
    t = be.create_texture(w, h, 1, fmt)
    be.upload_texture(t, img)
    if mips:
        be.gen_mipmap(t, A.FILTER_LINEAR)
        if defined_mips and fmt == A.PF_RGBA8:
            chain = mip_chain_rgba8(img)
            assert be.level_count(t) == len(chain)
            # the first undefined level is the child of the first parent with an odd height (whole last row reads past the
            # allocation) or an odd width (its bottom-right texel does)
            first = next((l + 1 for l, c in enumerate(chain[:-1]) if (c.shape[0] & 1) or (c.shape[1] & 1)), None)
            if first is not None:
                for l in range(first, len(chain)):
                    be.upload_texture(t, chain[l], level=l)
    return t


# ===========================================================================================================
# C2: TextureAndBlending
# ===========================================================================================================
class TextureAndBlending:
    """samples/TextureAndBlending/TextureAndBlending.cpp:196-347: plane (wrap chessboard, replace) then the
    box twice (cull front, cull back) with lerp(dst, src, 0.5); color target bgra8; trilinear samplers."""

    def __init__(self, w=1920, h=1080, samples=1, ps_program=A.PS_TEX_ALPHA, mip_filter=A.FILTER_LINEAR,
                 max_aniso=0, boxes=True):
        self.w, self.h, self.samples, self.boxes = w, h, samples, boxes
        self.ps_program, self.mip_filter, self.max_aniso = ps_program, mip_filter, max_aniso
        # optional override (tests of run-time compiled SASL shaders): (reg, alpha, sampler) -> ShaderBinding
        self.ps_binding = None
        self.plane = create_planar((-3.0, -1.0, -3.0), (6, 0, 0), (0, 0, 6), 1, 1, True)
        box = create_box()
        # vs_box binds POSITION -> reg0 and TEXCOORD -> reg1 (the uv stream, slot 2)
        box.elements = [(0, _V4, 0, 0, 1.0), (1, _V4, 2, 0, 0.0)]
        self.box = box
        self.plane.elements = [(0, _V4, 0, 0, 1.0)]
        self.n_frames = 5
        # the sample's own chessboard.png (32x32 RGBA); Dirt.jpg is a Git-LFS pointer upstream -> seeded noise
        self.chess = asset_texture("chessboard.png") if os.path.exists(os.path.join(ASSET_DIR, "chessboard.png")) else chessboard_texture()
        self.noise = noise_texture()

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_BGRA8)
        self.plane.upload(be)
        self.box.upload(be)
        self.plane_tex = make_texture(be, self.chess)
        self.box_tex = make_texture(be, self.noise)
        self.plane_samp = be.create_sampler(
            A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, self.mip_filter, addr_u=A.ADDR_WRAP, addr_v=A.ADDR_WRAP,
                           max_anisotropy=self.max_aniso), self.plane_tex)
        self.box_samp = be.create_sampler(
            A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, self.mip_filter, addr_u=A.ADDR_CLAMP, addr_v=A.ADDR_CLAMP,
                           max_anisotropy=self.max_aniso), self.box_tex)

    def render(self, be: A.Backend, frame: int):
        t = self.t
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        scene_sec = float(f32(frame * 6) / f32(self.n_frames - 1))
        angle = -scene_sec * 60.0 * (2 * math.pi / 360.0)
        camera = (math.cos(angle) * 1.5, 1.5, math.sin(angle) * 1.5)
        view = mat_lookat(camera, (0, 0, 0), (0, 1, 0))
        proj = mat_perspective_fov(math.pi / 2, f32(self.w) / f32(self.h), 0.1, 100.0)
        wvp = mat_mul(mat_translate(-0.5, 0, -0.5), mat_mul(view, proj))

        d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
        self.plane.fill_desc(be, d)
        d.vs = A.shader_binding(A.VS_PLANE_XZ, pack_vs_plane_xz(wvp))
        d.ps = self.ps_binding(0, 1.0, self.plane_samp) if self.ps_binding else \
            A.shader_binding(self.ps_program, pack_ps_tex_alpha(0, 1.0), [self.plane_samp])
        d.bs = A.shader_binding(A.BS_REPLACE)
        be.draw(d)

        for cull in ((A.CULL_FRONT, A.CULL_BACK) if self.boxes else ()):
            d = base_desc(t, self.w, self.h, cull=cull)
            self.box.fill_desc(be, d)
            d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, pack_vs_mvp_passthrough(wvp, [0, 1]))
            d.ps = self.ps_binding(1, 0.5, self.box_samp) if self.ps_binding else \
                A.shader_binding(self.ps_program, pack_ps_tex_alpha(1, 0.5), [self.box_samp])
            d.bs = A.shader_binding(A.BS_LERP_SRC_ALPHA)
            be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


# ===========================================================================================================
# C3b: AnisotropicFilter
# ===========================================================================================================
class AnisotropicFilter:
    """samples/AnisotropicFilter/AnisotropicFilter.cpp:44-49,186-339 (BASELINE configs[2]): a tunnel of 96 cylinder segments,
    each one draw of create_planar((-hw, 0, -10), (hw, 0, 0), (0, 0, 1), 2, 50) = 200 triangles rotated about z by i * 3.75 degrees
    at radius 5, camera at (0, 0, -10) looking down the axis; vs_plane (uv = pos.xz), the SASL pixel shader
    `tex2D(samp, tex); color.w = 1` (sample_2d_grad with the SASL per-row / per-column derivatives: SLV_PS_TEX_GRAD_ALPHA with
    sasl_derivatives), replace blend, cull back; texture = the sample's own font/font_enu.png (400x400 RGB -> rgba8 with
    alpha 0, NOT a power of two: mip chain 400, 200, 100, 50, 25, 13, 7, 4, 2 with the odd-size steps of SURVEY App. B #9,
    wrap addressing through the float modulo).  Frame f uses filter row f % 7 of :222-230 (three trilinear qualities,
    then 2x / 4x / 8x / 16x anisotropic).  The sample renders 1 sample per pixel into rgba8; BASELINE configs[2] asks for
    4x MSAA + resolve."""

    FILTER_ROWS = [(A.FILTER_LINEAR, A.MIP_LO, 0), (A.FILTER_LINEAR, A.MIP_MI, 0), (A.FILTER_LINEAR, A.MIP_HI, 0),
                   (A.FILTER_ANISOTROPIC, A.MIP_MI, 2), (A.FILTER_ANISOTROPIC, A.MIP_MI, 4),
                   (A.FILTER_ANISOTROPIC, A.MIP_MI, 8), (A.FILTER_ANISOTROPIC, A.MIP_MI, 16)]
    RADIUS, SEGMENTS = 5.0, 96

    def __init__(self, w=1920, h=1080, samples=4, sasl_derivatives=True, segments=None):
        self.w, self.h, self.samples, self.sasl_derivatives = w, h, samples, sasl_derivatives
        self.segments = self.SEGMENTS if segments is None else segments
        self.seg_angle = f32(360.0) / f32(self.SEGMENTS)
        hw = f32(math.tan(math.radians(float(self.seg_angle) / 2.0))) * f32(self.RADIUS)
        self.plane = create_planar((-hw, 0.0, -10.0), (hw, 0.0, 0.0), (0.0, 0.0, 1.0), 2, 50, True)
        self.plane.elements = [(0, _V4, 0, 0, 1.0)]
        self.n_frames = len(self.FILTER_ROWS)
        self.ps_binding = None  # optional override (run-time compiled SASL shader): sampler -> ShaderBinding
        self.texels = asset_texture("font_enu.png")
        self._draw_cache = {}

    def setup(self, be: A.Backend):
        self._draw_cache = {}
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_RGBA8)
        self.plane.upload(be)
        self.tex = make_texture(be, self.texels, defined_mips=True)
        self.samplers = [be.create_sampler(A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, mip, mip_qual=q, addr_u=A.ADDR_WRAP,
                                                          addr_v=A.ADDR_WRAP, max_anisotropy=an), self.tex)
                         for mip, q, an in self.FILTER_ROWS]

    def frame_draws(self, be: A.Backend, frame: int):
        key = (id(be), frame)
        if key in self._draw_cache:
            return self._draw_cache[key]
        samp = self.samplers[frame % len(self.FILTER_ROWS)]
        view = mat_lookat((0.0, 0.0, -10.0), (0.0, 0.0, 0.0), (0, 1, 0))
        proj = mat_perspective_fov(math.pi / 2, f32(self.w) / f32(self.h), 0.1, 100.0)
        vp = mat_mul(view, proj)
        draws = []
        for i in range(self.segments):
            ang = math.radians(float(f32(i) * self.seg_angle))
            s_, c_ = f32(math.sin(ang)), f32(math.cos(ang))
            rot = np.eye(4, dtype=f32)  # mat_rotZ (eflib/src/math.cpp:405-416)
            rot[0, 0], rot[1, 0], rot[0, 1], rot[1, 1] = c_, -s_, s_, c_
            wvp = mat_mul(mat_mul(mat_translate(0.0, -self.RADIUS, 0.0), rot), vp)
            d = base_desc(self.t, self.w, self.h, cull=A.CULL_BACK)
            self.plane.fill_desc(be, d)
            d.vs = A.shader_binding(A.VS_PLANE_XZ, pack_vs_plane_xz(wvp))
            d.ps = self.ps_binding(samp) if self.ps_binding else \
                A.shader_binding(A.PS_TEX_GRAD_ALPHA, pack_ps_tex_alpha(0, 1.0, self.sasl_derivatives), [samp])
            d.bs = A.shader_binding(A.BS_REPLACE)
            draws.append(d)
        self._draw_cache[key] = draws
        return draws

    def render(self, be: A.Backend, frame: int):
        t = self.t
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        for d in self.frame_draws(be, frame):
            be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


# ===========================================================================================================
# VertexTextureFetch (scope row f-4): a plane displaced in the VERTEX shader by a height map
# ===========================================================================================================
class TerrainVTF:
    """samples/VertexTextureFetch/VertexTextureFetch.cpp:128-252: create_planar grid (2B x 2B quads of 0.5), vertex shader
    tex2Dlod of a height map (mirror addressing, linear filters, no mip chain) displacing y by 20 * height, pixel shader = colour
    ramp over the height; camera (0, 32, -7), the terrain scrolls with the frame.  The sample's plasma terrain (r32f) is replaced
    by a seeded smooth height field in an rg32f texture (.x = height in [0, 1])."""

    def __init__(self, w=640, h=360, samples=1, block=32, tex_size=64, seed=77, vs_binding=None):
        self.w, self.h, self.samples, self.block = w, h, samples, block
        self.vs_binding = vs_binding   # optional override: (wvp, offset, scale, sampler) -> ShaderBinding (SASL vertex shader)
        self.plane = create_planar((-block / 2.0, 0.0, -block / 2.0), (0.5, 0, 0), (0, 0, 0.5), block * 2, block * 2, False)
        self.plane.elements = [(0, _V4, 0, 0, 1.0), (1, _V4, 2, 0, 0.0)]   # POSITION -> reg 0, TEXCOORD (uv stream) -> reg 1
        rng = np.random.default_rng(seed)
        f = rng.uniform(0, 1, size=(tex_size, tex_size))
        for _ in range(3):  # smooth, periodic
            f = (f + np.roll(f, 1, 0) + np.roll(f, -1, 0) + np.roll(f, 1, 1) + np.roll(f, -1, 1)) / 5.0
        f = (f - f.min()) / (f.max() - f.min())
        self.height = np.stack([f, np.zeros_like(f)], -1).astype(f32)
        self.n_frames = 5

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_BGRA8)
        self.plane.upload(be)
        self.tex = be.create_texture(self.height.shape[1], self.height.shape[0], 1, A.PF_RG32F)
        be.upload_texture(self.tex, self.height)
        self.samp = be.create_sampler(A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, A.FILTER_LINEAR, addr_u=A.ADDR_MIRROR,
                                                     addr_v=A.ADDR_MIRROR), self.tex)

    def frame_uniforms(self, frame):
        view = mat_lookat((0.0, 32.0, -7.0), (0, 0, 0), (0, 1, 0))
        proj = mat_perspective_fov(math.pi / 2, f32(self.w) / f32(self.h), 0.1, 1000.0)
        wvp = mat_mul(view, proj)
        scene_sec = float(f32(frame * 161.2) / f32(self.n_frames - 1))
        scale = 1.0 / 32
        return wvp, (0.006 * scene_sec, 0.0088 * scene_sec), (scale, scale)

    def render(self, be: A.Backend, frame: int):
        t = self.t
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        wvp, off, scale = self.frame_uniforms(frame)
        d = base_desc(t, self.w, self.h, cull=A.CULL_NONE)
        self.plane.fill_desc(be, d)
        if self.vs_binding is not None:
            d.vs = self.vs_binding(wvp, off, scale, self.samp)
        else:
            d.vs = A.shader_binding(A.VS_TERRAIN_VTF, np.asarray(wvp, f32).tobytes() + struct.pack("<4f", *off, *scale), [self.samp])
        d.ps = A.shader_binding(A.PS_HEIGHT_COLOR)
        d.bs = A.shader_binding(A.BS_REPLACE)
        be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


# ===========================================================================================================
# C5: two-pass height field (the mesh and pass structure of BASELINE configs[4]: StandardShadowMap + 10 M triangles)
# ===========================================================================================================
class HeightFieldTwoPass:
    """a seeded (10000019) jittered height field of nx x nz quads (shared vertices, u32 indices; xz jitter and
    height uniform in +-0.2 cell) — 2500 x 2000 quads = 10,000,000 triangles at full size — drawn in two passes like
    samples/StandardShadowMap/StandardShadowMap.cpp:212-260: pass 1 depth-only from the light into an rg32f texture of the
    screen size with NO colour target (the shadow map), pass 2 from the camera with colour + depth.  The binning / raster
    stress of the config.  Pass 2 is lit with the ColorizedTriangle shaders: the sample's exponential-shadow pixel shader
    (exp / log / pow on a 9-tap tex2dlod) is not part of this scene.  FrameResult.count carries the shadow map's depth."""

    def __init__(self, w=7680, h=4320, samples=1, nx=2500, nz=2000, seed=10000019, shadowed=False):
        self.w, self.h, self.samples, self.nx, self.nz = w, h, samples, nx, nz
        # shadowed: pass 2 runs the StandardShadowMap sample's colour-pass shaders (SLV_VS_SSM_DRAW / SLV_PS_SSM_DRAW) and
        # reads pass 1's depth through the sample's point / border sampler; the shadow map is then single-sample
        self.shadowed = shadowed
        rng = np.random.default_rng(seed)
        gx, gz = np.meshgrid(np.arange(nx + 1, dtype=np.float64), np.arange(nz + 1, dtype=np.float64))
        jx = rng.uniform(-0.2, 0.2, size=gx.shape)
        jz = rng.uniform(-0.2, 0.2, size=gx.shape)
        hy = rng.uniform(-0.2, 0.2, size=gx.shape)
        cell = 20.0 / max(nx, nz)
        px = ((gx + jx) - nx / 2.0) * cell
        pz = ((gz + jz) - nz / 2.0) * cell
        py = hy * cell
        pos = np.stack([px, py, pz, np.ones_like(px)], -1).reshape(-1, 4).astype(f32)
        # normals from the un-jittered height differences (any fixed normal field does: it is an input, not a result)
        ny_ = np.ones_like(px)
        nxv = np.gradient(py, axis=1) / cell
        nzv = np.gradient(py, axis=0) / cell
        nrm = np.stack([-nxv, ny_, -nzv, np.zeros_like(px)], -1).reshape(-1, 4).astype(f32)
        i0 = (np.arange(nz)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1).astype(np.uint32)
        tri = np.stack([i0, i0 + (nx + 1), i0 + 1, i0 + 1, i0 + (nx + 1), i0 + (nx + 2)], -1).reshape(-1)
        self.mesh = Mesh([pos, nrm], [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0)], tri.astype(np.uint32), 2 * nx * nz)
        self.n_frames = 3

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_BGRA8)
        self.shadow = be.create_texture(self.w, self.h, 1 if self.shadowed else self.samples, A.PF_RG32F)
        if self.shadowed:
            self.sm_samp = be.create_sampler(A.sampler_desc(A.FILTER_POINT, A.FILTER_POINT, A.FILTER_POINT, addr_u=A.ADDR_BORDER,
                                                            addr_v=A.ADDR_BORDER, border=(1.0, 0.0, 0.0, 0.0)), self.shadow)
        self.mesh.upload(be)

    def frame_uniforms(self, frame):
        ang = 0.3 + 0.45 * frame
        camera = (math.cos(ang) * 6.0, 4.5, math.sin(ang) * 6.0)
        aspect = f32(self.w) / f32(self.h)
        cam = mat_mul(mat_lookat(camera, (0, 0, 0), (0, 1, 0)), mat_perspective_fov(math.pi / 3, aspect, 0.1, 100.0))
        light_pos = (3.0, 9.0, -2.0)
        light = mat_mul(mat_lookat(light_pos, (0, 0, 0), (0, 0, 1)), mat_perspective_fov(math.pi / 2, aspect, 0.1, 100.0))
        lights = [(light_pos[0], light_pos[1], light_pos[2], 0.0), (-4.0, 3.0, 4.0, 0.0), (0.0, 2.0, -6.0, 0.0)]
        return cam, light, lights

    def render(self, be: A.Backend, frame: int):
        t = self.t
        cam, light, lights = self.frame_uniforms(frame)
        # ---- pass 1: depth only, from the light, into the shadow map (no colour target)
        be.clear_depth_stencil(self.shadow, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        d = base_desc(t, self.w, self.h, cull=A.CULL_NONE)
        d.n_color_targets = 0
        d.ds_target = self.shadow.handle
        self.mesh.fill_desc(be, d)
        d.vs = A.shader_binding(A.VS_LIGHTS3, pack_vs_lights3(light, lights))
        d.ps = A.shader_binding(A.PS_LIGHTS3)
        d.bs = A.shader_binding(A.BS_REPLACE)
        be.draw(d)
        # ---- pass 2: colour + depth from the camera
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
        self.mesh.fill_desc(be, d)
        if self.shadowed:
            ang = 0.3 + 0.45 * frame
            d.vs = A.shader_binding(A.VS_SSM_DRAW, pack_vs_ssm_draw(cam, light, (*lights[0][:3], 1.0),
                                                                    (math.cos(ang) * 6.0, 4.5, math.sin(ang) * 6.0, 1.0)))
            d.ps = A.shader_binding(A.PS_SSM_DRAW, pack_ps_ssm_draw((0.1, 0.1, 0.1, 0.1), (0.8, 0.8, 0.8, 0.1), (0.4, 0.4, 0.4, 0.1), 32,
                                                                    False, True), [0, self.sm_samp])
        else:
            d.vs = A.shader_binding(A.VS_LIGHTS3, pack_vs_lights3(cam, lights))
            d.ps = A.shader_binding(A.PS_LIGHTS3)
        d.bs = A.shader_binding(A.BS_REPLACE)
        be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        res = read_frame(be, self.t, stats)
        sm = be.read_texture(self.shadow)
        res.count = sm.view(np.float32).reshape(sm.shape[0], sm.shape[1], sm.shape[2], 2)[..., 0].copy()
        return res


# ===========================================================================================================
# StandardShadowMap (BASELINE configs[4], the sample itself): shadow pass from the light + colour pass that SAMPLES the map
# ===========================================================================================================
class StandardShadowMap:
    """samples/StandardShadowMap/StandardShadowMap.cpp:160-330.  Pass 1 (gen_sm): depth only from the light into an rg32f
    texture of the screen size, no colour target, cull back.  Pass 2 (draw): Draw.savs + draw_cpp_ps — diffuse texture,
    Phong terms and an exponential shadow map read through a SECOND sampler (point filter, border colour (1, 0, 0, 0)) with
    nine tex2dlod taps.  The slanted ground plane, camera, light orbit and material constants are the sample's; cup.obj (an
    asset that does not travel) is replaced by a bulged cylinder and a sphere from this file's generators, textured with the
    seeded brick texture through a linear / clamp sampler like the sample's materials."""

    def __init__(self, w=640, h=360, samples=1, tex_size=128, detail=1, textured_plane=False, ps_binding=None):
        self.w, self.h, self.samples = w, h, samples
        # optional override (tests of run-time compiled SASL shaders): (ambient, diffuse, specular, shininess, tex_sampler,
        # shadow_sampler) -> ShaderBinding; textured_plane binds the brick texture to the ground plane as well
        self.textured_plane, self.ps_binding = textured_plane, ps_binding
        self.plane = create_planar((-3.0, 0.0, -3.0), (6.0, -1.0, 0.0), (0.0, -1.0, 6.0), 1, 1, False)   # :199-205
        cyl_vb, cyl_ib = _cylinder((0.2, -1.0, 0.1), 0.7, 1.7, 24 * detail, 6 * detail, uvscale=(2, 1), bulge=0.25)
        sph_vb, sph_ib = _sphere((1.5, -0.55, -1.3), 0.55, 20 * detail, 10 * detail)
        # interleaved 48-byte vertices (pos, uv, normal): POSITION -> reg 0, NORMAL -> reg 1, TEXCOORD0 -> reg 2 (Draw.savs VSIn)
        el = [(0, _V4, 0, 0, 1.0), (1, _V4, 0, 32, 0.0), (2, _V4, 0, 16, 0.0)]
        self.objects = [Mesh([cyl_vb], el, cyl_ib.reshape(-1), len(cyl_ib)), Mesh([sph_vb], el, sph_ib.reshape(-1), len(sph_ib))]
        self.brick = brick_texture(tex_size)
        self.n_frames = 8                                                                                   # TEST_FRAME_COUNT

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_BGRA8)
        self.shadow = be.create_texture(self.w, self.h, 1, A.PF_RG32F)                                      # :174-175
        self.sm_samp = be.create_sampler(A.sampler_desc(A.FILTER_POINT, A.FILTER_POINT, A.FILTER_POINT, addr_u=A.ADDR_BORDER,
                                                        addr_v=A.ADDR_BORDER, border=(1.0, 0.0, 0.0, 0.0)), self.shadow)  # :177-185
        self.tex = make_texture(be, self.brick)
        self.tex_samp = be.create_sampler(A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, A.FILTER_LINEAR, addr_u=A.ADDR_CLAMP,
                                                         addr_v=A.ADDR_CLAMP), self.tex)                    # :272-278
        self.plane.upload(be)
        for m in self.objects:
            m.upload(be)

    def frame_uniforms(self, frame):
        aspect = f32(self.w) / f32(self.h)
        camera_pos = (6.0, 3.1, 3.0, 1.0)                                                                   # :305-310
        cam = mat_mul(mat_lookat(camera_pos[:3], (0.0, 0.6, 0.0), (0, 1, 0)), mat_perspective_fov(math.pi / 4, aspect, 0.1, 100.0))
        scene_sec = float(f32(frame * 3.3) / f32(self.n_frames - 1))                                        # :317-319
        theta = 0.3 * scene_sec
        light_pos = (-4.0 * math.sin(theta), 6.1, 3.5 * math.cos(theta), 1.0)                               # :323-327
        light = mat_mul(mat_lookat(light_pos[:3], (0.0, 0.6, 0.0), (0, 1, 0)), mat_perspective_fov(math.pi / 4, aspect, 0.1, 40.0))
        return cam, light, light_pos, camera_pos

    def render(self, be: A.Backend, frame: int):
        t = self.t
        cam, light, light_pos, camera_pos = self.frame_uniforms(frame)
        # ---- gen_sm (:212-239): every mesh from the light, depth only
        be.clear_depth_stencil(self.shadow, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        for m in [self.plane] + self.objects:
            d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
            d.n_color_targets = 0
            d.ds_target = self.shadow.handle
            m.fill_desc(be, d)
            d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, pack_vs_mvp_passthrough(light, [0]))
            d.ps = A.shader_binding(A.PS_ATTR0_COLOR)
            d.bs = A.shader_binding(A.BS_REPLACE)
            be.draw(d)
        # ---- draw (:241-300)
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        vs = A.shader_binding(A.VS_SSM_DRAW, pack_vs_ssm_draw(cam, light, light_pos, camera_pos))
        mats = [((0.1, 0.1, 0.1, 0.1), (0.8, 0.8, 0.8, 0.1), (0.4, 0.4, 0.4, 0.1), 32, self.textured_plane),   # the plane (:262-270)
                ((0.18, 0.14, 0.12, 1.0), (0.9, 0.8, 0.7, 1.0), (0.5, 0.5, 0.5, 1.0), 16, True),
                ((0.10, 0.14, 0.20, 1.0), (0.6, 0.8, 0.95, 1.0), (0.9, 0.9, 0.9, 1.0), 48, True)]
        for m, (amb, dif, spe, shin, textured) in zip([self.plane] + self.objects, mats):
            d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
            m.fill_desc(be, d)
            d.vs = vs
            if self.ps_binding is not None:
                d.ps = self.ps_binding(amb, dif, spe, shin, self.tex_samp, self.sm_samp)
            else:
                d.ps = A.shader_binding(A.PS_SSM_DRAW, pack_ps_ssm_draw(amb, dif, spe, shin, textured, True),
                                        [self.tex_samp if textured else 0, self.sm_samp])
            d.bs = A.shader_binding(A.BS_REPLACE)
            be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        res = read_frame(be, self.t, stats)
        sm = be.read_texture(self.shadow)
        res.count = sm.view(np.float32).reshape(sm.shape[0], sm.shape[1], sm.shape[2], 2)[..., 0].copy()
        return res


# ===========================================================================================================
# triangle soup: seeded random triangles that cross the near/far planes, every cull/depth/stencil variant
# ===========================================================================================================
class TriangleSoup:
    """Parity torture scene (no reference twin): `n` random triangles in clip space, a good share of which
    straddle z=0 / z=w, drawn with PS_ATTR0_COLOR.  Exercises clipper.cpp:103-228, cull modes, depth
    functions, the stencil path and the coverage counter."""

    def __init__(self, w=256, h=192, samples=1, n=300, seed=7, cull=A.CULL_NONE, ds=None, stencil_ref=0,
                 bs=A.BS_REPLACE_AND_COUNT, index_dtype=np.uint16, modifiers=None, strip=False, size=1.0,
                 color_fmt=A.PF_RGBA8, ps=A.PS_ATTR0_COLOR, base_vertex=0, indexed=True, split=1, viewport=None, front_ccw=False,
                 stream_pad=(0, 0), color_element=None):
        """`base_vertex`: the index buffer holds (index - base_vertex) and draw_index adds it back (index_fetcher.cpp:26-115;
        negative values wrap through uint32 exactly as upstream).  `indexed=False`: renderer::draw - the vertex buffers are
        expanded in index order and drawn without an index buffer.  `split`: the primitives are drawn in that many draws with
        start != 0.  Upstream quirk, mirrored by every backend and pinned by the non-indexed cases: renderer::draw's `startpos`
        only offsets the INDEX buffer (index_fetcher.cpp:23,85), so a non-indexed draw ignores it and always begins at vertex 0
        - a split non-indexed soup draws its first range `split` times."""
        self.base_vertex, self.indexed, self.split = base_vertex, indexed, split
        self.viewport = viewport  # (x, y, w, h, minz, maxz) instead of the whole target with depth range 0..1 (viewport.h:5-12)
        # front_ccw: raster_desc::front_ccw (raster_state.cpp:10-31).  stream_pad: vertices of padding in front of each vertex
        # stream, skipped through the stream's byte OFFSET (set_vertex_buffers' offsets, stream_assembler.cpp:88-93)
        self.front_ccw, self.stream_pad = front_ccw, stream_pad
        # color_element: (format, default_w) of the colour stream's input element instead of four floats - get_vec4 reads 1 / 2 / 3
        # floats and fills in 0 and the element's default w (stream_assembler.cpp:26-45)
        self.color_element = color_element
        self.w, self.h, self.samples, self.n = w, h, samples, n
        self.cull, self.ds, self.stencil_ref, self.bs, self.modifiers = cull, ds, stencil_ref, bs, modifiers
        self.color_fmt, self.ps = color_fmt, ps
        rng = np.random.default_rng(seed)
        nv = n + 2 if strip else n * 3
        c = rng.uniform(-1.2, 1.2, size=(nv, 2)).astype(f32)
        if strip:
            pos_xy = c + rng.uniform(-0.1, 0.1, size=(nv, 2)).astype(f32)
        else:
            ctr = np.repeat(rng.uniform(-1.1, 1.1, size=(n, 2)), 3, axis=0)
            pos_xy = (ctr + rng.normal(0, 0.25 * size, size=(nv, 2))).astype(f32)
        wv = rng.uniform(0.2, 3.0, size=(nv, 1)).astype(f32)
        z = (rng.uniform(-0.3, 1.3, size=(nv, 1)) * wv).astype(f32)
        # a few exactly-degenerate and w<=0 vertices
        neg = rng.random(nv) < 0.05
        wv[neg] *= -1
        pos = np.concatenate([pos_xy * wv, z, wv], axis=1).astype(f32)
        col = rng.uniform(0, 1, size=(nv, 4)).astype(f32)
        if strip:
            idx = np.arange(nv, dtype=index_dtype)
            self.mesh = Mesh([pos, col], [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0)], idx, n, A.TOPO_TRIANGLE_STRIP)
        else:
            idx = rng.permutation(nv).astype(index_dtype)
            self.mesh = Mesh([pos, col], [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0)], idx, n)
        if not indexed:  # non-indexed draw(): vertex k of the stream IS vertex k of the primitive stream
            order = self.mesh.indices.astype(np.int64)
            self.mesh.streams = [np.ascontiguousarray(st[order]) for st in self.mesh.streams]
            self.mesh.indices = None
        elif base_vertex:
            # shift the vertex data by base_vertex slots (padding in front when positive) and keep the stored indices: the
            # fetch adds base_vertex back.  A negative base needs indices >= -base_vertex, so they are stored shifted up.
            if base_vertex > 0:
                pad = [np.zeros((base_vertex, st.shape[1]), f32) for st in self.mesh.streams]
                self.mesh.streams = [np.concatenate([p_, st]) for p_, st in zip(pad, self.mesh.streams)]
            else:
                self.mesh.indices = (self.mesh.indices.astype(np.int64) - base_vertex).astype(index_dtype)
        if color_element is not None:
            reg, _, slot, off, _ = self.mesh.elements[1]
            self.mesh.elements[1] = (reg, color_element[0], slot, off, color_element[1])
        if any(stream_pad):
            rng_pad = np.random.default_rng(seed + 99)
            self.mesh.streams = [np.concatenate([rng_pad.uniform(-5, 5, size=(k, st.shape[1])).astype(f32), st]) if k else st
                                 for k, st in zip(stream_pad, self.mesh.streams)]
        self.n_frames = 1

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, self.color_fmt,
                                with_count=self.bs == A.BS_REPLACE_AND_COUNT)
        self.mesh.upload(be)

    def render(self, be: A.Backend, frame: int = 0):
        t = self.t
        be.clear_color(t.color, (0.1, 0.2, 0.3, 1.0))
        if t.count is not None:
            be.clear_color(t.count, (0, 0, 0, 0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 0.6, 3)
        per = (self.mesh.prim_count + self.split - 1) // self.split
        strip = self.mesh.topology == A.TOPO_TRIANGLE_STRIP
        for first in range(0, max(self.mesh.prim_count, 1), max(per, 1)):
            d = base_desc(t, self.w, self.h, cull=self.cull, ds=self.ds)
            d.stencil_ref = self.stencil_ref
            if self.viewport is not None:
                d.viewport.x, d.viewport.y, d.viewport.w, d.viewport.h, d.viewport.minz, d.viewport.maxz = self.viewport
            d.raster.front_ccw = 1 if self.front_ccw else 0
            # start = first index (draw_index) / first vertex (draw) of the range: 3 per list primitive, 1 per strip primitive
            # (an odd strip start would flip the winding parity, so strips are only split at even primitives)
            self.mesh.fill_desc(be, d, start=first if strip else first * 3, prim_count=min(per, self.mesh.prim_count - first),
                                base_vertex=self.base_vertex)
            for i, k in enumerate(self.stream_pad):
                d.streams[i].offset = k * d.streams[i].stride
            if self.modifiers:
                for i, m in enumerate(self.modifiers):
                    d.vs_attr_modifiers[i] = m
            d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, pack_vs_mvp_passthrough(mat_identity(), [1]))
            d.ps = A.shader_binding(self.ps)
            d.bs = A.shader_binding(self.bs)
            be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int = 0) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


class OverlayQuads:
    """Robustness scene (no reference twin): `n` full-target quads (2 triangles each), front to back and back to front
    interleaved, each with its own colour: every triangle trivially accepts nearly every 64x64 tile it touches, the worst case
    for the per-tile / per-region work lists (a fully covered tile costs a list entry in each of its 16 regions)."""

    def __init__(self, w=3840, h=2160, samples=1, n=100, seed=3, draws=4):
        self.w, self.h, self.samples, self.n, self.draws = w, h, samples, n, draws
        rng = np.random.default_rng(seed)
        pos, col, idx = [], [], []
        for k in range(n):
            z = f32(0.05 + 0.9 * (((k * 37) % n) / n))      # a permutation of depths: some quads win, most lose early-Z
            x0, x1 = (-1.0, 1.0) if k % 3 else (-1.0 + 0.3 * rng.random(), 1.0 - 0.3 * rng.random())
            c = rng.uniform(0, 1, size=4).astype(f32)
            b = len(pos)
            pos += [(x0, -1.0, z, 1.0), (x1, -1.0, z, 1.0), (x1, 1.0, z, 1.0), (x0, 1.0, z, 1.0)]
            col += [c, c * f32(0.5), c, c * f32(0.25)]
            idx += [b, b + 1, b + 2, b + 2, b + 3, b]
        self.mesh = Mesh([np.array(pos, f32), np.array(col, f32)], [(0, _V4, 0, 0, 1.0), (1, _V4, 1, 0, 0.0)],
                         np.array(idx, np.uint32), 2 * n)
        self.n_frames = 1

    def setup(self, be: A.Backend):
        self.t = create_targets(be, self.w, self.h, self.samples, A.PF_RGBA8)
        self.mesh.upload(be)

    def render(self, be: A.Backend, frame: int = 0):
        t = self.t
        be.clear_color(t.color, (0.1, 0.2, 0.3, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        per = (self.mesh.prim_count + self.draws - 1) // self.draws
        for first in range(0, self.mesh.prim_count, per):
            d = base_desc(t, self.w, self.h, cull=A.CULL_NONE)
            self.mesh.fill_desc(be, d, start=first * 3, prim_count=min(per, self.mesh.prim_count - first))
            d.vs = A.shader_binding(A.VS_MVP_PASSTHROUGH, pack_vs_mvp_passthrough(mat_identity(), [1]))
            d.ps = A.shader_binding(A.PS_ATTR0_COLOR)
            d.bs = A.shader_binding(A.BS_REPLACE)
            be.draw(d)
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int = 0) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)


# ===========================================================================================================
# C4: Sponza-like atrium (procedural stand-in for resources/sponza_lq/sponza.obj, a Git-LFS pointer upstream)
# ===========================================================================================================
def _grid(fn, nu, nv, flip=False):
    """Samples a parametric surface fn(u, v) -> (pos[...,3], normal[...,3], uv[...,2]) on an (nu x nv) quad grid.
    Returns interleaved 48-byte vertices (pos.xyz1, uv00, normal.xyz0 — mesh_io_obj.cpp:398-431 layout) and
    a (2*nu*nv, 3) index array."""
    u, v = np.meshgrid(np.linspace(0, 1, nu + 1, dtype=np.float64), np.linspace(0, 1, nv + 1, dtype=np.float64), indexing="ij")
    pos, nrm, uv = fn(u, v)
    n = (nu + 1) * (nv + 1)
    vb = np.zeros((n, 12), dtype=f32)
    vb[:, 0:3] = pos.reshape(n, 3)
    vb[:, 3] = 1.0
    vb[:, 4:6] = uv.reshape(n, 2)
    nn = nrm.reshape(n, 3)
    nn = nn / np.maximum(np.linalg.norm(nn, axis=1, keepdims=True), 1e-12)
    vb[:, 8:11] = nn
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    q0 = (i * (nv + 1) + j).reshape(-1)
    q1, q2, q3 = q0 + 1, q0 + nv + 2, q0 + nv + 1
    tri = np.stack([np.stack([q0, q1, q2], 1), np.stack([q2, q3, q0], 1)], 1).reshape(-1, 3)
    if flip:
        tri = tri[:, ::-1]
    return vb, tri.astype(np.uint32)


def _plane(origin, du, dv, uvscale, nu, nv, flip=False, bump=0.0, seed=0):
    origin, du, dv = (np.asarray(a, dtype=np.float64) for a in (origin, du, dv))
    # with index order (q0, q0+v, q0+u+v) the front (un-culled) side is the one dv x du points to
    nrm0 = np.cross(du, dv) if flip else np.cross(dv, du)
    nrm0 = nrm0 / np.linalg.norm(nrm0)

    def fn(u, v):
        pos = origin + u[..., None] * du + v[..., None] * dv
        nrm = np.broadcast_to(nrm0, pos.shape).copy()
        if bump:
            h = bump * (np.sin(u * 37.0 + seed) * np.cos(v * 29.0 + seed * 0.7) + 0.5 * np.sin(u * 91.0) * np.sin(v * 83.0))
            pos = pos + h[..., None] * nrm0
            nrm = nrm + 0.3 * np.stack([np.cos(u * 37.0 + seed), np.sin(v * 29.0), np.zeros_like(u)], -1)
        uv = np.stack([u * uvscale[0], v * uvscale[1]], -1)
        return pos, nrm, uv

    return _grid(fn, nu, nv, flip)


def _cylinder(base, radius, height, nu, nv, uvscale=(4, 4), bulge=0.0):
    base = np.asarray(base, dtype=np.float64)

    def fn(u, v):
        a = u * 2 * np.pi
        r = radius * (1.0 + bulge * np.sin(v * np.pi))
        pos = base + np.stack([r * np.cos(a), v * height, r * np.sin(a)], -1)
        nrm = np.stack([np.cos(a), np.zeros_like(a), np.sin(a)], -1)
        return pos, nrm, np.stack([u * uvscale[0], v * uvscale[1]], -1)

    return _grid(fn, nu, nv, flip=True)


def _sphere(center, radius, nu, nv, squash=1.0):
    center = np.asarray(center, dtype=np.float64)

    def fn(u, v):
        a, b = u * 2 * np.pi, (v - 0.5) * np.pi
        d = np.stack([np.cos(b) * np.cos(a), squash * np.sin(b), np.cos(b) * np.sin(a)], -1)
        return center + radius * d, d, np.stack([u * 2, v * 2], -1)

    return _grid(fn, nu, nv, flip=True)


def _arch(center, span, rise, depth, nu, nv):
    """Half-ring (arch) in the xy plane extruded along z, between two columns."""
    center = np.asarray(center, dtype=np.float64)

    def fn(u, v):
        a = u * np.pi
        ring = np.stack([0.5 * span * np.cos(a), rise * np.sin(a), np.zeros_like(a)], -1)
        b = v * 2 * np.pi
        tube = 0.6 * np.stack([np.cos(a) * np.cos(b), np.sin(a) * np.cos(b), (depth / 0.6) * 0.5 * np.sin(b)], -1)
        pos = center + ring + tube
        return pos, tube, np.stack([u * 6, v * 2], -1)

    return _grid(fn, nu, nv, flip=True)


def _drape(origin, width, height, nu, nv, seed, flip=False):
    origin = np.asarray(origin, dtype=np.float64)
    sgn = -1.0 if flip else 1.0

    def fn(u, v):
        wave = 0.6 * np.sin(u * 14.0 + seed) * (0.3 + v) + 0.25 * np.sin(v * 9.0 + 2 * seed)
        pos = origin + np.stack([u * width, -v * height, wave], -1)
        nrm = sgn * np.stack([-0.6 * 14.0 / width * np.cos(u * 14.0 + seed) * (0.3 + v), np.zeros_like(u), np.ones_like(u)], -1)
        return pos, nrm, np.stack([u * 3, v * 3], -1)

    return _grid(fn, nu, nv, flip=flip)


class SponzaLike:
    """Deterministic 262,249-triangle atrium in 24 material groups, one shared 48-byte-stride vertex buffer and
    one u32 index buffer, one draw per material, trilinear wrap samplers, PS = diffuse x clamp(N.L)
    (samples/Sponza/Sponza.cpp:143-278; counters to match: benchmark.db.txt ia_primitives 262,249)."""

    N_TRIS = 262249
    N_MATERIALS = 24

    def __init__(self, w=3840, h=2160, samples=4, tex_size=1024, max_aniso=0, color_fmt=A.PF_BGRA8, ps_program=A.PS_SPONZA,
                 textured=True, sasl_derivatives=True):
        self.w, self.h, self.samples, self.tex_size, self.max_aniso = w, h, samples, tex_size, max_aniso
        self.sasl_derivatives = sasl_derivatives  # PS_SPONZA_GRAD: the derivative convention of the tex2D fetch
        self.color_fmt, self.ps_program, self.textured = color_fmt, ps_program, textured
        self.n_frames = 8
        self._draw_cache = {}
        # optional override (tests of run-time compiled SASL shaders): (wvp, light, eye) -> ShaderBinding
        self.vs_binding = None
        self._build()

    def _build(self):
        parts = []  # (material, vb, tris)
        add = lambda m, g: parts.append((m, g[0], g[1]))  # noqa: E731
        X0, X1, Y1, Z = -45.0, 45.0, 32.0, 14.0
        add(0, _plane((X0, 0, -Z), (X1 - X0, 0, 0), (0, 0, 2 * Z), (12, 4), 128, 48, flip=False))          # floor
        add(1, _plane((X0, Y1, -Z), (X1 - X0, 0, 0), (0, 0, 2 * Z), (6, 2), 64, 24, flip=True))            # ceiling
        add(2, _plane((X0, 0, -Z), (X1 - X0, 0, 0), (0, Y1, 0), (10, 4), 128, 32, flip=True, bump=0.05))   # wall -z
        add(3, _plane((X0, 0, Z), (X1 - X0, 0, 0), (0, Y1, 0), (10, 4), 128, 32, flip=False, bump=0.05, seed=3))
        add(4, _plane((X0, 0, -Z), (0, 0, 2 * Z), (0, Y1, 0), (4, 4), 32, 32, flip=False))                 # end wall
        add(5, _plane((X1, 0, -Z), (0, 0, 2 * Z), (0, Y1, 0), (4, 4), 32, 32, flip=True))
        xs = np.linspace(-40, 40, 12)
        for k, x in enumerate(xs):
            for side in (-1, 1):
                add(6 + (k % 2), _cylinder((x, 0, side * 8.0), 0.9, 12.0, 32, 24, bulge=0.08))
                add(8, _sphere((x, 12.4, side * 8.0), 1.3, 24, 12, squash=0.5))
        for k in range(11):
            xc = 0.5 * (xs[k] + xs[k + 1])
            for side in (-1, 1):
                add(9 + (k % 2), _arch((xc, 12.0, side * 8.0), xs[1] - xs[0] - 1.8, 3.0, 1.2, 32, 16))
                add(11 + (k % 2), _arch((xc, 22.0, side * 8.0), xs[1] - xs[0] - 1.8, 2.5, 1.0, 32, 16))
        add(13, _plane((X0, 16.0, -Z), (X1 - X0, 0, 0), (0, 0, Z - 8.0), (12, 1), 128, 8, flip=False))     # balconies
        add(14, _plane((X0, 16.0, 8.0), (X1 - X0, 0, 0), (0, 0, Z - 8.0), (12, 1), 128, 8, flip=False))
        for k in range(8):
            add(15 + (k % 3), _drape((-38 + k * 10.0, 15.5, (-1) ** k * 7.0), 6.0, 9.0, 64, 64, seed=k, flip=(k % 2 == 0)))
        for k in range(16):
            add(18 + (k % 2), _sphere((-37.5 + k * 5.0, 1.2, (-1) ** k * 4.5), 1.2, 48, 24))
        add(20, _plane((-30, 4.0, -Z + 0.2), (20, 0, 0), (0, 12, 0), (2, 2), 64, 64, flip=True, bump=0.35, seed=5))
        add(21, _plane((10, 4.0, Z - 0.2), (20, 0, 0), (0, 12, 0), (2, 2), 64, 64, flip=False, bump=0.35, seed=6))
        add(22, _plane((-35, 0.02, -2.0), (70, 0, 0), (0, 0, 4.0), (20, 1), 139, 28, flip=False, bump=0.02, seed=8))  # carpet
        n = sum(p[2].shape[0] for p in parts)
        assert n == self.N_TRIS - 1, n
        vb1 = np.zeros((3, 12), dtype=f32)  # one lone triangle (banner) completes the published count
        vb1[:, 0:3] = [(-1, 26, 0), (1, 26, 0), (0, 28, 0)]
        vb1[:, 3] = 1
        vb1[:, 4:6] = [(0, 0), (1, 0), (0.5, 1)]
        vb1[:, 8:11] = (0, 0, -1)
        parts.append((23, vb1, np.array([[0, 1, 2]], dtype=np.uint32)))
        # concatenate by material
        parts.sort(key=lambda p: p[0])
        vbs, ibs, self.groups = [], [], []
        base = 0
        tri_cursor = 0
        cur_m, cur_start = None, 0
        for m, vb, tri in parts:
            if m != cur_m:
                if cur_m is not None:
                    self.groups.append((cur_m, cur_start, tri_cursor - cur_start))
                cur_m, cur_start = m, tri_cursor
            vbs.append(vb)
            ibs.append(tri + base)
            base += vb.shape[0]
            tri_cursor += tri.shape[0]
        self.groups.append((cur_m, cur_start, tri_cursor - cur_start))
        vb = np.concatenate(vbs).astype(f32)
        ib = np.concatenate(ibs).astype(np.uint32).reshape(-1)
        assert ib.size == self.N_TRIS * 3 and len(self.groups) == self.N_MATERIALS
        self.mesh = Mesh([vb], [(0, _V4, 0, 0, 1.0), (1, _V4, 0, 16, 0.0), (2, _V4, 0, 32, 0.0)], ib, self.N_TRIS)

    def setup(self, be: A.Backend):
        self._draw_cache = {}
        self.t = create_targets(be, self.w, self.h, self.samples, self.color_fmt)
        self.mesh.upload(be)
        self.textures, self.samplers = [], []
        mipf = A.FILTER_ANISOTROPIC if self.max_aniso > 1 else A.FILTER_LINEAR
        for m in range(self.N_MATERIALS):
            tex = make_texture(be, brick_texture(self.tex_size, seed=100 + m))
            self.textures.append(tex)
            self.samplers.append(be.create_sampler(
                A.sampler_desc(A.FILTER_LINEAR, A.FILTER_LINEAR, mipf, addr_u=A.ADDR_WRAP, addr_v=A.ADDR_WRAP,
                               max_anisotropy=self.max_aniso), tex))

    def frame_uniforms(self, frame):
        scene_sec = float(f32(frame * 18) / f32(self.n_frames - 1))
        xpos = -36.0 + math.fmod(3.0 * scene_sec, 66.0)
        camera = (xpos, 8.0, 0.0)
        view = mat_lookat(camera, (40.0, 15.0, 0.0), (0, 1, 0))
        proj = mat_perspective_fov(math.pi / 2, f32(self.w) / f32(self.h), 0.1, 1000.0)
        wvp = mat_mul(mat_translate(-0.5, 0, -0.5), mat_mul(view, proj))
        ypos = 10.0 + math.fmod(8.0 * scene_sec, 40.0)
        return wvp, (0.0, ypos, 0.0, 1.0), (*camera, 1.0)

    def frame_draws(self, be: A.Backend, frame: int):
        """The frame's draw descriptors, built once per (backend, frame) and reused: a host application holds its
        render state in long-lived objects too (the reference's render_state pool, async_renderer.cpp:42-74), so the
        per-frame host cost is the C ABI calls, not descriptor marshalling in Python."""
        key = (id(be), frame)
        cached = self._draw_cache.get(key)
        if cached is not None:
            return cached
        t = self.t
        wvp, light, eye = self.frame_uniforms(frame)
        vs = self.vs_binding(wvp, light, eye) if self.vs_binding else A.shader_binding(A.VS_SPONZA, pack_vs_sponza(wvp, light, eye))
        bs = A.shader_binding(A.BS_REPLACE)
        draws = []
        # the SASL tex2D path is required for anisotropic filtering; the cpp tex2d path
        # (LOD once per quad) is what samples/Sponza uses for trilinear
        for m, start, count in self.groups:
            d = base_desc(t, self.w, self.h, cull=A.CULL_BACK)
            self.mesh.fill_desc(be, d, start=start * 3, prim_count=count)
            d.vs = vs
            if self.ps_program == A.PS_SPONZA:
                d.ps = A.shader_binding(A.PS_SPONZA, pack_ps_sponza(self.textured), [self.samplers[m]])
            elif self.ps_program == A.PS_SPONZA_GRAD:
                d.ps = A.shader_binding(A.PS_SPONZA_GRAD, pack_ps_sponza_grad(self.textured, self.sasl_derivatives), [self.samplers[m]])
            else:
                d.ps = A.shader_binding(self.ps_program)
            d.bs = bs
            draws.append(d)
        self._draw_cache[key] = draws
        return draws

    def render(self, be: A.Backend, frame: int, before_resolve=None):
        """`before_resolve`: hook of the sort-first frame assembly (sortfirst.FrameGather.before_resolve)."""
        t = self.t
        be.clear_color(t.color, (0.2, 0.2, 0.5, 1.0))
        be.clear_depth_stencil(t.ds, A.CLEAR_DEPTH | A.CLEAR_STENCIL, 1.0, 0)
        for d in self.frame_draws(be, frame):
            be.draw(d)
        if before_resolve is not None:
            before_resolve()
        if t.resolved is not None:
            be.resolve(t.color, t.resolved)

    def run(self, be: A.Backend, frame: int) -> FrameResult:
        be.query_begin()
        self.render(be, frame)
        stats = be.query_get()
        return read_frame(be, self.t, stats)
