"""ctypes mirror of include/salvia_b200.h.

The same binding drives three shared libraries that export the identical slv_* table:
the CUDA product (`libsalvia_b200.so`), the CPU restatement (`oracle/libsalvia_oracle.so`) and the
unmodified reference behind the ABI (`oracle/_ref/libsalvia_ref.so`).  The last two are test
infrastructure; the package itself only ever loads the product (see `salviarenderer_b200.load`).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

# ---- constants (values of include/salvia_b200.h) --------------------------------------------------
OK, FAILED, OUT_OF_MEMORY, INVALID_PARAMETER = 0, 1, 2, 3
TOPO_TRIANGLE_LIST, TOPO_TRIANGLE_STRIP = 2, 4
CULL_NONE, CULL_FRONT, CULL_BACK = 0, 1, 2
ADDR_WRAP, ADDR_MIRROR, ADDR_CLAMP, ADDR_BORDER = 0, 1, 2, 3
FILTER_POINT, FILTER_LINEAR, FILTER_ANISOTROPIC = 0, 1, 2
MIP_LO, MIP_MI, MIP_HI = 0, 1, 2
CMP_NEVER, CMP_LESS, CMP_EQUAL, CMP_LESS_EQUAL, CMP_GREATER, CMP_NOT_EQUAL, CMP_GREATER_EQUAL, CMP_ALWAYS = range(8)
SOP_KEEP, SOP_ZERO, SOP_REPLACE, SOP_INCR_SAT, SOP_DECR_SAT, SOP_INVERT, SOP_INCR_WRAP, SOP_DECR_WRAP = range(1, 9)
CLEAR_DEPTH, CLEAR_STENCIL = 1, 2
INDEX_NONE, INDEX_R16_UINT, INDEX_R32_UINT = 0, 57, 42
FMT_R32_FLOAT, FMT_R32G32_FLOAT, FMT_R32G32B32_FLOAT, FMT_R32G32B32A32_FLOAT = 41, 16, 6, 2
PF_RGBA32F, PF_BGRA8, PF_RGBA8, PF_RG32F = 0, 2, 3, 5
PF_BYTES = {PF_RGBA32F: 16, PF_BGRA8: 4, PF_RGBA8: 4, PF_RG32F: 8}
AM_LINEAR, AM_CENTROID, AM_NOINTERPOLATION, AM_NOPERSPECTIVE = 1, 2, 4, 8

VS_MVP_PASSTHROUGH, VS_PLANE_XZ, VS_LIGHTS3, VS_SPONZA, VS_TERRAIN_VTF, VS_SSM_DRAW = 1, 2, 3, 4, 5, 6
PS_ATTR0_COLOR, PS_LIGHTS3, PS_TEX_ALPHA, PS_SPONZA, PS_TEX_GRAD_ALPHA, PS_DISCARD_ALL, PS_HEIGHT_COLOR, PS_SSM_DRAW = 1, 2, 3, 4, 5, 6, 7, 8
PS_SPONZA_GRAD = 9
BS_REPLACE, BS_LERP_SRC_ALPHA, BS_REPLACE_AND_COUNT = 1, 2, 3

MAX_VS_INPUT_ATTRS = 8
MAX_VS_OUTPUT_ATTRS = 5
MAX_RENDER_TARGETS = 8
MAX_SAMPLERS = 4
MAX_UNIFORM_BYTES = 256


# ---- POD descriptors ------------------------------------------------------------------------------
class SamplerDesc(C.Structure):
    _fields_ = [
        ("min_filter", C.c_uint32), ("mag_filter", C.c_uint32), ("mip_filter", C.c_uint32),
        ("mip_qual", C.c_uint32),
        ("addr_mode_u", C.c_uint32), ("addr_mode_v", C.c_uint32), ("addr_mode_w", C.c_uint32),
        ("mip_lod_bias", C.c_float), ("max_anisotropy", C.c_uint32), ("comparison_func", C.c_uint32),
        ("border_color", C.c_float * 4), ("min_lod", C.c_float), ("max_lod", C.c_float),
    ]


class Viewport(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("x", "y", "w", "h", "minz", "maxz")]


class StencilOpDesc(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("stencil_fail_op", "stencil_depth_fail_op", "stencil_pass_op", "stencil_func")]


class DepthStencilDesc(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("depth_enable", "depth_write_mask", "depth_func", "stencil_enable",
                                          "stencil_read_mask", "stencil_write_mask")] + [
        ("front_face", StencilOpDesc), ("back_face", StencilOpDesc)]


class RasterDesc(C.Structure):
    _fields_ = [("cull_mode", C.c_uint32), ("front_ccw", C.c_uint32)]


class VertexStream(C.Structure):
    _fields_ = [("buffer", C.c_uint32), ("stride", C.c_uint32), ("offset", C.c_uint32)]


class InputElement(C.Structure):
    _fields_ = [("reg", C.c_uint32), ("format", C.c_uint32), ("slot", C.c_uint32),
                ("aligned_byte_offset", C.c_uint32), ("default_w", C.c_float)]


class ShaderBinding(C.Structure):
    _fields_ = [("program", C.c_uint32), ("uniform_bytes", C.c_uint32),
                ("uniforms", C.c_uint8 * MAX_UNIFORM_BYTES), ("samplers", C.c_uint32 * MAX_SAMPLERS)]


class DrawDesc(C.Structure):
    _fields_ = [
        ("n_streams", C.c_uint32), ("streams", VertexStream * 8),
        ("n_elements", C.c_uint32), ("elements", InputElement * MAX_VS_INPUT_ATTRS),
        ("index_buffer", C.c_uint32), ("index_format", C.c_uint32), ("topology", C.c_uint32),
        ("start", C.c_uint32), ("prim_count", C.c_uint32), ("base_vertex", C.c_int32),
        ("vs", ShaderBinding), ("ps", ShaderBinding), ("bs", ShaderBinding),
        ("vs_attr_modifiers", C.c_uint32 * MAX_VS_OUTPUT_ATTRS),
        ("raster", RasterDesc), ("ds", DepthStencilDesc), ("stencil_ref", C.c_int32),
        ("viewport", Viewport),
        ("n_color_targets", C.c_uint32), ("color_targets", C.c_uint32 * MAX_RENDER_TARGETS),
        ("ds_target", C.c_uint32),
    ]


class PipelineStatistics(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("ia_vertices", "ia_primitives", "vs_invocations", "gs_invocations",
                                          "gs_primitives", "cinvocations", "cprimitives", "ps_invocations",
                                          "backend_input_pixels")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class PipelineProfiles(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("gather_vtx", "vtx_proc", "clipping", "compact_clip", "vp_trans",
                                          "tri_dispatch", "ras")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class TrafficCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("z_tested", "z_written", "c_written", "c_read", "list_entries_scanned",
                                          "region_survivors", "warp_pairs", "quads_shaded", "ps_executed")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# every symbol include/salvia_b200.h declares (the CPU test-suite checks each library exports them all)
ENTRY_POINTS = [
    "slv_device_create", "slv_device_destroy", "slv_backend_name", "slv_abi_version",
    "slv_buffer_create", "slv_buffer_upload", "slv_buffer_readback",
    "slv_texture_create", "slv_texture_gen_mipmap", "slv_texture_level_count", "slv_texture_level_size",
    "slv_texture_upload", "slv_texture_readback", "slv_sampler_create", "slv_resource_release",
    "slv_draw", "slv_clear_color", "slv_clear_depth_stencil", "slv_resolve", "slv_flush",
    "slv_query_begin", "slv_query_get", "slv_sampler_probe",
    "slv_set_tile_shard", "slv_profile_get",
    "slv_traffic_get", "slv_kernel_launch_count", "slv_event_record", "slv_event_elapsed_ms", "slv_profile_enable",
    "slv_texture_device_ptr", "slv_pack_tiles", "slv_unpack_tiles", "slv_set_stream", "slv_profile_get_stages",
    "slv_peer_export_texture", "slv_peer_export_flags", "slv_peer_open", "slv_peer_close", "slv_resolve_target_peer",
    "slv_peer_signal", "slv_flags_wait", "slv_shader_module_load", "slv_shader_compile_cubin", "slv_shader_compile", "slv_free", "slv_sasl_translate", "slv_texture_level_tracking", "slv_texture_levels_touched", "slv_texture_readback_async", "slv_readback_wait",
    "slv_readback_fence", "slv_host_register", "slv_host_unregister", "slv_texture_export_tiles_async",
    "slv_assembly_wait", "slv_peer_signal_after_consumers",
    "slv_buffer_device_ptr", "slv_external_write_begin", "slv_external_write_end",
]


class SlvError(RuntimeError):
    pass


def _chk(rc, what):
    if rc != OK:
        raise SlvError(f"{what} failed with slv_result={rc}")


def program_jit(module: int) -> int:
    """SLV_PROGRAM_JIT(module): the program id that selects a loaded run-time shader module."""
    return 0x80000000 | int(module)


def shader_binding(program: int, uniforms: bytes = b"", samplers=()) -> ShaderBinding:
    b = ShaderBinding()
    b.program = program
    if len(uniforms) > MAX_UNIFORM_BYTES:
        raise ValueError("uniform block too large")
    b.uniform_bytes = len(uniforms)
    C.memmove(b.uniforms, uniforms, len(uniforms))
    for i, s in enumerate(samplers):
        b.samplers[i] = int(s)
    return b


def depth_stencil_desc(depth_enable=True, depth_write=True, depth_func=CMP_LESS, stencil_enable=False,
                       read_mask=0xFF, write_mask=0xFF, front=None, back=None) -> DepthStencilDesc:
    d = DepthStencilDesc()
    d.depth_enable, d.depth_write_mask, d.depth_func = int(depth_enable), int(depth_write), depth_func
    d.stencil_enable, d.stencil_read_mask, d.stencil_write_mask = int(stencil_enable), read_mask, write_mask
    for name, v in (("front_face", front), ("back_face", back)):
        o = getattr(d, name)
        fail, dfail, pas, func = v if v is not None else (SOP_KEEP, SOP_KEEP, SOP_KEEP, CMP_ALWAYS)
        o.stencil_fail_op, o.stencil_depth_fail_op, o.stencil_pass_op, o.stencil_func = fail, dfail, pas, func
    return d


def sampler_desc(min_filter=FILTER_POINT, mag_filter=FILTER_POINT, mip_filter=FILTER_POINT, mip_qual=MIP_MI,
                 addr_u=ADDR_WRAP, addr_v=ADDR_WRAP, max_anisotropy=0, border=(0, 0, 0, 0)) -> SamplerDesc:
    s = SamplerDesc()
    s.min_filter, s.mag_filter, s.mip_filter, s.mip_qual = min_filter, mag_filter, mip_filter, mip_qual
    s.addr_mode_u, s.addr_mode_v, s.addr_mode_w = addr_u, addr_v, ADDR_WRAP
    s.mip_lod_bias, s.max_anisotropy, s.comparison_func = 0.0, max_anisotropy, CMP_ALWAYS
    for i in range(4):
        s.border_color[i] = border[i]
    s.min_lod, s.max_lod = -1e20, 1e20
    return s


@dataclass
class Texture:
    handle: int
    width: int
    height: int
    samples: int
    fmt: int


class Backend:
    """One loaded slv_* library + one device. Thin, explicit, no hidden fallbacks."""

    def __init__(self, lib_path: str, ordinal: int = 0):
        if not os.path.exists(lib_path):
            raise SlvError(f"shared library not found: {lib_path}")
        self.lib_path = lib_path
        self.lib = C.CDLL(lib_path)
        L = self.lib
        L.slv_backend_name.restype = C.c_char_p
        L.slv_abi_version.restype = C.c_uint32
        L.slv_device_destroy.restype = None
        L.slv_device_destroy.argtypes = [C.c_void_p]
        for n in ENTRY_POINTS:
            f = getattr(L, n)
            if n not in ("slv_backend_name", "slv_abi_version", "slv_device_destroy", "slv_free"):
                f.restype = C.c_int32
        L.slv_buffer_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
        L.slv_buffer_upload.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t]
        L.slv_buffer_readback.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t]
        L.slv_texture_create.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.slv_texture_gen_mipmap.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.slv_texture_level_count.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.slv_texture_level_size.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.slv_texture_upload.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        L.slv_texture_readback.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        L.slv_sampler_create.argtypes = [C.c_void_p, C.POINTER(SamplerDesc), C.c_uint32, C.POINTER(C.c_uint32)]
        L.slv_resource_release.argtypes = [C.c_void_p, C.c_uint32]
        L.slv_draw.argtypes = [C.c_void_p, C.POINTER(DrawDesc)]
        L.slv_clear_color.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float)]
        L.slv_clear_depth_stencil.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32]
        L.slv_resolve.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.slv_flush.argtypes = [C.c_void_p]
        L.slv_query_begin.argtypes = [C.c_void_p]
        L.slv_query_get.argtypes = [C.c_void_p, C.POINTER(PipelineStatistics)]
        L.slv_sampler_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_uint32, C.c_void_p]
        L.slv_set_tile_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.slv_profile_get.argtypes = [C.c_void_p, C.POINTER(PipelineProfiles)]
        L.slv_traffic_get.argtypes = [C.c_void_p, C.POINTER(TrafficCounters)]
        L.slv_kernel_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.slv_event_record.argtypes = [C.c_void_p, C.c_uint32]
        L.slv_event_elapsed_ms.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        L.slv_profile_enable.argtypes = [C.c_void_p, C.c_uint32]
        L.slv_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.slv_texture_device_ptr.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.slv_pack_tiles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_size_t)]
        L.slv_unpack_tiles.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.slv_profile_get_stages.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint32]
        L.slv_texture_readback_async.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        L.slv_readback_wait.argtypes = [C.c_void_p]
        L.slv_readback_fence.argtypes = [C.c_void_p, C.c_uint32]
        L.slv_assembly_wait.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.slv_peer_signal_after_consumers.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.slv_buffer_device_ptr.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.slv_external_write_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.slv_external_write_end.argtypes = [C.c_void_p, C.c_void_p]
        L.slv_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.slv_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        L.slv_texture_export_tiles_async.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]
        L.slv_shader_module_load.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
        L.slv_shader_compile_cubin.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                               C.c_char_p, C.c_size_t]
        L.slv_shader_compile.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]
        L.slv_free.argtypes = [C.c_void_p]
        L.slv_free.restype = None
        L.slv_peer_export_texture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.slv_peer_export_flags.argtypes = [C.c_void_p, C.c_void_p]
        L.slv_peer_open.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.slv_peer_close.argtypes = [C.c_void_p, C.c_void_p]
        L.slv_resolve_target_peer.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.slv_peer_signal.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.slv_flags_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        self.name = L.slv_backend_name().decode()
        if L.slv_abi_version() != 1:
            raise SlvError("ABI version mismatch")
        dev = C.c_void_p()
        _chk(L.slv_device_create(ordinal, C.byref(dev)), f"slv_device_create({self.name})")
        self.dev = dev
        self._tex: dict[int, Texture] = {}

    def close(self):
        if getattr(self, "dev", None):
            self.lib.slv_device_destroy(self.dev)
            self.dev = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- resources ------------------------------------------------------------------------------
    def create_buffer(self, data: np.ndarray | bytes) -> int:
        raw = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
        h = C.c_uint32()
        _chk(self.lib.slv_buffer_create(self.dev, len(raw), C.byref(h)), "slv_buffer_create")
        _chk(self.lib.slv_buffer_upload(self.dev, h.value, 0, raw, len(raw)), "slv_buffer_upload")
        return h.value

    def create_texture(self, width, height, samples, fmt) -> Texture:
        h = C.c_uint32()
        _chk(self.lib.slv_texture_create(self.dev, width, height, samples, fmt, C.byref(h)), "slv_texture_create")
        t = Texture(h.value, width, height, samples, fmt)
        self._tex[h.value] = t
        return t

    def upload_texture(self, tex: Texture, data: np.ndarray, level: int = 0):
        raw = np.ascontiguousarray(data).tobytes()
        _chk(self.lib.slv_texture_upload(self.dev, tex.handle, level, raw, len(raw)), "slv_texture_upload")

    def gen_mipmap(self, tex: Texture, filt=FILTER_LINEAR):
        _chk(self.lib.slv_texture_gen_mipmap(self.dev, tex.handle, filt), "slv_texture_gen_mipmap")

    def level_count(self, tex: Texture) -> int:
        n = C.c_uint32()
        _chk(self.lib.slv_texture_level_count(self.dev, tex.handle, C.byref(n)), "slv_texture_level_count")
        return n.value

    def level_size(self, tex: Texture, level: int):
        w, h = C.c_uint32(), C.c_uint32()
        _chk(self.lib.slv_texture_level_size(self.dev, tex.handle, level, C.byref(w), C.byref(h)), "level_size")
        return w.value, h.value

    def read_texture(self, tex: Texture, level: int = 0) -> np.ndarray:
        """Returns the raw level as uint8[h, w, samples, bpp] (reference layout, surface.cpp:277-295)."""
        w, h = self.level_size(tex, level)
        bpp = PF_BYTES[tex.fmt]
        out = np.empty((h, w, tex.samples, bpp), dtype=np.uint8)
        _chk(self.lib.slv_texture_readback(self.dev, tex.handle, level, out.ctypes.data, out.nbytes), "slv_texture_readback")
        return out

    def create_sampler(self, desc: SamplerDesc, tex: Texture) -> int:
        h = C.c_uint32()
        _chk(self.lib.slv_sampler_create(self.dev, C.byref(desc), tex.handle, C.byref(h)), "slv_sampler_create")
        return h.value

    def release(self, handle: int):
        _chk(self.lib.slv_resource_release(self.dev, handle), "slv_resource_release")

    # ---- commands -------------------------------------------------------------------------------
    def draw(self, desc: DrawDesc):
        _chk(self.lib.slv_draw(self.dev, C.byref(desc)), "slv_draw")

    def clear_color(self, tex: Texture, rgba):
        arr = (C.c_float * 4)(*rgba)
        _chk(self.lib.slv_clear_color(self.dev, tex.handle, arr), "slv_clear_color")

    def clear_depth_stencil(self, tex: Texture, flags, depth, stencil):
        _chk(self.lib.slv_clear_depth_stencil(self.dev, tex.handle, flags, depth, stencil), "slv_clear_depth_stencil")

    def resolve(self, src: Texture, dst: Texture):
        _chk(self.lib.slv_resolve(self.dev, src.handle, dst.handle), "slv_resolve")

    def flush(self):
        _chk(self.lib.slv_flush(self.dev), "slv_flush")

    def query_begin(self):
        _chk(self.lib.slv_query_begin(self.dev), "slv_query_begin")

    def query_get(self) -> dict:
        st = PipelineStatistics()
        _chk(self.lib.slv_query_get(self.dev, C.byref(st)), "slv_query_get")
        return st.as_dict()

    def sampler_probe(self, sampler: int, coords, ddx=None, ddy=None, lod=None) -> np.ndarray:
        coords = np.ascontiguousarray(coords, dtype=np.float32)
        n = coords.shape[0]
        out = np.empty((n, 4), dtype=np.float32)
        if lod is not None:
            lod = np.ascontiguousarray(lod, dtype=np.float32)
            z = np.zeros((n, 2), dtype=np.float32)
            rc = self.lib.slv_sampler_probe(self.dev, sampler, n, coords.ctypes.data, z.ctypes.data, z.ctypes.data,
                                            lod.ctypes.data, 1, out.ctypes.data)
        else:
            ddx = np.ascontiguousarray(ddx, dtype=np.float32)
            ddy = np.ascontiguousarray(ddy, dtype=np.float32)
            z = np.zeros((n,), dtype=np.float32)
            rc = self.lib.slv_sampler_probe(self.dev, sampler, n, coords.ctypes.data, ddx.ctypes.data, ddy.ctypes.data,
                                            z.ctypes.data, 0, out.ctypes.data)
        _chk(rc, "slv_sampler_probe")
        return out

    def set_tile_shard(self, rank: int, nranks: int):
        _chk(self.lib.slv_set_tile_shard(self.dev, rank, nranks), "slv_set_tile_shard")

    # ---- measurement / multi-GPU plumbing -------------------------------------------------------------
    def traffic(self) -> dict:
        t = TrafficCounters()
        _chk(self.lib.slv_traffic_get(self.dev, C.byref(t)), "slv_traffic_get")
        return t.as_dict()

    def texture_level_tracking(self, on: bool):
        """B_tex accounting: sampler calls of draws issued from now on record the mip levels they read (resets the masks)."""
        self.lib.slv_texture_level_tracking.argtypes = [C.c_void_p, C.c_uint32]
        _chk(self.lib.slv_texture_level_tracking(self.dev, 1 if on else 0), "slv_texture_level_tracking")

    def texture_levels_touched(self, tex: Texture) -> int:
        self.lib.slv_texture_levels_touched.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        m = C.c_uint32()
        _chk(self.lib.slv_texture_levels_touched(self.dev, tex.handle, C.byref(m)), "slv_texture_levels_touched")
        return m.value

    def launch_count(self) -> int:
        n = C.c_uint64()
        _chk(self.lib.slv_kernel_launch_count(self.dev, C.byref(n)), "slv_kernel_launch_count")
        return n.value

    def event_record(self, slot: int):
        _chk(self.lib.slv_event_record(self.dev, slot), "slv_event_record")

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_float()
        _chk(self.lib.slv_event_elapsed_ms(self.dev, a, b, C.byref(ms)), "slv_event_elapsed_ms")
        return ms.value

    def set_stream(self, cuda_stream: int | None):
        _chk(self.lib.slv_set_stream(self.dev, C.c_void_p(cuda_stream or 0)), "slv_set_stream")

    def profile_enable(self, on: bool):
        _chk(self.lib.slv_profile_enable(self.dev, int(on)), "slv_profile_enable")

    def profile_get(self) -> dict:
        p = PipelineProfiles()
        _chk(self.lib.slv_profile_get(self.dev, C.byref(p)), "slv_profile_get")
        return p.as_dict()

    def profile_stages(self) -> dict:
        """Event time per kernel stage of the product since query_begin, in milliseconds."""
        ms = (C.c_double * 6)()
        _chk(self.lib.slv_profile_get_stages(self.dev, ms, 6), "slv_profile_get_stages")
        return dict(zip(("geometry", "bin", "sort", "raster_or_cover", "shade", "region_bin"), (float(v) for v in ms)))

    def texture_ptr(self, tex: Texture, level: int = 0):
        p, n = C.c_void_p(), C.c_size_t()
        _chk(self.lib.slv_texture_device_ptr(self.dev, tex.handle, level, C.byref(p), C.byref(n)), "slv_texture_device_ptr")
        return p.value, n.value

    # ---- SASL shaders compiled at run time (product only; see salviarenderer_b200/sasl) ----
    def shader_module_load(self, stage: str, image: bytes, n_vs_output_attrs: int = 0) -> int:
        h = C.c_uint32()
        _chk(self.lib.slv_shader_module_load(self.dev, 0 if stage == "vs" else 1, image, len(image), n_vs_output_attrs, C.byref(h)),
             "slv_shader_module_load")
        return h.value

    def shader_compile(self, stage: str, device_code: str, n_vs_output_attrs: int = 0, deriv_cpp: bool = False) -> int:
        """slv_shader_compile: NVRTC in process + module load; returns the module handle."""
        h = C.c_uint32()
        log = C.create_string_buffer(16384)
        rc = self.lib.slv_shader_compile(self.dev, 0 if stage == "vs" else 1, device_code.encode(), n_vs_output_attrs, 1 if deriv_cpp else 0,
                                         C.byref(h), log, len(log))
        if rc != 0:
            raise SlvError(f"slv_shader_compile failed ({rc}):\n{log.value.decode(errors='replace')[-4000:]}")
        return h.value

    # ---- peer-memory frame assembly (CUDA IPC; product only) ----
    def peer_export_texture(self, tex: Texture, level: int = 0) -> bytes:
        buf = (C.c_uint8 * 64)()
        _chk(self.lib.slv_peer_export_texture(self.dev, tex.handle, level, buf), "slv_peer_export_texture")
        return bytes(buf)

    def peer_export_flags(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        _chk(self.lib.slv_peer_export_flags(self.dev, buf), "slv_peer_export_flags")
        return bytes(buf)

    def peer_open(self, handle: bytes) -> int:
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        out = C.c_void_p()
        _chk(self.lib.slv_peer_open(self.dev, buf, C.byref(out)), "slv_peer_open")
        return out.value

    def peer_close(self, dptr: int):
        _chk(self.lib.slv_peer_close(self.dev, C.c_void_p(dptr)), "slv_peer_close")

    def resolve_target_peer(self, dst: Texture, peer_surface: int | None):
        _chk(self.lib.slv_resolve_target_peer(self.dev, dst.handle, C.c_void_p(peer_surface or 0)), "slv_resolve_target_peer")

    def assembly_wait(self, tex: Texture, flags: int | None, first: int, count: int, value: int):
        _chk(self.lib.slv_assembly_wait(self.dev, tex.handle, C.c_void_p(flags or 0), first, count, value & 0xFFFFFFFF), "slv_assembly_wait")

    def peer_signal_after_consumers(self, tex: Texture, peer_flags: int | None, index: int, value: int):
        _chk(self.lib.slv_peer_signal_after_consumers(self.dev, tex.handle, C.c_void_p(peer_flags or 0), index, value & 0xFFFFFFFF),
             "slv_peer_signal_after_consumers")

    def peer_signal(self, peer_flags: int | None, index: int, value: int):
        _chk(self.lib.slv_peer_signal(self.dev, C.c_void_p(peer_flags or 0), index, value & 0xFFFFFFFF), "slv_peer_signal")

    def flags_wait(self, flags: int | None, first: int, count: int, value: int):
        _chk(self.lib.slv_flags_wait(self.dev, C.c_void_p(flags or 0), first, count, value & 0xFFFFFFFF), "slv_flags_wait")

    def packed_tiles_bytes(self, tex: Texture, rank: int, nranks: int) -> int:
        n = C.c_size_t()
        _chk(self.lib.slv_pack_tiles(self.dev, tex.handle, rank, nranks, None, C.byref(n)), "slv_pack_tiles(size)")
        return n.value

    def pack_tiles(self, tex: Texture, rank: int, nranks: int, staging_ptr: int) -> int:
        n = C.c_size_t()
        _chk(self.lib.slv_pack_tiles(self.dev, tex.handle, rank, nranks, C.c_void_p(staging_ptr), C.byref(n)), "slv_pack_tiles")
        return n.value

    def unpack_tiles(self, tex: Texture, rank: int, nranks: int, staging_ptr: int):
        _chk(self.lib.slv_unpack_tiles(self.dev, tex.handle, rank, nranks, C.c_void_p(staging_ptr)), "slv_unpack_tiles")

    def upload_from_ptr(self, handle: int, host_ptr: int, nbytes: int, offset: int = 0):
        """slv_buffer_upload from a raw host address (e.g. pinned memory) — no intermediate copy."""
        _chk(self.lib.slv_buffer_upload(self.dev, handle, offset, C.c_void_p(host_ptr), nbytes), "slv_buffer_upload")

    def read_texture_into_async(self, tex: Texture, host_ptr: int, nbytes: int, level: int = 0):
        """Enqueues the copy on the library's copy stream; the bytes are valid after readback_wait() / flush()."""
        _chk(self.lib.slv_texture_readback_async(self.dev, tex.handle, level, C.c_void_p(host_ptr), nbytes), "slv_texture_readback_async")

    def buffer_device_ptr(self, buf: int):
        """(pointer, bytes) of a buffer's allocation (device memory for the product, host memory for the CPU checkers)."""
        p, n = C.c_void_p(), C.c_size_t()
        _chk(self.lib.slv_buffer_device_ptr(self.dev, buf, C.byref(p), C.byref(n)), "slv_buffer_device_ptr")
        return p.value, n.value

    def external_write_begin(self, cuda_stream: int, skip_latest: int = 0):
        _chk(self.lib.slv_external_write_begin(self.dev, C.c_void_p(cuda_stream), skip_latest), "slv_external_write_begin")

    def external_write_end(self, cuda_stream: int):
        _chk(self.lib.slv_external_write_end(self.dev, C.c_void_p(cuda_stream)), "slv_external_write_end")

    def host_register(self, host_ptr: int, nbytes: int):
        _chk(self.lib.slv_host_register(self.dev, C.c_void_p(host_ptr), nbytes), "slv_host_register")

    def host_unregister(self, host_ptr: int):
        _chk(self.lib.slv_host_unregister(self.dev, C.c_void_p(host_ptr)), "slv_host_unregister")

    def export_tiles_async(self, tex: Texture, host_ptr: int, nbytes: int):
        """This rank's owned tiles of the resolved surface -> the shared host frame (see sortfirst.HostFrame)."""
        _chk(self.lib.slv_texture_export_tiles_async(self.dev, tex.handle, C.c_void_p(host_ptr), nbytes), "slv_texture_export_tiles_async")

    def readback_fence(self, tex: Texture):
        _chk(self.lib.slv_readback_fence(self.dev, tex.handle), "slv_readback_fence")

    def readback_wait(self):
        _chk(self.lib.slv_readback_wait(self.dev), "slv_readback_wait")

    def read_texture_into(self, tex: Texture, host_ptr: int, nbytes: int, level: int = 0):
        _chk(self.lib.slv_texture_readback(self.dev, tex.handle, level, C.c_void_p(host_ptr), nbytes), "slv_texture_readback")
