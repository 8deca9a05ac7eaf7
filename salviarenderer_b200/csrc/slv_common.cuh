// slv_common.cuh — device-side data layout and exact-arithmetic helpers of the B200 SALVIA pipeline.
//
// Numerics contract: every float op is issued un-fused and in the reference's
// association order.  This translation unit is compiled with
//   -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
// so `a*b + c` is a rounded multiply followed by a rounded add, `/` and sqrtf are IEEE-correct and
// denormals are kept — the same results as the reference's x86-64 SSE2 build (no FMA).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "salvia_b200.h"

namespace slv {

constexpr int MAX_REGS = 1 + SLV_MAX_VS_OUTPUT_ATTRS;  // position + attributes of a vs_output
constexpr int TILE = SLV_TILE_SIZE;                    // reference tile, 64x64 px (rasterizer.cpp:32)
constexpr int REGION = 16;                             // one raster CTA = one level-16 child of a tile
constexpr int RASTER_THREADS = REGION * REGION;        // thread == pixel
constexpr int MAX_LEVELS = 14;                         // 8192 -> 1

// ---- resources -------------------------------------------------------------------------------------
struct SurfaceRef {  // reference layout: ((y*W + x)*S + s)*bpp (surface.cpp:277-295)
  uint8_t* data;
  uint32_t w, h, samples, fmt, bpp;
  uint32_t wmask, hmask;  // size-1 when the size is a power of two <= 1024 (exact integer wrap), else 0
  size_t bytes;
};

struct TextureRef {
  SurfaceRef level[MAX_LEVELS];
  uint32_t n_levels;
};

struct SamplerRef {
  slv_sampler_desc d;
  TextureRef tex;
  // 1 when min = mag = linear, u and v both wrap, texels are rgba8 and every level is a power of two <= 1024 (exact
  // integer wrap): the specialised tap sample_wrap_rgba8_linear applies (same arithmetic, no mode / format dispatch)
  uint32_t fast_wrap_rgba8;
  uint32_t* touched;  // measurement aid (slv_texture_level_tracking): bit l is set when mip level l is sampled; nullptr = off
};

// ---- per-draw parameter block (passed by value to the kernels; < 4 KB) ------------------------------
struct StreamRef {
  const uint8_t* data;
  uint32_t stride, offset;
};

struct GeomParams {
  StreamRef streams[8];
  slv_input_element elements[SLV_MAX_VS_INPUT_ATTRS];
  uint32_t n_elements;
  uint32_t fast_layout;    // every element is a 16-byte aligned rgba32f feeding register == element index: direct 128-bit loads
  const uint8_t* indices;  // nullptr for draw()
  uint32_t index_stride;   // 2 or 4
  uint32_t topology, start, prim_count;
  int32_t base_vertex;
  uint32_t vs_program;
  alignas(16) uint8_t vs_uniforms[SLV_MAX_UNIFORM_BYTES];  // built-in programs: their slv_vs_*_uniforms; SASL: the packed globals
  uint32_t n_attrs;
  uint32_t mods[SLV_MAX_VS_OUTPUT_ATTRS];
  uint32_t cull_mode, front_ccw;
  slv_viewport vp;
  uint32_t tiles_x, tiles_y;
  uint32_t shard_rank, shard_n;
  uint32_t slot_base;      // first global triangle slot of this draw inside the batch
  uint32_t draw_id;        // index of this draw inside the batch
  // outputs
  float4* tris;            // triangle records, slot = slot_base + prim*3 + k, stride tri_stride float4
  uint32_t tri_stride;
  uint32_t* tile_count;    // [tiles]
  unsigned long long* stats;  // slv_pipeline_statistics as 9 x u64
  uint32_t* valid_slots;      // compact list of the slots that hold a triangle binned on this rank
  uint32_t* valid_count;
  // triangles that span many tiles: the thread that sets one up does not walk its tile range itself (a full-screen triangle is
  // 2,040 tile tests at 4K - one thread's loop was the critical path of the whole kernel); it appends the slot here and
  // k_big_tiles counts the tiles with a warp per triangle
  uint32_t* big_slots;
  uint32_t* big_count;
  // two-kernel geometry (k_geometry_cull -> k_geometry): when surv != nullptr, k_geometry_cull has already run the position pass
  // for every primitive of the draw and left the ids of the ones that need set-up on this rank (clipped by the near / far plane,
  // or un-culled and reaching a tile this rank owns) in surv[slot_base / 3 ...], their number in surv_count[draw_id]; thread i of
  // the draw's CTAs then works on primitive surv[i] instead of primitive i
  uint32_t* surv;
  uint32_t* surv_count;
  // post-transform vertex cache (default_vertex_cache.cpp:128-197): when vc_pos != nullptr, k_vertex_mark / k_vertex_shade have
  // run the vertex shader ONCE for every vertex index < vc_cap that the draws sharing this cache reference; k_geometry gathers
  // the clip-space position (vc_pos[v]) and the attributes (vc_attr[v * n_attrs + i]) instead of re-running the shader per
  // corner.  Indices >= vc_cap (outside the bound buffers: undefined upstream) fall back to the per-corner run.
  const float4* vc_pos;
  const float4* vc_attr;
  uint8_t* vc_flags;          // [vc_cap] "referenced" marks (set by k_vertex_mark, consumed and cleared by k_vertex_shade)
  uint32_t vc_cap;
  SamplerRef sampler0;        // vertex texture fetch (vs.samplers[0]); tex.n_levels == 0 when the draw binds none
};

// bytes of GeomParams in front of the trailing sampler block (NVRTC has no offsetof: derived from the sizes; may include a few
// bytes of the sampler when the struct has tail padding, which is harmless for the staging copy that uses it)
constexpr size_t GEOM_PARAMS_HEAD_BYTES = sizeof(GeomParams) - sizeof(SamplerRef);
#ifndef __CUDACC_RTC__
static_assert(GEOM_PARAMS_HEAD_BYTES >= offsetof(GeomParams, sampler0) && GEOM_PARAMS_HEAD_BYTES < offsetof(GeomParams, sampler0) + 16 &&
              GEOM_PARAMS_HEAD_BYTES % 4 == 0, "sampler0 must stay the last member of GeomParams");
#endif

// geometry of all queued draws in ONE launch: CTA b works on draw draw_of[g] where cta_prefix[g] <= b < cta_prefix[g+1]
constexpr uint32_t MAX_BATCH_DRAWS = 64;
struct GeomBatch {
  uint32_t n;
  uint32_t cta_prefix[MAX_BATCH_DRAWS + 1];
  uint32_t draw_of[MAX_BATCH_DRAWS];
};

// triangle record (float4 units): [0..2] edge A,B,C,0  [3] bbox xmin,xmax,ymin,ymax
// [4] misc: x = as_uint(valid | front<<1), y = sx | ex<<16, z = sy | ey<<16 (tile range), w = draw id in the batch
// then one (v0, ddx, ddy) triple per register r (0 = position, 1.. = attributes) at TRI_HEADER + 3r: the layout does
// not depend on the draw's register count, so readers need no per-draw state to address it
constexpr int BIG_TILE_RANGE = 48;  // tile ranges above this many tiles are walked by a warp (k_big_tiles, k_bin_fill)
constexpr int TRI_HEADER = 5;
constexpr int REC_V0 = TRI_HEADER, REC_DDX = TRI_HEADER + 1, REC_DDY = TRI_HEADER + 2;  // + 3 * reg

struct BinParams {
  const float4* tris;
  uint32_t tri_stride, n_slots;  // n_slots = prim_count * 3
  uint32_t tiles_x, tiles_y;
  uint32_t shard_rank, shard_n;
  const uint32_t* tile_offset;  // [tiles + 1]
  uint32_t* tile_cursor;        // [tiles]
  uint32_t* list;               // entries (slot << 1) | accept
  uint32_t list_capacity;
  uint32_t* overflow_flag;
  const uint32_t* valid_slots;
  const uint32_t* valid_count;
};

struct RasterParams {
  const float4* tris;
  uint32_t tri_stride;
  uint32_t tiles_x, tiles_y;
  uint32_t shard_rank, shard_n;
  const uint32_t* tile_offset;
  const uint32_t* active_tiles;  // [0] = count, [1..] = ids of the non-empty tiles of this draw
  uint32_t* work_counter;        // persistent-CTA work queue head (reset by k_scan_tiles)
  const uint32_t* list;
  uint32_t list_capacity;
  uint32_t n_attrs;
  uint32_t mods[SLV_MAX_VS_OUTPUT_ATTRS];
  uint32_t has_centroid;
  uint32_t slot_base;             // first global triangle slot of this draw inside the batch (slots are draw-ordered)
  // targets
  SurfaceRef color0, color1, ds;  // data == nullptr when unbound
  uint32_t target_w, target_h;    // min over colour targets (renderer_impl.cpp:159-238)
  // depth-stencil state (framebuffer.cpp:358-425 resolved on the host)
  uint32_t depth_enable, depth_func, read_depth, write_depth, stencil_enable, early_z;
  uint32_t stencil_ref, read_mask, write_mask;
  slv_stencil_op_desc front_face, back_face;
  // shaders
  uint32_t ps_program, bs_program;
  alignas(16) uint8_t ps_uniforms[SLV_MAX_UNIFORM_BYTES];
  SamplerRef sampler0;
  SamplerRef sampler1;  // second pixel-shader sampler (SLV_PS_SSM_DRAW: the shadow map); tex.n_levels == 0 when unbound
  unsigned long long* stats;
};

// ---- exact float helpers (eflib/include/eflib/math/math.h:89-164) ------------------------------------
__device__ __forceinline__ float fast_log2(float val) {
  int x = __float_as_int(val);
  int log_2 = ((x >> 23) & 255) - 128;
  x &= ~(255 << 23);
  x += 127 << 23;
  float f = __int_as_float(x);
  f = ((-1.0f / 3) * f + 2) * f - 2.0f / 3;
  return f + (float)log_2;
}

__device__ __forceinline__ float fast_round(float val) {
  int n = __float_as_int(val);
  float bias = __int_as_float(((23 + 127) << 23) + (n & 0x80000000));
  float t = __fadd_rn(val, bias);
  return __fsub_rn(t, bias);
}

__device__ __forceinline__ float fast_floor(float val) {
  float f = fast_round(val);
  return (f > val) ? f - 1 : f;
}

__device__ __forceinline__ int fast_roundi(double d) { return (int)floor(d + 0.5); }
#define SLV_MAGIC_EPS ((double)0.5f - 1.5e-8)
__device__ __forceinline__ int fast_ceili(double d) { return fast_roundi(d + SLV_MAGIC_EPS); }
__device__ __forceinline__ int fast_floori(double d) { return fast_roundi(d - SLV_MAGIC_EPS); }

__device__ __forceinline__ bool eq_eps(float a, float b) { return fabsf(a - b) <= 1.1920928955078125e-7f; }

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return ax * bx + ay * by + az * bz;
}
__device__ __forceinline__ float length3(float x, float y, float z) {
  float t = 0.0f;
  t += x * x;
  t += y * y;
  t += z * z;
  return sqrtf(t);
}
__device__ __forceinline__ float length2(float x, float y) {
  float t = 0.0f;
  t += x * x;
  t += y * y;
  return sqrtf(t);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float f4_get(const float4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// ---- packed IEEE fp32 pairs (Blackwell FADD2 / FMUL2: two results per lane per issue slot) ------------------------------------
// Each half is the correctly rounded fp32 result, so a packed op equals the two scalar ops bit for bit.  Explicit .rn in the
// PTX; all the same ptxas (12.9) contracts a mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2 - even with --fmad=false -
// which would round once instead of twice.  So a SUM THAT CONSUMES A PACKED PRODUCT IS ALWAYS WRITTEN WITH SCALAR ADDS
// (add2_after_mul); packed adds are only used on operands that are not products.  The SASS of the library is checked for
// FFMA2 by tests/test_abi.py.
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 add2_after_mul(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat4(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float2 sub2_after_mul(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
// a + (b - a) * t on a pair whose inputs may themselves be packed products (the two taps of a trilinear fetch end in
// * (1 / 255)): difference and sum in scalar ops, only the product packed
__device__ __forceinline__ float2 lerp2_of_products(float2 a, float2 b, float2 t) { return add2_after_mul(a, mul2(sub2_after_mul(b, a), t)); }

// ---- cache hints for data with no reuse inside a frame (L2 evict-first): SLV_STREAM_HINTS=0 builds plain accesses ------------------
#ifndef SLV_STREAM_HINTS
#define SLV_STREAM_HINTS 1
#endif
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
#if SLV_STREAM_HINTS
  __stcs(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) {
#if SLV_STREAM_HINTS
  __stcs(p, v);
#else
  *p = v;
#endif
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
#if SLV_STREAM_HINTS
  return __ldcs(p);
#else
  return *p;
#endif
}

// float -> unorm8: mul 255, max 0, min 255, cvtps2dq (round-to-nearest-even)  (colors.h:182-192)
__device__ __forceinline__ uint32_t unorm8_rne(float x) {
  float m = x * 255.0f;
  m = (m > 0.0f) ? m : 0.0f;
  m = (m < 255.0f) ? m : 255.0f;
  return (uint32_t)__float2int_rn(m);
}

__device__ __forceinline__ uint32_t pack_color(uint32_t fmt, float4 c) {
  uint32_t r = unorm8_rne(c.x), g = unorm8_rne(c.y), b = unorm8_rne(c.z), a = unorm8_rne(c.w);
  return fmt == SLV_PF_RGBA8 ? (r | (g << 8) | (b << 16) | (a << 24)) : (b | (g << 8) | (r << 16) | (a << 24));
}

__device__ __forceinline__ float4 unpack_color(uint32_t fmt, uint32_t p) {
  const float inv_255 = 1.0f / 255;
  float b0 = (float)(p & 0xFF) * inv_255, b1 = (float)((p >> 8) & 0xFF) * inv_255;
  float b2 = (float)((p >> 16) & 0xFF) * inv_255, b3 = (float)(p >> 24) * inv_255;
  return fmt == SLV_PF_RGBA8 ? make_float4(b0, b1, b2, b3) : make_float4(b2, b1, b0, b3);
}

// generic texel <-> rgba32f (colors.h:148-267)
__device__ __forceinline__ float4 load_texel_rgba32f(uint32_t fmt, const uint8_t* p) {
  switch (fmt) {
  case SLV_PF_RGBA32F: return *reinterpret_cast<const float4*>(p);
  case SLV_PF_RG32F: {
    float2 v = *reinterpret_cast<const float2*>(p);
    return make_float4(v.x, v.y, 0.0f, 0.0f);
  }
  default: return unpack_color(fmt, *reinterpret_cast<const uint32_t*>(p));
  }
}

__device__ __forceinline__ void store_texel_rgba32f(uint32_t fmt, uint8_t* p, float4 c) {
  switch (fmt) {
  case SLV_PF_RGBA32F: *reinterpret_cast<float4*>(p) = c; break;
  case SLV_PF_RG32F: *reinterpret_cast<float2*>(p) = make_float2(c.x, c.y); break;
  default: *reinterpret_cast<uint32_t*>(p) = pack_color(fmt, c); break;
  }
}

}  // namespace slv

static_assert(sizeof(slv_vs_lights3_uniforms) <= SLV_MAX_UNIFORM_BYTES && sizeof(slv_vs_sponza_uniforms) <= SLV_MAX_UNIFORM_BYTES &&
                  sizeof(slv_vs_mvp_passthrough_uniforms) <= SLV_MAX_UNIFORM_BYTES && sizeof(slv_vs_plane_xz_uniforms) <= SLV_MAX_UNIFORM_BYTES,
              "GeomParams::vs_uniforms too small");
static_assert(sizeof(slv_ps_tex_alpha_uniforms) <= SLV_MAX_UNIFORM_BYTES && sizeof(slv_ps_sponza_uniforms) <= SLV_MAX_UNIFORM_BYTES, "ps_uniforms too small");
