// slv_kernels.cuh — the sm_100a kernels of the draw pipeline.
//
//   k_geometry      one launch per batch (all queued draws), one thread per input primitive: index fetch, vertex fetch, vertex shader, near/far
//                   clip, cull, fan, viewport+project, triangle setup, 64x64 tile coverage count.
//                   (reference phases 1-4: geom_setup_engine.cpp:43-124, clipper.cpp:43-228,
//                   shader.cpp:499-511, rasterizer.cpp:864-945, 775-862)
//   k_scan_tiles    exclusive scan of the per-tile counts -> per-tile list offsets
//   k_bin_fill      one thread per triangle slot: scatter (slot<<1 | accept) into the tile lists
//   k_sort_lists    one CTA per tile: restore API order (the reference's std::sort, rasterizer.cpp:976)
//   k_raster<S,PS>  one CTA per 16x16-pixel region (= one level-16 child of the reference's tile
//                   hierarchy), thread == pixel, all S samples of depth/stencil/colour live in registers
//                   for the whole draw; triangles are streamed through shared memory in API order.
//                   (reference phase 5: rasterizer.cpp:617-773, 298-439, 1245-1421, framebuffer.cpp:445-614)
#pragma once

#include "slv_common.cuh"
#include "slv_sampler.cuh"

// SASL shaders compiled at run time (salviarenderer_b200/sasl): the JIT translation unit csrc/slv_jit_unit.cu includes
// this header with SLV_JIT_VS / SLV_JIT_PS defined and the generated entry points below defined after it, so the shader is
// inlined into k_geometry / k_raster like the built-in programs.  The library build knows nothing about them.
#ifdef SLV_JIT_VS
namespace slv { struct SamplerRef; }
__device__ void slv_jit_vs(const float4* in, const unsigned char* uniforms, float4* out, const slv::SamplerRef& s0);
#endif
#ifdef SLV_JIT_PS
namespace slv { struct RasterParams; }
template <class Ctx>
__device__ bool slv_jit_ps(const slv::RasterParams& p, const Ctx& px, float4& color);
#endif

namespace slv {

// =====================================================================================================
// geometry
// =====================================================================================================
template <int R>
struct VsOut {
  float4 r[R];
};

__device__ __forceinline__ float4 fetch_element(const GeomParams& p, const slv_input_element& el, uint32_t index) {
  const StreamRef& st = p.streams[el.slot];
  const float* f = reinterpret_cast<const float*>(st.data + el.aligned_byte_offset + (size_t)st.stride * index + st.offset);
  switch (el.format) {  // get_vec4 (stream_assembler.cpp:26-45)
  case SLV_FMT_R32_FLOAT: return make_float4(__ldg(f), 0.0f, 0.0f, el.default_w);
  case SLV_FMT_R32G32_FLOAT: return make_float4(__ldg(f), __ldg(f + 1), 0.0f, el.default_w);
  case SLV_FMT_R32G32B32_FLOAT: return make_float4(__ldg(f), __ldg(f + 1), __ldg(f + 2), el.default_w);
  default:
    if ((reinterpret_cast<uintptr_t>(f) & 15) == 0) return __ldg(reinterpret_cast<const float4*>(f));
    return make_float4(__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3));
  }
}

// pos = v · M (row vector × matrix; eflib/src/math.cpp:142-154)
__device__ __forceinline__ float4 transform(float4 v, const float* m) {
  float4 o;
  o.x = v.x * m[0] + v.y * m[4] + v.z * m[8] + v.w * m[12];
  o.y = v.x * m[1] + v.y * m[5] + v.z * m[9] + v.w * m[13];
  o.z = v.x * m[2] + v.y * m[6] + v.z * m[10] + v.w * m[14];
  o.w = v.x * m[3] + v.y * m[7] + v.z * m[11] + v.w * m[15];
  return o;
}

// vertex texture fetch: ONE out-of-line copy of the sampler for the whole geometry kernel (run_vs is instantiated six times per
// primitive - three corners, position-only and full - and the sampler is ~3 k instructions: inlined, it pushed the kernel out
// of the instruction cache and cost every draw 15 % whether it sampled or not)
__device__ __noinline__ float4 vs_sample_lod(const SamplerRef& sm, float u, float v, float lod) { return sample_impl(sm, u, v, lod, nullptr); }

// POS_ONLY: only input register 0 is fetched and only out.r[0] is meaningful (every shipped vertex program derives the
// position from in[0] alone); the attribute fetches and arithmetic are dead code in that instantiation.
template <int R, bool POS_ONLY = false>
__device__ __forceinline__ void run_vs(const GeomParams& p, uint32_t index, VsOut<R>& out) {
  float4 in[SLV_MAX_VS_INPUT_ATTRS];
#pragma unroll
  for (int i = 0; i < SLV_MAX_VS_INPUT_ATTRS; ++i) in[i] = make_float4(0, 0, 0, 0);
  // position-only pass: the built-in programs derive the position from in[0] alone, except the vertex-texture-fetch one (uv
  // in in[1]); a SASL vertex shader may read anything
  const bool pos_from_in0 = POS_ONLY && p.vs_program != SLV_VS_TERRAIN_VTF && p.vs_program != SLV_VS_JIT;
  if (p.fast_layout) {  // element e -> register e, one aligned 128-bit load each (the interleaved layouts of the samples)
#pragma unroll
    for (int e = 0; e < SLV_MAX_VS_INPUT_ATTRS; ++e)
      if ((uint32_t)e < p.n_elements && (e == 0 || !pos_from_in0)) {
        const slv_input_element& el = p.elements[e];
        const StreamRef& st = p.streams[el.slot];
        in[e] = __ldg(reinterpret_cast<const float4*>(st.data + el.aligned_byte_offset + (size_t)st.stride * index + st.offset));
      }
  } else {
    for (uint32_t e = 0; e < p.n_elements; ++e) {
      uint32_t reg = p.elements[e].reg;
      if (pos_from_in0 && reg != 0) continue;
      float4 v = fetch_element(p, p.elements[e], index);
#pragma unroll
      for (int i = 0; i < SLV_MAX_VS_INPUT_ATTRS; ++i)
        if (reg == (uint32_t)i) in[i] = v;
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) out.r[i] = make_float4(0, 0, 0, 0);
  switch (p.vs_program) {
  case SLV_VS_MVP_PASSTHROUGH: {
    auto u = reinterpret_cast<const slv_vs_mvp_passthrough_uniforms*>(p.vs_uniforms);
    out.r[0] = transform(in[0], u->wvp);
#pragma unroll
    for (int i = 1; i < R; ++i) {
      uint32_t s = u->src[i - 1];
      float4 v = in[0];
#pragma unroll
      for (int k = 1; k < SLV_MAX_VS_INPUT_ATTRS; ++k)
        if (s == (uint32_t)k) v = in[k];
      out.r[i] = v;
    }
  } break;
  case SLV_VS_PLANE_XZ: {
    auto u = reinterpret_cast<const slv_vs_plane_xz_uniforms*>(p.vs_uniforms);
    out.r[0] = transform(in[0], u->wvp);
    if (R > 1) out.r[R > 1 ? 1 : 0] = make_float4(in[0].x, in[0].z, 0.0f, 0.0f);
  } break;
  case SLV_VS_LIGHTS3: {
    auto u = reinterpret_cast<const slv_vs_lights3_uniforms*>(p.vs_uniforms);
    out.r[0] = transform(in[0], u->wvp);
    if (R == 5) {
      out.r[R > 1 ? 1 : 0] = in[1];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float4 l = make_float4(u->light_pos[k][0], u->light_pos[k][1], u->light_pos[k][2], u->light_pos[k][3]);
        out.r[R > 4 ? 2 + k : 0] = f4_sub(l, in[0]);
      }
    }
  } break;
  case SLV_VS_TERRAIN_VTF: {  // VertexTextureFetch.cpp:38-61; tex2Dlod = sampler::sample_2d_lod (sampler_api.cpp:50-52)
    auto u = reinterpret_cast<const slv_vs_terrain_vtf_uniforms*>(p.vs_uniforms);
    const float tu = u->offset[0] + in[1].x * u->scale[0], tv = u->offset[1] + in[1].y * u->scale[1];
    const float disp = vs_sample_lod(p.sampler0, tu, tv, 0.0f).x;
    out.r[0] = transform(make_float4(in[0].x + 0.0f, in[0].y + disp * 20.0f, in[0].z + 0.0f, 1.0f), u->wvp);
    if (R > 1) out.r[R > 1 ? 1 : 0] = make_float4(disp, 0.0f, 0.0f, 0.0f);
  } break;
#ifdef SLV_JIT_VS
  case SLV_VS_JIT:
    if (R == SLV_JIT_R) slv_jit_vs(in, p.vs_uniforms, out.r, p.sampler0);
    break;
#endif
  case SLV_VS_SSM_DRAW: {  // resources/ssm/Draw.savs (StandardShadowMap colour pass)
    auto u = reinterpret_cast<const slv_vs_ssm_draw_uniforms*>(p.vs_uniforms);
    out.r[0] = transform(in[0], u->camera_wvp);
    if (R == 6) {
      out.r[R > 5 ? 1 : 0] = in[2];
      out.r[R > 5 ? 2 : 0] = in[1];
      out.r[R > 5 ? 3 : 0] = f4_sub(make_float4(u->light_pos[0], u->light_pos[1], u->light_pos[2], u->light_pos[3]), in[0]);
      out.r[R > 5 ? 4 : 0] = f4_sub(make_float4(u->camera_pos[0], u->camera_pos[1], u->camera_pos[2], u->camera_pos[3]), in[0]);
      out.r[R > 5 ? 5 : 0] = transform(in[0], u->light_wvp);
    }
  } break;
  case SLV_VS_SPONZA: {
    auto u = reinterpret_cast<const slv_vs_sponza_uniforms*>(p.vs_uniforms);
    out.r[0] = transform(in[0], u->wvp);
    if (R == 5) {
      out.r[R > 4 ? 1 : 0] = in[1];
      out.r[R > 4 ? 2 : 0] = in[2];
      out.r[R > 4 ? 3 : 0] = f4_sub(make_float4(u->light_pos[0], u->light_pos[1], u->light_pos[2], u->light_pos[3]), in[0]);
      out.r[R > 4 ? 4 : 0] = f4_sub(make_float4(u->eye_pos[0], u->eye_pos[1], u->eye_pos[2], u->eye_pos[3]), in[0]);
    }
  } break;
  }
}

__device__ __forceinline__ bool cull_tri(uint32_t cull_mode, uint32_t front_ccw, float area) {  // raster_state.cpp:10-31
  switch (cull_mode) {
  case SLV_CULL_FRONT: return front_ccw ? (area <= 0) : (area >= 0);
  case SLV_CULL_BACK: return front_ccw ? (area >= 0) : (area <= 0);
  default: return false;
  }
}

template <int R>
__device__ __forceinline__ void lerp_vso(const GeomParams& p, VsOut<R>& out, const VsOut<R>& a, const VsOut<R>& b, float t) {
  // shader.cpp:170-180: start + (end - start) * step; nointerpolation attributes copied from start
  out.r[0] = make_float4(a.r[0].x + (b.r[0].x - a.r[0].x) * t, a.r[0].y + (b.r[0].y - a.r[0].y) * t,
                         a.r[0].z + (b.r[0].z - a.r[0].z) * t, a.r[0].w + (b.r[0].w - a.r[0].w) * t);
#pragma unroll
  for (int i = 1; i < R; ++i) {
    float4 s = a.r[i];
    if (!(p.mods[i - 1] & SLV_AM_NOINTERPOLATION)) {
      s.x += (b.r[i].x - a.r[i].x) * t;
      s.y += (b.r[i].y - a.r[i].y) * t;
      s.z += (b.r[i].z - a.r[i].z) * t;
      s.w += (b.r[i].w - a.r[i].w) * t;
    }
    out.r[i] = s;
  }
}

__device__ __forceinline__ float plane_dist(int plane, float4 pos) {
  // dot_prod4 with (0,0,1,0) / (0,0,-1,1), zero terms included (clipper.cpp:21-27,111)
  return plane == 0 ? (0.0f * pos.x + 0.0f * pos.y + 1.0f * pos.z + 0.0f * pos.w)
                    : (0.0f * pos.x + 0.0f * pos.y + -1.0f * pos.z + 1.0f * pos.w);
}

// viewport_transform + project_n (shader.cpp:499-511, 116-134), position part: returns the screen position, .w = 1/w
__device__ __forceinline__ float4 viewport_project_pos(const GeomParams& p, float4 pos) {
  const slv_viewport& vp = p.vp;
  float w = pos.w;
  float invw = eq_eps(w, 0.0f) ? 1.0f : 1.0f / w;
  float px = pos.x * invw, py = pos.y * invw, pz = pos.z * invw;
  float ox = (vp.x + vp.w) * 0.5f;
  float oy = (vp.y + vp.h) * 0.5f;
  return make_float4((vp.w * 0.5f) * px + ox, (vp.h * 0.5f) * -py + oy, (vp.maxz - vp.minz) * pz + vp.minz, invw);
}
// attribute part: attributes * (1/w) unless noperspective
template <int R>
__device__ __forceinline__ void project_attrs(const GeomParams& p, VsOut<R>& v, float invw) {
#pragma unroll
  for (int i = 1; i < R; ++i)
    if (!(p.mods[i - 1] & SLV_AM_NOPERSPECTIVE)) {
      v.r[i].x *= invw; v.r[i].y *= invw; v.r[i].z *= invw; v.r[i].w *= invw;
    }
}
template <int R>
__device__ __forceinline__ void viewport_project(const GeomParams& p, VsOut<R>& v) {
  v.r[0] = viewport_project_pos(p, v.r[0]);
  project_attrs<R>(p, v, v.r[0].w);
}

// the reference's tile-level test (rasterizer.cpp:831-848): returns 0 = rejected, 1 = partial, 3 = accepted
__device__ __forceinline__ int tile_test(const float4 edge[3], int x, int y) {
  int rejection = 0, acceptance = 1;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    float A = edge[e].x, B = edge[e].y, C = edge[e].z;
    int mark_x = A > 0, mark_y = B > 0;
    float step_x = TILE * A, step_y = TILE * B;
    float rej_to_acc = -fabsf(step_x) - fabsf(step_y);
    float ev = C - ((float)(x + mark_x) * TILE * A + (float)(y + mark_y) * TILE * B);
    rejection |= (0 < ev);
    acceptance &= (rej_to_acc >= ev);
  }
  return rejection ? 0 : (acceptance ? 3 : 1);
}

struct TileRange { int sx, sy, ex, ey; };

__device__ __forceinline__ TileRange tile_range(const float bbox[4], uint32_t tiles_x, uint32_t tiles_y) {
  // rasterizer.cpp:800-807 (fast_floori / fast_ceili evaluated in double)
  TileRange r;
  r.sx = min(fast_floori((double)(std_max(0.0f, bbox[0]) / TILE)), (int)tiles_x);
  r.sy = min(fast_floori((double)(std_max(0.0f, bbox[2]) / TILE)), (int)tiles_y);
  r.ex = min(fast_ceili((double)(std_max(0.0f, bbox[1]) / TILE)) + 1, (int)tiles_x);
  r.ey = min(fast_ceili((double)(std_max(0.0f, bbox[3]) / TILE)) + 1, (int)tiles_y);
  return r;
}

__device__ __forceinline__ bool tile_owned(uint32_t tx, uint32_t ty, uint32_t rank, uint32_t n) {
  return n <= 1 || ((tx + 3 * ty) % n) == rank;
}

// rasterizer::compute_triangle_info (rasterizer.cpp:864-945) + record store + tile coverage count, in two steps:
// setup_position needs the three screen positions only (rotation, area, bounding box, edge equations, tile range, tile
// counting); setup_store computes the derivatives of every register and writes the record.  Triangles that are
// degenerate or reach none of this rank's tiles stop after the first step, before their attributes are even fetched.
struct TriPos {
  float4 edge[3];
  float bbox[4];
  TileRange tr;
  int r0;                         // vertex nearest the origin: the record's v0 (rasterizer.cpp:873-890)
  float e01x, e01y, e02x, e02y, inv_area;
  bool front;
  bool big;  // the tile range is left to k_big_tiles (the caller appends the slot to p.big_slots once the record is stored)
};

// COUNT = false: the ownership / coverage test only (k_geometry_cull): no tile counter is touched, the first hit ends the search
template <bool COUNT = true>
__device__ __forceinline__ bool setup_position(const GeomParams& p, const float4 v[3], TriPos& s) {  // true: binned somewhere
  double d0 = (double)fabsf(v[0].x) + (double)fabsf(v[0].y);
  double d1 = (double)fabsf(v[1].x) + (double)fabsf(v[1].y);
  double d2 = (double)fabsf(v[2].x) + (double)fabsf(v[2].y);
  int r0;
  if (d0 < d1) r0 = (d0 < d2) ? 0 : 2;
  else r0 = (d1 < d2) ? 1 : 2;
  s.r0 = r0;
  const float4 a = r0 == 0 ? v[0] : (r0 == 1 ? v[1] : v[2]);
  const float4 b = r0 == 0 ? v[1] : (r0 == 1 ? v[2] : v[0]);
  const float4 c = r0 == 0 ? v[2] : (r0 == 1 ? v[0] : v[1]);
  s.e01x = b.x - a.x; s.e01y = b.y - a.y;
  s.e02x = c.x - a.x; s.e02y = c.y - a.y;
  float area = s.e02x * s.e01y - s.e02y * s.e01x;  // cross_prod2(e02.xy, e01.xy)
  if (eq_eps(area, 0.0f)) return false;            // invalid (v0 == nullptr upstream)
  s.front = area > 0.0f;
  s.inv_area = 1.0f / area;
  s.bbox[0] = std_min(std_min(v[0].x, v[1].x), v[2].x);
  s.bbox[1] = std_max(std_max(v[0].x, v[1].x), v[2].x);
  s.bbox[2] = std_min(std_min(v[0].y, v[1].y), v[2].y);
  s.bbox[3] = std_max(std_max(v[0].y, v[1].y), v[2].y);
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // original vertex order (rasterizer.cpp:928-939)
    float4 st = v[i], e = v[(i + 1) % 3];
    s.edge[i] = make_float4(st.y - e.y, e.x - st.x, e.x * st.y - e.y * st.x, 0.0f);
  }
  // tile coverage count (rasterizer.cpp:809-857); under sort-first sharding only this rank's tiles count
  s.tr = tile_range(s.bbox, p.tiles_x, p.tiles_y);
  const TileRange& tr = s.tr;
  s.big = false;
  if (COUNT && p.big_slots && (tr.ex - tr.sx) * (tr.ey - tr.sy) > BIG_TILE_RANGE) {
    s.big = true;  // counted by a warp of k_big_tiles; with this many tiles some are this rank's in practice (if none is, the
    return true;   // record is stored for nothing and k_bin_fill finds no tile for it: harmless)
  }
  bool any_owned = false;
  if ((tr.sx + 1 == tr.ex) && (tr.sy + 1 == tr.ey)) {
    if (tile_owned(tr.sx, tr.sy, p.shard_rank, p.shard_n)) {
      if (COUNT) {
        // warp-aggregated count: neighbouring primitives of a fine mesh fall into the same tile, and one atomic per distinct
        // tile among the lanes that are converged here replaces up to 32 same-address atomics (the 10 M-triangle stress config)
        const uint32_t t = tr.sy * p.tiles_x + tr.sx;
        const uint32_t peers = __match_any_sync(__activemask(), t);
        if ((threadIdx.x & 31u) == (uint32_t)__ffs(peers) - 1u) atomicAdd(&p.tile_count[t], (uint32_t)__popc(peers));
      }
      any_owned = true;
    }
  } else {
    for (int y = tr.sy; y < tr.ey && (COUNT || !any_owned); ++y)
      for (int x = tr.sx; x < tr.ex && (COUNT || !any_owned); ++x)
        if (tile_owned(x, y, p.shard_rank, p.shard_n) && tile_test(s.edge, x, y)) {
          if (COUNT) atomicAdd(&p.tile_count[y * p.tiles_x + x], 1u);
          any_owned = true;
        }
  }
  return any_owned;
}

template <int R>
__device__ __forceinline__ void setup_store(const GeomParams& p, const VsOut<R> v[3], const TriPos& s, float4* rec) {
  const int r0 = s.r0;
  const float e01x = s.e01x, e01y = s.e01y, e02x = s.e02x, e02y = s.e02y, inv_area = s.inv_area;
  rec[0] = s.edge[0];
  rec[1] = s.edge[1];
  rec[2] = s.edge[2];
  rec[3] = make_float4(s.bbox[0], s.bbox[1], s.bbox[2], s.bbox[3]);
  float4 misc;
  misc.x = __uint_as_float(1u | (s.front ? 2u : 0u));
  misc.y = __uint_as_float((uint32_t)s.tr.sx | ((uint32_t)s.tr.ex << 16));
  misc.z = __uint_as_float((uint32_t)s.tr.sy | ((uint32_t)s.tr.ey << 16));
  misc.w = __uint_as_float(p.draw_id);
  rec[4] = misc;
  // compute_derivative_n (shader.cpp:413-449), registers rotated so that v0 is the vertex nearest the origin
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const float4 a = r0 == 0 ? v[0].r[i] : (r0 == 1 ? v[1].r[i] : v[2].r[i]);
    const float4 b = r0 == 0 ? v[1].r[i] : (r0 == 1 ? v[2].r[i] : v[0].r[i]);
    const float4 c = r0 == 0 ? v[2].r[i] : (r0 == 1 ? v[0].r[i] : v[1].r[i]);
    float4 e01 = f4_sub(b, a), e02 = f4_sub(c, a);
    float4 ddx, ddy;
    ddx.x = (e02.x * e01y - e01.x * e02y) * inv_area;
    ddx.y = (e02.y * e01y - e01.y * e02y) * inv_area;
    ddx.z = (e02.z * e01y - e01.z * e02y) * inv_area;
    ddx.w = (e02.w * e01y - e01.w * e02y) * inv_area;
    ddy.x = (e01.x * e02x - e02.x * e01x) * inv_area;
    ddy.y = (e01.y * e02x - e02.y * e01x) * inv_area;
    ddy.z = (e01.z * e02x - e02.z * e01x) * inv_area;
    ddy.w = (e01.w * e02x - e02.w * e01x) * inv_area;
    // the x / y derivatives of the screen position itself are never read (they are 1 / 0): the slot carries the draw id,
    // which k_cover / k_shade pick up with the depth / w derivatives they load anyway
    if (i == 0) ddx.x = __uint_as_float(p.draw_id);
    rec[REC_V0 + 3 * i] = a;
    rec[REC_DDX + 3 * i] = ddx;
    rec[REC_DDY + 3 * i] = ddy;
  }
}

template <int R>
__device__ __forceinline__ bool setup_triangle(const GeomParams& p, const VsOut<R> v[3], float4* rec, uint32_t slot) {  // true: binned somewhere
  const float4 pos[3] = {v[0].r[0], v[1].r[0], v[2].r[0]};
  TriPos s;
  if (!setup_position(p, pos, s)) return false;
  setup_store<R>(p, v, s, rec);
  if (s.big) p.big_slots[atomicAdd(p.big_count, 1u)] = slot;
  return true;
}

// the CTA's entry of a GeomBatch: CTA b works on entry g with cta_prefix[g] <= b < cta_prefix[g + 1]
__device__ __forceinline__ uint32_t batch_entry_of_cta(const GeomBatch& hb) {
  uint32_t lo = 0, hi = hb.n;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (blockIdx.x >= hb.cta_prefix[mid]) lo = mid; else hi = mid;
  }
  return lo;
}
// the draw's parameter block, staged once per CTA: every later field access is a shared-memory broadcast
__device__ __forceinline__ void stage_geom_params(GeomParams& s_params, const GeomParams* gsrc) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(gsrc);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&s_params);
  // the sampler block (vertex texture fetch) is the tail of the struct: staged only for draws that bind one
  const bool has_sampler = gsrc->sampler0.tex.n_levels != 0;
  const uint32_t n_words = (uint32_t)((has_sampler ? sizeof(GeomParams) : GEOM_PARAMS_HEAD_BYTES) / 4);
  for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = __ldg(src + i);
  if (!has_sampler && threadIdx.x == 0) s_params.sampler0.tex.n_levels = 0;
  __syncthreads();
}

// index_fetcher.cpp:26-115: the i-th index of the draw (already offset by start), plus base_vertex
__device__ __forceinline__ uint32_t fetch_index(const GeomParams& p, uint32_t i) {
  uint32_t v = i;
  if (p.indices) {
    const uint8_t* base = p.indices + (size_t)p.start * p.index_stride;
    v = p.index_stride == 2 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(base) + i)
                            : __ldg(reinterpret_cast<const uint32_t*>(base) + i);
  }
  return v + (uint32_t)p.base_vertex;
}

// ---- post-transform vertex cache (default_vertex_cache.cpp:128-197, the precomputed flavour: the vertex shader runs once per
// distinct index of the draw).  k_vertex_mark flags the vertices the queued draws reference, k_vertex_shade runs the shader on
// the flagged ones and stores clip-space position + attributes, k_geometry gathers.  Draws of a batch with the same vertex
// state (streams, layout, program, uniforms) share ONE cache: a mesh drawn per material group is transformed once.
constexpr int VC_MARK_PER_THREAD = 4;
__global__ void __launch_bounds__(256) k_vertex_mark(const GeomParams* __restrict__ draws, GeomBatch hb) {
  const uint32_t g = batch_entry_of_cta(hb);
  const GeomParams& p = draws[hb.draw_of[g]];
  const uint32_t n_idx = p.topology == SLV_TOPO_TRIANGLE_LIST ? p.prim_count * 3 : p.prim_count + 2;
  uint8_t* flags = p.vc_flags;
  const uint32_t cap = p.vc_cap;
  const uint32_t base = (blockIdx.x - hb.cta_prefix[g]) * (256 * VC_MARK_PER_THREAD) + threadIdx.x;
#pragma unroll
  for (int k = 0; k < VC_MARK_PER_THREAD; ++k) {
    const uint32_t i = base + k * 256;
    if (i < n_idx) {
      const uint32_t v = fetch_index(p, i);
      if (v < cap) flags[v] = 1;  // same value from every writer: no atomic needed
    }
  }
}

template <int R>
__device__ __forceinline__ void vertex_shade_main(const GeomParams* __restrict__ draws, const GeomBatch& hb) {
  const uint32_t g = batch_entry_of_cta(hb);
  __shared__ GeomParams s_params;
  stage_geom_params(s_params, draws + hb.draw_of[g]);
  const GeomParams& p = s_params;
  const uint32_t v = (blockIdx.x - hb.cta_prefix[g]) * blockDim.x + threadIdx.x;
  uint32_t ran = 0;
  if (v < p.vc_cap && p.vc_flags[v]) {
    p.vc_flags[v] = 0;  // self-cleaning: the next batch that uses this scratch set starts from zeroed marks
    VsOut<R> o;
    run_vs<R>(p, v, o);
    const_cast<float4*>(p.vc_pos)[v] = o.r[0];
    float4* a = const_cast<float4*>(p.vc_attr) + (size_t)v * (R - 1);
#pragma unroll
    for (int i = 1; i < R; ++i) a[i - 1] = o.r[i];
    ran = 1;
  }
  const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, ran));
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(&p.stats[17], (unsigned long long)n);  // vs_invocations of cached draws
}
template <int R>
__global__ void __launch_bounds__(128) k_vertex_shade(const GeomParams* __restrict__ draws, GeomBatch hb) {
  vertex_shade_main<R>(draws, hb);
}

template <int R>
__device__ __forceinline__ void geometry_main(const GeomParams* __restrict__ draws, const GeomBatch& hb) {
  // every queued draw with R registers shares this launch; a CTA belongs to exactly one draw
  const uint32_t lo = batch_entry_of_cta(hb);
  __shared__ GeomParams s_params;
  stage_geom_params(s_params, draws + hb.draw_of[lo]);
  const GeomParams& p = s_params;
  uint32_t prim = (blockIdx.x - hb.cta_prefix[lo]) * blockDim.x + threadIdx.x;
  if (p.surv) {  // second kernel of the two-kernel geometry: the i-th survivor of k_geometry_cull instead of primitive i
    const uint32_t n_surv = __ldg(p.surv_count + p.draw_id);
    if ((blockIdx.x - hb.cta_prefix[lo]) * blockDim.x >= n_surv) return;  // whole CTA beyond the list (uniform)
    prim = prim < n_surv ? p.surv[p.slot_base / 3 + prim] : 0xFFFFFFFFu;
  }
  uint32_t n_out = 0, valid_mask = 0;  // valid_mask bit k: slot prim*3+k holds a triangle binned on this rank
  if (prim < p.prim_count) {
    // ---- index fetch (index_fetcher.cpp:26-115)
    uint32_t ids[3];
    if (p.topology == SLV_TOPO_TRIANGLE_LIST) {
      ids[0] = prim * 3; ids[1] = prim * 3 + 1; ids[2] = prim * 3 + 2;
    } else {
      ids[0] = prim; ids[1] = prim + 1; ids[2] = prim + 2;
      if (prim & 1) { uint32_t t = ids[0]; ids[0] = ids[2]; ids[2] = t; }
    }
    uint32_t idx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) idx[i] = fetch_index(p, ids[i]);
    // ---- post-transform vertices: gathered from the vertex cache when the draw has one (k_vertex_shade ran the shader once per
    //      referenced vertex), else vertex fetch + vertex shader recomputed per corner (the VS is pure, so both equal the
    //      reference's cache hit: default_vertex_cache.cpp:354-390).  Positions first: most primitives are culled, degenerate
    //      or (sort-first) outside this rank's tiles and never need their attributes.
    const bool use_vc = p.vc_pos != nullptr && idx[0] < p.vc_cap && idx[1] < p.vc_cap && idx[2] < p.vc_cap;
    float4 cpos[3];
    if (use_vc) {
#pragma unroll
      for (int i = 0; i < 3; ++i) cpos[i] = __ldg(p.vc_pos + idx[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        VsOut<R> t;
        run_vs<R, true>(p, idx[i], t);
        cpos[i] = t.r[0];
      }
    }

    float4* rec = p.tris + ((size_t)p.slot_base + (size_t)prim * 3) * p.tri_stride;
    // ---- clip (clipper.cpp:103-228)
    bool in_frustum = true;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
      for (int v = 0; v < 3; ++v)
        if (plane_dist(pl, cpos[v]) < 0) in_frustum = false;

    if (in_frustum) {
      float px[3], py[3];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        float iw = 1.0f / cpos[v].w;
        px[v] = cpos[v].x * iw;
        py[v] = cpos[v].y * iw;
      }
      float area = (px[2] - px[0]) * (py[1] - py[0]) - (py[2] - py[0]) * (px[1] - px[0]);
      bool front = area > 0.0f;
      if (!cull_tri(p.cull_mode, p.front_ccw, front ? 1.0f : -1.0f)) {
        // un-clipped back faces get v1 <-> v2 swapped (clipper.cpp:59-65)
        const uint32_t oi[3] = {idx[0], front ? idx[1] : idx[2], front ? idx[2] : idx[1]};
        float4 spos[3];
        spos[0] = viewport_project_pos(p, cpos[0]);
        spos[1] = viewport_project_pos(p, front ? cpos[1] : cpos[2]);
        spos[2] = viewport_project_pos(p, front ? cpos[2] : cpos[1]);
        TriPos ts;
        if (setup_position(p, spos, ts)) {
          VsOut<R> o[3];
          if (use_vc) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float4* a = p.vc_attr + (size_t)oi[k] * (R - 1);
#pragma unroll
              for (int i = 1; i < R; ++i) o[k].r[i] = __ldg(a + (i - 1));
            }
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) run_vs<R>(p, oi[k], o[k]);
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            o[k].r[0] = spos[k];
            project_attrs<R>(p, o[k], spos[k].w);
          }
          setup_store<R>(p, o, ts, rec);
          if (ts.big) p.big_slots[atomicAdd(p.big_count, 1u)] = p.slot_base + prim * 3;
          valid_mask = 1u;
        }
        n_out = p.surv ? 0 : 1;  // two-kernel geometry: k_geometry_cull has counted the primitives that need no clipping
      }
    } else {
      VsOut<R> pool[2][5];
      int n[2] = {3, 0};
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        if (use_vc) {
          pool[0][i].r[0] = __ldg(p.vc_pos + idx[i]);
          const float4* a = p.vc_attr + (size_t)idx[i] * (R - 1);
#pragma unroll
          for (int k = 1; k < R; ++k) pool[0][i].r[k] = __ldg(a + (k - 1));
        } else {
          run_vs<R>(p, idx[i], pool[0][i]);
        }
      }
      int src = 0, dst = 1;
      bool is_front = false, culled = false;
      for (int pl = 0; pl < 2 && !culled; ++pl) {
        n[dst] = 0;
        float dd0 = 0.0f, dd1;
        if (n[src] != 0) dd0 = plane_dist(pl, pool[src][0].r[0]);
        for (int i = 0, j = 1; i < n[src]; ++i, ++j) {
          j %= n[src];
          dd1 = plane_dist(pl, pool[src][j].r[0]);
          if (dd0 >= 0.0f) {
            pool[dst][n[dst]++] = pool[src][i];
            if (dd1 < 0.0f) {
              lerp_vso<R>(p, pool[dst][n[dst]], pool[src][i], pool[src][j], dd0 / (dd0 - dd1));
              ++n[dst];
            }
          } else if (dd1 >= 0.0f) {
            lerp_vso<R>(p, pool[dst][n[dst]], pool[src][j], pool[src][i], dd1 / (dd1 - dd0));
            ++n[dst];
          }
          dd0 = dd1;
        }
        if (pl == 0 && n[dst] >= 3) {  // facing after the near plane (clipper.cpp:191-211)
          float px[3], py[3];
          for (int i = 0; i < 3; ++i) {
            float inv_abs_w = 1 / fabsf(pool[dst][i].r[0].w);
            px[i] = pool[dst][i].r[0].x * inv_abs_w;
            py[i] = pool[dst][i].r[0].y * inv_abs_w;
          }
          float area = (px[2] - px[0]) * (py[1] - py[0]) - (py[2] - py[0]) * (px[1] - px[0]);
          is_front = area > 0.0f;
          if (cull_tri(p.cull_mode, p.front_ccw, area)) culled = true;
        }
        src ^= 1;
        dst ^= 1;
      }
      int nv = culled ? 0 : n[src];
      if (nv >= 3) {
        for (int t = 1; t <= nv - 2; ++t) {  // fan (0, t, t+1); back faces reversed (clipper.cpp:75-89)
          VsOut<R> o[3];
          o[0] = pool[src][0];
          o[1] = is_front ? pool[src][t] : pool[src][t + 1];
          o[2] = is_front ? pool[src][t + 1] : pool[src][t];
#pragma unroll
          for (int k = 0; k < 3; ++k) viewport_project<R>(p, o[k]);
          if (setup_triangle<R>(p, o, rec + (size_t)(t - 1) * p.tri_stride, p.slot_base + prim * 3 + (uint32_t)(t - 1))) valid_mask |= 1u << (t - 1);
        }
        n_out = nv - 2;
      }
    }
  }
  // compact list of the slots k_bin_fill has to look at (order is irrelevant: the tile lists are sorted afterwards);
  // one atomicAdd per warp
  {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t nv = __popc(valid_mask), incl = nv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && warp_total) base = atomicAdd(p.valid_count, warp_total);
    base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - nv;
    const uint32_t slot0 = p.slot_base + prim * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (valid_mask & (1u << k)) p.valid_slots[base++] = slot0 + k;
  }
  // cprimitives (rasterizer.cpp:1138): warp-aggregated
  uint32_t total = n_out;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
  if ((threadIdx.x & 31) == 0 && total) atomicAdd(&p.stats[6], (unsigned long long)total);
}

// ---- two-kernel geometry, first kernel: the POSITION pass alone, one thread per input primitive.  Index fetch, position (cache
// or position-only vertex shader), frustum test, facing / cull, projection, and whether the triangle reaches a tile this rank
// owns - the operations of geometry_main up to the point where a primitive turns out to need set-up - at a quarter of
// geometry_main's registers, so at full occupancy.  Survivors (primitives to clip, and un-culled ones that are binned here) are
// appended to the draw's range of p.surv; cprimitives of the primitives that need no clipping is counted here.
template <int R>
__device__ __forceinline__ void geometry_cull_main(const GeomParams* __restrict__ draws, const GeomBatch& hb) {
  const uint32_t lo = batch_entry_of_cta(hb);
  __shared__ GeomParams s_params;
  stage_geom_params(s_params, draws + hb.draw_of[lo]);
  const GeomParams& p = s_params;
  const uint32_t prim = (blockIdx.x - hb.cta_prefix[lo]) * blockDim.x + threadIdx.x;
  bool survive = false;
  uint32_t n_out = 0;
  if (prim < p.prim_count) {
    uint32_t ids[3];
    if (p.topology == SLV_TOPO_TRIANGLE_LIST) {
      ids[0] = prim * 3; ids[1] = prim * 3 + 1; ids[2] = prim * 3 + 2;
    } else {
      ids[0] = prim; ids[1] = prim + 1; ids[2] = prim + 2;
      if (prim & 1) { uint32_t t = ids[0]; ids[0] = ids[2]; ids[2] = t; }
    }
    uint32_t idx[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) idx[i] = fetch_index(p, ids[i]);
    const bool use_vc = p.vc_pos != nullptr && idx[0] < p.vc_cap && idx[1] < p.vc_cap && idx[2] < p.vc_cap;
    float4 cpos[3];
    if (use_vc) {
#pragma unroll
      for (int i = 0; i < 3; ++i) cpos[i] = __ldg(p.vc_pos + idx[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        VsOut<R> t;
        run_vs<R, true>(p, idx[i], t);
        cpos[i] = t.r[0];
      }
    }
    bool in_frustum = true;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl)
#pragma unroll
      for (int v = 0; v < 3; ++v)
        if (plane_dist(pl, cpos[v]) < 0) in_frustum = false;
    if (!in_frustum) {
      survive = true;  // clipped (or rejected) by geometry_main, which also counts what the clipper emits
    } else {
      float px[3], py[3];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        float iw = 1.0f / cpos[v].w;
        px[v] = cpos[v].x * iw;
        py[v] = cpos[v].y * iw;
      }
      const float area = (px[2] - px[0]) * (py[1] - py[0]) - (py[2] - py[0]) * (px[1] - px[0]);
      const bool front = area > 0.0f;
      if (!cull_tri(p.cull_mode, p.front_ccw, front ? 1.0f : -1.0f)) {
        float4 spos[3];
        spos[0] = viewport_project_pos(p, cpos[0]);
        spos[1] = viewport_project_pos(p, front ? cpos[1] : cpos[2]);
        spos[2] = viewport_project_pos(p, front ? cpos[2] : cpos[1]);
        TriPos ts;
        survive = setup_position<false>(p, spos, ts);
        n_out = 1;
      }
    }
  }
  {  // append the survivors of the warp to the draw's list: one atomic per warp
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, survive);
    uint32_t base = 0;
    if (lane == 0 && bal) base = atomicAdd(p.surv_count + p.draw_id, (uint32_t)__popc(bal));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (survive) p.surv[p.slot_base / 3 + base + __popc(bal & ((1u << lane) - 1u))] = prim;
    const uint32_t total = __popc(__ballot_sync(0xFFFFFFFFu, n_out != 0));
    if (lane == 0 && total) atomicAdd(&p.stats[6], (unsigned long long)total);
  }
}
template <int R>
__global__ void __launch_bounds__(128, 8) k_geometry_cull(const GeomParams* __restrict__ draws, GeomBatch hb) {
  geometry_cull_main<R>(draws, hb);
}

#ifndef SLV_GEOM_CTAS_PER_SM
#define SLV_GEOM_CTAS_PER_SM 4
#endif
template <int R>
__global__ void __launch_bounds__(128, SLV_GEOM_CTAS_PER_SM) k_geometry(const GeomParams* __restrict__ draws, GeomBatch hb) {
  geometry_main<R>(draws, hb);
}

// =====================================================================================================
// binning
// =====================================================================================================
constexpr int SORT_SMEM = 4096;         // list length k_sort_lists sorts in static shared memory
constexpr int SORT_THREADS = 1024;      // CTA of k_sort_lists: one long list, or four short ones (one per SORT_GROUP threads)
constexpr int SORT_GROUP = 256;
constexpr int SORT_LONG = SORT_SMEM / (SORT_THREADS / SORT_GROUP);  // lists longer than this (1024) get a whole CTA
constexpr int SORT_LARGE_SMEM = 49152;  // list length k_sort_lists_large sorts in dynamic shared memory (192 KB)

// single CTA, 1024 threads: exclusive scan of tile_count -> tile_offset[0..n]; zeroes count and cursor
// also compacts the ids of the non-empty tiles into active_tiles[1..] (count in active_tiles[0]) and resets the
// raster work counter.
// Runs once per batch over the counts accumulated by every queued draw's k_geometry.  Also compacts the ids of the
// non-empty tiles into active_tiles[1..] (count in [0]) and resets the raster work-queue head.
__global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t* tile_count, uint32_t* tile_offset, uint32_t* tile_cursor,
                                                     uint32_t n_tiles, uint32_t* active_tiles, uint32_t* work_counter,
                                                     uint32_t* large_tiles, uint32_t* arena_need, uint32_t* surv_count, uint32_t* big_count) {
  __shared__ uint32_t s_active, s_large;
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < MAX_BATCH_DRAWS) surv_count[tid] = 0;  // two-kernel geometry: the survivor counts of the batch have been consumed
  if (tid == 0) big_count[0] = 0;                   // ... and so has the big-triangle queue (k_big_tiles runs ahead of this kernel)
  if (tid == 0) { s_carry = 0; s_active = 0; s_large = 0; work_counter[0] = 0; work_counter[1] = 0; work_counter[2] = 0; work_counter[3] = 0; work_counter[4] = 0; }
  __syncthreads();
  // active-tile compaction, longest lists first (classes >= 2048, >= 512, >= 128, rest): the raster work queue hands
  // items out in this order, so the few very long in-order chains start early instead of forming the kernel's tail
  for (int cls = 0; cls < 4; ++cls) {
    const uint32_t lo = cls == 0 ? (uint32_t)SORT_LONG + 1 : (cls == 1 ? 512u : (cls == 2 ? 128u : 1u));
    const uint32_t hi = cls == 0 ? 0xFFFFFFFFu : (cls == 1 ? (uint32_t)SORT_LONG + 1 : (cls == 2 ? 512u : 128u));
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
      const uint32_t i = base + tid;
      const uint32_t v = i < n_tiles ? tile_count[i] : 0;
      const bool in_cls = v >= lo && v < hi;
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, in_cls);
      uint32_t wbase = 0;
      if (lane == 0 && bal) wbase = atomicAdd(&s_active, (uint32_t)__popc(bal));
      wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
      if (in_cls) active_tiles[1 + wbase + __popc(bal & ((1u << lane) - 1))] = i;
      if (cls == 0 && v > (uint32_t)SORT_SMEM) large_tiles[1 + atomicAdd(&s_large, 1u)] = i;  // sorted by k_sort_lists_large
    }
    __syncthreads();  // classes must not interleave in active_tiles
    if (cls == 0) {  // number of long lists (k_sort_lists gives each a whole CTA)
      if (tid == 0) work_counter[3] = s_active;
      __syncthreads();
    }
  }
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + tid;
    const uint32_t v = i < n_tiles ? tile_count[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = s_warp[lane];
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
        if (lane >= (uint32_t)o) wi += t;
      }
      s_warp[lane] = wi - w;  // exclusive offset of each warp
    }
    __syncthreads();
    const uint32_t excl = s_carry + s_warp[warp] + incl - v;
    if (i < n_tiles) {
      tile_offset[i] = excl;
      tile_cursor[i] = 0;
      tile_count[i] = 0;
    }
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) {
    tile_offset[n_tiles] = s_carry; active_tiles[0] = s_active; large_tiles[0] = s_large;
    atomicMax(arena_need + 1, s_carry);  // tile-list entries this batch needs (the host sizes the arena from it after an overflow)
  }
}

// Tile counting of the triangles k_geometry left in the big-triangle queue: a WARP per triangle, lanes stride over its tile
// range with the reference's tile test (rasterizer.cpp:831-848) - the same counts the setting-up thread would have produced.
__global__ void __launch_bounds__(128) k_big_tiles(const float4* __restrict__ tris, uint32_t tri_stride, uint32_t tiles_x, uint32_t shard_rank,
                                                   uint32_t shard_n, uint32_t* tile_count, const uint32_t* __restrict__ big_slots,
                                                   const uint32_t* __restrict__ big_count) {
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_big = *big_count;
  for (uint32_t i = warp; i < n_big; i += n_warps) {
    const float4* rec = tris + (size_t)big_slots[i] * tri_stride;
    const float4 edge[3] = {__ldg(rec), __ldg(rec + 1), __ldg(rec + 2)};
    const float4 misc = __ldg(rec + 4);
    const uint32_t xr = __float_as_uint(misc.y), yr = __float_as_uint(misc.z);
    const int sx = xr & 0xFFFF, ex = xr >> 16, sy = yr & 0xFFFF, ey = yr >> 16;
    const int ntx = ex - sx, n = ntx * (ey - sy);
    for (int t = (int)lane; t < n; t += 32) {
      const int x = sx + t % ntx, y = sy + t / ntx;
      if (tile_owned(x, y, shard_rank, shard_n) && tile_test(edge, x, y)) atomicAdd(&tile_count[y * tiles_x + x], 1u);
    }
  }
}

// One valid slot per lane; a triangle whose tile range is large is walked by the whole warp (lanes stride over the range) instead
// of by the one lane that owns the slot - on the Sponza-like scene a handful of wall / floor triangles span up to the whole
// 60 x 34 tile grid and their serial loops were the kernel's critical path.  Order inside a tile list is irrelevant here: the
// lists are sorted afterwards.
__device__ __forceinline__ void bin_fill_warp(const BinParams& p, uint32_t slot, bool have) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t xr = 0, yr = 0;
  const float4* rec = p.tris + (size_t)(have ? slot : 0) * p.tri_stride;
  if (have) {
    const float4 misc = __ldg(rec + 4);
    have = (__float_as_uint(misc.x) & 1u) != 0;
    xr = __float_as_uint(misc.y);
    yr = __float_as_uint(misc.z);
  }
  const int sx = xr & 0xFFFF, ex = xr >> 16, sy = yr & 0xFFFF, ey = yr >> 16;
  const int n_tiles = (ex - sx) * (ey - sy);
  const bool big = have && n_tiles > BIG_TILE_RANGE;
  if (have && !big) {
    if (n_tiles == 1) {
      if (tile_owned(sx, sy, p.shard_rank, p.shard_n)) {
        // warp-aggregated cursor bump: one atomic per distinct tile among the converged lanes, consecutive entries for the group
        const uint32_t t = sy * p.tiles_x + sx;
        const uint32_t peers = __match_any_sync(__activemask(), t);
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&p.tile_cursor[t], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint32_t at = p.tile_offset[t] + base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (at < p.list_capacity) p.list[at] = slot << 1;
        else *p.overflow_flag = 1;
      }
    } else {
      const float4 edge[3] = {__ldg(rec), __ldg(rec + 1), __ldg(rec + 2)};
      for (int y = sy; y < ey; ++y)
        for (int x = sx; x < ex; ++x) {
          if (!tile_owned(x, y, p.shard_rank, p.shard_n)) continue;
          const int st = tile_test(edge, x, y);
          if (!st) continue;
          const uint32_t t = y * p.tiles_x + x;
          const uint32_t at = p.tile_offset[t] + atomicAdd(&p.tile_cursor[t], 1u);
          if (at < p.list_capacity) p.list[at] = (slot << 1) | (st == 3 ? 1u : 0u);
          else *p.overflow_flag = 1;
        }
    }
  }
  // the warp's large triangles, one after the other, 32 tiles at a time
  uint32_t todo = __ballot_sync(0xFFFFFFFFu, big);
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1;
    const uint32_t bslot = __shfl_sync(0xFFFFFFFFu, slot, l);
    const uint32_t bxr = __shfl_sync(0xFFFFFFFFu, xr, l), byr = __shfl_sync(0xFFFFFFFFu, yr, l);
    const int bsx = bxr & 0xFFFF, bex = bxr >> 16, bsy = byr & 0xFFFF, bey = byr >> 16;
    const int ntx = bex - bsx, n = ntx * (bey - bsy);
    const float4* brec = p.tris + (size_t)bslot * p.tri_stride;
    const float4 edge[3] = {__ldg(brec), __ldg(brec + 1), __ldg(brec + 2)};
    for (int t = (int)lane; t < n; t += 32) {
      const int x = bsx + t % ntx, y = bsy + t / ntx;
      if (!tile_owned(x, y, p.shard_rank, p.shard_n)) continue;
      const int st = tile_test(edge, x, y);
      if (!st) continue;
      const uint32_t tt = y * p.tiles_x + x;
      const uint32_t at = p.tile_offset[tt] + atomicAdd(&p.tile_cursor[tt], 1u);
      if (at < p.list_capacity) p.list[at] = (bslot << 1) | (st == 3 ? 1u : 0u);
      else *p.overflow_flag = 1;
    }
  }
}
// grid-stride over the compact list of valid slots (warp-uniform trip count): the list holds what this RANK bins (an N-th of
// the frame's triangles on a sort-first rank, far fewer than the slots), so the grid is a few CTAs per SM, not a thread per slot
__global__ void __launch_bounds__(256) k_bin_fill(BinParams p) {
  const uint32_t n_valid = *p.valid_count;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_valid; base += gridDim.x * blockDim.x) {
    const uint32_t vi = base + lane;
    const bool have = vi < n_valid;
    bin_fill_warp(p, have ? p.valid_slots[vi] : 0u, have);
  }
}

// Bitonic network over buf[0..N) (N a power of two, entries past the list padded with 0xFFFFFFFF) by NTH threads that
// share named barrier `bar_id`.  Pair t of a pass sits at i = index with bit j clear; for j <= 32 both ends of every
// pair a warp handles lie in that warp's own 64-entry blocks (the same blocks in every pass), so those passes need
// __syncwarp() only: a 4096-entry list takes 33 CTA-wide barriers instead of 78.
template <int NTH>
__device__ __forceinline__ void bitonic_sort_padded(uint32_t* buf, uint32_t N, uint32_t tid, int bar_id) {
  for (uint32_t k = 2; k <= N; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const bool flip = (j == (k >> 1));
      for (uint32_t t = tid; t < (N >> 1); t += NTH) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const uint32_t l = flip ? (i ^ (k - 1)) : (i | j);
        const uint32_t x = buf[i], y = buf[l];
        if (x > y) { buf[i] = y; buf[l] = x; }
      }
      if (j > 32) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTH) : "memory");
      else __syncwarp();
    }
    if (k >= 64) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTH) : "memory");
  }
}

// generic in-place variant (any n, any buffer): only used for lists that do not fit shared memory
__device__ __forceinline__ void bitonic_sort_block(uint32_t* buf, uint32_t n) {
  uint32_t N = 1;
  while (N < n) N <<= 1;
  for (uint32_t k = 2; k <= N; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      const bool first = (j == (k >> 1));
      for (uint32_t t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
        uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        uint32_t l = first ? (i ^ (k - 1)) : (i | j);
        if (l < n && i < n) {
          uint32_t x = buf[i], y = buf[l];
          if (x > y) { buf[i] = y; buf[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// Restores API order in every tile list (the reference's std::sort, rasterizer.cpp:976).  1024-thread CTAs: the first
// *n_long CTAs take one long list (1025..4096 entries, the head of active_tiles) with all their threads, the others four
// short lists each, one per 256-thread group with its own named barrier - the few long lists no longer form the
// kernel's tail while the short ones keep their low barrier cost.
template <int NTH>
__device__ __forceinline__ void sort_one_list(uint32_t* a, uint32_t n, uint32_t* s, uint32_t tid, int bar_id) {
  uint32_t N = 2;
  while (N < n) N <<= 1;
  for (uint32_t i = tid; i < N; i += NTH) s[i] = i < n ? a[i] : 0xFFFFFFFFu;
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTH) : "memory");
  bitonic_sort_padded<NTH>(s, N, tid, bar_id);
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTH) : "memory");
  for (uint32_t i = tid; i < n; i += NTH) a[i] = s[i];
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_lists(const uint32_t* tile_offset, uint32_t* list, uint32_t capacity,
                                                             const uint32_t* active_tiles, const uint32_t* n_long_ptr) {
  __shared__ uint32_t s[SORT_SMEM];
  const uint32_t n_active = active_tiles[0], n_long = min(*n_long_ptr, n_active);
  if (blockIdx.x < n_long) {
    const uint32_t tile = active_tiles[1 + blockIdx.x];
    uint32_t beg = tile_offset[tile], end = tile_offset[tile + 1];
    if (end > capacity) end = capacity;
    if (beg >= end) return;
    const uint32_t n = end - beg;
    if (n < 2 || n > (uint32_t)SORT_SMEM) return;  // longer lists: k_sort_lists_large
    sort_one_list<SORT_THREADS>(list + beg, n, s, threadIdx.x, 0);
  } else {
    const uint32_t grp = threadIdx.x / SORT_GROUP;
    const uint32_t idx = n_long + (blockIdx.x - n_long) * (SORT_THREADS / SORT_GROUP) + grp;
    if (idx >= n_active) return;
    const uint32_t tile = active_tiles[1 + idx];
    uint32_t beg = tile_offset[tile], end = tile_offset[tile + 1];
    if (end > capacity) end = capacity;
    if (beg >= end) return;
    const uint32_t n = end - beg;
    if (n < 2 || n > (uint32_t)SORT_LONG) return;
    sort_one_list<SORT_GROUP>(list + beg, n, s + grp * SORT_LONG, threadIdx.x % SORT_GROUP, 1 + (int)grp);
  }
}

// the few tiles whose list exceeds SORT_SMEM entries (ids compacted by k_scan_tiles): 1024 threads, dynamic shared memory
__global__ void __launch_bounds__(1024) k_sort_lists_large(const uint32_t* tile_offset, uint32_t* list, uint32_t capacity,
                                                           const uint32_t* large_tiles, uint32_t* valid_count) {
  extern __shared__ uint32_t s_dyn[];
  // last kernel of the binning chain (k_bin_fill, the only reader, is done): reset the valid-slot counter for the next batch
  // built in this scratch set, so that no memset sits at the head of the next frame's chain
  if (blockIdx.x == 0 && threadIdx.x == 0) *valid_count = 0;
  const uint32_t n_large = large_tiles[0];
  for (uint32_t k = blockIdx.x; k < n_large; k += gridDim.x) {
    const uint32_t tile = large_tiles[1 + k];
    uint32_t beg = tile_offset[tile], end = tile_offset[tile + 1];
    if (end > capacity) end = capacity;
    if (beg >= end) continue;
    const uint32_t n = end - beg;
    uint32_t* a = list + beg;
    uint32_t N = 2;
    while (N < n) N <<= 1;
    if (N <= (uint32_t)SORT_LARGE_SMEM) {
      // up to 32768 entries: the padded network of k_sort_lists (passes with partner distance <= 32 are warp-local, about a
      // third of the CTA barriers of the generic one) - these lists sit on the critical path of the front half
      for (uint32_t i = threadIdx.x; i < N; i += blockDim.x) s_dyn[i] = i < n ? a[i] : 0xFFFFFFFFu;
      __syncthreads();
      bitonic_sort_padded<1024>(s_dyn, N, threadIdx.x, 0);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) a[i] = s_dyn[i];
      __syncthreads();
    } else if (n <= (uint32_t)SORT_LARGE_SMEM) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_dyn[i] = a[i];
      __syncthreads();
      bitonic_sort_block(s_dyn, n);
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) a[i] = s_dyn[i];
      __syncthreads();
    } else {
      bitonic_sort_block(a, n);  // in place in global memory
    }
  }
}

// =====================================================================================================
// raster + shade + merge
// =====================================================================================================
__device__ __forceinline__ bool compare_f(uint32_t fn, float l, float r) {  // framebuffer.cpp:96-134
  switch (fn) {
  case SLV_CMP_NEVER: return false;
  case SLV_CMP_LESS: return l < r;
  case SLV_CMP_EQUAL: return l == r;
  case SLV_CMP_LESS_EQUAL: return l <= r;
  case SLV_CMP_GREATER: return l > r;
  case SLV_CMP_NOT_EQUAL: return l != r;
  case SLV_CMP_GREATER_EQUAL: return l >= r;
  default: return true;
  }
}
// branch-free form for the hot early-Z loop: bit 0 = result when l < r, 1 = l == r, 2 = l > r, 3 = unordered
__device__ __forceinline__ uint32_t compare_lut(uint32_t fn) {
  switch (fn) {
  case SLV_CMP_NEVER: return 0x0u;
  case SLV_CMP_LESS: return 0x1u;
  case SLV_CMP_EQUAL: return 0x2u;
  case SLV_CMP_LESS_EQUAL: return 0x3u;
  case SLV_CMP_GREATER: return 0x4u;
  case SLV_CMP_NOT_EQUAL: return 0xDu;
  case SLV_CMP_GREATER_EQUAL: return 0x6u;
  default: return 0xFu;
  }
}
__device__ __forceinline__ bool compare_with_lut(uint32_t lut, float l, float r) {
  const uint32_t idx = (l < r) ? 0u : ((l == r) ? 1u : ((l > r) ? 2u : 3u));
  return (lut >> idx) & 1u;
}
__device__ __forceinline__ bool compare_u(uint32_t fn, uint32_t l, uint32_t r) {
  switch (fn) {
  case SLV_CMP_NEVER: return false;
  case SLV_CMP_LESS: return l < r;
  case SLV_CMP_EQUAL: return l == r;
  case SLV_CMP_LESS_EQUAL: return l <= r;
  case SLV_CMP_GREATER: return l > r;
  case SLV_CMP_NOT_EQUAL: return l != r;
  case SLV_CMP_GREATER_EQUAL: return l >= r;
  default: return true;
  }
}
__device__ __forceinline__ uint32_t stencil_op_apply(uint32_t op, uint32_t ref, uint32_t cur) {  // framebuffer.cpp:136-166
  switch (op) {
  case SLV_SOP_ZERO: return 0;
  case SLV_SOP_REPLACE: return ref;
  case SLV_SOP_INCR_SAT: return min(0xFFu, cur + 1);
  case SLV_SOP_DECR_SAT: return cur - 1;  // max<uint32>(0, cur-1): underflows (Appendix B #5)
  case SLV_SOP_INVERT: return ~cur;
  case SLV_SOP_INCR_WRAP: return (cur + 1) & 0xFF;
  case SLV_SOP_DECR_WRAP: return (cur - 1 + 256) & 0xFF;
  default: return cur;
  }
}

struct TriEntry {  // one surviving triangle of the current chunk, staged in shared memory
  float A[3], B[3], C[3];
  float bbox[4];
  uint32_t slot_flags;  // global slot << 2 | front
  uint32_t draw;        // index of the triangle's draw in the batch
  uint32_t pad;
};

template <int S>
struct SamplePattern;
template <>
struct SamplePattern<1> {
  __device__ static float x(int) { return 0.5f; }
  __device__ static float y(int) { return 0.5f; }
};
template <>
struct SamplePattern<2> {
  __device__ static float x(int s) { return s == 0 ? 0.25f : 0.75f; }
  __device__ static float y(int s) { return s == 0 ? 0.25f : 0.75f; }
};
template <>
struct SamplePattern<4> {  // rasterizer.cpp:1095-1100
  __device__ static float x(int s) { return s == 0 ? 0.375f : (s == 1 ? 0.875f : (s == 2 ? 0.125f : 0.625f)); }
  __device__ static float y(int s) { return s == 0 ? 0.125f : (s == 1 ? 0.375f : (s == 2 ? 0.625f : 0.875f)); }
};

// interpolated attribute register of this pixel: quad stepping (shader.cpp:289-367) or the centroid
// variant (rasterizer.cpp:1366-1397 + shader.cpp:208-255)
__device__ __forceinline__ float4 interp_attr(const float4* rec, int R, int reg, uint32_t mod, float dx, float dy, bool odd_x,
                                              bool odd_y, bool centroid_path, float pdx, float pdy, float inv_w) {
  float4 a0 = __ldg(rec + REC_V0 + 3 * reg);
  float2 rl = lo2(a0), rh = hi2(a0);  // two channels per instruction (FMUL2 / FADD2); every half is the scalar result
  if (!(mod & SLV_AM_NOINTERPOLATION)) {
    const float4 gx = __ldg(rec + REC_DDX + 3 * reg), gy = __ldg(rec + REC_DDY + 3 * reg);
    const float sx = centroid_path ? pdx : dx, sy = centroid_path ? pdy : dy;
    // a0 + (gx * s + gy * t): the sum of the two products in scalar adds (see add2_after_mul), the rest packed
    const float2 pl = mul2(lo2(gx), splat2(sx)), ph = mul2(hi2(gx), splat2(sx));
    const float2 ql = mul2(lo2(gy), splat2(sy)), qh = mul2(hi2(gy), splat2(sy));
    rl = add2(rl, add2_after_mul(pl, ql));
    rh = add2(rh, add2_after_mul(ph, qh));
    if (!centroid_path) {
      if (odd_x) { rl = add2(rl, lo2(gx)); rh = add2(rh, hi2(gx)); }
      if (odd_y) { rl = add2(rl, lo2(gy)); rh = add2(rh, hi2(gy)); }
    }
  }
  if (!(mod & SLV_AM_NOPERSPECTIVE)) { rl = mul2(rl, splat2(inv_w)); rh = mul2(rh, splat2(inv_w)); }
  return cat4(rl, rh);
}

// expf / logf of the host C library, on the device: evaluated in double and rounded once to float, i.e. the correctly
// rounded result, which is what glibc's expf (<= 0.502 ULP) and logf return in all but rare last-bit cases (documented
// tolerance of the programs that use them, tests/cases.py TRANSCENDENTAL_CASES)
__device__ __forceinline__ float exp_f32(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log_f32(float x) { return (float)log((double)x); }

struct PixelCtx {  // what a pixel shader may read (immediate path: the four lanes of a quad run together)
  const float4* rec;
  int R;
  const uint32_t* mods;
  float dx, dy, pdx, pdy, inv_w;
  bool odd_x, odd_y, centroid_path;
  uint32_t quad_base;  // lane of pixel 0 of this quad
  __device__ __forceinline__ float4 attr(int i) const {
    return interp_attr(rec, R, 1 + i, mods[i], dx, dy, odd_x, odd_y, centroid_path, pdx, pdy, inv_w);
  }
  // .xy of attribute register i (value `a` on this lane) at pixels 0, 1, 2 of the quad: shuffled from the quad's lanes
  // (all 32 lanes converge here)
  __device__ __forceinline__ void quad_xy(int, float4 a, float& u0, float& v0, float& u1, float& v1, float& u2, float& v2) const {
    u0 = __shfl_sync(0xFFFFFFFFu, a.x, quad_base); v0 = __shfl_sync(0xFFFFFFFFu, a.y, quad_base);
    u1 = __shfl_sync(0xFFFFFFFFu, a.x, quad_base + 1); v1 = __shfl_sync(0xFFFFFFFFu, a.y, quad_base + 1);
    u2 = __shfl_sync(0xFFFFFFFFu, a.x, quad_base + 2); v2 = __shfl_sync(0xFFFFFFFFu, a.y, quad_base + 2);
  }
  // ... at all four pixels of the quad
  __device__ __forceinline__ void quad_xy4(int, float4 a, float (&u)[4], float (&v)[4]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      u[i] = __shfl_sync(0xFFFFFFFFu, a.x, quad_base + i);
      v[i] = __shfl_sync(0xFFFFFFFFu, a.y, quad_base + i);
    }
  }
  __device__ __forceinline__ int quad_index() const { return (int)((threadIdx.x & 31u) - quad_base); }
};

// cpp_pixel_shader::tex2d (cpp_pixel_shader.cpp:13-31): ddx = q1 - q0, ddy = q2 - q0 for the whole quad,
// LOD once per quad (computed redundantly by the lanes of the quad)
template <class Ctx>
__device__ __forceinline__ float4 ps_tex2d(const SamplerRef& sm, const Ctx& px, int reg, float4 a) {
  float u0, v0, u1, v1, u2, v2;
  px.quad_xy(reg, a, u0, v0, u1, v1, u2, v2);
  float lod = calc_lod_2d(sm, u1 - u0, v1 - v0, u2 - u0, v2 - v0);
  return sample_impl(sm, a.x, a.y, lod, nullptr);
}

// tex2D of a SASL pixel shader: sample_2d_grad with the derivatives of attribute `reg`.xy over the quad
// (sasl/src/codegen/cg_impl.cpp:902-909 -> sampler_api.cpp:13-18).  sasl: ddx = right - left pixel of the pixel's own quad
// ROW, ddy = lower - upper pixel of its COLUMN (cgs_simd.cpp:275-313); else the cpp_pixel_shader convention, q1 - q0 and
// q2 - q0 for all four pixels (cpp_pixel_shader.cpp:13-19).
template <class Ctx>
__device__ __forceinline__ float4 ps_tex2d_grad(const SamplerRef& sm, const Ctx& px, int reg, float4 a, bool sasl) {
  float qu[4], qv[4];
  px.quad_xy4(reg, a, qu, qv);
  const int pi = sasl ? px.quad_index() : 0;
  const bool row1 = pi & 2, col1 = pi & 1;
  const float ulx = row1 ? qu[2] : qu[0], vlx = row1 ? qv[2] : qv[0], uhx = row1 ? qu[3] : qu[1], vhx = row1 ? qv[3] : qv[1];
  const float uly = col1 ? qu[1] : qu[0], vly = col1 ? qv[1] : qv[0], uhy = col1 ? qu[3] : qu[2], vhy = col1 ? qv[3] : qv[2];
  return sample_2d_grad(sm, a.x, a.y, uhx - ulx, vhx - vlx, uhy - uly, vhy - vly, 0.0f);
}

template <int PS, class Ctx>
__device__ __forceinline__ bool run_ps(const RasterParams& p, const Ctx& px, float4& color) {
#ifdef SLV_JIT_PS
  if (PS == SLV_PS_JIT) return slv_jit_ps(p, px, color);
#endif
  if (PS == SLV_PS_ATTR0_COLOR) {
    color = px.attr(0);
    return true;
  }
  if (PS == SLV_PS_DISCARD_ALL) {
    color = px.attr(0);
    return false;
  }
  if (PS == SLV_PS_HEIGHT_COLOR) {  // VertexTextureFetch.cpp:70-113
    const float height = px.attr(0).x;
    const float colors[6][4] = {{0.0f, 0.0f, 0.5f, 1.0f}, {0.7f, 0.6f, 0.0f, 1.0f}, {0.45f, 0.38f, 0.26f, 1.0f},
                                {0.0f, 0.7f, 0.8f, 1.0f}, {0.9f, 0.9f, 1.0f, 1.0f}, {0.9f, 0.9f, 1.0f, 1.0f}};
    const float boundary[6] = {0.0f, 0.62f, 0.75f, 0.88f, 1.0f, 1.0f};
    int lower = -1;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      if (lower == i - 1 && !(height < boundary[i])) lower = i;
    if (lower == -1) {
      color = make_float4(colors[0][0], colors[0][1], colors[0][2], colors[0][3]);
    } else {
      float c0[4], c1[4], lv = 0.0f, hv = 0.0f;
#pragma unroll
      for (int i = 0; i < 5; ++i)
        if (lower == i) {
          lv = boundary[i]; hv = boundary[i + 1];
#pragma unroll
          for (int k = 0; k < 4; ++k) { c0[k] = colors[i][k]; c1[k] = colors[i + 1][k]; }
        }
      const float t = (height - lv) / (hv - lv);
      color = make_float4(c0[0] + (c1[0] - c0[0]) * t, c0[1] + (c1[1] - c0[1]) * t, c0[2] + (c1[2] - c0[2]) * t, c0[3] + (c1[3] - c0[3]) * t);
    }
    return true;
  }
  if (PS == SLV_PS_LIGHTS3) {  // ColorizedTriangle.cpp:55-92
    float4 nrm = px.attr(0), l0 = px.attr(1), l1 = px.attr(2), l2 = px.attr(3);
    float i0 = 1.0f / length3(l0.x, l0.y, l0.z);
    float i1 = 1.0f / length3(l1.x, l1.y, l1.z);
    float i2 = 1.0f / length3(l2.x, l2.y, l2.z);
    float nl = length3(nrm.x, nrm.y, nrm.z);
    if (eq_eps(nl, 0.0f)) nl = 1.0f;
    float ninv = 1.0f / nl;
    float nx = nrm.x * ninv, ny = nrm.y * ninv, nz = nrm.z * ninv;
    float r0 = dot3(nx, ny, nz, l0.x * i0, l0.y * i0, l0.z * i0);
    float r1 = dot3(nx, ny, nz, l1.x * i1, l1.y * i1, l1.z * i1);
    float r2 = dot3(nx, ny, nz, l2.x * i2, l2.y * i2, l2.z * i2);
    const float A[4] = {0.7f, 0.1f, 0.3f, 1.0f}, B[4] = {0.1f, 0.3f, 0.7f, 1.0f}, Cc[4] = {0.3f, 0.7f, 0.1f, 1.0f};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = ((A[k] * r0) * i0) * i0;
      float b = ((B[k] * r1) * i1) * i1;
      float c = ((Cc[k] * r2) * i2) * i2;
      o[k] = clampf((a + b) + c, 0.0f, 1.0f);
    }
    color = make_float4(o[0], o[1], o[2], 1.0f);
    return true;
  }
  if (PS == SLV_PS_TEX_ALPHA) {  // TextureAndBlending.cpp:96-166
    auto u = reinterpret_cast<const slv_ps_tex_alpha_uniforms*>(p.ps_uniforms);
    float4 a = px.attr((int)u->reg);
    color = ps_tex2d(p.sampler0, px, (int)u->reg, a);
    color.w = u->alpha;
    return true;
  }
  if (PS == SLV_PS_TEX_GRAD_ALPHA) {  // SASL tex2D == sample_2d_grad with the quad derivatives
    auto u = reinterpret_cast<const slv_ps_tex_alpha_uniforms*>(p.ps_uniforms);
    float4 a = px.attr((int)u->reg);
    color = ps_tex2d_grad(p.sampler0, px, (int)u->reg, a, u->sasl_derivatives != 0);
    color.w = u->alpha;
    return true;
  }
  if (PS == SLV_PS_SPONZA) {  // Sponza.cpp:117-136
    auto u = reinterpret_cast<const slv_ps_sponza_uniforms*>(p.ps_uniforms);
    float4 diff = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    float4 uv = px.attr(0);
    if (u->has_sampler) diff = ps_tex2d(p.sampler0, px, 0, uv);
    float4 n = px.attr(1), l = px.attr(2);
    float nl = length3(n.x, n.y, n.z);
    if (eq_eps(nl, 0.0f)) nl = 1.0f;
    float ninv = 1.0f / nl;
    float ll = length3(l.x, l.y, l.z);
    if (eq_eps(ll, 0.0f)) ll = 1.0f;
    float linv = 1.0f / ll;
    float illum = clampf(dot3(l.x * linv, l.y * linv, l.z * linv, n.x * ninv, n.y * ninv, n.z * ninv), 0.0f, 1.0f);
    color = make_float4(diff.x * illum, diff.y * illum, diff.z * illum, 1.0f);
    return true;
  }
  if (PS == SLV_PS_SPONZA_GRAD) {  // Sponza.cpp:117-136 with the SASL tex2D fetch (sample_2d_grad, quad derivatives)
    auto u = reinterpret_cast<const slv_ps_sponza_grad_uniforms*>(p.ps_uniforms);
    float4 diff = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    float4 uv = px.attr(0);
    if (u->has_sampler) diff = ps_tex2d_grad(p.sampler0, px, 0, uv, u->sasl_derivatives != 0);
    float4 n = px.attr(1), l = px.attr(2);
    float nl = length3(n.x, n.y, n.z);
    if (eq_eps(nl, 0.0f)) nl = 1.0f;
    float ninv = 1.0f / nl;
    float ll = length3(l.x, l.y, l.z);
    if (eq_eps(ll, 0.0f)) ll = 1.0f;
    float linv = 1.0f / ll;
    float illum = clampf(dot3(l.x * linv, l.y * linv, l.z * linv, n.x * ninv, n.y * ninv, n.z * ninv), 0.0f, 1.0f);
    color = make_float4(diff.x * illum, diff.y * illum, diff.z * illum, 1.0f);
    return true;
  }
  if (PS == SLV_PS_SSM_DRAW) {  // draw_cpp_ps, StandardShadowMap.cpp:62-142
    auto u = reinterpret_cast<const slv_ps_ssm_draw_uniforms*>(p.ps_uniforms);
    const float esm = 25000.0f;
    float occlusion = 0.0f;
    if (u->has_depth_sampler) {
      const float4 a4 = px.attr(4);
      const float lx = a4.x / a4.w, ly = a4.y / a4.w, lz = a4.z / a4.w;  // vec3 / float divides per component
      const float cx = (lx + 1.0f) * 0.5f, cy = 1.0f - (ly + 1.0f) * 0.5f;
      const float off = 1 / 512.0f;
      // nine taps, row-major from (-off, -off); tap 0 is the reference depth of the exponential filter.  One out-of-line
      // sampler (tex2dlod = sampler::sample, cpp_pixel_shader.cpp:33-35) serves all of them.
      const float sd0 = vs_sample_lod(p.sampler1, cx + -off, cy + -off, 0.0f).x;
      float occluder = 0.0f;
#pragma unroll 1
      for (int i = 1; i < 9; ++i) {
        const int ix = i % 3, iy = i / 3;
        const float ox = ix == 0 ? -off : (ix == 1 ? 0.0f : off), oy = iy == 0 ? -off : (iy == 1 ? 0.0f : off);
        const float sd = vs_sample_lod(p.sampler1, cx + ox, cy + oy, 0.0f).x;
        const float gw = i == 4 ? 0.445213f : ((i & 1) ? 0.111014f : 0.027681f);
        occluder += gw * exp_f32(esm * (sd - sd0));
      }
      occluder += 0.027681f;
      occluder = log_f32(occluder);
      occluder += esm * sd0;
      occlusion = clampf(exp_f32(occluder - esm * lz), 0.0f, 1.0f);
    }
    float4 tex = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    if (u->has_tex_sampler) tex = ps_tex2d(p.sampler0, px, 0, px.attr(0));
    const float4 nr = px.attr(1), lr = px.attr(2), er = px.attr(3);
    float nl = length3(nr.x, nr.y, nr.z); if (eq_eps(nl, 0.0f)) nl = 1.0f;
    float ll = length3(lr.x, lr.y, lr.z); if (eq_eps(ll, 0.0f)) ll = 1.0f;
    float el = length3(er.x, er.y, er.z); if (eq_eps(el, 0.0f)) el = 1.0f;
    const float ninv = 1.0f / nl, linv = 1.0f / ll, einv = 1.0f / el;
    const float nx = nr.x * ninv, ny = nr.y * ninv, nz = nr.z * ninv;
    const float Lx = lr.x * linv, Ly = lr.y * linv, Lz = lr.z * linv;
    const float ex = er.x * einv, ey = er.y * einv, ez = er.z * einv;
    const float illum_diffuse = clampf(dot3(Lx, Ly, Lz, nx, ny, nz), 0.0f, 1.0f);
    const float k2 = 2.0f * dot3(Lx, Ly, Lz, nx, ny, nz);  // reflect3: i - n * (2 * dot(i, n))  (eflib/src/math.cpp:85-87)
    const float rx = -(Lx - nx * k2), ry = -(Ly - ny * k2), rz = -(Lz - nz * k2);
    const float illum_specular = clampf(dot3(rx, ry, rz, ex, ey, ez), 0.0f, 1.0f);
    const float sp = (float)pow((double)illum_specular, (double)u->shininess);  // pow(float, int) promotes to double
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      o[k] = u->ambient[k] + (u->diffuse[k] * illum_diffuse + u->specular[k] * sp) * occlusion;
    color = make_float4(tex.x * o[0], tex.y * o[1], tex.z * o[2], 1.0f);
    return true;
  }
  color = make_float4(0, 0, 0, 0);
  return true;
}

// level-4 decision of the reference hierarchy for one 4x4 block of a region (subdivide_tile at the 4-px
// level, rasterizer.cpp:441-602): 0 rejected, 1 partial, 2 full.  (rx, ry) = tile-relative block origin,
// (left_f, top_f) = origin of the 16-px region the block belongs to, (bx, by) = block index inside it.
__device__ __forceinline__ int block_test(const TriEntry& t, float x_min, float x_max, float y_min, float y_max, int rx,
                                          int ry, float left_f, float top_f, int bx, int by) {
  bool rej = (x_min >= (float)(rx + 4)) || (x_max < (float)rx) || (y_min >= (float)(ry + 4)) || (y_max < (float)ry);
  bool acc = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float A = t.A[k], B = t.B[k], C = t.C[k];
    float step_x = TILE * A, step_y = TILE * B;
    float r2a = -fabsf(step_x) - fabsf(step_y);
    float part = (float)((A > 0) * TILE) * A + (float)((B > 0) * TILE) * B;
    step_x *= 0.25f; step_y *= 0.25f; r2a *= 0.25f; part *= 0.25f;
    step_x *= 0.25f; step_y *= 0.25f; r2a *= 0.25f; part *= 0.25f;
    float ev = C - part;
    float ev1 = ev - (left_f * A + top_f * B);
    float step = step_x * (float)bx + step_y * (float)by;
    rej |= (step < ev1);
    acc &= !((step + r2a) < ev1);
  }
  return rej ? 0 : (acc ? 2 : 1);
}

constexpr int RASTER_WARPS = RASTER_THREADS / 32;
#ifndef SLV_RASTER_CTAS_PER_SM
#define SLV_RASTER_CTAS_PER_SM 3
#endif
constexpr int RASTER_CTAS_PER_SM = SLV_RASTER_CTAS_PER_SM;  // resident CTAs per SM the kernel is compiled for
constexpr int QCAP = 32;  // quads a warp may queue per round (a triangle adds at most 8)

// Persistent CTAs: work item = (active tile, 16x16 region).  Per item the tile's sorted triangle list is
// filtered in chunks of 256 entries: one thread per entry evaluates the reference's level-16 decision for the
// region and the level-4 decision of all 16 blocks ONCE, and the survivors are compacted (order preserving) into
// one shared-memory triangle array plus one ordered work list per warp holding only the triangles that touch
// that warp's two 4x4 blocks.  Each warp then walks its own list: thread == pixel, all S samples of
// depth / stencil / colour stay in registers until the item is finished.
template <int S, int PS>
__device__ __forceinline__ void raster_main(const RasterParams& c, const RasterParams* __restrict__ batch, uint32_t n_draws) {
  __shared__ TriEntry s_tri[RASTER_THREADS];
  __shared__ uint16_t s_wlist[RASTER_WARPS][RASTER_THREADS];
  __shared__ uint16_t s_cnt[RASTER_WARPS + 1][RASTER_WARPS];  // [list][filter warp]; list RASTER_WARPS = survivors
  __shared__ uint2 s_items[RASTER_WARPS][QCAP];                // quad queue of each warp
  // the warp's 8x4-pixel framebuffer tile, [sample][pixel == owner lane]: loaded once per work item, merged into by
  // whichever lane shades the pixel, stored once
  __shared__ float s_z[RASTER_WARPS][S][32];
  __shared__ uint32_t s_st[RASTER_WARPS][S][32];
  __shared__ uint32_t s_c[RASTER_WARPS][S][32];
  __shared__ uint32_t s_item;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp = 8x4 pixels = two 4x4 blocks; lanes 4q..4q+3 form the 2x2 quad q
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int q = lane >> 2, pi = lane & 3;
  const int lx = wx + (q & 3) * 2 + (pi & 1), ly = wy + (q >> 2) * 2 + (pi >> 1);  // region-relative
  const int bx = lx >> 2, by = ly >> 2;  // 4x4 block inside the region
  const int ix = lx & 3, iy = ly & 3;    // pixel inside the block
  const uint32_t quad_base = lane & ~3u;
  const uint32_t fullmask = (1u << S) - 1;
  const bool c0_packed = c.color0.data && c.color0.bpp == 4;

  uint32_t n_ps_quads = 0, n_backend_quads = 0;
  uint32_t n_ztest = 0, n_zwrite = 0, n_cwrite = 0, n_cread = 0;  // algorithmic traffic (SURVEY §8d B_frag)
  uint32_t n_scanned = 0, n_surv = 0, n_pairs = 0;                 // work counters

  const uint32_t n_items = c.active_tiles[0] * 16u;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(c.work_counter, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t tile = c.active_tiles[1 + (item >> 4)], sub = item & 15;
    const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
    const int X16 = (sub & 3) * REGION, Y16 = (sub >> 2) * REGION;  // tile-relative origin of this region
    const int gx0 = tile_x * TILE + X16, gy0 = tile_y * TILE + Y16;
    // "Sub tile is out of screen" (rasterizer.cpp:721-724)
    if ((float)gx0 >= (float)c.target_w || (float)gy0 >= (float)c.target_h) continue;
    const int x = gx0 + lx, y = gy0 + ly;
    const bool odd_x = x & 1, odd_y = y & 1;
    const bool in_target = (uint32_t)x < c.target_w && (uint32_t)y < c.target_h;
    const float vpx = (float)(tile_x * TILE), vpy = (float)(tile_y * TILE);

    // ---- the warp's framebuffer tile lives in shared memory (loaded when the first triangle reaches the warp) ----
    bool fb_loaded = false;
    uint32_t wdirty = 0;  // bit 0: depth/stencil modified, bit 1: colour modified (by this lane, for any pixel)

    // ONE list per tile for the whole batch: entries are global triangle slots, sorted = submission order of the
    // draws and API order inside each draw; the per-draw state is looked up through the triangle's draw id.
    const uint32_t list_beg = c.tile_offset[tile];
    uint32_t list_end = c.tile_offset[tile + 1];
    if (list_end > c.list_capacity) list_end = c.list_capacity;
    for (uint32_t chunk = list_beg; chunk < list_end; chunk += RASTER_THREADS) {
      // ================= filter: level-16 decision for the region + level-4 decision of its 16 blocks =================
      const uint32_t ei = chunk + tid;
      bool keep = false;
      uint32_t st_bits = 0;  // 2 bits per block: 0 rejected, 1 partial, 2 full
      TriEntry ent;
      if (ei < list_end) {
        ++n_scanned;
        const uint32_t e = __ldg(c.list + ei);
        const uint32_t slot = e >> 1;
        const float4* rec = c.tris + (size_t)slot * c.tri_stride;
        const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2), bb = __ldg(rec + 3);
        const float4 misc = __ldg(rec + 4);
        const uint32_t flags = __float_as_uint(misc.x);
        ent.draw = __float_as_uint(misc.w);
        ent.A[0] = e0.x; ent.B[0] = e0.y; ent.C[0] = e0.z;
        ent.A[1] = e1.x; ent.B[1] = e1.y; ent.C[1] = e1.z;
        ent.A[2] = e2.x; ent.B[2] = e2.y; ent.C[2] = e2.z;
        ent.bbox[0] = bb.x; ent.bbox[1] = bb.y; ent.bbox[2] = bb.z; ent.bbox[3] = bb.w;
        uint32_t full16;
        const float x_min = bb.x - vpx, x_max = bb.y - vpx, y_min = bb.z - vpy, y_max = bb.w - vpy;
        if (e & 1) {  // the whole 64x64 tile is inside the triangle (rasterizer.cpp:736-743)
          keep = true;
          full16 = 1;
        } else {
          // subdivide_tile at the 16-px level (rasterizer.cpp:441-602, 698-772)
          bool rej = (x_min >= (float)(X16 + REGION)) || (x_max < (float)X16) || (y_min >= (float)(Y16 + REGION)) ||
                     (y_max < (float)Y16);
          bool acc = true;
          const float ftx = (float)(X16 / REGION), fty = (float)(Y16 / REGION);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            float A = ent.A[k], B = ent.B[k], C = ent.C[k];
            float step_x = TILE * A, step_y = TILE * B;
            float r2a = -fabsf(step_x) - fabsf(step_y);
            float part = (float)((A > 0) * TILE) * A + (float)((B > 0) * TILE) * B;
            step_x *= 0.25f; step_y *= 0.25f; r2a *= 0.25f; part *= 0.25f;
            float ev = C - part;
            float ev1 = ev - (vpx * A + vpy * B);
            float step = step_x * ftx + step_y * fty;
            rej |= (step < ev1);
            acc &= !((step + r2a) < ev1);
          }
          keep = !rej;
          full16 = acc ? 1 : 0;
        }
        if (keep) {
          if (full16) {
            st_bits = 0xAAAAAAAAu;  // every block full
          } else {
            const float left_f = (float)gx0, top_f = (float)gy0;
            for (int b = 0; b < 16; ++b) {
              const int bbx = b & 3, bby = b >> 2;
              st_bits |= (uint32_t)block_test(ent, x_min, x_max, y_min, y_max, X16 + bbx * 4, Y16 + bby * 4, left_f, top_f,
                                              bbx, bby) << (2 * b);
            }
            keep = st_bits != 0;
          }
        }
        ent.slot_flags = (slot << 2) | ((flags >> 1) & 1);
      }
      // ---- order-preserving compaction: survivors -> s_tri, and per target warp -> s_wlist[w] ----
      // target warp w owns blocks (by = w >> 1, bx = (w & 1) * 2 + {0, 1}): 4 status bits at 2 * (by * 4 + bx)
      uint32_t hit = 0;  // bit w: this triangle touches warp w
#pragma unroll
      for (int w = 0; w < RASTER_WARPS; ++w) {
        const uint32_t four = (st_bits >> (2 * ((w >> 1) * 4 + (w & 1) * 2))) & 0xFu;
        hit |= (keep && four) ? (1u << w) : 0u;
      }
      uint32_t bal[RASTER_WARPS + 1];
#pragma unroll
      for (int w = 0; w < RASTER_WARPS; ++w) bal[w] = __ballot_sync(0xFFFFFFFFu, (hit >> w) & 1u);
      bal[RASTER_WARPS] = __ballot_sync(0xFFFFFFFFu, keep);
      if (lane <= RASTER_WARPS) {
        uint32_t mine = bal[0];
#pragma unroll
        for (int w = 1; w <= RASTER_WARPS; ++w) mine = (lane == (uint32_t)w) ? bal[w] : mine;
        s_cnt[lane][warp] = (uint16_t)__popc(mine);
      }
      __syncthreads();
      uint32_t my_cnt = 0;  // entries in this warp's list
      {
        const uint32_t below = (1u << lane) - 1;
        uint32_t sbase = 0;
#pragma unroll
        for (int fw = 0; fw < RASTER_WARPS; ++fw) {
          if ((uint32_t)fw < warp) sbase += s_cnt[RASTER_WARPS][fw];
          my_cnt += s_cnt[warp][fw];
        }
        if (keep) {
          ++n_surv;
          const uint32_t sidx = sbase + __popc(bal[RASTER_WARPS] & below);
          s_tri[sidx] = ent;
#pragma unroll
          for (int w = 0; w < RASTER_WARPS; ++w) {
            if ((hit >> w) & 1u) {
              uint32_t wbase = 0;
#pragma unroll
              for (int fw = 0; fw < RASTER_WARPS; ++fw)
                if ((uint32_t)fw < warp) wbase += s_cnt[w][fw];
              const uint32_t four = (st_bits >> (2 * ((w >> 1) * 4 + (w & 1) * 2))) & 0xFu;
              s_wlist[w][wbase + __popc(bal[w] & below)] = (uint16_t)(sidx | (four << 8));
            }
          }
        }
      }
      __syncthreads();

      // ================= per-warp loop over this warp's triangles of the chunk, in API order =================
      // (A) coverage (+ early-Z) by the pixel owners -> quads are pushed to the warp's queue;
      // (B) when the queue is full (or the list ends) 8 queued quads at a time are shaded DENSELY, whatever
      //     triangles they come from, and merged into the warp's depth/stencil/colour tile in shared memory in
      //     queue (= API) order.  No CTA-wide barrier inside this loop.
      if (my_cnt && !fb_loaded) {
        fb_loaded = true;
        if (in_target) {
          if (c.ds.data) {
            const float2* dp = reinterpret_cast<const float2*>(c.ds.data + ((size_t)y * c.ds.w + x) * S * 8);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float2 v = dp[s];
              s_z[warp][s][lane] = v.x;
              s_st[warp][s][lane] = __float_as_uint(v.y);
            }
          }
          if (c0_packed) {
            const uint32_t* cp = reinterpret_cast<const uint32_t*>(c.color0.data + ((size_t)y * c.color0.w + x) * S * 4);
#pragma unroll
            for (int s = 0; s < S; ++s) s_c[warp][s][lane] = cp[s];
          }
        }
        __syncwarp();
      }
      uint32_t wi = 0, qn = 0;
      bool queue_has_late = false;
      if (lane == 0) n_pairs += my_cnt;
      for (;;) {
        const bool have_entry = wi < my_cnt;
        uint32_t we = 0;
        bool entry_early_z = true;
        if (have_entry) {
          we = s_wlist[warp][wi];
          entry_early_z = batch[s_tri[we & 0xFF].draw].early_z != 0;
        }
        const bool need_flush = qn > 0 && (!have_entry || qn + 8 > (uint32_t)QCAP || (entry_early_z && queue_has_late));
        if (need_flush) {
          // ---------------- (B) dense shading + ordered merge of the queued quads ----------------
          for (uint32_t jb = 0; jb < qn; jb += 8) {
            const uint32_t j_raw = jb + (lane >> 2);
            const bool valid = j_raw < qn;
            const uint2 it = s_items[warp][valid ? j_raw : qn - 1];
            const TriEntry& t = s_tri[it.x & 0xFF];
            const RasterParams& p = batch[t.draw];
            const int R = 1 + (int)p.n_attrs;
            const uint32_t oq = (it.x >> 8) & 7;
            const bool quad_full = (it.x >> 11) & 1;
            const uint32_t pm = (it.x >> (16 + 4 * pi)) & 0xF;
            const uint32_t tested = (it.y >> (4 * pi)) & 0xF;
            // the pixel this lane shades: pixel pi of quad oq of this warp (its owner is lane oq*4 + pi)
            const uint32_t own = oq * 4 + pi;
            const int plx = wx + (int)(oq & 3) * 2 + (pi & 1), ply = wy + (int)(oq >> 2) * 2 + (pi >> 1);
            const int sx_ = gx0 + plx, sy_ = gy0 + ply;
            const float4* rec = c.tris + (size_t)(t.slot_flags >> 2) * c.tri_stride;
            // step_2d_unproj_pos_quad (shader.cpp:257-287)
            const float4 v0p = __ldg(rec + REC_V0), gxp = __ldg(rec + REC_DDX), gyp = __ldg(rec + REC_DDY);
            PixelCtx px;
            px.rec = rec; px.R = R; px.mods = p.mods;
            px.dx = 0.5f + (float)(uint32_t)(sx_ & ~1) - v0p.x;
            px.dy = 0.5f + (float)(uint32_t)(sy_ & ~1) - v0p.y;
            px.odd_x = sx_ & 1; px.odd_y = sy_ & 1;
            float pz = v0p.z + (gxp.z * px.dx + gyp.z * px.dy);
            float pw = v0p.w + (gxp.w * px.dx + gyp.w * px.dy);
            if (px.odd_x) { pz += gxp.z; pw += gxp.w; }
            if (px.odd_y) { pz += gyp.z; pw += gyp.w; }
            px.inv_w = 1.0f / pw;
            px.quad_base = quad_base;
            px.centroid_path = p.has_centroid && !quad_full;
            px.pdx = px.dx + (float)(int)px.odd_x;
            px.pdy = px.dy + (float)(int)px.odd_y;
            if (px.centroid_path && pm != fullmask && pm != 0) {
              float cx = 0.0f, cy = 0.0f;
              int n = 0;
#pragma unroll
              for (int s = 0; s < S; ++s)
                if (pm & (1u << s)) { cx += SamplePattern<S>::x(s); cy += SamplePattern<S>::y(s); ++n; }
              float inv = 1 / (float)n;
              cx *= inv; cy *= inv;
              px.pdx += cx - 0.5f;
              px.pdy += cy - 0.5f;
            }
            float4 color;
            const bool keep_px = run_ps<PS>(p, px, color);
            uint32_t fin = (keep_px && valid) ? tested : 0u;
            const uint32_t f0 = __shfl_sync(0xFFFFFFFFu, fin, quad_base), f1 = __shfl_sync(0xFFFFFFFFu, fin, quad_base + 1);
            const uint32_t f2 = __shfl_sync(0xFFFFFFFFu, fin, quad_base + 2), f3 = __shfl_sync(0xFFFFFFFFu, fin, quad_base + 3);
            // draw_full_quad tests the post-PS mask, draw_quad the pre-Z mask (rasterizer.cpp:1311,1409)
            const bool to_backend = valid && (quad_full ? ((f0 | f1 | f2 | f3) != 0) : true);
            if (to_backend && pi == 0) ++n_backend_quads;
            if (!to_backend) fin = 0;
            // two queued quads of this group may be the same screen quad (different triangles): merge those in
            // queue order -> rank = number of earlier group members with the same quad
            uint32_t rank = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint32_t oq_k = __shfl_sync(0xFFFFFFFFu, oq, k * 4);
              const bool valid_k = (jb + k) < qn;
              if (valid_k && oq_k == oq && (uint32_t)k < (lane >> 2)) ++rank;
            }
            uint32_t max_rank = rank;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) max_rank = max(max_rank, __shfl_xor_sync(0xFFFFFFFFu, max_rank, o));
            const bool front = t.slot_flags & 1;
            const bool sx_in = (uint32_t)sx_ < c.target_w && (uint32_t)sy_ < c.target_h;
            for (uint32_t r = 0; r <= max_rank; ++r) {
              if (rank == r && fin && sx_in) {
                // ---- output merger (framebuffer.cpp:445-520) ----
                if (p.early_z && c0_packed && p.bs_program == SLV_BS_REPLACE) {  // the common case, hoisted
                  const uint32_t packed = pack_color(c.color0.fmt, color);
#pragma unroll
                  for (int s = 0; s < S; ++s)
                    if (fin & (1u << s)) s_c[warp][s][own] = packed;
                  n_cwrite += __popc(fin);
                  wdirty |= 2u;
                } else
#pragma unroll
                for (int s = 0; s < S; ++s) {
                  if (!(fin & (1u << s))) continue;
                  if (!p.early_z) {
                    const float aa = (S > 1) ? (SamplePattern<S>::x(s) - 0.5f) * gxp.z + (SamplePattern<S>::y(s) - 0.5f) * gyp.z : 0.0f;
                    const float sd = (S == 1) ? pz : pz + aa;
                    const float od = p.read_depth ? s_z[warp][s][own] : 0.0f;
                    const uint32_t os = p.stencil_enable ? (s_st[warp][s][own] & p.read_mask) : 0u;
                    const bool dp = p.depth_enable ? compare_f(p.depth_func, sd, od) : true;
                    n_ztest += (p.read_depth | p.stencil_enable) ? 1u : 0u;
                    const slv_stencil_op_desc& face = front ? p.front_face : p.back_face;
                    const bool sp = p.stencil_enable ? compare_u(face.stencil_func, p.stencil_ref, os) : true;
                    if (!(dp && sp)) continue;
                    const uint32_t ns = p.stencil_enable ? stencil_op_apply(face.stencil_pass_op, p.stencil_ref, os) : os;
                    if (p.write_depth) s_z[warp][s][own] = sd;
                    if (p.stencil_enable) s_st[warp][s][own] = ns & p.write_mask;
                    if (p.write_depth | p.stencil_enable) { ++n_zwrite; wdirty |= 1u; }
                  }
                  // blend shader
                  if (c.color0.data) {
                    ++n_cwrite;
                    n_cread += (p.bs_program == SLV_BS_LERP_SRC_ALPHA) ? 1u : 0u;
                    if (c0_packed) {
                      if (p.bs_program == SLV_BS_LERP_SRC_ALPHA) {
                        const float4 d = unpack_color(c.color0.fmt, s_c[warp][s][own]);
                        const float4 rr = make_float4(d.x + (color.x - d.x) * color.w, d.y + (color.y - d.y) * color.w,
                                                      d.z + (color.z - d.z) * color.w, d.w + (color.w - d.w) * color.w);
                        s_c[warp][s][own] = pack_color(c.color0.fmt, rr);
                      } else {
                        s_c[warp][s][own] = pack_color(c.color0.fmt, color);
                      }
                      wdirty |= 2u;
                    } else {
                      uint8_t* cp = c.color0.data + (((size_t)sy_ * c.color0.w + sx_) * S + s) * c.color0.bpp;
                      if (p.bs_program == SLV_BS_LERP_SRC_ALPHA) {
                        const float4 d = load_texel_rgba32f(c.color0.fmt, cp);
                        const float4 rr = make_float4(d.x + (color.x - d.x) * color.w, d.y + (color.y - d.y) * color.w,
                                                      d.z + (color.z - d.z) * color.w, d.w + (color.w - d.w) * color.w);
                        store_texel_rgba32f(c.color0.fmt, cp, rr);
                      } else {
                        store_texel_rgba32f(c.color0.fmt, cp, color);
                      }
                    }
                  }
                  if (p.bs_program == SLV_BS_REPLACE_AND_COUNT && c.color1.data) {
                    uint8_t* cp = c.color1.data + (((size_t)sy_ * c.color1.w + sx_) * S + s) * c.color1.bpp;
                    float4 v = load_texel_rgba32f(c.color1.fmt, cp);
                    v.x += 1.0f;
                    store_texel_rgba32f(c.color1.fmt, cp, v);
                  }
                }
              }
              __syncwarp();
            }
          }
          qn = 0;
          queue_has_late = false;
          continue;
        }
        if (!have_entry) break;
        ++wi;
        // ---------------- (A) coverage + early-Z of the next triangle of this warp's list ----------------
        const TriEntry& t = s_tri[we & 0xFF];
        const RasterParams& p = batch[t.draw];
        const int R = 1 + (int)p.n_attrs;
        const int blk = (we >> (8 + 2 * (bx & 1))) & 3;  // 0 rejected, 1 partial, 2 full
        // per-sample coverage (draw_partial_tile, rasterizer.cpp:298-439)
        uint32_t pm = 0;
        if (in_target) {
          if (blk == 2) {
            pm = fullmask;
          } else if (blk == 1) {
            const float left_f = (float)(gx0 + bx * 4), top_f = (float)(gy0 + by * 4);
            float ev[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) ev[k] = t.C[k] - (left_f * t.A[k] + top_f * t.B[k]);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              float fx = SamplePattern<S>::x(s) + (float)ix, fy = SamplePattern<S>::y(s) + (float)iy;
              bool rj = false;
#pragma unroll
              for (int k = 0; k < 3; ++k) rj |= (fx * t.A[k] + fy * t.B[k]) < ev[k];
              if (!rj) pm |= 1u << s;
            }
          }
        }
        if (!__any_sync(0xFFFFFFFFu, pm != 0)) continue;
        // early-Z: test and WRITE depth now (framebuffer.cpp:522-614; Appendix B #3)
        uint32_t tested = pm;
        if (entry_early_z) {
          tested = 0;
          if (pm) {
            const float4* rec = c.tris + (size_t)(t.slot_flags >> 2) * c.tri_stride;
            const float4 v0p = __ldg(rec + REC_V0), gxp = __ldg(rec + REC_DDX), gyp = __ldg(rec + REC_DDY);
            const float dx = 0.5f + (float)(uint32_t)(x & ~1) - v0p.x;
            const float dy = 0.5f + (float)(uint32_t)(y & ~1) - v0p.y;
            float depth = v0p.z + (gxp.z * dx + gyp.z * dy);
            if (odd_x) depth += gxp.z;
            if (odd_y) depth += gyp.z;
            const uint32_t cmp_lut = p.depth_enable ? compare_lut(p.depth_func) : 0xFu;
#pragma unroll
            for (int s = 0; s < S; ++s) {
              if (pm & (1u << s)) {
                const float aa = (S > 1) ? (SamplePattern<S>::x(s) - 0.5f) * gxp.z + (SamplePattern<S>::y(s) - 0.5f) * gyp.z : 0.0f;
                const float nd = (S == 1) ? depth : aa + depth;
                const float od = p.read_depth ? s_z[warp][s][lane] : 0.0f;
                const bool pass = compare_with_lut(cmp_lut, nd, od);
                n_ztest += p.read_depth;
                if (pass) {
                  tested |= 1u << s;
                  if (p.write_depth) { s_z[warp][s][lane] = nd; wdirty |= 1u; ++n_zwrite; }
                }
              }
            }
          }
        }
        // quad assembly: one queue item per quad that still has a live sample
        const uint32_t m0 = __shfl_sync(0xFFFFFFFFu, pm, quad_base), m1 = __shfl_sync(0xFFFFFFFFu, pm, quad_base + 1);
        const uint32_t m2 = __shfl_sync(0xFFFFFFFFu, pm, quad_base + 2), m3 = __shfl_sync(0xFFFFFFFFu, pm, quad_base + 3);
        const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, tested, quad_base), t1 = __shfl_sync(0xFFFFFFFFu, tested, quad_base + 1);
        const uint32_t t2 = __shfl_sync(0xFFFFFFFFu, tested, quad_base + 2), t3 = __shfl_sync(0xFFFFFFFFu, tested, quad_base + 3);
        const bool quad_shade = ((m0 | m1 | m2 | m3) != 0) && ((t0 | t1 | t2 | t3) != 0);
        const uint32_t qbal = __ballot_sync(0xFFFFFFFFu, quad_shade && pi == 0);
        if (quad_shade && pi == 0) {
          const uint32_t quad_full = ((m0 & m1 & m2 & m3) == fullmask) ? 1u : 0u;
          uint2 it;
          it.x = (we & 0xFF) | ((uint32_t)q << 8) | (quad_full << 11) | ((m0 | (m1 << 4) | (m2 << 8) | (m3 << 12)) << 16);
          it.y = t0 | (t1 << 4) | (t2 << 8) | (t3 << 12);
          s_items[warp][qn + __popc(qbal & ((1u << lane) - 1))] = it;
        }
        const uint32_t pushed = __popc(qbal);
        if (lane == 0) n_ps_quads += pushed;
        qn += pushed;
        if (pushed && !entry_early_z) queue_has_late = true;
        __syncwarp();
      }
      // (the barrier at the top of the next chunk / item protects s_tri, s_wlist and s_cnt)
      __syncthreads();
    }

    // ---- write the warp's tile back once, 128-bit stores at 4x MSAA ----
    if (fb_loaded) {
      __syncwarp();
      const bool any_ds = __any_sync(0xFFFFFFFFu, (wdirty & 1u) != 0), any_c = __any_sync(0xFFFFFFFFu, (wdirty & 2u) != 0);
      if (in_target) {
        if (any_ds && c.ds.data) {
          uint8_t* ds_ptr = c.ds.data + ((size_t)y * c.ds.w + x) * S * 8;
          if (S == 4) {
            *reinterpret_cast<float4*>(ds_ptr) = make_float4(s_z[warp][0][lane], __uint_as_float(s_st[warp][0][lane]),
                                                             s_z[warp][1 % S][lane], __uint_as_float(s_st[warp][1 % S][lane]));
            *reinterpret_cast<float4*>(ds_ptr + 16) = make_float4(s_z[warp][2 % S][lane], __uint_as_float(s_st[warp][2 % S][lane]),
                                                                  s_z[warp][3 % S][lane], __uint_as_float(s_st[warp][3 % S][lane]));
          } else if (S == 2) {
            *reinterpret_cast<float4*>(ds_ptr) = make_float4(s_z[warp][0][lane], __uint_as_float(s_st[warp][0][lane]),
                                                             s_z[warp][1 % S][lane], __uint_as_float(s_st[warp][1 % S][lane]));
          } else {
            *reinterpret_cast<float2*>(ds_ptr) = make_float2(s_z[warp][0][lane], __uint_as_float(s_st[warp][0][lane]));
          }
        }
        if (any_c && c0_packed) {
          uint8_t* c_ptr = c.color0.data + ((size_t)y * c.color0.w + x) * S * 4;
          if (S == 4) *reinterpret_cast<uint4*>(c_ptr) = make_uint4(s_c[warp][0][lane], s_c[warp][1 % S][lane], s_c[warp][2 % S][lane], s_c[warp][3 % S][lane]);
          else if (S == 2) *reinterpret_cast<uint2*>(c_ptr) = make_uint2(s_c[warp][0][lane], s_c[warp][1 % S][lane]);
          else *reinterpret_cast<uint32_t*>(c_ptr) = s_c[warp][0][lane];
        }
      }
      __syncwarp();
    }
  }

  // ---- statistics: ps_invocations / backend_input_pixels count 4 per quad ----
  uint32_t a = n_ps_quads, b = n_backend_quads;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
    n_ztest += __shfl_xor_sync(0xFFFFFFFFu, n_ztest, o);
    n_zwrite += __shfl_xor_sync(0xFFFFFFFFu, n_zwrite, o);
    n_cwrite += __shfl_xor_sync(0xFFFFFFFFu, n_cwrite, o);
    n_cread += __shfl_xor_sync(0xFFFFFFFFu, n_cread, o);
    n_scanned += __shfl_xor_sync(0xFFFFFFFFu, n_scanned, o);
    n_surv += __shfl_xor_sync(0xFFFFFFFFu, n_surv, o);
    n_pairs += __shfl_xor_sync(0xFFFFFFFFu, n_pairs, o);
  }
  if (lane == 0) {
    if (a) atomicAdd(&c.stats[7], (unsigned long long)a * 4ull);
    if (a) atomicAdd(&c.stats[16], (unsigned long long)a * 4ull);  // the immediate path shades every quad it counts
    if (b) atomicAdd(&c.stats[8], (unsigned long long)b * 4ull);
    if (n_ztest) atomicAdd(&c.stats[9], (unsigned long long)n_ztest);
    if (n_zwrite) atomicAdd(&c.stats[10], (unsigned long long)n_zwrite);
    if (n_cwrite) atomicAdd(&c.stats[11], (unsigned long long)n_cwrite);
    if (n_cread) atomicAdd(&c.stats[12], (unsigned long long)n_cread);
    if (n_scanned) atomicAdd(&c.stats[13], (unsigned long long)n_scanned);
    if (n_surv) atomicAdd(&c.stats[14], (unsigned long long)n_surv);
    if (n_pairs) atomicAdd(&c.stats[15], (unsigned long long)n_pairs);
  }
}

template <int S, int PS>
__global__ void __launch_bounds__(RASTER_THREADS, RASTER_CTAS_PER_SM) k_raster(RasterParams c, const RasterParams* __restrict__ batch,
                                                               uint32_t n_draws) {
  raster_main<S, PS>(c, batch, n_draws);
}

// =====================================================================================================
// clears / resolve / mip generation / sampler probe
// =====================================================================================================
// surface::fill (surface.cpp:170-271): every texel = one converted pattern; 128-bit stores
__global__ void k_fill(uint4* dst, size_t n_vec, uint4 pattern) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n_vec; i += stride) dst[i] = pattern;
}
__global__ void k_fill_words(uint32_t* dst, size_t first_word, size_t n_words, uint4 pattern) {
  size_t i = first_word + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_words) dst[i] = (i & 3) == 0 ? pattern.x : ((i & 3) == 1 ? pattern.y : ((i & 3) == 2 ? pattern.z : pattern.w));
}
// one 64x64 tile of a surface <- pattern (all threads of the CTA; 128-bit stores along the tile's rows)
__device__ __forceinline__ void fill_tile(const SurfaceRef& s, uint4 pattern, uint32_t tx, uint32_t ty) {
  const uint32_t texel = s.samples * s.bpp;                      // bytes per pixel (all samples)
  const uint32_t x0 = tx * TILE, y0 = ty * TILE;
  if (x0 >= s.w || y0 >= s.h) return;
  const uint32_t cols = min((uint32_t)TILE, s.w - x0), rows = min((uint32_t)TILE, s.h - y0);
  const uint32_t row_bytes = cols * texel;                       // multiple of 4; 16-byte aligned when texel*x0 is
  for (uint32_t r = 0; r < rows; ++r) {
    uint8_t* row = s.data + ((size_t)(y0 + r) * s.w + x0) * texel;
    if (((reinterpret_cast<uintptr_t>(row) | row_bytes) & 15) == 0) {
      for (uint32_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(row)[i] = pattern;
    } else {  // rows that are not 16-byte aligned: word stores, pattern word chosen by absolute position
      for (uint32_t i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) {
        const uint32_t wabs = (uint32_t)(((row - s.data) / 4 + i) & 3);
        reinterpret_cast<uint32_t*>(row)[i] = wabs == 0 ? pattern.x : (wabs == 1 ? pattern.y : (wabs == 2 ? pattern.z : pattern.w));
      }
    }
  }
}

// sort-first: a rank only ever reads and writes its own 64x64 tiles, so it only clears those (CTA per owned tile)
__global__ void __launch_bounds__(256) k_fill_tiles(SurfaceRef s, uint4 pattern, uint32_t tiles_x, uint32_t rank, uint32_t n) {
  const uint32_t tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  if (!tile_owned(tx, ty, rank, n)) return;
  fill_tile(s, pattern, tx, ty);
}

// (k_inactive_tiles, after k_resolve below, handles the tiles a lazily cleared / fused-resolve batch never visits)

// framebuffer::clear_depth_stencil with a single flag (framebuffer.cpp:616-644)
__global__ void k_clear_ds_partial(float2* dst, size_t n, uint32_t flags, float depth, uint32_t stencil) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float2 v = dst[i];
    if (flags & SLV_CLEAR_DEPTH) v.x = depth;
    if (flags & SLV_CLEAR_STENCIL) v.y = __uint_as_float(stencil);
    dst[i] = v;
  }
}

// ---- sort-first frame assembly over peer memory (NVLink / NVSwitch): flags that order one rank's stream after
// another rank's, without the host.  A rank resolves its tiles straight into the root's surface (k_resolve with a peer
// destination), then raises its flag there; the root's stream waits for every rank's flag.
__global__ void k_peer_signal(uint32_t* flag, uint32_t value) {
  __threadfence_system();  // the preceding kernels' peer stores are performed before the flag becomes visible
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// one warp; lane i polls flags[first + i] until it reaches `value`.  Gives up after ~10 s (a peer died): *err = 2, which the
// next flush point reports, instead of hanging the stream forever.
__global__ void __launch_bounds__(32) k_flags_wait(const uint32_t* flags, uint32_t first, uint32_t count, uint32_t value, uint32_t* err) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t i = threadIdx.x; i < count; i += 32) {
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + first + i) : "memory");
      if ((int32_t)(v - value) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { *err = 2; break; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

// surface::resolve (surface.cpp:123-140): sum of to_rgba32f(sample) in sample order, * (1/S), convert (RNE)
__device__ __forceinline__ void resolve_pixel(const SurfaceRef& src, const SurfaceRef& dst, uint32_t x, uint32_t y) {
  float4 clr = make_float4(0, 0, 0, 0);
  const uint8_t* sp = src.data + ((size_t)y * src.w + x) * src.samples * src.bpp;
  const float inv = 1 / (float)src.samples;
  if (src.bpp == 4 && src.samples == 4) {  // one 128-bit load per pixel
    const uint4 v = *reinterpret_cast<const uint4*>(sp);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float4 t = unpack_color(src.fmt, w[s]);
      clr.x += t.x; clr.y += t.y; clr.z += t.z; clr.w += t.w;
    }
  } else {
    for (uint32_t s = 0; s < src.samples; ++s) {
      float4 t = load_texel_rgba32f(src.fmt, sp + (size_t)s * src.bpp);
      clr.x += t.x; clr.y += t.y; clr.z += t.z; clr.w += t.w;
    }
  }
  clr.x *= inv; clr.y *= inv; clr.z *= inv; clr.w *= inv;
  store_texel_rgba32f(dst.fmt, dst.data + ((size_t)y * dst.w + x) * dst.bpp, clr);
}

__global__ void k_resolve(SurfaceRef src, SurfaceRef dst, uint32_t rank, uint32_t n) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.w || y >= src.h) return;
  if (!tile_owned(x / TILE, y / TILE, rank, n)) return;  // sort-first: other ranks resolve their own tiles
  resolve_pixel(src, dst, x, y);
}

// Lazy clears / fused resolve: the batch's k_cover / k_shade write (and resolve) every pixel of the tiles that have
// triangles; the owned tiles WITHOUT triangles (empty list: tile_offset[t + 1] == tile_offset[t]) receive the clear values
// and are resolved here, one CTA per tile (CTAs of the other tiles exit at once).  color / ds / resolve_dst .data ==
// nullptr: nothing to do for that surface.  Independent of k_cover / k_shade (disjoint tiles): runs after them.
__global__ void __launch_bounds__(256) k_inactive_tiles(SurfaceRef color, uint32_t fill_color, uint4 color_pattern, SurfaceRef ds,
                                                        uint4 ds_pattern, SurfaceRef resolve_dst,
                                                        const uint32_t* __restrict__ tile_offset, uint32_t tiles_x, uint32_t rank,
                                                        uint32_t n) {
  const uint32_t t = blockIdx.x;
  const uint32_t tx = t % tiles_x, ty = t / tiles_x;
  if (!tile_owned(tx, ty, rank, n) || tile_offset[t + 1] != tile_offset[t]) return;
  if (fill_color) fill_tile(color, color_pattern, tx, ty);
  if (ds.data) fill_tile(ds, ds_pattern, tx, ty);
  if (resolve_dst.data) {
    __syncthreads();  // the colour fill above is read back below (same CTA)
    for (uint32_t i = threadIdx.x; i < TILE * TILE; i += blockDim.x) {
      const uint32_t x = tx * TILE + (i % TILE), y = ty * TILE + (i / TILE);
      if (x < color.w && y < color.h) resolve_pixel(color, resolve_dst, x, y);
    }
  }
}

// surface::make_mip_surface (surface.cpp:53-92): box filter ((c0+c1)+c2)+c3 then *0.25; reads of texel
// 2x+1 / 2y+1 are not bounds-checked upstream: an x overflow wraps into the next row (mirrored), a read
// past the allocation is undefined upstream and returns zeros here.
__global__ void k_mipgen(SurfaceRef src, SurfaceRef dst, uint32_t filter) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dst.w || y >= dst.h) return;
  for (uint32_t s = 0; s < src.samples; ++s) {
    auto rd = [&](uint32_t xx, uint32_t yy) {
      size_t off = (((size_t)yy * src.w + xx) * src.samples + s) * src.bpp;
      if (off + src.bpp > src.bytes) return make_float4(0, 0, 0, 0);
      return load_texel_rgba32f(src.fmt, src.data + off);
    };
    float4 o;
    if (filter == SLV_FILTER_POINT) {
      o = rd(x * 2, y * 2);
    } else {
      float4 c0 = rd(x * 2, y * 2), c1 = rd(x * 2 + 1, y * 2), c2 = rd(x * 2, y * 2 + 1), c3 = rd(x * 2 + 1, y * 2 + 1);
      o = make_float4((((c0.x + c1.x) + c2.x) + c3.x) * 0.25f, (((c0.y + c1.y) + c2.y) + c3.y) * 0.25f,
                      (((c0.z + c1.z) + c2.z) + c3.z) * 0.25f, (((c0.w + c1.w) + c2.w) + c3.w) * 0.25f);
    }
    store_texel_rgba32f(dst.fmt, dst.data + (((size_t)y * dst.w + x) * dst.samples + s) * dst.bpp, o);
  }
}

// sort-first gather helpers: owned 64x64 tiles of a single-sampled surface <-> dense staging buffer
__global__ void k_pack_tiles(SurfaceRef s, uint32_t tiles_x, uint32_t tiles_y, uint32_t rank, uint32_t n, uint8_t* staging,
                             const uint32_t* tile_slot, int unpack) {
  uint32_t tile = blockIdx.x;
  uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  if (!tile_owned(tx, ty, rank, n)) return;
  uint32_t slot = tile_slot[tile];
  uint32_t words_per_texel = s.bpp / 4;
  uint32_t row_words = TILE * words_per_texel;
  for (uint32_t i = threadIdx.x; i < TILE * row_words; i += blockDim.x) {
    uint32_t row = i / row_words, w = i % row_words;
    uint32_t x = tx * TILE + w / words_per_texel, y = ty * TILE + row;
    if (x >= s.w || y >= s.h) continue;
    uint32_t* g = reinterpret_cast<uint32_t*>(s.data + ((size_t)y * s.w + x) * s.bpp) + (w % words_per_texel);
    uint32_t* st = reinterpret_cast<uint32_t*>(staging) + (size_t)slot * TILE * row_words + i;
    if (unpack) *g = *st; else *st = *g;
  }
}

// sort-first, end to end: the owned 64x64 tiles of a single-sampled surface -> a host frame of the same linear layout (mapped
// pinned memory).  A tile row is 64 px * bpp contiguous bytes written with 128-bit stores by consecutive lanes, i.e. full-size
// PCIe write transactions.  The kernel is bound by the host link (~50 GB/s), not by the SMs: a SMALL persistent grid walks the
// tiles, so that the stores in flight saturate the link while the render kernels of the next frame keep (almost) every SM slot -
// one CTA per tile would park a thousand CTAs behind the link's back-pressure for the whole copy.
__global__ void __launch_bounds__(256) k_export_tiles(SurfaceRef s, uint32_t tiles_x, uint32_t n_tiles, uint32_t rank, uint32_t n, uint8_t* host) {
  const size_t pitch = (size_t)s.w * s.bpp;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    if (!tile_owned(tx, ty, rank, n)) continue;
    const uint32_t x0 = tx * TILE, y0 = ty * TILE;
    const uint32_t w = min((uint32_t)TILE, s.w - x0), h = min((uint32_t)TILE, s.h - y0);
    const uint32_t row_bytes = w * s.bpp;
    if (((pitch | ((size_t)x0 * s.bpp) | row_bytes) & 15) == 0) {
      const uint32_t vec_per_row = row_bytes / 16;
      for (uint32_t i = threadIdx.x; i < h * vec_per_row; i += blockDim.x) {
        const uint32_t r = i / vec_per_row, v = i % vec_per_row;
        const size_t off = (size_t)(y0 + r) * pitch + (size_t)x0 * s.bpp + (size_t)v * 16;
        *reinterpret_cast<uint4*>(host + off) = *reinterpret_cast<const uint4*>(s.data + off);
      }
    } else {
      const uint32_t words_per_row = row_bytes / 4;
      for (uint32_t i = threadIdx.x; i < h * words_per_row; i += blockDim.x) {
        const uint32_t r = i / words_per_row, v = i % words_per_row;
        const size_t off = (size_t)(y0 + r) * pitch + (size_t)x0 * s.bpp + (size_t)v * 4;
        *reinterpret_cast<uint32_t*>(host + off) = *reinterpret_cast<const uint32_t*>(s.data + off);
      }
    }
  }
}

__global__ void k_sampler_probe(SamplerRef sm, uint32_t n, const float* coords, const float* ddx, const float* ddy,
                                const float* lod, uint32_t use_lod, float4* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 c;
  if (use_lod) {
    bool aniso = sm.d.mip_filter == SLV_FILTER_ANISOTROPIC;
    AfInfo af = {0, 1, 0, 0, 0};
    c = sample_impl(sm, coords[2 * i], coords[2 * i + 1], lod[i], aniso ? &af : nullptr);
  } else {
    c = sample_2d_grad(sm, coords[2 * i], coords[2 * i + 1], ddx[2 * i], ddx[2 * i + 1], ddy[2 * i], ddy[2 * i + 1], 0.0f);
  }
  out[i] = c;
}

}  // namespace slv
