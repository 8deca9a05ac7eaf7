// slv_sampler.cuh — the texture sampler on the device: address modes, point / bilinear in the native texel
// format, mip selection (three LOD qualities) kept in registers, trilinear and anisotropic (EWA) probes.
// Follows salvia/src/resource/sampler.cpp op for op; rgba8 texels of a
// bilinear footprint are fetched as 32-bit words and expanded in registers, float formats with 64/128-bit
// vector loads.
#pragma once

#include "slv_common.cuh"

namespace slv {

__constant__ float c_ewa_wts[256] = {
#include "ewa_weights.inc"
};

__device__ __forceinline__ float trunc_f(float x) { return (float)(int)x; }  // cvttps2dq ; cvtdq2ps
__device__ __forceinline__ float floor_fix(float x) {                          // sampler.cpp:36-40
  float ip = trunc_f(x);
  if (ip > x) ip = ip - 1.0f;
  return ip;
}
__device__ __forceinline__ int wrap_index(float ipart_plus, float fsize) {  // sampler.cpp:42-44, 90-92
  float v = ipart_plus + fsize * 8192.0f;
  float dv = trunc_f(v / fsize);
  return (int)(v - dv * fsize);
}
__device__ __forceinline__ float clamp01_idx(float v, float hi) {  // _mm_min_ps(_mm_max_ps(v, 0), hi)
  float m = v > 0.0f ? v : 0.0f;
  return m < hi ? m : hi;
}

// same addresser on u and v: the SIMD branches of addresser::*::do_coordi_point_2d
__device__ inline int point_coord_2d(uint32_t mode, float c, int size) {
  float fs = (float)size;
  switch (mode) {
  case SLV_ADDR_WRAP: {
    float f = c - trunc_f(c);
    f = fs * f;
    return wrap_index(floor_fix(f), fs);
  }
  case SLV_ADDR_MIRROR: {
    int sel = fast_floori((double)c);
    float o = ((sel & 1) ? (float)(1 + sel) - c : c - (float)sel) * fs;
    return (int)clamp01_idx(floor_fix(o), fs - 1.0f);
  }
  case SLV_ADDR_CLAMP: {
    float o = c * fs;
    o = o > 0.5f ? o : 0.5f;
    float hi = fs - 0.5f;
    o = o < hi ? o : hi;
    return (int)clamp01_idx(floor_fix(o), fs - 1.0f);
  }
  default: {  // border
    float o = c * fs;
    o = o > -0.5f ? o : -0.5f;
    float hi = fs - (-0.5f);
    o = o < hi ? o : hi;
    int ip = (int)floor_fix(o);
    return ip >= size ? -1 : ip;
  }
  }
}

// `mask` = size-1 when size is a power of two <= 1024: then ip + size*8192 < 2^24 and every step of the reference's
// float modulo (sampler.cpp:88-92) is exact, so it equals the integer (ip + k) & mask bit for bit.
__device__ inline void linear_coord_2d(uint32_t mode, float c, int size, uint32_t mask, int& lo, int& up, float& frac) {
  float fs = (float)size;
  switch (mode) {
  case SLV_ADDR_WRAP: {
    float f = c - trunc_f(c);
    f = fs * f;
    f = f - 0.5f;
    float ip = floor_fix(f);
    frac = f - ip;
    if (mask) {
      const int i = (int)ip;
      lo = i & (int)mask;
      up = (i + 1) & (int)mask;
    } else {
      lo = wrap_index(ip + 0.0f, fs);
      up = wrap_index(ip + 1.0f, fs);
    }
    return;
  }
  case SLV_ADDR_MIRROR: {
    int sel = fast_floori((double)c);
    float o = ((sel & 1) ? (float)(1 + sel) - c : c - (float)sel) * fs - 0.5f;
    float ip = floor_fix(o);
    frac = o - ip;
    lo = (int)clamp01_idx(ip + 0.0f, fs - 1.0f);
    up = (int)clamp01_idx(ip + 1.0f, fs - 1.0f);
    return;
  }
  case SLV_ADDR_CLAMP: {
    float o = c * fs;
    o = o > 0.5f ? o : 0.5f;
    float hi = fs - 0.5f;
    o = o < hi ? o : hi;
    o = o - 0.5f;
    float ip = floor_fix(o);
    frac = o - ip;
    lo = (int)clamp01_idx(ip + 0.0f, fs - 1.0f);
    up = (int)clamp01_idx(ip + 1.0f, fs - 1.0f);
    return;
  }
  default: {  // border: the reference reads texel -1 (undefined); indices are clamped at the fetch
    float o = c * fs;
    o = o > -0.5f ? o : -0.5f;
    float hi = fs - (-0.5f);
    o = o < hi ? o : hi;
    o = o + (-0.5f);
    float ip = floor_fix(o);
    frac = o - ip;
    int i = (int)ip;
    lo = i >= size ? -1 : i;
    up = i + 1 >= size ? -1 : i + 1;
    return;
  }
  }
}

// mixed addressers: the scalar coord_calculator path (sampler.cpp:396-410)
__device__ inline float do_coordf(uint32_t mode, float coord, int size) {
  float fs = (float)size;
  switch (mode) {
  case SLV_ADDR_WRAP: return (coord - fast_floor(coord)) * fs - 0.5f;
  case SLV_ADDR_MIRROR: {
    int sel = fast_floori((double)coord);
    return ((sel & 1) ? (float)(1 + sel) - coord : coord - (float)sel) * fs - 0.5f;
  }
  case SLV_ADDR_CLAMP: return clampf(coord * fs, 0.5f, fs - 0.5f) - 0.5f;
  default: return clampf(coord * fs, -0.5f, fs + 0.5f) - 0.5f;
  }
}
__device__ inline int do_coordi_point_1d(uint32_t mode, int coord, int size) {
  switch (mode) {
  case SLV_ADDR_WRAP: return (size * 8192 + coord) % size;
  case SLV_ADDR_MIRROR:
  case SLV_ADDR_CLAMP: return min(max(coord, 0), size - 1);
  default: return coord >= size ? -1 : coord;
  }
}

__device__ __forceinline__ const uint8_t* texel_ptr(const SurfaceRef& s, int x, int y) {
  x = min(max(x, 0), (int)s.w - 1);
  y = min(max(y, 0), (int)s.h - 1);
  return s.data + ((size_t)y * s.w + x) * s.bpp;  // textures are single-sampled
}

__device__ __forceinline__ float lerp1(float a, float b, float t) { return a + (b - a) * t; }

// Two channels (lo, lo + 1) of an rgba8 texel as floats BIASED by 2^23: the byte goes into the mantissa of 0x4B000000
// (= 8388608.0f), one PRMT per channel on the ALU pipe instead of an I2F.U8 on the quarter-rate XU pipe.  8388608 + b is exact.
__device__ __forceinline__ float2 biased_pair(uint32_t t, int lo) {
  return make_float2(__uint_as_float(__byte_perm(t, 0x4B000000u, 0x7440 | lo)), __uint_as_float(__byte_perm(t, 0x4B000000u, 0x7440 | (lo + 1))));
}
// lerp(lerp(c0, c1, tx), lerp(c2, c3, tx), ty) * (1 / 255) per channel with c = (float)byte (colors.h:341-488), two channels per
// instruction.  c1 - c0 is taken on the biased values: both are integers below 2^24, so the difference is the same exact float.
__device__ __forceinline__ float4 bilinear_rgba8(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, float tx, float ty) {
  const float2 bias = splat2(8388608.0f), txx = splat2(tx), tyy = splat2(ty), k = splat2(1.0f / 255);
  float2 o[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float2 m0 = biased_pair(t0, 2 * h), m1 = biased_pair(t1, 2 * h), m2 = biased_pair(t2, 2 * h), m3 = biased_pair(t3, 2 * h);
    const float2 c0 = sub2(m0, bias), c2 = sub2(m2, bias);
    const float2 c01 = add2_after_mul(c0, mul2(sub2(m1, m0), txx)), c23 = add2_after_mul(c2, mul2(sub2(m3, m2), txx));
    o[h] = mul2(add2_after_mul(c01, mul2(sub2(c23, c01), tyy)), k);
  }
  return cat4(o[0], o[1]);
}

// surface::get_texel(x0,y0,x1,y1,tx,ty): bilinear in the NATIVE format (colors.h:341-488)
__device__ inline float4 bilinear(const SurfaceRef& s, int x0, int y0, int x1, int y1, float tx, float ty, bool in_range) {
  if (s.fmt == SLV_PF_RGBA8) {
    // `in_range`: the addresser guarantees 0 <= index < size (every mode but border): 32-bit index arithmetic, no clamps
    if (!in_range) {
      x0 = min(max(x0, 0), (int)s.w - 1); x1 = min(max(x1, 0), (int)s.w - 1);
      y0 = min(max(y0, 0), (int)s.h - 1); y1 = min(max(y1, 0), (int)s.h - 1);
    }
    const uint32_t* base = reinterpret_cast<const uint32_t*>(s.data);
    const uint32_t r0 = (uint32_t)y0 * s.w, r1 = (uint32_t)y1 * s.w;
    uint32_t t0 = __ldg(base + (r0 + (uint32_t)x0)), t1 = __ldg(base + (r0 + (uint32_t)x1));
    uint32_t t2 = __ldg(base + (r1 + (uint32_t)x0)), t3 = __ldg(base + (r1 + (uint32_t)x1));
    return bilinear_rgba8(t0, t1, t2, t3, tx, ty);
  }
  const uint8_t* p0 = texel_ptr(s, x0, y0);
  const uint8_t* p1 = texel_ptr(s, x1, y0);
  const uint8_t* p2 = texel_ptr(s, x0, y1);
  const uint8_t* p3 = texel_ptr(s, x1, y1);
  if (s.fmt == SLV_PF_RGBA32F) {
    float4 c0 = __ldg(reinterpret_cast<const float4*>(p0)), c1 = __ldg(reinterpret_cast<const float4*>(p1));
    float4 c2 = __ldg(reinterpret_cast<const float4*>(p2)), c3 = __ldg(reinterpret_cast<const float4*>(p3));
    return make_float4(lerp1(lerp1(c0.x, c1.x, tx), lerp1(c2.x, c3.x, tx), ty),
                       lerp1(lerp1(c0.y, c1.y, tx), lerp1(c2.y, c3.y, tx), ty),
                       lerp1(lerp1(c0.z, c1.z, tx), lerp1(c2.z, c3.z, tx), ty),
                       lerp1(lerp1(c0.w, c1.w, tx), lerp1(c2.w, c3.w, tx), ty));
  }
  if (s.fmt == SLV_PF_RG32F) {  // upstream bases every channel on c0.r (Appendix B #10) — mirrored
    float2 c0 = __ldg(reinterpret_cast<const float2*>(p0)), c1 = __ldg(reinterpret_cast<const float2*>(p1));
    float2 c2 = __ldg(reinterpret_cast<const float2*>(p2)), c3 = __ldg(reinterpret_cast<const float2*>(p3));
    float c01r = c0.x + (c1.x - c0.x) * tx, c01g = c0.x + (c1.y - c0.y) * tx;
    float c23r = c2.x + (c3.x - c2.x) * tx, c23g = c2.x + (c3.y - c2.y) * tx;
    return make_float4(c01r + (c23r - c01r) * ty, c01r + (c23g - c01g) * ty, 0.0f, 0.0f);
  }
  // bgra8 textures: upstream loads from byte offset 2 of the texel (Appendix B #10) — mirrored
  const uint8_t* end = s.data + s.bytes;
  const uint8_t* ps[4] = {p0, p1, p2, p3};
  float c[4][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint8_t* q = ps[k] + 2;
    uint32_t b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (q + j < end) ? q[j] : 0;
    c[k][0] = (float)b[2]; c[k][1] = (float)b[1]; c[k][2] = (float)b[0]; c[k][3] = (float)b[3];
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = lerp1(lerp1(c[0][j], c[1][j], tx), lerp1(c[2][j], c[3][j], tx), ty) * (1.0f / 255);
  return make_float4(o[0], o[1], o[2], o[3]);
}

// surface_sampler::point / linear ::op (sampler.cpp:423-483)
__device__ inline float4 sample_surface(const SurfaceRef& s, const slv_sampler_desc& d, uint32_t filter, float x, float y) {
  int W = (int)s.w, H = (int)s.h;
  bool same = d.addr_mode_u == d.addr_mode_v;
  if (filter == SLV_FILTER_POINT) {
    int ix, iy;
    bool inside;
    if (same) {
      ix = point_coord_2d(d.addr_mode_u, x, W);
      iy = point_coord_2d(d.addr_mode_v, y, H);
      inside = (0 <= ix && ix < W && 0 <= iy && iy < H);
    } else {
      ix = do_coordi_point_1d(d.addr_mode_u, fast_floori((double)(do_coordf(d.addr_mode_u, x, W) + 0.5f)), W);
      iy = do_coordi_point_1d(d.addr_mode_v, fast_floori((double)(do_coordf(d.addr_mode_v, y, H) + 0.5f)), H);
      inside = !(ix < 0 || iy < 0);
    }
    if (!inside) return make_float4(d.border_color[0], d.border_color[1], d.border_color[2], d.border_color[3]);
    return load_texel_rgba32f(s.fmt, texel_ptr(s, ix, iy));
  }
  int x0, x1, y0, y1;
  float tx, ty;
  if (same) {
    linear_coord_2d(d.addr_mode_u, x, W, s.wmask, x0, x1, tx);
    linear_coord_2d(d.addr_mode_v, y, H, s.hmask, y0, y1, ty);
  } else {
    float ox = do_coordf(d.addr_mode_u, x, W);
    int ipx = fast_floori((double)ox);
    x0 = do_coordi_point_1d(d.addr_mode_u, ipx, W);
    x1 = do_coordi_point_1d(d.addr_mode_u, ipx + 1, W);
    tx = ox - (float)ipx;
    float oy = do_coordf(d.addr_mode_v, y, H);
    int ipy = fast_floori((double)oy);
    y0 = do_coordi_point_1d(d.addr_mode_v, ipy, H);
    y1 = do_coordi_point_1d(d.addr_mode_v, ipy + 1, H);
    ty = oy - (float)ipy;
  }
  // indices are provably inside the level for clamp / mirror, and for wrap when the exact integer wrap is used; the
  // reference's float modulo (other sizes) and border addressing can step outside: those keep the clamped fetch
  const bool safe_u = d.addr_mode_u == SLV_ADDR_CLAMP || d.addr_mode_u == SLV_ADDR_MIRROR || (d.addr_mode_u == SLV_ADDR_WRAP && s.wmask);
  const bool safe_v = d.addr_mode_v == SLV_ADDR_CLAMP || d.addr_mode_v == SLV_ADDR_MIRROR || (d.addr_mode_v == SLV_ADDR_WRAP && s.hmask);
  return bilinear(s, x0, y0, x1, y1, tx, ty, safe_u && safe_v);
}

// One bilinear tap for the commonest sampler state (SamplerRef::fast_wrap_rgba8): wrap addressing with the exact
// integer modulo, rgba8 texels.  Operation for operation the WRAP branch of linear_coord_2d + the rgba8 branch of
// bilinear above, minus the per-tap mode / format dispatch.
struct WrapTap { uint32_t i00, i01, i10, i11; float tx, ty; };  // the four texel indices and the two weights of one tap
__device__ __forceinline__ WrapTap wrap_tap_address(const SurfaceRef& s, float x, float y) {
  const float fw = (float)(int)s.w, fh = (float)(int)s.h;
  float fx = x - trunc_f(x);
  fx = fw * fx;
  fx = fx - 0.5f;
  const float ipx = floor_fix(fx);
  WrapTap t;
  t.tx = fx - ipx;
  float fy = y - trunc_f(y);
  fy = fh * fy;
  fy = fy - 0.5f;
  const float ipy = floor_fix(fy);
  t.ty = fy - ipy;
  const int ix = (int)ipx, iy = (int)ipy;
  const uint32_t wm = s.w - 1, hm = s.h - 1;
  const uint32_t x0 = (uint32_t)ix & wm, x1 = (uint32_t)(ix + 1) & wm;
  const uint32_t r0 = ((uint32_t)iy & hm) * s.w, r1 = ((uint32_t)(iy + 1) & hm) * s.w;
  t.i00 = r0 + x0; t.i01 = r0 + x1; t.i10 = r1 + x0; t.i11 = r1 + x1;
  return t;
}
__device__ __forceinline__ float4 sample_wrap_rgba8_linear(const SurfaceRef& s, float x, float y) {
  const WrapTap t = wrap_tap_address(s, x, y);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(s.data);
  const uint32_t t0 = __ldg(base + t.i00), t1 = __ldg(base + t.i01);
  const uint32_t t2 = __ldg(base + t.i10), t3 = __ldg(base + t.i11);
  return bilinear_rgba8(t0, t1, t2, t3, t.tx, t.ty);
}

#ifndef SLV_EWA_PAIR
#define SLV_EWA_PAIR 1
#endif
struct AfInfo { float lod, probe_count, weight_D, du, dv; };

// sampler::calc_lod (sampler.cpp:521-603)
__device__ inline float calc_lod(const slv_sampler_desc& d, float sw, float sh, float ddx0, float ddx1, float ddy0,
                                 float ddy1, float bias) {
  if (d.mip_qual == SLV_MIP_LO_QUALITY) {
    float m0 = std_max(fabsf(ddx0), fabsf(ddy0));
    float m1 = std_max(fabsf(ddx1), fabsf(ddy1));
    float m2 = 0.0f;
    m0 *= sw; m1 *= sh; m2 *= 1.0f;
    float rho = std_max(std_max(m0, m1), m2);
    return fast_log2(rho) + bias;
  }
  float dxs0 = ddx0 * sw, dxs1 = ddx1 * sh, dxs2 = 0.0f;
  float dys0 = ddy0 * sw, dys1 = ddy1 * sh, dys2 = 0.0f;
  float rho;
  if (d.mip_qual == SLV_MIP_HI_QUALITY) {
    float A = dxs0 * dxs0 + dys0 * dys0;
    float B = -2.0f * (dxs0 * dxs1 + dys0 * dys1);
    float Cc = dxs1 * dxs1 + dys1 * dys1;
    float F = A * Cc - B * B * 0.25f;
    float invF = 1.0f / F;
    A *= invF; B *= invF; Cc *= invF;
    float AsubC = A - Cc;
    float R = sqrtf(AsubC * AsubC + B * B);
    rho = sqrtf(2.0f / (A + Cc - R));
  } else {
    rho = std_max(length3(dxs0, dxs1, dxs2), length3(dys0, dys1, dys2));
  }
  if (rho == 0.0f) rho = 0.000001f;
  return fast_log2(rho) + bias;
}

// sampler::calc_anisotropic_info (sampler.cpp:875-958)
__device__ inline void calc_af(const slv_sampler_desc& d, float sw, float sh, float ddx0, float ddx1, float ddy0,
                               float ddy1, float bias, AfInfo& o) {
  float dxs0 = ddx0 * sw, dxs1 = ddx1 * sh;
  float dys0 = ddy0 * sw, dys1 = ddy1 * sh;
  float ddx_len = length2(dxs0, dxs1);
  float ddy_len = length2(dys0, dys1);
  float diag0 = length2(dxs0 - dys0, dxs1 - dys1);
  float diag1 = length2(dxs0 + dys0, dxs1 + dys1);
  float minor = std_min(std_min(diag0, diag1), std_min(ddx_len, ddy_len));
  if (minor == 0.0f) minor = 0.000001f;
  float la0, la1, la_len;
  if (ddx_len > ddy_len) { la_len = ddx_len; la0 = dxs0; la1 = dxs1; } else { la_len = ddy_len; la0 = dys0; la1 = dys1; }
  float probe = (2.0f * la_len / minor) - 1.0f;
  float rp = fast_round(probe);
  rp = std_min((float)d.max_anisotropy, rp);
  if (rp < probe) minor = 2.0f * la_len / (rp + 1.0f);
  o.lod = fast_log2(minor) + bias;
  o.probe_count = rp;
  if (rp <= 1.0f) {
    o.du = o.dv = 0.0f;
    o.weight_D = 0.0f;
  } else {
    float r = minor / la_len;
    float k0 = (1.0f - r), k2 = (1.0f / (rp - 1.0f));
    float dx = ((la0 * k0) * 2.0f) * k2, dy = ((la1 * k0) * 2.0f) * k2;
    float dz = ((0.0f * k0) * 2.0f) * k2;
    float lsq = 0.0f;
    lsq += dx * dx; lsq += dy * dy; lsq += dz * dz; lsq += dz * dz;
    o.weight_D = 256.0f * lsq * 0.25f / (la_len * la_len);
    o.du = dx / sw;
    o.dv = dy / sh;
  }
}

// sampler::sample_impl<false> (sampler.cpp:680-764).  The magnification / point-mip / linear-mip cases differ only in
// which level(s) and which filter are used, so they share ONE sample_surface call site (a 1- or 2-trip loop) and the
// anisotropic probes another: the sampler is the bulk of the shading kernels' code and must stay inside the
// instruction cache.
__device__ inline float4 sample_impl(const SamplerRef& sm, float cx, float cy, float miplevel, const AfInfo* af) {
  const slv_sampler_desc& d = sm.d;
  const TextureRef& t = sm.tex;
  const int max_lod = 0, min_lod = (int)t.n_levels - 1;
  const bool is_mag = (d.mip_filter == SLV_FILTER_POINT) ? (miplevel < 0.5f) : (miplevel < 0.0f);
  const bool aniso = !is_mag && d.mip_filter == SLV_FILTER_ANISOTROPIC;
  int lv0 = max_lod, lv1 = max_lod, n = 1;
  uint32_t filter = d.mag_filter;
  float frac = 0.0f, sx = cx, sy = cy, du = 0.0f, dv = 0.0f, weight_D = 0.0f;
  int tap_i = 0;  // anisotropic: i = -N+1, -N+3, ... (sampler.cpp:741)
  if (!is_mag) {
    filter = d.min_filter;
    if (d.mip_filter == SLV_FILTER_POINT) {
      const int ml = fast_floori(0.5 + (double)miplevel);
      lv0 = min(max(ml, max_lod), min_lod);
    } else if (d.mip_filter == SLV_FILTER_LINEAR) {
      // miplevel >= 0 here (the mag case is handled above): for non-negative floats the reference's double-precision
      // fast_floori(d) = floor(d + 1.5e-8) equals floorf(d) exactly (no float lies within 1.5e-8 below an integer >= 1)
      const int lo = (int)floorf(miplevel);
      frac = miplevel - (float)lo;
      lv0 = min(max(lo, max_lod), min_lod);
      lv1 = min(max(lo + 1, max_lod), min_lod);
      n = 2;
    } else {  // anisotropic: N probes along the major axis, EWA weights (sampler.cpp:730-760)
      const float start = -0.5f * (af->probe_count - 1.0f);
      du = af->du; dv = af->dv; weight_D = af->weight_D;
      sx = cx + du * start;
      sy = cy + dv * start;
      const int lo = fast_roundi((double)miplevel);
      lv0 = lv1 = lo < 0 ? min_lod : min(max(lo, max_lod), min_lod);  // size_t cast of a negative int clamps to min_lod
      const int pc = (int)af->probe_count;
      n = pc > 0 ? pc : 0;  // taps i = -pc+1, -pc+3, ..., pc-1
      tap_i = -pc + 1;
    }
  }
  if (sm.touched) {  // B_tex accounting (off in timed regions): set the bits once - after that every call only reads the word
    const uint32_t bits = (1u << lv0) | (n == 2 ? (1u << lv1) : 0u);
    if ((__ldcg(sm.touched) & bits) != bits) atomicOr(sm.touched, bits);
  }
  float4 c0 = make_float4(0, 0, 0, 0), c1 = c0;  // aniso: c0 accumulates the weighted colour
  float w_sum = 0.0f;
  if (sm.fast_wrap_rgba8 && !aniso) {
    // straight-line: the texel loads of both mip levels are in flight together (the tap is ~100 instructions, so
    // duplicating it costs little code; the generic loop below keeps ONE copy of the big mode/format dispatch)
    c0 = sample_wrap_rgba8_linear(t.level[lv0], sx, sy);
    if (n == 1) return c0;
    c1 = sample_wrap_rgba8_linear(t.level[lv1], sx, sy);
    return cat4(lerp2_of_products(lo2(c0), lo2(c1), splat2(frac)), lerp2_of_products(hi2(c0), hi2(c1), splat2(frac)));
  }
#if SLV_EWA_PAIR
  // anisotropic probes of the common sampler state, TWO per trip: both probes' eight texel loads are issued before either is
  // filtered, so their latencies overlap (the probes of one pixel are independent; the weighted sum keeps its order: probe k,
  // then probe k + 1, exactly the operations of the one-probe loop below)
  if (aniso && sm.fast_wrap_rgba8) {
    const SurfaceRef& lvl = t.level[lv0];
    const uint32_t* base = reinterpret_cast<const uint32_t*>(lvl.data);
    int k = 0;
#pragma unroll 1
    for (; k + 1 < n; k += 2) {
      const float sx1 = sx + du, sy1 = sy + dv;
      const WrapTap a = wrap_tap_address(lvl, sx, sy), b = wrap_tap_address(lvl, sx1, sy1);
      const uint32_t a0 = __ldg(base + a.i00), a1 = __ldg(base + a.i01), a2 = __ldg(base + a.i10), a3 = __ldg(base + a.i11);
      const uint32_t b0 = __ldg(base + b.i00), b1 = __ldg(base + b.i01), b2 = __ldg(base + b.i10), b3 = __ldg(base + b.i11);
      const int wia = (int)((float)(tap_i * tap_i) * weight_D), wib = (int)((float)((tap_i + 2) * (tap_i + 2)) * weight_D);
      const float wa = c_ewa_wts[min(max(wia, 0), 255)], wb = c_ewa_wts[min(max(wib, 0), 255)];
      const float4 va = bilinear_rgba8(a0, a1, a2, a3, a.tx, a.ty);
      c0 = cat4(add2_after_mul(lo2(c0), mul2(lo2(va), splat2(wa))), add2_after_mul(hi2(c0), mul2(hi2(va), splat2(wa))));
      w_sum += wa;
      const float4 vb = bilinear_rgba8(b0, b1, b2, b3, b.tx, b.ty);
      c0 = cat4(add2_after_mul(lo2(c0), mul2(lo2(vb), splat2(wb))), add2_after_mul(hi2(c0), mul2(hi2(vb), splat2(wb))));
      w_sum += wb;
      sx = sx1 + du;
      sy = sy1 + dv;
      tap_i += 4;
    }
    if (k < n) {
      const float4 v = sample_wrap_rgba8_linear(lvl, sx, sy);
      const int wi = (int)((float)(tap_i * tap_i) * weight_D);
      const float w = c_ewa_wts[min(max(wi, 0), 255)];
      c0 = cat4(add2_after_mul(lo2(c0), mul2(lo2(v), splat2(w))), add2_after_mul(hi2(c0), mul2(hi2(v), splat2(w))));
      w_sum += w;
    }
    const float inv = 1 / w_sum;
    return cat4(mul2(lo2(c0), splat2(inv)), mul2(hi2(c0), splat2(inv)));
  }
#endif
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    const SurfaceRef& lvl = t.level[k ? lv1 : lv0];
    const float4 v = sm.fast_wrap_rgba8 ? sample_wrap_rgba8_linear(lvl, sx, sy) : sample_surface(lvl, d, filter, sx, sy);
    if (aniso) {
      const int wi = (int)((float)(tap_i * tap_i) * weight_D);
      const float w = c_ewa_wts[min(max(wi, 0), 255)];
      c0 = cat4(add2_after_mul(lo2(c0), mul2(lo2(v), splat2(w))), add2_after_mul(hi2(c0), mul2(hi2(v), splat2(w))));
      w_sum += w;
      sx += du;
      sy += dv;
      tap_i += 2;
    } else if (k) {
      c1 = v;
    } else {
      c0 = v;
    }
  }
  if (aniso) {
    const float inv = 1 / w_sum;
    return cat4(mul2(lo2(c0), splat2(inv)), mul2(hi2(c0), splat2(inv)));
  }
  if (n == 1) return c0;
  return cat4(lerp2_of_products(lo2(c0), lo2(c1), splat2(frac)), lerp2_of_products(hi2(c0), hi2(c1), splat2(frac)));
}

// sampler::calc_lod_2d (sampler.cpp:831-848)
__device__ inline float calc_lod_2d(const SamplerRef& sm, float ddx0, float ddx1, float ddy0, float ddy1) {
  float sw = (float)sm.tex.level[0].w, sh = (float)sm.tex.level[0].h;
  if (sm.d.mip_filter == SLV_FILTER_ANISOTROPIC && sm.d.max_anisotropy > 1) {
    AfInfo af;
    calc_af(sm.d, sw, sh, ddx0, ddx1, ddy0, ddy1, 0.0f, af);
    return af.lod;
  }
  return calc_lod(sm.d, sw, sh, ddx0, ddx1, ddy0, ddy1, 0.0f);
}

// sampler::sample_2d_grad (sampler.cpp:854-873)
__device__ inline float4 sample_2d_grad(const SamplerRef& sm, float u, float v, float ddx0, float ddx1, float ddy0,
                                        float ddy1, float bias) {
  float sw = (float)sm.tex.level[0].w, sh = (float)sm.tex.level[0].h;
  AfInfo af = {0, 0, 0, 0, 0};
  float lod;
  if (sm.d.mip_filter == SLV_FILTER_ANISOTROPIC && sm.d.max_anisotropy > 1) {
    calc_af(sm.d, sw, sh, ddx0, ddx1, ddy0, ddy1, bias, af);
    lod = af.lod;
  } else {
    lod = calc_lod(sm.d, sw, sh, ddx0, ddx1, ddy0, ddy1, bias);
  }
  return sample_impl(sm, u, v, lod, &af);
}

}  // namespace slv
