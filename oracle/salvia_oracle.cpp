// oracle/salvia_oracle.cpp — CPU restatement of SALVIA's draw pipeline.  TEST INFRASTRUCTURE ONLY.
//
// Plain scalar, single-threaded C++ that re-states, operation for operation in float32, what the
// reference computes on the hot path `rasterizer::draw()` (salvia/src/core/rasterizer.cpp:1111-1201)
// and the sampler / output merger it calls.  It exports the same slv_* C ABI as the product
// (include/salvia_b200.h) so that tests drive both with identical calls.
//
// PARITY PIN: this file is checked bit-for-bit against the UNMODIFIED reference (oracle/_ref,
// compiled in place from /root/reference) by tests/test_oracle_vs_reference.py and against the frozen
// fixtures under tests/golden/ (generated from the reference by tests/golden/make_golden.py).
//
// Beyond the restatement (test infrastructure on top of it, marked where it appears): slv_shader_compile builds the code the
// product's SASL front end generates for the HOST and run_vs / draw_quad call it (oracle/slv_host_shader.h), so that SASL frames
// can be compared with the samples' cpp twins without a GPU; slv_sasl_translate is the product's own front end (header-only).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (salviarenderer_b200/csrc) never links, imports or falls back to it.
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction, matching the reference's x86-64 SSE2 build,
// SURVEY Appendix A).  Every function cites the reference file:line it follows.

#include "salvia_b200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include "sasl_frontend.hpp"  // salviarenderer_b200/host: the product's SASL front end, for slv_sasl_translate

namespace {

// ------------------------------------------------------------------------------------------------
// eflib math (eflib/include/eflib/math/math.h:89-164)
struct V4 {
  float v[4];
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};

inline V4 mk4(float a, float b, float c, float d) { return V4{{a, b, c, d}}; }

inline bool eq_eps(float a, float b) { return std::fabs(a - b) <= FLT_EPSILON; }  // math.h:21-26

inline float fast_log2(float val) {  // math.h:89-105
  union { int i; float f; } u;
  u.f = val;
  int x = u.i;
  int log_2 = ((x >> 23) & 255) - 128;
  x &= ~(255 << 23);
  x += 127 << 23;
  u.i = x;
  u.f = ((-1.0f / 3) * u.f + 2) * u.f - 2.0f / 3;
  return u.f + log_2;
}

inline float fast_round(float val) {  // math.h:111-124
  union { int i; float f; } n, bias;
  n.f = val;
  bias.i = ((23 + 127) << 23) + (n.i & 0x80000000);
  volatile float t = n.f + bias.f;  // keep the two roundings
  t = t - bias.f;
  return t;
}

inline float fast_floor(float val) {  // math.h:131-134
  float f = fast_round(val);
  return (f > val) ? f - 1 : f;
}

inline int fast_roundi(double d) { return static_cast<int>(std::floor(d + 0.5)); }  // math.h:137-153
const double MAGIC_EPS = (0.5f - 1.5e-8);                                              // math.h:156,162
inline int fast_ceili(double d) { return fast_roundi(d + MAGIC_EPS); }
inline int fast_floori(double d) { return fast_roundi(d - MAGIC_EPS); }

inline float dot4(float const* a, float const* b) {  // eflib/src/math.cpp:43-45
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
}
inline float dot3(float const* a, float const* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline float length3(float const* a) {  // vector_generic.h:109-118
  float t = 0.0f;
  for (int i = 0; i < 3; ++i) t += a[i] * a[i];
  return std::sqrt(t);
}
inline float length2(float x, float y) {
  float t = 0.0f;
  t += x * x;
  t += y * y;
  return std::sqrt(t);
}
inline void normalize3(float* out, float const* v) {  // eflib/src/math.cpp:17-24
  float len = length3(v);
  if (eq_eps(len, 0.0f)) len = 1.0f;
  float inv = 1.0f / len;
  for (int i = 0; i < 3; ++i) out[i] = v[i] * inv;
}
inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }  // std::clamp

// ------------------------------------------------------------------------------------------------
// resources
constexpr int MAX_REGS = 1 + SLV_MAX_VS_OUTPUT_ATTRS;

struct VsOut {  // vs_output: position + attributes (shader_regs.h:30-83)
  V4 r[MAX_REGS];
};

struct Surface {  // surface.cpp:15-51,277-295: linear layout ((y*W+x)*S+s)*bpp
  uint32_t w = 0, h = 0, samples = 1, fmt = 0, bpp = 0;
  std::vector<uint8_t> data;
  uint8_t* addr(size_t x, size_t y, size_t s) { return data.data() + ((y * w + x) * samples + s) * bpp; }
  uint8_t const* addr(size_t x, size_t y, size_t s) const {
    return data.data() + ((y * w + x) * samples + s) * bpp;
  }
};

uint32_t bpp_of(uint32_t fmt) {
  switch (fmt) {
  case SLV_PF_RGBA32F: return 16;
  case SLV_PF_RG32F: return 8;
  case SLV_PF_RGBA8:
  case SLV_PF_BGRA8: return 4;
  }
  return 0;
}

// color conversions (salvia/include/salvia/common/colors.h:148-267)
inline V4 to_rgba32f(uint32_t fmt, uint8_t const* p) {
  const float inv_255 = 1.0f / 255;
  switch (fmt) {
  case SLV_PF_RGBA32F: { V4 c; memcpy(c.v, p, 16); return c; }
  case SLV_PF_RG32F: { V4 c = mk4(0, 0, 0, 0); memcpy(c.v, p, 8); return c; }
  case SLV_PF_RGBA8: return mk4(p[0] * inv_255, p[1] * inv_255, p[2] * inv_255, p[3] * inv_255);
  case SLV_PF_BGRA8: return mk4(p[2] * inv_255, p[1] * inv_255, p[0] * inv_255, p[3] * inv_255);
  }
  return mk4(0, 0, 0, 0);
}

inline uint8_t unorm8_rne(float x) {  // colors.h:182-192: mul, max 0, min 255, cvtps2dq (RNE)
  float m = x * 255.0f;
  m = (m > 0.0f) ? m : 0.0f;   // _mm_max_ps(m, 0): returns 0 when m is NaN
  m = (m < 255.0f) ? m : 255.0f;
  return (uint8_t)std::lrintf(m);  // default rounding mode = nearest even
}

inline void from_rgba32f(uint32_t fmt, uint8_t* p, V4 const& c) {
  switch (fmt) {
  case SLV_PF_RGBA32F: memcpy(p, c.v, 16); break;
  case SLV_PF_RG32F: memcpy(p, c.v, 8); break;
  case SLV_PF_RGBA8:
    p[0] = unorm8_rne(c[0]); p[1] = unorm8_rne(c[1]); p[2] = unorm8_rne(c[2]); p[3] = unorm8_rne(c[3]);
    break;
  case SLV_PF_BGRA8:
    p[2] = unorm8_rne(c[0]); p[1] = unorm8_rne(c[1]); p[0] = unorm8_rne(c[2]); p[3] = unorm8_rne(c[3]);
    break;
  }
}

struct Texture {
  uint32_t fmt = 0, samples = 1;
  std::vector<Surface> levels;  // max_lod = 0, min_lod = levels.size()-1 (texture.h:17-35)
};

struct Sampler {
  slv_sampler_desc d;
  slv_handle tex;
};

// A SASL shader compiled for the HOST (slv_shader_compile below; runtime: oracle/slv_host_shader.h) - test infrastructure on
// top of the restatement: the generated code the product hands to NVRTC, run inside this pipeline.
struct HostApi {  // = SlvHostApi of slv_host_shader.h
  void* ctx;
  void (*sample_grad)(void*, int, float, float, float, float, float, float, float, float*);
  void (*sample_lod)(void*, int, float, float, float, float*);
  void (*vs_sample_lod)(void*, float, float, float, float*);
};
struct Module {
  uint32_t stage = 0, n_vs_output_attrs = 0;
  void* dl = nullptr;
  void (*vs)(const float*, const unsigned char*, float*, const HostApi*) = nullptr;
  void (*ps_quad)(const float*, const unsigned char*, size_t, const HostApi*, float*, int*) = nullptr;
};

struct Resource {
  enum Kind { NONE, BUFFER, TEXTURE, SAMPLER, MODULE } kind = NONE;
  std::vector<uint8_t> buf;
  Texture tex;
  Sampler samp;
  Module mod;
};

// ------------------------------------------------------------------------------------------------
// sampler (salvia/src/resource/sampler.cpp)
inline float trunc_f(float x) { return (float)(int)x; }  // cvttps2dq + cvtdq2ps
inline float floor_fix(float x) {                          // sampler.cpp:36-40
  float ip = trunc_f(x);
  if (ip > x) ip = ip - 1.0f;
  return ip;
}

// 2-D (same addresser on u and v) paths = the SIMD branches of addresser::*::do_coordi_*_2d.
// Returns false when point sampling must return the border colour.
struct LinearCoord { int x0, y0, x1, y1; float tx, ty; };

inline int wrap_index(float ipart_plus, float fsize) {  // sampler.cpp:42-44 / 90-92, float modulo
  float v = ipart_plus + fsize * 8192.0f;
  float dv = trunc_f(v / fsize);
  return (int)(v - dv * fsize);
}

int point_coord_1axis_2d(uint32_t mode, float c, int size) {
  float fs = (float)size;
  switch (mode) {
  case SLV_ADDR_WRAP: {  // sampler.cpp:27-47
    float f = c - trunc_f(c);
    f = fs * f;
    float ip = floor_fix(f);
    return wrap_index(ip, fs);
  }
  case SLV_ADDR_MIRROR: {  // sampler.cpp:131-157
    int sel = fast_floori(c);
    float o = ((sel & 1) ? (float)(1 + sel) - c : c - (float)sel) * (float)size;
    float ip = floor_fix(o);
    float m = ip > 0.0f ? ip : 0.0f;
    m = m < (fs - 1.0f) ? m : (fs - 1.0f);
    return (int)m;
  }
  case SLV_ADDR_CLAMP: {  // sampler.cpp:226-249
    float o = c * fs;
    o = o > 0.5f ? o : 0.5f;
    float hi = fs - 0.5f;
    o = o < hi ? o : hi;
    float ip = floor_fix(o);
    float m = ip > 0.0f ? ip : 0.0f;
    m = m < (fs - 1.0f) ? m : (fs - 1.0f);
    return (int)m;
  }
  case SLV_ADDR_BORDER: {  // sampler.cpp:320-350
    float o = c * fs;
    o = o > -0.5f ? o : -0.5f;
    float hi = fs - (-0.5f);
    o = o < hi ? o : hi;
    int ip = (int)floor_fix(o);
    return ip >= size ? -1 : ip;
  }
  }
  return 0;
}

void linear_coord_1axis_2d(uint32_t mode, float c, int size, int& lo, int& up, float& frac) {
  float fs = (float)size;
  switch (mode) {
  case SLV_ADDR_WRAP: {  // sampler.cpp:67-95
    float f = c - trunc_f(c);
    f = fs * f;
    f = f - 0.5f;
    float ip = floor_fix(f);
    frac = f - ip;
    lo = wrap_index(ip + 0.0f, fs);
    up = wrap_index(ip + 1.0f, fs);
    return;
  }
  case SLV_ADDR_MIRROR: {  // sampler.cpp:168-202
    int sel = fast_floori(c);
    float o = ((sel & 1) ? (float)(1 + sel) - c : c - (float)sel) * (float)size - 0.5f;
    float ip = floor_fix(o);
    frac = o - ip;
    float a = ip + 0.0f, b = ip + 1.0f;
    a = a > 0.0f ? a : 0.0f; a = a < (fs - 1.0f) ? a : (fs - 1.0f);
    b = b > 0.0f ? b : 0.0f; b = b < (fs - 1.0f) ? b : (fs - 1.0f);
    lo = (int)a; up = (int)b;
    return;
  }
  case SLV_ADDR_CLAMP: {  // sampler.cpp:264-292
    float o = c * fs;
    o = o > 0.5f ? o : 0.5f;
    float hi = fs - 0.5f;
    o = o < hi ? o : hi;
    o = o - 0.5f;
    float ip = floor_fix(o);
    frac = o - ip;
    float a = ip + 0.0f, b = ip + 1.0f;
    a = a > 0.0f ? a : 0.0f; a = a < (fs - 1.0f) ? a : (fs - 1.0f);
    b = b > 0.0f ? b : 0.0f; b = b < (fs - 1.0f) ? b : (fs - 1.0f);
    lo = (int)a; up = (int)b;
    return;
  }
  case SLV_ADDR_BORDER: {  // sampler.cpp:352-391 (up/low may be -1; the reference then reads out of bounds:
                           // undefined upstream, we clamp the read and document it)
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

// 1-D scalar paths, used when addr_mode_u != addr_mode_v (coord_calculator::point_cc / linear_cc,
// sampler.cpp:396-410 with addresser::*::do_coordf / do_coordi_point_1d).
float do_coordf(uint32_t mode, float coord, int size) {
  switch (mode) {
  case SLV_ADDR_WRAP: return (coord - fast_floor(coord)) * size - 0.5f;
  case SLV_ADDR_MIRROR: {
    int sel = fast_floori(coord);
    return ((sel & 1) ? 1 + sel - coord : coord - sel) * size - 0.5f;
  }
  case SLV_ADDR_CLAMP: return clampf(coord * size, 0.5f, size - 0.5f) - 0.5f;
  case SLV_ADDR_BORDER: return clampf(coord * size, -0.5f, size + 0.5f) - 0.5f;
  }
  return 0;
}
int do_coordi_point_1d(uint32_t mode, int coord, int size) {
  switch (mode) {
  case SLV_ADDR_WRAP: return (size * 8192 + coord) % size;
  case SLV_ADDR_MIRROR:
  case SLV_ADDR_CLAMP: return std::min(std::max(coord, 0), size - 1);
  case SLV_ADDR_BORDER: return coord >= size ? -1 : coord;
  }
  return 0;
}

V4 texel(Surface const& s, int x, int y) {
  x = std::min(std::max(x, 0), (int)s.w - 1);  // only differs from the reference where it reads OOB
  y = std::min(std::max(y, 0), (int)s.h - 1);
  return to_rgba32f(s.fmt, s.addr(x, y, 0));
}

// surface::get_texel(x0,y0,x1,y1,tx,ty) -> lerp_2d in native format (surface.cpp:152-160, colors.h:341-488)
V4 bilinear(Surface const& s, int x0, int y0, int x1, int y1, float tx, float ty) {
  auto cl = [&](int& x, int& y) {
    x = std::min(std::max(x, 0), (int)s.w - 1);
    y = std::min(std::max(y, 0), (int)s.h - 1);
  };
  cl(x0, y0);
  cl(x1, y1);
  uint8_t const* p0 = s.addr(x0, y0, 0);
  uint8_t const* p1 = s.addr(x1, y0, 0);
  uint8_t const* p2 = s.addr(x0, y1, 0);
  uint8_t const* p3 = s.addr(x1, y1, 0);
  V4 out;
  switch (s.fmt) {
  case SLV_PF_RGBA8:
  case SLV_PF_BGRA8: {
    // colors.h:420-447 (rgba8).  bgra8 (colors.h:383-400) loads from &c0.r == byte offset 2 of the texel
    // (Appendix B #10): channels come from bytes [2,3,4,5] shuffled (3,0,1,2) — mirrored here.
    float c[4][4];
    uint8_t const* ps[4] = {p0, p1, p2, p3};
    for (int k = 0; k < 4; ++k) {
      if (s.fmt == SLV_PF_RGBA8) {
        for (int j = 0; j < 4; ++j) c[k][j] = (float)ps[k][j];
      } else {
        uint8_t const* q = ps[k] + 2;  // may read 2 bytes past the texel, like the reference
        uint8_t b[4];
        size_t remain = (s.data.data() + s.data.size()) - q;
        for (int j = 0; j < 4; ++j) b[j] = (size_t)j < remain ? q[j] : 0;
        c[k][0] = (float)b[2]; c[k][1] = (float)b[1]; c[k][2] = (float)b[0]; c[k][3] = (float)b[3];
      }
    }
    for (int j = 0; j < 4; ++j) {
      float c01 = c[0][j] + (c[1][j] - c[0][j]) * tx;
      float c23 = c[2][j] + (c[3][j] - c[2][j]) * tx;
      float r = c01 + (c23 - c01) * ty;
      out[j] = r * (1.0f / 255);
    }
    return out;
  }
  case SLV_PF_RGBA32F: {  // colors.h:341-363
    float c[4][4];
    memcpy(c[0], p0, 16); memcpy(c[1], p1, 16); memcpy(c[2], p2, 16); memcpy(c[3], p3, 16);
    for (int j = 0; j < 4; ++j) {
      float c01 = c[0][j] + (c[1][j] - c[0][j]) * tx;
      float c23 = c[2][j] + (c[3][j] - c[2][j]) * tx;
      out[j] = c01 + (c23 - c01) * ty;
    }
    return out;
  }
  case SLV_PF_RG32F: {  // colors.h:469-478: every channel is based on c0.r (Appendix B #10) — mirrored
    float c[4][2];
    memcpy(c[0], p0, 8); memcpy(c[1], p1, 8); memcpy(c[2], p2, 8); memcpy(c[3], p3, 8);
    float c01r = c[0][0] + (c[1][0] - c[0][0]) * tx, c01g = c[0][0] + (c[1][1] - c[0][1]) * tx;
    float c23r = c[2][0] + (c[3][0] - c[2][0]) * tx, c23g = c[2][0] + (c[3][1] - c[2][1]) * tx;
    out = mk4(c01r + (c23r - c01r) * ty, c01r + (c23g - c01g) * ty, 0.0f, 0.0f);
    return out;
  }
  }
  return mk4(0, 0, 0, 0);
}

// surface_sampler::point / linear ::op (sampler.cpp:423-483)
V4 sample_surface(Surface const& s, slv_sampler_desc const& d, uint32_t filter, float x, float y) {
  int W = (int)s.w, H = (int)s.h;
  V4 border = mk4(d.border_color[0], d.border_color[1], d.border_color[2], d.border_color[3]);
  bool same = d.addr_mode_u == d.addr_mode_v;
  if (filter == SLV_FILTER_POINT) {
    if (same) {
      int ix = point_coord_1axis_2d(d.addr_mode_u, x, W);
      int iy = point_coord_1axis_2d(d.addr_mode_v, y, H);
      if (0 <= ix && ix < W && 0 <= iy && iy < H) return texel(s, ix, iy);
      return border;
    }
    int ix = do_coordi_point_1d(d.addr_mode_u, fast_floori(do_coordf(d.addr_mode_u, x, W) + 0.5f), W);
    int iy = do_coordi_point_1d(d.addr_mode_v, fast_floori(do_coordf(d.addr_mode_v, y, H) + 0.5f), H);
    if (ix < 0 || iy < 0) return border;
    return texel(s, ix, iy);
  }
  // linear (anisotropic min/mag filters index past filter_table: not used by any config)
  int x0, x1, y0, y1;
  float tx, ty;
  if (same) {
    linear_coord_1axis_2d(d.addr_mode_u, x, W, x0, x1, tx);
    linear_coord_1axis_2d(d.addr_mode_v, y, H, y0, y1, ty);
  } else {
    float ox = do_coordf(d.addr_mode_u, x, W);
    int ipx = fast_floori(ox);
    x0 = do_coordi_point_1d(d.addr_mode_u, ipx, W);
    x1 = do_coordi_point_1d(d.addr_mode_u, ipx + 1, W);
    tx = ox - ipx;
    float oy = do_coordf(d.addr_mode_v, y, H);
    int ipy = fast_floori(oy);
    y0 = do_coordi_point_1d(d.addr_mode_v, ipy, H);
    y1 = do_coordi_point_1d(d.addr_mode_v, ipy + 1, H);
    ty = oy - ipy;
  }
  return bilinear(s, x0, y0, x1, y1, tx, ty);
}

struct AfInfo { float lod, probe_count, weight_D, du, dv; };

extern const float EWA_WTS[256];

// sampler::calc_lod (sampler.cpp:521-603); size = level-0 size, ddx/ddy in uv units
float calc_lod(slv_sampler_desc const& d, float sw, float sh, float ddx0, float ddx1, float ddy0, float ddy1,
               float bias) {
  const float sz2 = 1.0f;  // size[2] == 1 for 2-D textures (texture2d.cpp:21)
  if (d.mip_qual == SLV_MIP_LO_QUALITY) {
    float m0 = std::max(std::fabs(ddx0), std::fabs(ddy0));
    float m1 = std::max(std::fabs(ddx1), std::fabs(ddy1));
    float m2 = std::max(std::fabs(0.0f), std::fabs(0.0f));
    m0 *= sw; m1 *= sh; m2 *= sz2;
    float rho = std::max(std::max(m0, m1), m2);
    return fast_log2(rho) + bias;
  }
  float dxs0 = ddx0 * sw, dxs1 = ddx1 * sh, dxs2 = 0.0f * sz2;
  float dys0 = ddy0 * sw, dys1 = ddy1 * sh, dys2 = 0.0f * sz2;
  float rho;
  if (d.mip_qual == SLV_MIP_HI_QUALITY) {
    float A = dxs0 * dxs0 + dys0 * dys0;
    float B = -2.0f * (dxs0 * dxs1 + dys0 * dys1);
    float Cc = dxs1 * dxs1 + dys1 * dys1;
    float F = A * Cc - B * B * 0.25f;
    float invF = 1.0f / F;
    A *= invF; B *= invF; Cc *= invF;
    float AsubC = A - Cc;
    float R = std::sqrt(AsubC * AsubC + B * B);
    rho = std::sqrt(2.0f / (A + Cc - R));
  } else {
    float vx[3] = {dxs0, dxs1, dxs2}, vy[3] = {dys0, dys1, dys2};
    rho = std::max(length3(vx), length3(vy));
  }
  if (rho == 0.0f) rho = 0.000001f;
  return fast_log2(rho) + bias;
}

// sampler::calc_anisotropic_info (sampler.cpp:875-958)
void calc_af(slv_sampler_desc const& d, float sw, float sh, float ddx0, float ddx1, float ddy0, float ddy1,
             float bias, AfInfo& o) {
  float dxs[2] = {ddx0 * sw, ddx1 * sh};
  float dys[2] = {ddy0 * sw, ddy1 * sh};
  float ddx_len = length2(dxs[0], dxs[1]);
  float ddy_len = length2(dys[0], dys[1]);
  float diag0 = length2(dxs[0] - dys[0], dxs[1] - dys[1]);
  float diag1 = length2(dxs[0] + dys[0], dxs[1] + dys[1]);
  float minor = std::min(std::min(diag0, diag1), std::min(ddx_len, ddy_len));
  if (minor == 0.0f) minor = 0.000001f;
  float const* la;
  float la_len;
  if (ddx_len > ddy_len) { la_len = ddx_len; la = dxs; } else { la_len = ddy_len; la = dys; }
  float probe = (2.0f * la_len / minor) - 1.0f;
  float rp = fast_round(probe);
  rp = std::min(static_cast<float>(d.max_anisotropy), rp);
  if (rp < probe) minor = 2.0f * la_len / (rp + 1.0f);
  o.lod = fast_log2(minor) + bias;
  o.probe_count = rp;
  if (rp <= 1.0f) {
    o.du = o.dv = 0.0f;
    o.weight_D = 0.0f;
  } else {
    float r = minor / la_len;
    // (*long_axis) * (1 - r) * 2 * (1/(rp-1)): vec4 * float, left to right
    float k0 = (1.0f - r), k2 = (1.0f / (rp - 1.0f));
    float dx = ((la[0] * k0) * 2.0f) * k2, dy = ((la[1] * k0) * 2.0f) * k2;
    float dz = ((0.0f * k0) * 2.0f) * k2;
    float lsq = 0.0f;  // vec4::length_sqr over 4 comps
    lsq += dx * dx; lsq += dy * dy; lsq += dz * dz; lsq += dz * dz;
    o.weight_D = static_cast<float>(256) * lsq * 0.25f / (la_len * la_len);
    o.du = dx / sw;
    o.dv = dy / sh;
  }
}

// sampler::sample_impl<false> (sampler.cpp:680-764)
V4 sample_impl(Texture const& t, slv_sampler_desc const& d, float cx, float cy, float miplevel, AfInfo const* af) {
  int max_lod = 0, min_lod = (int)t.levels.size() - 1;
  bool is_mag = (d.mip_filter == SLV_FILTER_POINT) ? (miplevel < 0.5f) : (miplevel < 0.0f);
  if (is_mag) return sample_surface(t.levels[max_lod], d, d.mag_filter, cx, cy);
  if (d.mip_filter == SLV_FILTER_POINT) {
    int ml = fast_floori(0.5 + miplevel);
    ml = std::min(std::max(ml, max_lod), min_lod);
    return sample_surface(t.levels[ml], d, d.min_filter, cx, cy);
  }
  if (d.mip_filter == SLV_FILTER_LINEAR) {
    int lo = fast_floori(miplevel);
    int hi = lo + 1;
    float frac = miplevel - lo;
    int lo_sz = std::min(std::max(lo, max_lod), min_lod);
    int hi_sz = std::min(std::max(hi, max_lod), min_lod);
    V4 c0 = sample_surface(t.levels[lo_sz], d, d.min_filter, cx, cy);
    V4 c1 = sample_surface(t.levels[hi_sz], d, d.min_filter, cx, cy);
    V4 r;
    for (int j = 0; j < 4; ++j) r[j] = c0[j] + (c1[j] - c0[j]) * frac;  // colors.h:303-314
    return r;
  }
  // anisotropic (sampler.cpp:730-760)
  float start = -0.5f * (af->probe_count - 1.0f);
  float sx = cx + af->du * start;
  float sy = cy + af->dv * start;
  int lo = fast_roundi(miplevel);
  size_t lo_sz = std::min(std::max(static_cast<size_t>(lo), (size_t)max_lod), (size_t)min_lod);
  V4 color = mk4(0, 0, 0, 0);
  float w_sum = 0.0f;
  int pc = static_cast<int>(af->probe_count);
  for (int i = -pc + 1; i < pc; i += 2) {
    V4 c0 = sample_surface(t.levels[lo_sz], d, d.min_filter, sx, sy);
    float w = EWA_WTS[static_cast<int>(i * i * af->weight_D)];
    for (int j = 0; j < 4; ++j) color[j] += c0[j] * w;
    w_sum += w;
    sx += af->du;
    sy += af->dv;
  }
  float inv = 1 / w_sum;  // vector_generic.h:75-79
  for (int j = 0; j < 4; ++j) color[j] *= inv;
  return color;
}

struct Device;
Texture const* sampler_texture(Device const& dev, Sampler const& s);

// sampler::calc_lod_2d (sampler.cpp:831-848)
float calc_lod_2d(Texture const& t, slv_sampler_desc const& d, float ddx0, float ddx1, float ddy0, float ddy1) {
  float sw = (float)t.levels[0].w, sh = (float)t.levels[0].h;
  if (d.mip_filter == SLV_FILTER_ANISOTROPIC && d.max_anisotropy > 1) {
    AfInfo af;
    calc_af(d, sw, sh, ddx0, ddx1, ddy0, ddy1, 0.0f, af);
    return af.lod;
  }
  return calc_lod(d, sw, sh, ddx0, ddx1, ddy0, ddy1, 0.0f);
}

// sampler::sample_2d_grad (sampler.cpp:854-873)
V4 sample_2d_grad(Texture const& t, slv_sampler_desc const& d, float u, float v, float ddx0, float ddx1, float ddy0,
                  float ddy1, float bias) {
  float sw = (float)t.levels[0].w, sh = (float)t.levels[0].h;
  AfInfo af = {0, 0, 0, 0, 0};
  float lod;
  if (d.mip_filter == SLV_FILTER_ANISOTROPIC && d.max_anisotropy > 1) {
    calc_af(d, sw, sh, ddx0, ddx1, ddy0, ddy1, bias, af);
    lod = af.lod;
  } else {
    lod = calc_lod(d, sw, sh, ddx0, ddx1, ddy0, ddy1, bias);
  }
  return sample_impl(t, d, u, v, lod, &af);
}

// ------------------------------------------------------------------------------------------------
struct Device {
  std::vector<Resource> res;
  slv_pipeline_statistics stats{};
  slv_traffic_counters traffic{};
  std::chrono::steady_clock::time_point events[16];
  uint32_t shard_rank = 0, shard_n = 1;
  Resource* get(slv_handle h, Resource::Kind k) {
    if (h == 0 || h >= res.size() || res[h].kind != k) return nullptr;
    return &res[h];
  }
  Resource const* get(slv_handle h, Resource::Kind k) const {
    if (h == 0 || h >= res.size() || res[h].kind != k) return nullptr;
    return &res[h];
  }
};

Texture const* sampler_texture(Device const& dev, Sampler const& s) {
  auto r = dev.get(s.tex, Resource::TEXTURE);
  return r ? &r->tex : nullptr;
}

// ------------------------------------------------------------------------------------------------
// draw state
struct TriInfo {  // triangle_info (shader_regs.h:93-104) + the rotated v0
  VsOut v0, ddx, ddy;
  float bbox[4];
  float edge[3][3];
  bool front_face;
  bool valid;
};

struct DrawCtx {
  Device* dev;
  slv_draw_desc const* d;
  uint32_t n_attrs;
  uint32_t mods[SLV_MAX_VS_OUTPUT_ATTRS];
  bool has_centroid;
  uint32_t S;
  float sp[4][2];
  std::vector<Surface*> colors;
  Surface* ds;
  float target_w, target_h;
  // depth/stencil function selection (framebuffer.cpp:358-425)
  bool read_depth, read_stencil, write_depth, write_stencil, early_z;
  uint32_t stencil_ref, read_mask, write_mask;
  Sampler const* samplers[SLV_MAX_SAMPLERS];
  Texture const* sampler_tex[SLV_MAX_SAMPLERS];
  Sampler const* vs_sampler = nullptr;       // vertex texture fetch: vs.samplers[0]
  Texture const* vs_sampler_tex = nullptr;
  Module const* vs_module = nullptr;         // SLV_PROGRAM_JIT(module): host-compiled SASL shaders
  Module const* ps_module = nullptr;
  HostApi host_api{};
  uint64_t ps_invocations = 0, backend_input_pixels = 0;
};

// ---- input assembler: index_fetcher.cpp:26-115, stream_assembler.cpp:26-93
bool fetch_indices(DrawCtx& c, uint32_t prim, uint32_t out[3]) {
  auto d = c.d;
  uint32_t ids[3];
  if (d->topology == SLV_TOPO_TRIANGLE_LIST) {
    ids[0] = prim * 3; ids[1] = prim * 3 + 1; ids[2] = prim * 3 + 2;
  } else {
    ids[0] = prim; ids[1] = prim + 1; ids[2] = prim + 2;
    if (prim & 1) std::swap(ids[0], ids[2]);
  }
  if (d->index_buffer) {
    auto r = c.dev->get(d->index_buffer, Resource::BUFFER);
    if (!r) return false;
    size_t stride = d->index_format == SLV_INDEX_R16_UINT ? 2 : 4;
    uint8_t const* base = r->buf.data() + (size_t)d->start * stride;
    for (int i = 0; i < 3; ++i) {
      uint32_t v;
      if (stride == 2) { uint16_t t; memcpy(&t, base + (size_t)ids[i] * 2, 2); v = t; }
      else memcpy(&v, base + (size_t)ids[i] * 4, 4);
      out[i] = v + (uint32_t)d->base_vertex;
    }
  } else {
    // NOTE: for non-indexed draws the reference ignores start_index (index_fetcher.cpp:107-111)
    for (int i = 0; i < 3; ++i) out[i] = ids[i] + (uint32_t)d->base_vertex;
  }
  return true;
}

void fetch_vertex(DrawCtx& c, uint32_t index, V4 in[SLV_MAX_VS_INPUT_ATTRS]) {
  auto d = c.d;
  for (int i = 0; i < SLV_MAX_VS_INPUT_ATTRS; ++i) in[i] = mk4(0, 0, 0, 0);
  for (uint32_t e = 0; e < d->n_elements; ++e) {
    auto const& el = d->elements[e];
    auto const& st = d->streams[el.slot];
    auto r = c.dev->get(st.buffer, Resource::BUFFER);
    float f[4] = {0, 0, 0, 0};
    uint8_t const* p = r->buf.data() + el.aligned_byte_offset + (size_t)st.stride * index + st.offset;
    switch (el.format) {  // get_vec4, stream_assembler.cpp:26-45
    case SLV_FMT_R32_FLOAT: memcpy(f, p, 4); in[el.reg] = mk4(f[0], 0, 0, el.default_w); break;
    case SLV_FMT_R32G32_FLOAT: memcpy(f, p, 8); in[el.reg] = mk4(f[0], f[1], 0, el.default_w); break;
    case SLV_FMT_R32G32B32_FLOAT: memcpy(f, p, 12); in[el.reg] = mk4(f[0], f[1], f[2], el.default_w); break;
    default: memcpy(f, p, 16); in[el.reg] = mk4(f[0], f[1], f[2], f[3]); break;
    }
  }
}

// pos = v · M: out[i] = dot4(v, column i)  (eflib/src/math.cpp:142-154)
V4 transform(V4 const& v, float const* m) {
  V4 o;
  for (int i = 0; i < 4; ++i) {
    float col[4] = {m[0 * 4 + i], m[1 * 4 + i], m[2 * 4 + i], m[3 * 4 + i]};
    o[i] = dot4(v.v, col);
  }
  return o;
}
V4 sub4(V4 const& a, V4 const& b) { return mk4(a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]); }

uint32_t vs_num_attrs(slv_shader_binding const& vs) {
  switch (vs.program) {
  case SLV_VS_MVP_PASSTHROUGH: return ((slv_vs_mvp_passthrough_uniforms const*)vs.uniforms)->n_attrs;
  case SLV_VS_PLANE_XZ: return 1;
  case SLV_VS_LIGHTS3: return 4;
  case SLV_VS_SPONZA: return 4;
  case SLV_VS_TERRAIN_VTF: return 1;
  case SLV_VS_SSM_DRAW: return 5;
  }
  return 0;
}

void run_vs(DrawCtx& c, V4 const in[SLV_MAX_VS_INPUT_ATTRS], VsOut& out) {
  auto const& vs = c.d->vs;
  for (auto& r : out.r) r = mk4(0, 0, 0, 0);
  if (c.vs_module) {  // a SASL vertex shader compiled for the host: input registers in, position + attributes out
    float regs[8][4] = {}, o[6][4] = {};
    for (int k = 0; k < 8 && k < SLV_MAX_VS_INPUT_ATTRS; ++k) for (int j = 0; j < 4; ++j) regs[k][j] = in[k][j];
    c.vs_module->vs(&regs[0][0], vs.uniforms, &o[0][0], &c.host_api);
    for (int k = 0; k < 6 && k < (int)(sizeof(out.r) / sizeof(out.r[0])); ++k) out.r[k] = mk4(o[k][0], o[k][1], o[k][2], o[k][3]);
    return;
  }
  switch (vs.program) {
  case SLV_VS_MVP_PASSTHROUGH: {
    auto u = (slv_vs_mvp_passthrough_uniforms const*)vs.uniforms;
    out.r[0] = transform(in[0], u->wvp);
    for (uint32_t i = 0; i < u->n_attrs; ++i) out.r[1 + i] = in[u->src[i]];
  } break;
  case SLV_VS_PLANE_XZ: {
    auto u = (slv_vs_plane_xz_uniforms const*)vs.uniforms;
    out.r[0] = transform(in[0], u->wvp);
    out.r[1] = mk4(in[0][0], in[0][2], 0, 0);
  } break;
  case SLV_VS_LIGHTS3: {
    auto u = (slv_vs_lights3_uniforms const*)vs.uniforms;
    out.r[0] = transform(in[0], u->wvp);
    out.r[1] = in[1];
    for (int k = 0; k < 3; ++k) {
      V4 l = mk4(u->light_pos[k][0], u->light_pos[k][1], u->light_pos[k][2], u->light_pos[k][3]);
      out.r[2 + k] = sub4(l, in[0]);
    }
  } break;
  case SLV_VS_SPONZA: {
    auto u = (slv_vs_sponza_uniforms const*)vs.uniforms;
    out.r[0] = transform(in[0], u->wvp);
    out.r[1] = in[1];
    out.r[2] = in[2];
    out.r[3] = sub4(mk4(u->light_pos[0], u->light_pos[1], u->light_pos[2], u->light_pos[3]), in[0]);
    out.r[4] = sub4(mk4(u->eye_pos[0], u->eye_pos[1], u->eye_pos[2], u->eye_pos[3]), in[0]);
  } break;
  case SLV_VS_SSM_DRAW: {  // resources/ssm/Draw.savs:25-36 (inputs: POSITION, NORMAL, TEXCOORD0)
    auto u = (slv_vs_ssm_draw_uniforms const*)vs.uniforms;
    out.r[0] = transform(in[0], u->camera_wvp);
    out.r[1] = in[2];                                                                                       // tex
    out.r[2] = in[1];                                                                                       // norm
    out.r[3] = sub4(mk4(u->light_pos[0], u->light_pos[1], u->light_pos[2], u->light_pos[3]), in[0]);        // lightDir
    out.r[4] = sub4(mk4(u->camera_pos[0], u->camera_pos[1], u->camera_pos[2], u->camera_pos[3]), in[0]);    // cameraDir
    out.r[5] = transform(in[0], u->light_wvp);                                                              // lightSpacePos
  } break;
  case SLV_VS_TERRAIN_VTF: {  // VertexTextureFetch.cpp:38-61; tex2Dlod = sample_2d_lod (sampler_api.cpp:50-52, sampler.cpp:850-852)
    auto u = (slv_vs_terrain_vtf_uniforms const*)vs.uniforms;
    float tu = u->offset[0] + in[1][0] * u->scale[0], tv = u->offset[1] + in[1][1] * u->scale[1];
    float disp = sample_impl(*c.vs_sampler_tex, c.vs_sampler->d, tu, tv, 0.0f, nullptr)[0];
    V4 displaced = mk4(in[0][0] + 0.0f, in[0][1] + disp * 20.0f, in[0][2] + 0.0f, 1.0f);
    out.r[0] = transform(displaced, u->wvp);
    out.r[1] = mk4(disp, 0, 0, 0);
  } break;
  }
}

// ---- clipper (clipper.cpp:43-228) ----------------------------------------------------------------
bool cull(slv_raster_desc const& r, float area) {  // raster_state.cpp:10-31
  switch (r.cull_mode) {
  case SLV_CULL_NONE: return false;
  case SLV_CULL_FRONT: return r.front_ccw ? (area <= 0) : (area >= 0);
  case SLV_CULL_BACK: return r.front_ccw ? (area >= 0) : (area <= 0);
  }
  return false;
}

void lerp_vso(DrawCtx& c, VsOut& out, VsOut const& a, VsOut const& b, float t) {  // shader.cpp:170-180
  for (int j = 0; j < 4; ++j) out.r[0][j] = a.r[0][j] + (b.r[0][j] - a.r[0][j]) * t;
  for (uint32_t i = 0; i < c.n_attrs; ++i) {
    out.r[1 + i] = a.r[1 + i];
    if (!(c.mods[i] & SLV_AM_NOINTERPOLATION)) {
      for (int j = 0; j < 4; ++j) out.r[1 + i][j] += (b.r[1 + i][j] - a.r[1 + i][j]) * t;
    }
  }
}

const float PLANES[2][4] = {{0.0f, 0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, -1.0f, 1.0f}};  // clipper.cpp:21-27

// Emits 0..3 triangles (clip-space vertices) for one input primitive.
int clip_triangle(DrawCtx& c, VsOut const tri[3], VsOut out[9]) {
  bool in_frustum = true;
  for (int p = 0; p < 2 && in_frustum; ++p)
    for (int v = 0; v < 3; ++v)
      if (dot4(PLANES[p], tri[v].r[0].v) < 0) { in_frustum = false; break; }

  if (in_frustum) {  // clipper.cpp:92-101,121-132 and :55-68
    float px[3], py[3];
    for (int v = 0; v < 3; ++v) {
      float iw = 1.0f / tri[v].r[0][3];
      px[v] = tri[v].r[0][0] * iw;
      py[v] = tri[v].r[0][1] * iw;
    }
    float area = (px[2] - px[0]) * (py[1] - py[0]) - (py[2] - py[0]) * (px[1] - px[0]);
    bool front = area > 0.0f;
    if (cull(c.d->raster, front ? 1.0f : -1.0f)) return 0;
    int off = front ? 0 : 1;
    out[0] = tri[0];
    out[1] = tri[1 + off];
    out[2] = tri[2 - off];
    return 1;
  }

  VsOut pool[2][5];
  int n[2] = {3, 0};
  pool[0][0] = tri[0]; pool[0][1] = tri[1]; pool[0][2] = tri[2];
  int src = 0, dst = 1;
  bool is_front = false;
  for (int p = 0; p < 2; ++p) {
    n[dst] = 0;
    float d0 = 0, d1;
    if (n[src] != 0) d0 = dot4(PLANES[p], pool[src][0].r[0].v);
    for (int i = 0, j = 1; i < n[src]; ++i, ++j) {
      j %= n[src];
      d1 = dot4(PLANES[p], pool[src][j].r[0].v);
      if (d0 >= 0.0f) {
        pool[dst][n[dst]++] = pool[src][i];
        if (d1 < 0.0f) {
          lerp_vso(c, pool[dst][n[dst]], pool[src][i], pool[src][j], d0 / (d0 - d1));
          ++n[dst];
        }
      } else if (d1 >= 0.0f) {
        lerp_vso(c, pool[dst][n[dst]], pool[src][j], pool[src][i], d1 / (d1 - d0));
        ++n[dst];
      }
      d0 = d1;
    }
    if (p == 0 && n[dst] >= 3) {  // clipper.cpp:191-211
      float px[3], py[3];
      for (int i = 0; i < 3; ++i) {
        float inv_abs_w = 1 / std::fabs(pool[dst][i].r[0][3]);
        px[i] = pool[dst][i].r[0][0] * inv_abs_w;
        py[i] = pool[dst][i].r[0][1] * inv_abs_w;
      }
      float area = (px[2] - px[0]) * (py[1] - py[0]) - (py[2] - py[0]) * (px[1] - px[0]);
      is_front = area > 0.0f;
      if (cull(c.d->raster, area)) return 0;
    }
    src ^= 1;
    dst ^= 1;
  }
  int nv = n[src];
  if (nv < 3) return 0;
  // fan (clipper.cpp:75-89); only the first nv-2 triangles are counted/consumed (Appendix B #8)
  for (int t = 1; t <= nv - 2; ++t) {
    VsOut* o = out + (t - 1) * 3;
    o[0] = pool[src][0];
    if (is_front) { o[1] = pool[src][t]; o[2] = pool[src][t + 1]; }
    else { o[1] = pool[src][t + 1]; o[2] = pool[src][t]; }
  }
  return nv - 2;
}

// ---- viewport transform + project (shader.cpp:499-511, 116-134)
void viewport_project(DrawCtx& c, VsOut& v) {
  auto const& vp = c.d->viewport;
  float w = v.r[0][3];
  float invw = eq_eps(w, 0.0f) ? 1.0f : 1.0f / w;
  float px = v.r[0][0] * invw, py = v.r[0][1] * invw, pz = v.r[0][2] * invw;
  float ox = (vp.x + vp.w) * 0.5f;
  float oy = (vp.y + vp.h) * 0.5f;
  v.r[0][0] = (vp.w * 0.5f) * px + ox;
  v.r[0][1] = (vp.h * 0.5f) * -py + oy;
  v.r[0][2] = (vp.maxz - vp.minz) * pz + vp.minz;
  v.r[0][3] = invw;
  for (uint32_t i = 0; i < c.n_attrs; ++i)
    if (!(c.mods[i] & SLV_AM_NOPERSPECTIVE))
      for (int j = 0; j < 4; ++j) v.r[1 + i][j] *= invw;
}

// ---- rasterizer::compute_triangle_info (rasterizer.cpp:864-945)
void compute_triangle_info(DrawCtx& c, VsOut const verts[3], TriInfo& ti) {
  ti.valid = false;
  double dist[3];
  for (int i = 0; i < 3; ++i)  // summed in double (rasterizer.cpp:880-882 compiles to cvtss2sd + addsd)
    dist[i] = (double)std::fabs(verts[i].r[0][0]) + (double)std::fabs(verts[i].r[0][1]);
  int ro[3];
  if (dist[0] < dist[1]) ro[0] = (dist[0] < dist[2]) ? 0 : 2;
  else ro[0] = (dist[1] < dist[2]) ? 1 : 2;
  ro[1] = (ro[0] + 1) % 3;
  ro[2] = (ro[1] + 1) % 3;
  VsOut const* rv[3] = {&verts[ro[0]], &verts[ro[1]], &verts[ro[2]]};
  int nreg = 1 + (int)c.n_attrs;
  VsOut e01, e02;
  for (int r = 0; r < nreg; ++r) {
    e01.r[r] = sub4(rv[1]->r[r], rv[0]->r[r]);
    e02.r[r] = sub4(rv[2]->r[r], rv[0]->r[r]);
  }
  float area = e02.r[0][0] * e01.r[0][1] - e02.r[0][1] * e01.r[0][0];  // cross_prod2(e02.xy, e01.xy)
  if (eq_eps(area, 0.0f)) return;
  ti.front_face = area > 0.0f;
  float inv_area = 1.0f / area;
  float const* p0 = verts[0].r[0].v; float const* p1 = verts[1].r[0].v; float const* p2 = verts[2].r[0].v;
  ti.bbox[0] = std::min(std::min(p0[0], p1[0]), p2[0]);
  ti.bbox[1] = std::max(std::max(p0[0], p1[0]), p2[0]);
  ti.bbox[2] = std::min(std::min(p0[1], p1[1]), p2[1]);
  ti.bbox[3] = std::max(std::max(p0[1], p1[1]), p2[1]);
  for (int i = 0; i < 3; ++i) {
    float const* s = verts[i].r[0].v;
    float const* e = verts[(i + 1) % 3].r[0].v;
    ti.edge[i][0] = s[1] - e[1];
    ti.edge[i][1] = e[0] - s[0];
    ti.edge[i][2] = e[0] * s[1] - e[1] * s[0];
  }
  // compute_derivative_n (shader.cpp:413-449, SSE path)
  float e01x = e01.r[0][0], e01y = e01.r[0][1], e02x = e02.r[0][0], e02y = e02.r[0][1];
  for (int r = 0; r < nreg; ++r)
    for (int j = 0; j < 4; ++j) {
      float xd = e02.r[r][j] * e01y - e01.r[r][j] * e02y;
      float yd = e01.r[r][j] * e02x - e02.r[r][j] * e01x;
      ti.ddx.r[r][j] = xd * inv_area;
      ti.ddy.r[r][j] = yd * inv_area;
    }
  ti.v0 = *rv[0];
  ti.valid = true;
}

// ---- depth / stencil (framebuffer.cpp:96-256)
bool compare_f(uint32_t fn, float l, float r) {
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
bool compare_u(uint32_t fn, uint32_t l, uint32_t r) {
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
uint32_t stencil_op(uint32_t op, uint32_t ref, uint32_t cur) {  // framebuffer.cpp:136-166 incl. quirks
  switch (op) {
  case SLV_SOP_KEEP: return cur;
  case SLV_SOP_ZERO: return 0;
  case SLV_SOP_REPLACE: return ref;
  case SLV_SOP_INCR_SAT: return std::min<uint32_t>(0xFF, cur + 1);
  case SLV_SOP_DECR_SAT: return std::max<uint32_t>(0, cur - 1);
  case SLV_SOP_INVERT: return ~cur;
  case SLV_SOP_INCR_WRAP: return (cur + 1) & 0xFF;
  case SLV_SOP_DECR_WRAP: return (cur - 1 + 256) & 0xFF;
  }
  return cur;
}

inline bool depth_test(DrawCtx& c, float ps_depth, float cur) {
  return c.d->ds.depth_enable ? compare_f(c.d->ds.depth_func, ps_depth, cur) : true;
}
inline void read_ds(DrawCtx& c, uint8_t const* p, float& depth, uint32_t& stencil) {
  depth = 0.0f;
  stencil = 0;
  if (c.read_depth) memcpy(&depth, p, 4);
  if (c.read_stencil) { memcpy(&stencil, p + 4, 4); stencil &= c.read_mask; }
}
inline void write_ds(DrawCtx& c, uint8_t* p, float depth, uint32_t stencil, uint32_t mask) {
  if (c.write_depth) memcpy(p, &depth, 4);
  if (c.write_stencil) { uint32_t s = stencil & mask; memcpy(p + 4, &s, 4); }
}

// framebuffer::early_z_test (framebuffer.cpp:522-590); px_mask = covered samples
uint32_t early_z_test(DrawCtx& c, uint32_t x, uint32_t y, uint32_t px_mask, float depth, float const* aa) {
  uint32_t full = (1u << c.S) - 1;
  uint32_t mask = 0;
  if (px_mask == full && c.S == 1) {
    uint8_t* p = c.ds->addr(x, y, 0);
    float od; uint32_t os;
    read_ds(c, p, od, os);
    c.dev->traffic.z_tested += c.read_depth;
    if (depth_test(c, depth, od)) { write_ds(c, p, depth, 0, 0); c.dev->traffic.z_written += c.write_depth; return 1; }
    return 0;
  }
  for (uint32_t s = 0; s < c.S; ++s) {
    if (!(px_mask & (1u << s))) continue;
    uint8_t* p = c.ds->addr(x, y, s);
    float od; uint32_t os;
    read_ds(c, p, od, os);
    float nd = aa[s] + depth;
    c.dev->traffic.z_tested += c.read_depth;
    if (depth_test(c, nd, od)) { mask |= 1u << s; write_ds(c, p, nd, 0, 0); c.dev->traffic.z_written += c.write_depth; }
  }
  return mask;
}

// ---- pixel shaders ---------------------------------------------------------------------------------
struct PsQuad {
  VsOut px[4];
  float lod[MAX_REGS];
  uint32_t lod_flag;
};

// cpp_pixel_shader::tex2d (cpp_pixel_shader.cpp:23-31): LOD once per quad per register
V4 ps_tex2d(DrawCtx& c, PsQuad& q, int pix, int samp, uint32_t reg) {
  Texture const& t = *c.sampler_tex[samp];
  auto const& d = c.samplers[samp]->d;
  if (!(q.lod_flag & (1u << reg))) {
    V4 const& a0 = q.px[0].r[1 + reg];
    V4 const& a1 = q.px[1].r[1 + reg];
    V4 const& a2 = q.px[2].r[1 + reg];
    q.lod[reg] = calc_lod_2d(t, d, a1[0] - a0[0], a1[1] - a0[1], a2[0] - a0[0], a2[1] - a0[1]);
    q.lod_flag |= 1u << reg;
  }
  V4 const& a = q.px[pix].r[1 + reg];
  return sample_impl(t, d, a[0], a[1], q.lod[reg], nullptr);
}

// tex2D of a SASL pixel shader: sample_2d_grad with the quad's derivatives (cg_impl.cpp:902-909 -> sampler_api.cpp:13-18).
// sasl: ddx per quad row, ddy per quad column (cgs_simd.cpp:275-313); else q1 - q0, q2 - q0 (cpp_pixel_shader.cpp:13-19)
V4 ps_tex2d_grad(DrawCtx& c, PsQuad& q, int pix, uint32_t reg, bool sasl) {
  int const pi = sasl ? pix : 0;
  V4 const& xl = q.px[pi & ~1].r[1 + reg]; V4 const& xh = q.px[pi | 1].r[1 + reg];
  V4 const& yl = q.px[pi & ~2].r[1 + reg]; V4 const& yh = q.px[pi | 2].r[1 + reg];
  V4 const& a = q.px[pix].r[1 + reg];
  return sample_2d_grad(*c.sampler_tex[0], c.samplers[0]->d, a[0], a[1], xh[0] - xl[0], xh[1] - xl[1], yh[0] - yl[0],
                        yh[1] - yl[1], 0.0f);
}

bool run_ps(DrawCtx& c, PsQuad& q, int pix, V4& color) {
  auto const& ps = c.d->ps;
  VsOut const& in = q.px[pix];
  switch (ps.program) {
  case SLV_PS_ATTR0_COLOR: color = in.r[1]; return true;
  case SLV_PS_DISCARD_ALL: color = in.r[1]; return false;
  case SLV_PS_HEIGHT_COLOR: {  // VertexTextureFetch.cpp:70-113
    float height = in.r[1][0];
    static const float colors[6][4] = {{0.0f, 0.0f, 0.5f, 1.0f}, {0.7f, 0.6f, 0.0f, 1.0f}, {0.45f, 0.38f, 0.26f, 1.0f},
                                       {0.0f, 0.7f, 0.8f, 1.0f}, {0.9f, 0.9f, 1.0f, 1.0f}, {0.9f, 0.9f, 1.0f, 1.0f}};
    static const float boundary[6] = {0.0f, 0.62f, 0.75f, 0.88f, 1.0f, 1.0f};
    int lower = -1;
    for (int i = 0; i < 5; ++i) {
      if (height < boundary[i]) break;
      lower = i;
    }
    if (lower == -1) {
      color = mk4(colors[0][0], colors[0][1], colors[0][2], colors[0][3]);
    } else {
      float lv = boundary[lower], interval = boundary[lower + 1] - lv;
      float t = (height - lv) / interval;
      for (int k = 0; k < 4; ++k) color[k] = colors[lower][k] + (colors[lower + 1][k] - colors[lower][k]) * t;  // eflib::lerp
    }
    return true;
  }
  case SLV_PS_LIGHTS3: {  // ColorizedTriangle.cpp:55-92
    float const* l0 = in.r[2].v; float const* l1 = in.r[3].v; float const* l2 = in.r[4].v;
    float const* norm = in.r[1].v;
    float i0 = 1.0f / length3(l0), i1 = 1.0f / length3(l1), i2 = 1.0f / length3(l2);
    float nn[3], n0[3], n1[3], n2[3];
    normalize3(nn, norm);
    for (int k = 0; k < 3; ++k) { n0[k] = l0[k] * i0; n1[k] = l1[k] * i1; n2[k] = l2[k] * i2; }
    float r0 = dot3(nn, n0), r1 = dot3(nn, n1), r2 = dot3(nn, n2);
    const float A[4] = {0.7f, 0.1f, 0.3f, 1.0f}, B[4] = {0.1f, 0.3f, 0.7f, 1.0f}, Cc[4] = {0.3f, 0.7f, 0.1f, 1.0f};
    for (int k = 0; k < 4; ++k) {
      float a = ((A[k] * r0) * i0) * i0;
      float b = ((B[k] * r1) * i1) * i1;
      float cc = ((Cc[k] * r2) * i2) * i2;
      color[k] = clampf((a + b) + cc, 0.0f, 1.0f);
    }
    color[3] = 1.0f;
    return true;
  }
  case SLV_PS_TEX_ALPHA: {
    auto u = (slv_ps_tex_alpha_uniforms const*)ps.uniforms;
    color = ps_tex2d(c, q, pix, 0, u->reg);
    color[3] = u->alpha;
    return true;
  }
  case SLV_PS_TEX_GRAD_ALPHA: {
    auto u = (slv_ps_tex_alpha_uniforms const*)ps.uniforms;
    color = ps_tex2d_grad(c, q, pix, u->reg, u->sasl_derivatives != 0);
    color[3] = u->alpha;
    return true;
  }
  case SLV_PS_SSM_DRAW: {  // draw_cpp_ps::shader_prog, samples/StandardShadowMap/StandardShadowMap.cpp:84-137
    auto u = (slv_ps_ssm_draw_uniforms const*)ps.uniforms;
    const float esm_constant = 25000.0f;                                                                   // :33
    static const float gaussian_weights[9] = {0.027681f, 0.111014f, 0.027681f, 0.111014f, 0.445213f,       // :34-42
                                              0.111014f, 0.027681f, 0.111014f, 0.027681f};
    float occlusion = 0.0f;
    if (u->has_depth_sampler) {
      V4 const& a4 = in.r[5];
      float lis[3] = {a4[0] / a4[3], a4[1] / a4[3], a4[2] / a4[3]};   // vec3 / float: per-component division (vector_generic.h:340)
      float cx = (lis[0] + 1.0f) * 0.5f, cy = (1.0f - (lis[1] + 1.0f) * 0.5f);
      float sm_offset = 1 / 512.0f;
      const float ox[3] = {-sm_offset, 0.0f, +sm_offset};
      float shadow_depth[9];
      for (int i = 0; i < 9; ++i)  // tex2dlod(s, coord) = s.sample(x, y, lod) (cpp_pixel_shader.cpp:33-35, sampler.cpp:767-769)
        shadow_depth[i] = sample_impl(*c.sampler_tex[1], c.samplers[1]->d, cx + ox[i % 3], cy + ox[i / 3], 0.0f, nullptr)[0];
      float occluder = 0.0f;
      for (int i = 1; i < 9; ++i) occluder += gaussian_weights[i] * std::exp(esm_constant * (shadow_depth[i] - shadow_depth[0]));
      occluder += gaussian_weights[0];
      occluder = std::log(occluder);
      occluder += esm_constant * shadow_depth[0];
      occlusion = clampf(std::exp(occluder - esm_constant * lis[2]), 0.0f, 1.0f);
    }
    V4 tex = mk4(1, 1, 1, 1);
    if (u->has_tex_sampler) tex = ps_tex2d(c, q, pix, 0, 0);
    float n[3], l[3], e[3];
    normalize3(n, in.r[2].v);
    normalize3(l, in.r[3].v);
    normalize3(e, in.r[4].v);
    float illum_diffuse = clampf(dot3(l, n), 0.0f, 1.0f);
    float k2 = 2.0f * dot3(l, n);  // reflect3(i, n) = i - n * (2 * dot(i, n))  (eflib/src/math.cpp:85-87)
    float r[3] = {-(l[0] - n[0] * k2), -(l[1] - n[1] * k2), -(l[2] - n[2] * k2)};
    float illum_specular = clampf(dot3(r, e), 0.0f, 1.0f);
    float sp = (float)std::pow((double)illum_specular, (double)u->shininess);  // pow(float, int): both promoted to double
    for (int k = 0; k < 3; ++k)
      color[k] = tex[k] * (u->ambient[k] + (u->diffuse[k] * illum_diffuse + u->specular[k] * sp) * occlusion);
    color[3] = 1.0f;
    return true;
  }
  case SLV_PS_SPONZA_GRAD: {  // Sponza.cpp:117-136 with the diffuse fetch of a SASL tex2D: sample_2d_grad (cg_impl.cpp:902-909)
    auto u = (slv_ps_sponza_grad_uniforms const*)ps.uniforms;
    V4 diff = mk4(1, 1, 1, 1);
    if (u->has_sampler) diff = ps_tex2d_grad(c, q, pix, 0, u->sasl_derivatives != 0);
    float n[3], l[3];
    normalize3(n, in.r[2].v);
    normalize3(l, in.r[3].v);
    float illum = clampf(dot3(l, n), 0.0f, 1.0f);
    for (int k = 0; k < 4; ++k) color[k] = diff[k] * illum;
    color[3] = 1.0f;
    return true;
  }
  case SLV_PS_SPONZA: {  // Sponza.cpp:117-136 (the dead specular/ambient terms are dropped)
    auto u = (slv_ps_sponza_uniforms const*)ps.uniforms;
    V4 diff = mk4(1, 1, 1, 1);
    if (u->has_sampler) diff = ps_tex2d(c, q, pix, 0, 0);
    float n[3], l[3];
    normalize3(n, in.r[2].v);
    normalize3(l, in.r[3].v);
    float illum = clampf(dot3(l, n), 0.0f, 1.0f);
    for (int k = 0; k < 4; ++k) color[k] = diff[k] * illum;
    color[3] = 1.0f;
    return true;
  }
  }
  color = mk4(0, 0, 0, 0);
  return true;
}

// ---- output merger (framebuffer.cpp:445-520) ----------------------------------------------------------
void blend(DrawCtx& c, uint32_t x, uint32_t y, uint32_t s, V4 const& src) {
  Surface* t0 = c.colors.size() > 0 ? c.colors[0] : nullptr;
  if (t0) {
    ++c.dev->traffic.c_written;
    if (c.d->bs.program == SLV_BS_LERP_SRC_ALPHA) ++c.dev->traffic.c_read;
  }
  switch (c.d->bs.program) {
  case SLV_BS_REPLACE:
    if (t0) from_rgba32f(t0->fmt, t0->addr(x, y, s), src);
    break;
  case SLV_BS_LERP_SRC_ALPHA:
    if (t0) {
      V4 dst = to_rgba32f(t0->fmt, t0->addr(x, y, s));
      V4 r;
      for (int j = 0; j < 4; ++j) r[j] = dst[j] + (src[j] - dst[j]) * src[3];
      from_rgba32f(t0->fmt, t0->addr(x, y, s), r);
    }
    break;
  case SLV_BS_REPLACE_AND_COUNT: {
    if (t0) from_rgba32f(t0->fmt, t0->addr(x, y, s), src);
    Surface* t1 = c.colors.size() > 1 ? c.colors[1] : nullptr;
    if (t1) {
      V4 v = to_rgba32f(t1->fmt, t1->addr(x, y, s));
      v[0] += 1.0f;
      from_rgba32f(t1->fmt, t1->addr(x, y, s), v);
    }
  } break;
  }
}

void render_sample(DrawCtx& c, uint32_t x, uint32_t y, uint32_t s, V4 const& color, float depth, bool front) {
  if (c.early_z) { blend(c, x, y, s, color); return; }
  uint8_t* p = c.ds->addr(x, y, s);
  float od; uint32_t os;
  read_ds(c, p, od, os);
  bool dp = depth_test(c, depth, od);
  c.dev->traffic.z_tested += (c.read_depth || c.read_stencil) ? 1 : 0;
  auto const& ds = c.d->ds;
  auto const& face = front ? ds.front_face : ds.back_face;
  bool sp = ds.stencil_enable ? compare_u(face.stencil_func, c.stencil_ref, os) : true;
  if (dp && sp) {
    // stencil_operation index (!front)*3 + !depth_pass + stencil_pass == the *pass* op (Appendix B #4)
    uint32_t ns = ds.stencil_enable ? stencil_op(face.stencil_pass_op, c.stencil_ref, os) : os;
    blend(c, x, y, s, color);
    write_ds(c, p, depth, ns, c.write_mask);
    c.dev->traffic.z_written += (c.write_depth || c.write_stencil) ? 1 : 0;
  }
}

// rasterizer::draw_full_quad / draw_quad (rasterizer.cpp:1245-1421); mask: 4 px x S bits
void draw_quad(DrawCtx& c, TriInfo const& ti, float const* aa, uint32_t left, uint32_t top, uint32_t const pxmask_in[4],
               bool full) {
  uint32_t fullm = (1u << c.S) - 1;
  int nreg = 1 + (int)c.n_attrs;
  PsQuad q;
  float dx = 0.5f + left - ti.v0.r[0][0];
  float dy = 0.5f + top - ti.v0.r[0][1];
  // step_2d_unproj_pos_quad (shader.cpp:257-287)
  for (int j = 0; j < 4; ++j) {
    float p00 = ti.v0.r[0][j] + (ti.ddx.r[0][j] * dx + ti.ddy.r[0][j] * dy);
    float p01 = p00 + ti.ddx.r[0][j];
    float p10 = p00 + ti.ddy.r[0][j];
    float p11 = p01 + ti.ddy.r[0][j];
    q.px[0].r[0][j] = p00; q.px[1].r[0][j] = p01; q.px[2].r[0][j] = p10; q.px[3].r[0][j] = p11;
  }
  float depth[4] = {q.px[0].r[0][2], q.px[1].r[0][2], q.px[2].r[0][2], q.px[3].r[0][2]};
  uint32_t tested[4];
  bool any = false;
  for (int i = 0; i < 4; ++i) {
    tested[i] = pxmask_in[i];
    if (c.early_z) {
      tested[i] = pxmask_in[i] == 0 ? 0 : early_z_test(c, left + (i & 1), top + (i >> 1), pxmask_in[i], depth[i], aa);
    }
    any |= tested[i] != 0;
  }
  if (!any) return;

  if (full || !c.has_centroid) {
    // step_2d_unproj_attr_n_quad (shader.cpp:289-367)
    float inv_w[4];
    for (int i = 0; i < 4; ++i) inv_w[i] = 1.0f / q.px[i].r[0][3];
    for (int r = 1; r < nreg; ++r) {
      uint32_t m = c.mods[r - 1];
      for (int j = 0; j < 4; ++j) {
        float a00, a01, a10, a11;
        if (m & SLV_AM_NOINTERPOLATION) {
          a00 = a01 = a10 = a11 = ti.v0.r[r][j];
        } else {
          a00 = ti.v0.r[r][j] + (ti.ddx.r[r][j] * dx + ti.ddy.r[r][j] * dy);
          a01 = a00 + ti.ddx.r[r][j];
          a10 = a00 + ti.ddy.r[r][j];
          a11 = a01 + ti.ddy.r[r][j];
        }
        if (!(m & SLV_AM_NOPERSPECTIVE)) { a00 *= inv_w[0]; a01 *= inv_w[1]; a10 *= inv_w[2]; a11 *= inv_w[3]; }
        q.px[0].r[r][j] = a00; q.px[1].r[r][j] = a01; q.px[2].r[r][j] = a10; q.px[3].r[r][j] = a11;
      }
    }
  } else {
    // centroid path (rasterizer.cpp:1366-1397) + step_2d_unproj_attr_n (shader.cpp:208-255)
    for (int i = 0; i < 4; ++i) {
      float pdx = dx + (i & 1), pdy = dy + ((i & 2) >> 1);
      uint32_t pm = pxmask_in[i];
      if (pm != fullm && pm != 0) {
        float cx = 0, cy = 0;
        int n = 0;
        for (uint32_t s = 0; s < c.S; ++s)
          if (pm & (1u << s)) { cx += c.sp[s][0]; cy += c.sp[s][1]; ++n; }
        float inv = 1 / (float)n;  // vec2 /= n -> *= 1/n
        cx *= inv; cy *= inv;
        pdx += cx - 0.5f;
        pdy += cy - 0.5f;
      }
      float inv_w = 1.0f / q.px[i].r[0][3];
      for (int r = 1; r < nreg; ++r) {
        uint32_t m = c.mods[r - 1];
        for (int j = 0; j < 4; ++j) {
          float a = (m & SLV_AM_NOINTERPOLATION) ? ti.v0.r[r][j]
                                                : ti.v0.r[r][j] + (ti.ddx.r[r][j] * pdx + ti.ddy.r[r][j] * pdy);
          if (!(m & SLV_AM_NOPERSPECTIVE)) a *= inv_w;
          q.px[i].r[r][j] = a;
        }
      }
    }
  }
  c.ps_invocations += 4;

  // cpp_pixel_shader::execute (cpp_pixel_shader.cpp:62-74)
  q.lod_flag = 0;
  V4 pso[4];
  bool keep[4];
  if (c.ps_module) {  // a SASL pixel shader compiled for the host runs the whole quad at once (derivatives need all four pixels)
    float attrs[4][5][4] = {}, colors[4][4] = {};
    int kept[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i)
      for (int r = 1; r < nreg && r <= 5; ++r) for (int j = 0; j < 4; ++j) attrs[i][r - 1][j] = q.px[i].r[r][j];
    c.ps_module->ps_quad(&attrs[0][0][0], c.d->ps.uniforms, c.d->ps.uniform_bytes, &c.host_api, &colors[0][0], kept);
    for (int i = 0; i < 4; ++i) { pso[i] = mk4(colors[i][0], colors[i][1], colors[i][2], colors[i][3]); keep[i] = kept[i] != 0; }
  } else {
    for (int i = 0; i < 4; ++i) keep[i] = run_ps(c, q, i, pso[i]);
  }
  bool any_out = false;
  uint32_t final_mask[4];
  for (int i = 0; i < 4; ++i) { final_mask[i] = keep[i] ? tested[i] : 0; any_out |= final_mask[i] != 0; }
  // full quad: `if (quad_mask != 0)` after &= ; partial quad: tests the pre-Z mask (always non-zero)
  if (full ? any_out : true) {
    c.backend_input_pixels += 4;
    for (int i = 0; i < 4; ++i) {
      uint32_t pm = final_mask[i];
      if (!pm) continue;
      uint32_t x = left + (i & 1), y = top + (i >> 1);
      if (c.S == 1) {
        render_sample(c, x, y, 0, pso[i], depth[i], ti.front_face);
      } else {
        for (uint32_t s = 0; s < c.S; ++s)
          if (pm & (1u << s)) render_sample(c, x, y, s, pso[i], depth[i] + aa[s], ti.front_face);
      }
    }
  }
}

// rasterizer::draw_partial_tile (rasterizer.cpp:298-439)
void draw_partial_tile(DrawCtx& c, TriInfo const& ti, float const* aa, int left, int top) {
  uint32_t fullm = (1u << c.S) - 1;
  float left_f = (float)left, top_f = (float)top;
  float ev[3];
  for (int e = 0; e < 3; ++e) ev[e] = ti.edge[e][2] - (left_f * ti.edge[e][0] + top_f * ti.edge[e][1]);
  uint32_t pm[16] = {0};
  for (uint32_t s = 0; s < c.S; ++s)
    for (int iy = 0; iy < 4; ++iy) {
      float fy = c.sp[s][1] + (float)iy;
      for (int ix = 0; ix < 4; ++ix) {
        float fx = c.sp[s][0] + (float)ix;
        bool rej = false;
        for (int e = 0; e < 3; ++e) rej |= (fx * ti.edge[e][0] + fy * ti.edge[e][1]) < ev[e];
        if (!rej) pm[iy * 4 + ix] |= 1u << s;
      }
    }
  for (int quad = 0; quad < 4; ++quad) {
    int qx = (quad & 1) << 1, qy = (quad & 2);
    int st = qx | (qy << 2);
    uint32_t m[4] = {pm[st], pm[st + 1], pm[st + 4], pm[st + 5]};
    if (!(m[0] | m[1] | m[2] | m[3])) continue;
    uint32_t x = left + qx, y = top + qy;
    // The reference does not clamp partial blocks to the target (Appendix B #16: targets are multiples of 4)
    if (x + 2 > (uint32_t)c.target_w || y + 2 > (uint32_t)c.target_h) continue;
    bool full = (m[0] == fullm && m[1] == fullm && m[2] == fullm && m[3] == fullm);
    draw_quad(c, ti, aa, x, y, m, full);
  }
}

void draw_full_tile(DrawCtx& c, TriInfo const& ti, float const* aa, int l, int t, int r, int b) {
  uint32_t fullm = (1u << c.S) - 1;
  uint32_t m[4] = {fullm, fullm, fullm, fullm};
  for (int y = t; y < b; y += 2)
    for (int x = l; x < r; x += 2) draw_quad(c, ti, aa, x, y, m, true);
}

// rasterizer::rasterize_triangle + subdivide_tile (rasterizer.cpp:617-773, 441-602)
void rasterize_triangle(DrawCtx& c, TriInfo const& ti, uint32_t full, int tile_x, int tile_y) {
  const int TS = SLV_TILE_SIZE;
  float vpx = (float)(tile_x * TS), vpy = (float)(tile_y * TS);
  bool mark_x[3], mark_y[3];
  for (int e = 0; e < 3; ++e) { mark_x[e] = ti.edge[e][0] > 0; mark_y[e] = ti.edge[e][1] > 0; }
  float x_min = ti.bbox[0] - vpx, x_max = ti.bbox[1] - vpx, y_min = ti.bbox[2] - vpy, y_max = ti.bbox[3] - vpy;

  std::vector<uint32_t> regions[2];
  regions[0].push_back(full << 31);
  int src = 0, dst = 1;
  int vpleft0 = fast_floori(vpx), vptop0 = fast_floori(vpy);
  uint32_t sub_w = fast_floori((float)TS), sub_h = fast_floori((float)TS);
  float step_x[3], step_y[3], rej_to_acc[3], e_value[3], part_e[3];
  for (int e = 0; e < 3; ++e) {
    step_x[e] = TS * ti.edge[e][0];
    step_y[e] = TS * ti.edge[e][1];
    rej_to_acc[e] = -std::fabs(step_x[e]) - std::fabs(step_y[e]);
    part_e[e] = mark_x[e] * TS * ti.edge[e][0] + mark_y[e] * TS * ti.edge[e][1];
    e_value[e] = ti.edge[e][2] - part_e[e];
  }
  float aa[4] = {0, 0, 0, 0};
  if (c.S > 1)
    for (uint32_t s = 0; s < c.S; ++s)
      aa[s] = (c.sp[s][0] - 0.5f) * ti.ddx.r[0][2] + (c.sp[s][1] - 0.5f) * ti.ddy.r[0][2];
  float tvr = c.target_w + 0.0f, tvb = c.target_h + 0.0f;

  while (!regions[src].empty()) {
    regions[dst].clear();
    sub_w /= 4;
    sub_h /= 4;
    for (int e = 0; e < 3; ++e) {
      step_x[e] *= 0.25f; step_y[e] *= 0.25f; rej_to_acc[e] *= 0.25f; part_e[e] *= 0.25f;
      e_value[e] = ti.edge[e][2] - part_e[e];
    }
    for (uint32_t packed : regions[src]) {
      uint32_t rx = packed & 0xFF, ry = (packed >> 8) & 0xFF;
      bool is_full = (packed >> 31) != 0;
      int vpleft = (int)std::max(0U, (unsigned)(vpleft0 + rx));
      int vptop = (int)std::max(0U, (unsigned)(vptop0 + ry));
      if (vpleft >= tvr || vptop >= tvb) continue;
      if (is_full) {
        int vpright = (int)std::min<uint32_t>(vpleft0 + rx + sub_w * 4, (uint32_t)tvr);
        int vpbottom = (int)std::min<uint32_t>(vptop0 + ry + sub_h * 4, (uint32_t)tvb);
        draw_full_tile(c, ti, aa, vpleft, vptop, vpright, vpbottom);
      } else if (sub_w <= 1 && sub_h <= 1) {
        draw_partial_tile(c, ti, aa, vpleft, vptop);
      } else {
        // subdivide_tile, scalar twin (rasterizer.cpp:573-601); left/top are converted to float
        float ev1[3];
        for (int e = 0; e < 3; ++e)
          ev1[e] = e_value[e] - ((float)vpleft * ti.edge[e][0] + (float)vptop * ti.edge[e][1]);
        for (int ty = 0; ty < 4; ++ty) {
          uint32_t y = ry + sub_h * ty;
          for (int tx = 0; tx < 4; ++tx) {
            uint32_t x = rx + sub_w * tx;
            // SSE path (rasterizer.cpp:554-559): reject if x_min >= x+w, x_max < x, y_min >= y+h, y_max < y
            if ((x_min >= (float)(int)(x + sub_w)) || (x_max < (float)(int)x) ||
                (y_min >= (float)(int)(y + sub_h)) || (y_max < (float)(int)y))
              continue;
            int rejection = 0, acception = 1;
            for (int e = 0; e < 3; ++e) {
              float step = step_x[e] * (float)tx + step_y[e] * (float)ty;
              rejection |= (step < ev1[e]);
              acception &= !((step + rej_to_acc[e]) < ev1[e]);
            }
            if (!rejection) regions[dst].push_back(x + (y << 8) + ((uint32_t)acception << 31));
          }
        }
      }
    }
    std::swap(src, dst);
  }
}

slv_result do_draw(Device& dev, slv_draw_desc const& d) {
  DrawCtx c;
  c.dev = &dev;
  c.d = &d;
  if (d.topology != SLV_TOPO_TRIANGLE_LIST && d.topology != SLV_TOPO_TRIANGLE_STRIP) return SLV_FAILED;
  c.n_attrs = vs_num_attrs(d.vs);
  if (d.vs.program & 0x80000000u) {
    auto r = dev.get(d.vs.program & 0x7FFFFFFFu, Resource::MODULE);
    if (!r || r->mod.stage != SLV_STAGE_VS) return SLV_INVALID_PARAMETER;
    c.vs_module = &r->mod;
    c.n_attrs = r->mod.n_vs_output_attrs;
  }
  if (d.ps.program & 0x80000000u) {
    auto r = dev.get(d.ps.program & 0x7FFFFFFFu, Resource::MODULE);
    if (!r || r->mod.stage != SLV_STAGE_PS) return SLV_INVALID_PARAMETER;
    c.ps_module = &r->mod;
  }
  c.host_api.ctx = &c;
  c.host_api.sample_grad = [](void* ctx, int slot, float u, float v, float dudx, float dvdx, float dudy, float dvdy, float bias, float* rgba) {
    auto& dc = *static_cast<DrawCtx*>(ctx);
    V4 r = mk4(0, 0, 0, 0);
    if (slot >= 0 && slot < SLV_MAX_SAMPLERS && dc.sampler_tex[slot]) r = sample_2d_grad(*dc.sampler_tex[slot], dc.samplers[slot]->d, u, v, dudx, dvdx, dudy, dvdy, bias);
    for (int j = 0; j < 4; ++j) rgba[j] = r[j];
  };
  c.host_api.sample_lod = [](void* ctx, int slot, float u, float v, float lod, float* rgba) {
    auto& dc = *static_cast<DrawCtx*>(ctx);
    V4 r = mk4(0, 0, 0, 0);
    if (slot >= 0 && slot < SLV_MAX_SAMPLERS && dc.sampler_tex[slot]) r = sample_impl(*dc.sampler_tex[slot], dc.samplers[slot]->d, u, v, lod, nullptr);
    for (int j = 0; j < 4; ++j) rgba[j] = r[j];
  };
  c.host_api.vs_sample_lod = [](void* ctx, float u, float v, float lod, float* rgba) {
    auto& dc = *static_cast<DrawCtx*>(ctx);
    V4 r = mk4(0, 0, 0, 0);
    if (dc.vs_sampler_tex) r = sample_impl(*dc.vs_sampler_tex, dc.vs_sampler->d, u, v, lod, nullptr);
    for (int j = 0; j < 4; ++j) rgba[j] = r[j];
  };
  if (c.n_attrs > SLV_MAX_VS_OUTPUT_ATTRS) return SLV_INVALID_PARAMETER;
  c.has_centroid = false;
  for (uint32_t i = 0; i < SLV_MAX_VS_OUTPUT_ATTRS; ++i) {
    c.mods[i] = d.vs_attr_modifiers[i] ? d.vs_attr_modifiers[i] : (uint32_t)SLV_AM_LINEAR;
    if (i < c.n_attrs && (c.mods[i] & SLV_AM_CENTROID)) c.has_centroid = true;
  }
  // render targets (renderer_impl.cpp:159-238)
  c.ds = nullptr;
  c.target_w = c.target_h = FLT_MAX;
  uint32_t S = 0;
  if (d.n_color_targets >= SLV_MAX_RENDER_TARGETS) return SLV_FAILED;
  for (uint32_t i = 0; i < d.n_color_targets; ++i) {
    auto r = dev.get(d.color_targets[i], Resource::TEXTURE);
    Surface* s = r ? &r->tex.levels[0] : nullptr;
    c.colors.push_back(s);
    if (s) {
      c.target_w = std::min((float)s->w, c.target_w);
      c.target_h = std::min((float)s->h, c.target_h);
      if (S == 0) S = s->samples; else if (S != s->samples) return SLV_FAILED;
    }
  }
  if (d.ds_target) {
    auto r = dev.get(d.ds_target, Resource::TEXTURE);
    if (!r || r->tex.fmt != SLV_PF_RG32F) return SLV_FAILED;
    c.ds = &r->tex.levels[0];
    if (d.n_color_targets == 0) { S = c.ds->samples; c.target_w = (float)c.ds->w; c.target_h = (float)c.ds->h; }
    if ((float)c.ds->w < c.target_w || (float)c.ds->h < c.target_h || c.ds->samples != S) return SLV_FAILED;
  }
  if (d.n_color_targets == 0 && !c.ds) return SLV_FAILED;
  if (S != 1 && S != 2 && S != 4) return SLV_INVALID_PARAMETER;
  c.S = S;
  switch (S) {  // rasterizer.cpp:1087-1103
  case 1: c.sp[0][0] = 0.5f; c.sp[0][1] = 0.5f; break;
  case 2: c.sp[0][0] = 0.25f; c.sp[0][1] = 0.25f; c.sp[1][0] = 0.75f; c.sp[1][1] = 0.75f; break;
  case 4:
    c.sp[0][0] = 0.375f; c.sp[0][1] = 0.125f; c.sp[1][0] = 0.875f; c.sp[1][1] = 0.375f;
    c.sp[2][0] = 0.125f; c.sp[2][1] = 0.625f; c.sp[3][0] = 0.625f; c.sp[3][1] = 0.875f;
    break;
  }
  // ds read/write selection (framebuffer.cpp:358-425)
  auto const& ds = d.ds;
  c.read_depth = c.read_stencil = c.write_depth = c.write_stencil = false;
  if (c.ds) {
    if (ds.depth_enable) {
      if (ds.depth_func != SLV_CMP_NEVER && ds.depth_func != SLV_CMP_ALWAYS) c.read_depth = true;
      if (ds.depth_write_mask && ds.depth_func != SLV_CMP_NEVER) c.write_depth = true;
    }
    c.read_stencil = c.write_stencil = ds.stencil_enable != 0;
  }
  c.early_z = !ds.stencil_enable;
  c.read_mask = ds.stencil_read_mask & 0xFF;
  c.write_mask = ds.stencil_write_mask & 0xFF;
  c.stencil_ref = ds.stencil_enable ? ((uint32_t)d.stencil_ref & c.read_mask) : 0;
  for (int i = 0; i < SLV_MAX_SAMPLERS; ++i) {
    c.samplers[i] = nullptr;
    c.sampler_tex[i] = nullptr;
    if (d.ps.samplers[i]) {
      auto r = dev.get(d.ps.samplers[i], Resource::SAMPLER);
      if (!r) return SLV_INVALID_PARAMETER;
      c.samplers[i] = &r->samp;
      c.sampler_tex[i] = sampler_texture(dev, r->samp);
    }
  }
  if (d.ps.program == SLV_PS_SSM_DRAW) {
    auto u = (slv_ps_ssm_draw_uniforms const*)d.ps.uniforms;
    if (d.ps.uniform_bytes < sizeof(slv_ps_ssm_draw_uniforms) || vs_num_attrs(d.vs) < 5) return SLV_INVALID_PARAMETER;
    if ((u->has_tex_sampler && !c.sampler_tex[0]) || (u->has_depth_sampler && !c.sampler_tex[1])) return SLV_INVALID_PARAMETER;
  }
  if (d.vs.program == SLV_VS_TERRAIN_VTF || (c.vs_module && d.vs.samplers[0])) {
    auto r = dev.get(d.vs.samplers[0], Resource::SAMPLER);
    if (!r) return SLV_INVALID_PARAMETER;
    c.vs_sampler = &r->samp;
    c.vs_sampler_tex = sampler_texture(dev, r->samp);
    if (!c.vs_sampler_tex) return SLV_INVALID_PARAMETER;
  }

  auto const& vp = d.viewport;
  size_t tile_x_count = static_cast<size_t>(vp.w + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE;  // rasterizer.cpp:1106
  size_t tile_y_count = static_cast<size_t>(vp.h + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE;

  // phase 1: VS + clip (geom_setup_engine.cpp:84-124); post-transform results memoised per index
  std::vector<TriInfo> tris;
  {
    struct CacheItem { uint32_t idx; bool valid; };
    CacheItem tls_cache[128];  // counts vs_invocations like one tls_vertex_cache (default_vertex_cache.cpp:354-390)
    for (auto& it : tls_cache) it = {0xFFFFFFFFu, false};
    std::vector<VsOut> memo;
    std::vector<uint8_t> have;
    for (uint32_t p = 0; p < d.prim_count; ++p) {
      uint32_t idx[3];
      if (!fetch_indices(c, p, idx)) return SLV_INVALID_PARAMETER;
      VsOut tri[3];
      dev.stats.ia_vertices += 3;
      for (int i = 0; i < 3; ++i) {
        auto& it = tls_cache[idx[i] % 128];
        if (!(it.valid && it.idx == idx[i])) { ++dev.stats.vs_invocations; it = {idx[i], true}; }
        if (idx[i] >= memo.size()) { memo.resize((size_t)idx[i] + 1024); have.resize(memo.size(), 0); }
        if (!have[idx[i]]) {
          V4 in[SLV_MAX_VS_INPUT_ATTRS];
          fetch_vertex(c, idx[i], in);
          run_vs(c, in, memo[idx[i]]);
          have[idx[i]] = 1;
        }
        tri[i] = memo[idx[i]];
      }
      ++dev.stats.cinvocations;
      VsOut outv[9];
      int nt = clip_triangle(c, tri, outv);
      for (int t = 0; t < nt; ++t) {
        VsOut v[3] = {outv[t * 3], outv[t * 3 + 1], outv[t * 3 + 2]};
        for (int k = 0; k < 3; ++k) viewport_project(c, v[k]);  // phase 3
        TriInfo ti;
        compute_triangle_info(c, v, ti);  // phase 4
        tris.push_back(ti);
      }
    }
  }
  dev.stats.ia_primitives += d.prim_count;
  dev.stats.cprimitives += tris.size();

  // phase 4: binning (rasterizer.cpp:775-862)
  std::vector<std::vector<uint32_t>> bins(tile_x_count * tile_y_count);
  for (size_t i = 0; i < tris.size(); ++i) {
    TriInfo const& ti = tris[i];
    if (!ti.valid) continue;
    const int TS = SLV_TILE_SIZE;
    int sx = std::min(fast_floori(std::max(0.0f, ti.bbox[0]) / TS), (int)tile_x_count);
    int sy = std::min(fast_floori(std::max(0.0f, ti.bbox[2]) / TS), (int)tile_y_count);
    int ex = std::min(fast_ceili(std::max(0.0f, ti.bbox[1]) / TS) + 1, (int)tile_x_count);
    int ey = std::min(fast_ceili(std::max(0.0f, ti.bbox[3]) / TS) + 1, (int)tile_y_count);
    if ((sx + 1 == ex) && (sy + 1 == ey)) {
      bins[sy * tile_x_count + sx].push_back((uint32_t)i << 1);
    } else {
      bool mark_x[3], mark_y[3];
      float step_x[3], step_y[3], rej_to_acc[3];
      for (int e = 0; e < 3; ++e) {
        mark_x[e] = ti.edge[e][0] > 0; mark_y[e] = ti.edge[e][1] > 0;
        step_x[e] = TS * ti.edge[e][0];
        step_y[e] = TS * ti.edge[e][1];
        rej_to_acc[e] = -std::fabs(step_x[e]) - std::fabs(step_y[e]);
      }
      for (int y = sy; y < ey; ++y)
        for (int x = sx; x < ex; ++x) {
          int rejection = 0, acceptance = 1;
          for (int e = 0; e < 3; ++e) {
            float ev = ti.edge[e][2] - (static_cast<float>(x + mark_x[e]) * TS * ti.edge[e][0] +
                                        static_cast<float>(y + mark_y[e]) * TS * ti.edge[e][1]);
            rejection |= (0 < ev);
            acceptance &= (rej_to_acc[e] >= ev);
          }
          if (!rejection) bins[y * tile_x_count + x].push_back(((uint32_t)i << 1) | acceptance);
        }
    }
  }

  // phase 5: per tile, primitives in API order (rasterizer.cpp:947-994)
  for (size_t ty = 0; ty < tile_y_count; ++ty)
    for (size_t tx = 0; tx < tile_x_count; ++tx) {
      if ((tx + 3 * ty) % dev.shard_n != dev.shard_rank) continue;
      for (uint32_t e : bins[ty * tile_x_count + tx]) rasterize_triangle(c, tris[e >> 1], e & 1, (int)tx, (int)ty);
    }
  dev.stats.ps_invocations += c.ps_invocations;
  dev.stats.backend_input_pixels += c.backend_input_pixels;
  return SLV_OK;
}

const float EWA_WTS[256] = {
#include "ewa_weights.inc"
};

}  // namespace

struct slv_device_t : Device {};

extern "C" {

const char* slv_backend_name(void) { return "oracle"; }
uint32_t slv_abi_version(void) { return SLV_ABI_VERSION; }

slv_result slv_device_create(int32_t, slv_device* out) {
  if (!out) return SLV_INVALID_PARAMETER;
  auto d = new slv_device_t;
  d->res.resize(1);
  *out = d;
  return SLV_OK;
}
void slv_device_destroy(slv_device dev) { delete dev; }

slv_result slv_buffer_create(slv_device dev, size_t bytes, slv_handle* out) {
  Resource r;
  r.kind = Resource::BUFFER;
  r.buf.resize(bytes);
  dev->res.push_back(std::move(r));
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}
slv_result slv_buffer_upload(slv_device dev, slv_handle h, size_t off, const void* src, size_t bytes) {
  auto r = dev->get(h, Resource::BUFFER);
  if (!r || off + bytes > r->buf.size()) return SLV_INVALID_PARAMETER;
  memcpy(r->buf.data() + off, src, bytes);
  return SLV_OK;
}
slv_result slv_buffer_readback(slv_device dev, slv_handle h, size_t off, void* dst, size_t bytes) {
  auto r = dev->get(h, Resource::BUFFER);
  if (!r || off + bytes > r->buf.size()) return SLV_INVALID_PARAMETER;
  memcpy(dst, r->buf.data() + off, bytes);
  return SLV_OK;
}

static Surface make_surface(uint32_t w, uint32_t h, uint32_t samples, uint32_t fmt) {
  Surface s;
  s.w = w; s.h = h; s.samples = samples; s.fmt = fmt; s.bpp = bpp_of(fmt);
  s.data.assign((size_t)w * h * samples * s.bpp, 0);
  return s;
}

slv_result slv_texture_create(slv_device dev, uint32_t w, uint32_t h, uint32_t samples, uint32_t fmt, slv_handle* out) {
  if (!bpp_of(fmt) || !w || !h || !samples) return SLV_INVALID_PARAMETER;
  Resource r;
  r.kind = Resource::TEXTURE;
  r.tex.fmt = fmt;
  r.tex.samples = samples;
  r.tex.levels.push_back(make_surface(w, h, samples, fmt));
  dev->res.push_back(std::move(r));
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

// texture_2d::gen_mipmap + surface::make_mip_surface (texture2d.cpp:25-36, surface.cpp:53-92)
slv_result slv_texture_gen_mipmap(slv_device dev, slv_handle h, uint32_t filter) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r) return SLV_INVALID_PARAMETER;
  Texture& t = r->tex;
  t.levels.resize(1);
  uint32_t m = std::max(std::max(t.levels[0].w, t.levels[0].h), 1u);
  size_t limit = 0;
  while (m > 0) { m >>= 1; ++limit; }
  for (size_t lvl = 0; lvl + 1 < limit; ++lvl) {
    Surface const& src = t.levels.back();
    uint32_t mw = (src.w + 1) / 2, mh = (src.h + 1) / 2;
    Surface dst = make_surface(mw, mh, src.samples, src.fmt);
    for (uint32_t y = 0; y < mh; ++y)
      for (uint32_t x = 0; x < mw; ++x)
        for (uint32_t s = 0; s < src.samples; ++s) {
          if (filter == SLV_FILTER_POINT) {
            from_rgba32f(dst.fmt, dst.addr(x, y, s), to_rgba32f(src.fmt, src.addr(x * 2, y * 2, s)));
          } else {
            // reads 2x+1 / 2y+1 without a bounds check (Appendix B #9): the x overflow wraps into the
            // next row (linear addressing) and is mirrored; reads past the allocation return zeros here.
            auto rd = [&](uint32_t xx, uint32_t yy) {
              size_t off = (((size_t)yy * src.w + xx) * src.samples + s) * src.bpp;
              if (off + src.bpp > src.data.size()) return mk4(0, 0, 0, 0);
              return to_rgba32f(src.fmt, src.data.data() + off);
            };
            V4 c0 = rd(x * 2, y * 2), c1 = rd(x * 2 + 1, y * 2), c2 = rd(x * 2, y * 2 + 1), c3 = rd(x * 2 + 1, y * 2 + 1);
            V4 o;
            for (int j = 0; j < 4; ++j) o[j] = (((c0[j] + c1[j]) + c2[j]) + c3[j]) * 0.25f;
            from_rgba32f(dst.fmt, dst.addr(x, y, s), o);
          }
        }
    t.levels.push_back(std::move(dst));
  }
  return SLV_OK;
}

slv_result slv_texture_level_count(slv_device dev, slv_handle h, uint32_t* out) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r) return SLV_INVALID_PARAMETER;
  *out = (uint32_t)r->tex.levels.size();
  return SLV_OK;
}
slv_result slv_texture_level_size(slv_device dev, slv_handle h, uint32_t level, uint32_t* w, uint32_t* hh) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || level >= r->tex.levels.size()) return SLV_INVALID_PARAMETER;
  *w = r->tex.levels[level].w;
  *hh = r->tex.levels[level].h;
  return SLV_OK;
}
slv_result slv_texture_upload(slv_device dev, slv_handle h, uint32_t level, const void* src, size_t bytes) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || level >= r->tex.levels.size() || bytes != r->tex.levels[level].data.size()) return SLV_INVALID_PARAMETER;
  memcpy(r->tex.levels[level].data.data(), src, bytes);
  return SLV_OK;
}
slv_result slv_texture_readback(slv_device dev, slv_handle h, uint32_t level, void* dst, size_t bytes) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || level >= r->tex.levels.size() || bytes != r->tex.levels[level].data.size()) return SLV_INVALID_PARAMETER;
  memcpy(dst, r->tex.levels[level].data.data(), bytes);
  return SLV_OK;
}
slv_result slv_sampler_create(slv_device dev, const slv_sampler_desc* d, slv_handle tex, slv_handle* out) {
  if (!dev->get(tex, Resource::TEXTURE)) return SLV_INVALID_PARAMETER;
  Resource r;
  r.kind = Resource::SAMPLER;
  r.samp.d = *d;
  r.samp.tex = tex;
  dev->res.push_back(std::move(r));
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}
slv_result slv_resource_release(slv_device dev, slv_handle h) {
  if (h == 0 || h >= dev->res.size()) return SLV_INVALID_PARAMETER;
  if (dev->res[h].kind == Resource::MODULE && dev->res[h].mod.dl) dlclose(dev->res[h].mod.dl);
  dev->res[h] = Resource();
  return SLV_OK;
}

slv_result slv_draw(slv_device dev, const slv_draw_desc* d) { return do_draw(*dev, *d); }

// surface::fill (surface.cpp:170-271)
slv_result slv_clear_color(slv_device dev, slv_handle h, const float rgba[4]) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r) return SLV_INVALID_PARAMETER;
  Surface& s = r->tex.levels[0];
  uint8_t px[16];
  from_rgba32f(s.fmt, px, mk4(rgba[0], rgba[1], rgba[2], rgba[3]));
  for (size_t i = 0; i < s.data.size(); i += s.bpp) memcpy(s.data.data() + i, px, s.bpp);
  return SLV_OK;
}

// framebuffer::clear_depth_stencil (framebuffer.cpp:616-644); both flags -> surface::fill (render_core.cpp:101-111)
slv_result slv_clear_depth_stencil(slv_device dev, slv_handle h, uint32_t flags, float depth, uint32_t stencil) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || r->tex.fmt != SLV_PF_RG32F) return SLV_INVALID_PARAMETER;
  if (!(flags & (SLV_CLEAR_DEPTH | SLV_CLEAR_STENCIL))) return SLV_INVALID_PARAMETER;
  Surface& s = r->tex.levels[0];
  for (size_t i = 0; i < s.data.size(); i += 8) {
    if (flags & SLV_CLEAR_DEPTH) memcpy(s.data.data() + i, &depth, 4);
    if (flags & SLV_CLEAR_STENCIL) memcpy(s.data.data() + i + 4, &stencil, 4);
  }
  return SLV_OK;
}

// surface::resolve (surface.cpp:123-140)
slv_result slv_resolve(slv_device dev, slv_handle src, slv_handle dst) {
  auto rs = dev->get(src, Resource::TEXTURE);
  auto rd = dev->get(dst, Resource::TEXTURE);
  if (!rs || !rd) return SLV_INVALID_PARAMETER;
  Surface& s = rs->tex.levels[0];
  Surface& t = rd->tex.levels[0];
  if (t.samples != 1 || t.w < s.w || t.h < s.h) return SLV_INVALID_PARAMETER;
  for (uint32_t y = 0; y < s.h; ++y)
    for (uint32_t x = 0; x < s.w; ++x) {
      V4 clr = mk4(0, 0, 0, 0);
      for (uint32_t k = 0; k < s.samples; ++k) {
        V4 tmp = to_rgba32f(s.fmt, s.addr(x, y, k));
        for (int j = 0; j < 4; ++j) clr[j] += tmp[j];
      }
      float inv = 1 / static_cast<float>(s.samples);
      for (int j = 0; j < 4; ++j) clr[j] *= inv;
      from_rgba32f(t.fmt, t.addr(x, y, 0), clr);
    }
  return SLV_OK;
}

slv_result slv_flush(slv_device) { return SLV_OK; }
slv_result slv_query_begin(slv_device dev) {
  dev->stats = slv_pipeline_statistics{};
  dev->traffic = slv_traffic_counters{};
  return SLV_OK;
}
slv_result slv_traffic_get(slv_device dev, slv_traffic_counters* out) { *out = dev->traffic; return SLV_OK; }
slv_result slv_kernel_launch_count(slv_device, uint64_t* out) { *out = 0; return SLV_OK; }
slv_result slv_event_record(slv_device dev, uint32_t slot) {
  if (slot >= 16) return SLV_INVALID_PARAMETER;
  dev->events[slot] = std::chrono::steady_clock::now();
  return SLV_OK;
}
slv_result slv_event_elapsed_ms(slv_device dev, uint32_t a, uint32_t b, float* ms) {
  if (a >= 16 || b >= 16) return SLV_INVALID_PARAMETER;
  *ms = std::chrono::duration<float, std::milli>(dev->events[b] - dev->events[a]).count();
  return SLV_OK;
}
slv_result slv_profile_enable(slv_device, uint32_t) { return SLV_OK; }
slv_result slv_set_stream(slv_device, void*) { return SLV_OK; }
slv_result slv_texture_readback_async(slv_device dev, slv_handle tex, uint32_t level, void* dst, size_t bytes) {
  return slv_texture_readback(dev, tex, level, dst, bytes);
}
slv_result slv_readback_wait(slv_device) { return SLV_OK; }
slv_result slv_buffer_device_ptr(slv_device dev, slv_handle h, void** out, size_t* bytes) {
  auto r = dev->get(h, Resource::BUFFER);
  if (!r || !out) return SLV_INVALID_PARAMETER;
  *out = r->buf.data();
  if (bytes) *bytes = r->buf.size();
  return SLV_OK;
}
slv_result slv_external_write_begin(slv_device, void*, uint32_t) { return SLV_OK; }
slv_result slv_external_write_end(slv_device, void*) { return SLV_OK; }
slv_result slv_assembly_wait(slv_device, slv_handle, const void*, uint32_t, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_peer_signal_after_consumers(slv_device, slv_handle, void*, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_host_register(slv_device, void*, size_t) { return SLV_OK; }
slv_result slv_host_unregister(slv_device, void*) { return SLV_OK; }
// the owned tiles of a single-sampled surface into a host frame of the same linear layout (synchronous here)
slv_result slv_texture_export_tiles_async(slv_device dev, slv_handle tex, void* host_frame, size_t bytes) {
  auto r = dev->get(tex, Resource::TEXTURE);
  if (!r || !host_frame) return SLV_INVALID_PARAMETER;
  Surface const& s = r->tex.levels[0];
  if (s.samples != 1 || bytes != s.data.size()) return SLV_INVALID_PARAMETER;
  uint32_t const tiles_x = (s.w + 63) / 64, tiles_y = (s.h + 63) / 64;
  for (uint32_t ty = 0; ty < tiles_y; ++ty)
    for (uint32_t tx = 0; tx < tiles_x; ++tx) {
      if (dev->shard_n > 1 && (tx + 3 * ty) % dev->shard_n != dev->shard_rank) continue;
      uint32_t const w = std::min(64u, s.w - tx * 64), h = std::min(64u, s.h - ty * 64);
      for (uint32_t y = 0; y < h; ++y) {
        size_t const off = ((size_t)(ty * 64 + y) * s.w + tx * 64) * s.bpp;
        memcpy((uint8_t*)host_frame + off, s.data.data() + off, (size_t)w * s.bpp);
      }
    }
  return SLV_OK;
}
slv_result slv_readback_fence(slv_device, slv_handle) { return SLV_OK; }
// peer-memory frame assembly is a property of the CUDA product (NVLink); the CPU checkers do not implement it
slv_result slv_peer_export_texture(slv_device, slv_handle, uint32_t, uint8_t*) { return SLV_FAILED; }
slv_result slv_peer_export_flags(slv_device, uint8_t*) { return SLV_FAILED; }
slv_result slv_texture_level_tracking(slv_device, uint32_t) { return SLV_OK; }
slv_result slv_texture_levels_touched(slv_device, slv_handle, uint32_t* mask) { if (!mask) return SLV_INVALID_PARAMETER; *mask = 0; return SLV_OK; }
slv_result slv_shader_module_load(slv_device, uint32_t, const void*, size_t, uint32_t, slv_handle*) { return SLV_FAILED; }
slv_result slv_shader_compile_cubin(uint32_t, const char*, uint32_t, uint32_t, void**, size_t*, char*, size_t) { return SLV_FAILED; }
// slv_shader_compile on the checker: the generated code is compiled for the HOST (g++ over sasl_rt.h's host build and
// oracle/slv_host_shader.h: quads as fibers, fetches through this library's sampler) into a shared object that run_vs /
// draw_quad call.  Test infrastructure - lets SASL scenes be compared with the cpp twins on the CPU.  The directories of the two
// headers are found next to this library ($SLV_ORACLE_SASL_RT overrides the one of sasl_rt.h).
slv_result slv_shader_compile(slv_device dev, uint32_t stage, const char* device_code, uint32_t n_vs_output_attrs, uint32_t flags, slv_handle* out,
                              char* log, size_t log_bytes) {
  if (log && log_bytes) log[0] = 0;
  if (!dev || !device_code || !out || (stage != SLV_STAGE_VS && stage != SLV_STAGE_PS)) return SLV_INVALID_PARAMETER;
  Dl_info info{};
  if (!dladdr((void*)&slv_shader_compile, &info) || !info.dli_fname) return SLV_FAILED;
  std::string here(info.dli_fname);
  here = here.find('/') == std::string::npos ? std::string(".") : here.substr(0, here.rfind('/'));
  const char* rt_env = getenv("SLV_ORACLE_SASL_RT");
  const std::string rt = rt_env ? rt_env : here + "/../salviarenderer_b200/sasl";
  char dir[] = "/tmp/slv_oracle_jit_XXXXXX";
  if (!mkdtemp(dir)) return SLV_FAILED;
  const std::string src = std::string(dir) + "/shader.cpp", so = std::string(dir) + "/shader.so", err = std::string(dir) + "/log.txt";
  FILE* f = fopen(src.c_str(), "w");
  if (!f) return SLV_FAILED;
  fprintf(f, "#include \"sasl_rt.h\"\n#include \"slv_host_shader.h\"\n%s\n%s\n", device_code, stage == SLV_STAGE_VS ? "SLV_HOST_VS_ENTRY" : "SLV_HOST_PS_ENTRY");
  fclose(f);
  const std::string cmd = std::string("g++ -std=c++17 -O1 -ffp-contract=off -fPIC -shared -w") + ((flags & SLV_COMPILE_DERIV_CPP) ? " -DSLV_JIT_DERIV_CPP=1" : "") +
                          " -I'" + rt + "' -I'" + here + "' -o '" + so + "' '" + src + "' > '" + err + "' 2>&1";
  const int rc = system(cmd.c_str());
  Module m;
  m.stage = stage;
  m.n_vs_output_attrs = n_vs_output_attrs;
  if (rc == 0) m.dl = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (m.dl) {
    m.vs = reinterpret_cast<decltype(m.vs)>(dlsym(m.dl, "slv_host_vs"));
    m.ps_quad = reinterpret_cast<decltype(m.ps_quad)>(dlsym(m.dl, "slv_host_ps_quad"));
  }
  const bool ok = m.dl && (stage == SLV_STAGE_VS ? m.vs != nullptr : m.ps_quad != nullptr);
  if (!ok && log && log_bytes) {
    FILE* e = fopen(err.c_str(), "r");
    size_t n = e ? fread(log, 1, log_bytes - 1, e) : 0;
    log[n] = 0;
    if (e) fclose(e);
    if (!n) snprintf(log, log_bytes, "%s", m.dl ? "entry point missing" : (rc ? "g++ failed" : dlerror()));
  }
  unlink(src.c_str()); unlink(so.c_str()); unlink(err.c_str()); rmdir(dir);  // the mapped object stays valid after the unlink
  if (!ok) { if (m.dl) dlclose(m.dl); return SLV_FAILED; }
  dev->res.emplace_back();
  dev->res.back().kind = Resource::MODULE;
  dev->res.back().mod = m;
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}
void slv_free(void* p) { free(p); }
// the first half of compile(code, profile): the product's own SASL front end (header-only, host code) - the restatement runs
// what it generates (slv_shader_compile above), it does not restate the translation
slv_result slv_sasl_translate(uint32_t stage, const char* source, const char* entry, char** unit, size_t* unit_bytes, char* log, size_t log_bytes) {
  if (log && log_bytes) log[0] = 0;
  if (!source || !unit || (stage != SLV_STAGE_VS && stage != SLV_STAGE_PS)) return SLV_INVALID_PARAMETER;
  *unit = nullptr;
  if (unit_bytes) *unit_bytes = 0;
  salvia_b200::sasl::unit u;
  std::string error;
  if (!salvia_b200::sasl::compile(source, stage == SLV_STAGE_VS ? "vs" : "ps", entry ? entry : "", salvia_b200::sasl::options(), u, error)) {
    if (log && log_bytes) snprintf(log, log_bytes, "%s", error.c_str());
    return SLV_FAILED;
  }
  const std::string text = salvia_b200::sasl::render(u);
  char* o = static_cast<char*>(malloc(text.size() + 1));
  if (!o) return SLV_OUT_OF_MEMORY;
  memcpy(o, text.data(), text.size());
  o[text.size()] = 0;
  *unit = o;
  if (unit_bytes) *unit_bytes = text.size();
  return SLV_OK;
}
slv_result slv_peer_open(slv_device, const uint8_t*, void**) { return SLV_FAILED; }
slv_result slv_peer_close(slv_device, void*) { return SLV_FAILED; }
slv_result slv_resolve_target_peer(slv_device, slv_handle, void*) { return SLV_FAILED; }
slv_result slv_peer_signal(slv_device, void*, uint32_t, uint32_t) { return SLV_FAILED; }
slv_result slv_flags_wait(slv_device, const void*, uint32_t, uint32_t, uint32_t) { return SLV_FAILED; }

slv_result slv_texture_device_ptr(slv_device dev, slv_handle h, uint32_t level, void** out, size_t* bytes) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || level >= r->tex.levels.size()) return SLV_INVALID_PARAMETER;
  *out = r->tex.levels[level].data.data();
  if (bytes) *bytes = r->tex.levels[level].data.size();
  return SLV_OK;
}
static slv_result pack_common(slv_device dev, slv_handle h, uint32_t rank, uint32_t n, uint8_t* staging, size_t* bytes, bool unpack) {
  auto r = dev->get(h, Resource::TEXTURE);
  if (!r || n == 0 || rank >= n || r->tex.samples != 1) return SLV_INVALID_PARAMETER;
  Surface& s = r->tex.levels[0];
  const uint32_t T = SLV_TILE_SIZE;
  uint32_t tiles_x = (s.w + T - 1) / T, tiles_y = (s.h + T - 1) / T, slot = 0;
  for (uint32_t ty = 0; ty < tiles_y; ++ty)
    for (uint32_t tx = 0; tx < tiles_x; ++tx) {
      if (n > 1 && (tx + 3 * ty) % n != rank) continue;
      if (staging)
        for (uint32_t row = 0; row < T; ++row) {
          uint32_t y = ty * T + row;
          if (y >= s.h) break;
          uint32_t x0 = tx * T, cnt = std::min(T, s.w - x0);
          uint8_t* g = s.addr(x0, y, 0);
          uint8_t* st = staging + ((size_t)slot * T * T + (size_t)row * T) * s.bpp;
          if (unpack) memcpy(g, st, (size_t)cnt * s.bpp); else memcpy(st, g, (size_t)cnt * s.bpp);
        }
      ++slot;
    }
  if (bytes) *bytes = (size_t)slot * T * T * s.bpp;
  return SLV_OK;
}
slv_result slv_pack_tiles(slv_device dev, slv_handle h, uint32_t rank, uint32_t n, void* staging, size_t* bytes) {
  return pack_common(dev, h, rank, n, (uint8_t*)staging, bytes, false);
}
slv_result slv_unpack_tiles(slv_device dev, slv_handle h, uint32_t rank, uint32_t n, const void* staging) {
  if (!staging) return SLV_INVALID_PARAMETER;
  return pack_common(dev, h, rank, n, (uint8_t*)staging, nullptr, true);
}
slv_result slv_query_get(slv_device dev, slv_pipeline_statistics* out) { *out = dev->stats; return SLV_OK; }
slv_result slv_profile_get(slv_device, slv_pipeline_profiles* out) { memset(out, 0, sizeof(*out)); return SLV_OK; }
slv_result slv_profile_get_stages(slv_device, double* ms, uint32_t n) {
  if (!ms || n < 6) return SLV_INVALID_PARAMETER;
  for (uint32_t i = 0; i < n; ++i) ms[i] = 0.0;
  return SLV_OK;
}
slv_result slv_set_tile_shard(slv_device dev, uint32_t rank, uint32_t nranks) {
  if (nranks == 0 || rank >= nranks) return SLV_INVALID_PARAMETER;
  dev->shard_rank = rank;
  dev->shard_n = nranks;
  return SLV_OK;
}

slv_result slv_sampler_probe(slv_device dev, slv_handle sh, uint32_t n, const float* coords, const float* ddx,
                             const float* ddy, const float* lod, uint32_t use_lod, float* out) {
  auto r = dev->get(sh, Resource::SAMPLER);
  if (!r) return SLV_INVALID_PARAMETER;
  Texture const* t = sampler_texture(*dev, r->samp);
  if (!t) return SLV_INVALID_PARAMETER;
  for (uint32_t i = 0; i < n; ++i) {
    V4 c;
    if (use_lod) c = sample_impl(*t, r->samp.d, coords[2 * i], coords[2 * i + 1], lod[i], nullptr);
    else
      c = sample_2d_grad(*t, r->samp.d, coords[2 * i], coords[2 * i + 1], ddx[2 * i], ddx[2 * i + 1], ddy[2 * i],
                         ddy[2 * i + 1], 0.0f);
    memcpy(out + 4 * i, c.v, 16);
  }
  return SLV_OK;
}

}  // extern "C"
