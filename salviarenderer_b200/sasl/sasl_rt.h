// sasl_rt.h — runtime support of the code salviarenderer_b200/sasl/frontend.py generates.
//
// Two builds: (1) CUDA, inside csrc/slv_jit_unit.cu, after slv_kernels.cuh: derivatives are quad shuffles, texture
// sampling is the product's sampler (slv_sampler.cuh); (2) host C++ (no CUDA), used by the CPU test-suite to execute
// generated shaders without a GPU: the arithmetic is the same scalar code, derivatives / sampling are not available.
#pragma once

#if defined(__CUDACC__)
#define SASL_FN __device__ __forceinline__
#define SASL_FN_REC __device__ __noinline__  /* functions on a call cycle (recursion) */
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define SASL_FN static inline
#define SASL_FN_REC static
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
namespace slv { struct RasterParams { unsigned char ps_uniforms[256]; }; }
struct SaslSampler {};  // host build: no texture sampling
#endif

SASL_FN float sasl_clamp(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }  // eflib::clamp
SASL_FN bool sasl_eq_eps(float a, float b) { return fabsf(a - b) <= 1.1920928955078125e-7f; }      // eflib equal<float>
SASL_FN float sasl_smoothstep(float lo, float hi, float x) {
  const float t = sasl_clamp((x - lo) / (hi - lo), 0.0f, 1.0f);
  return (t * t) * (3.0f - (2.0f * t));
}
SASL_FN float sasl_asfloat(int v) { float f; memcpy(&f, &v, 4); return f; }
SASL_FN float sasl_asfloat(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
SASL_FN float sasl_asfloat(float v) { return v; }
SASL_FN int sasl_asint(float v) { int i; memcpy(&i, &v, 4); return i; }
SASL_FN int sasl_asint(unsigned v) { return (int)v; }
SASL_FN int sasl_asint(int v) { return v; }
SASL_FN unsigned sasl_asuint(float v) { unsigned i; memcpy(&i, &v, 4); return i; }
SASL_FN unsigned sasl_asuint(int v) { return (unsigned)v; }
SASL_FN unsigned sasl_asuint(unsigned v) { return v; }
SASL_FN unsigned sasl_countbits(unsigned v) {
#if defined(__CUDACC__)
  return (unsigned)__popc(v);
#else
  return (unsigned)__builtin_popcount(v);
#endif
}

// sasl.firstbithigh.u32 = 31 - bsr(v), sasl.firstbitlow.u32 = bsf(v), sasl.reversebits.u32 (compiler_impl.cpp:414-432);
// bsr / bsf of 0 are undefined upstream: all ones here
SASL_FN unsigned sasl_firstbithigh(unsigned v) {
#if defined(__CUDACC__)
  return v ? (unsigned)__clz((int)v) : 0xFFFFFFFFu;
#else
  return v ? (unsigned)__builtin_clz(v) : 0xFFFFFFFFu;
#endif
}
SASL_FN unsigned sasl_firstbitlow(unsigned v) {
#if defined(__CUDACC__)
  return v ? (unsigned)(__ffs((int)v) - 1) : 0xFFFFFFFFu;
#else
  return v ? (unsigned)__builtin_ctz(v) : 0xFFFFFFFFu;
#endif
}
SASL_FN unsigned sasl_reversebits(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}

// math intrinsics (see frontend.py UNARY_MATH): what the reference binds its JIT-ed code to, sasl/src/drivers/compiler_impl.cpp:340-404
#if defined(__CUDACC__)
#define SASL_M1(name) SASL_FN float sasl_m_##name(float x) { return (float)name((double)x); }
#define SASL_M2(name) SASL_FN float sasl_m_##name(float x, float y) { return (float)name((double)x, (double)y); }
#define SASL_FADD(a, b) __fadd_rn(a, b)
#define SASL_FSUB(a, b) __fsub_rn(a, b)
#else
#define SASL_M1(name) SASL_FN float sasl_m_##name(float x) { return std::name(x); }
#define SASL_M2(name) SASL_FN float sasl_m_##name(float x, float y) { return std::name(x, y); }
SASL_FN float sasl_host_add(float a, float b) { volatile float r = a + b; return r; }
SASL_FN float sasl_host_sub(float a, float b) { volatile float r = a - b; return r; }
#define SASL_FADD(a, b) sasl_host_add(a, b)
#define SASL_FSUB(a, b) sasl_host_sub(a, b)
#endif
SASL_M1(exp) SASL_M1(log10) SASL_M1(sin) SASL_M1(cos) SASL_M1(tan) SASL_M1(asin)
SASL_M1(acos) SASL_M1(atan) SASL_M1(sinh) SASL_M1(cosh) SASL_M1(tanh) SASL_M2(pow) SASL_M2(atan2)
#undef SASL_M1
#undef SASL_M2
// eflib::fast_log2 / fast_log (eflib/include/eflib/math/math.h:89-109): exponent + a quadratic in the mantissa
SASL_FN float sasl_m_log2(float val) {
  int x;
  memcpy(&x, &val, 4);
  const int log_2 = ((x >> 23) & 255) - 128;
  x &= ~(255 << 23);
  x += 127 << 23;
  float f;
  memcpy(&f, &x, 4);
  f = ((-1.0f / 3) * f + 2) * f - 2.0f / 3;
  return f + (float)log_2;
}
SASL_FN float sasl_m_log(float val) { return sasl_m_log2(val) * 0.69314718f; }
// eflib::fast_round / fast_ceil / fast_floor / trunc (math.h:56-60, 111-134): round to nearest even through the 2^23 bias
SASL_FN float sasl_m_round(float val) {
  int n;
  memcpy(&n, &val, 4);
  const int bi = ((23 + 127) << 23) + (int)((unsigned)n & 0x80000000u);
  float bias;
  memcpy(&bias, &bi, 4);
  return SASL_FSUB(SASL_FADD(val, bias), bias);
}
SASL_FN float sasl_m_ceil(float val) { const float f = sasl_m_round(val); return (f < val) ? f + 1 : f; }
SASL_FN float sasl_m_floor(float val) { const float f = sasl_m_round(val); return (f > val) ? f - 1 : f; }
SASL_FN float sasl_m_trunc(float val) { return val > 0.0f ? sasl_m_floor(val) : sasl_m_ceil(val); }
// sasl.exp2.f32 = ldexpf(1, (int)v), sasl.ldexp.f32 = ldexpf(x, (int)e): the exponent is truncated to an integer
SASL_FN float sasl_m_exp2(float v) { return ldexpf(1.0f, (int)v); }
SASL_FN float sasl_m_ldexp(float x, float e) { return ldexpf(x, (int)e); }

#if defined(__CUDACC__)
typedef slv::SamplerRef SaslSampler;
// sasl.vs.tex2d.lod -> sampler::sample_2d_lod(coord.xy, coord.w) (salvia/src/resource/sampler_api.cpp:50-52)
SASL_FN void sasl_vs_tex2d_lod(const SaslSampler& s0, int slot, float u, float v, float lod, float& r, float& g, float& b, float& a) {
  (void)slot;  // one sampler per vertex shader (slot 0)
  const float4 c = slv::vs_sample_lod(s0, u, v, lod);
  r = c.x; g = c.y; b = c.z; a = c.w;
}
// Screen-space derivatives: the four pixels of a quad sit in four consecutive lanes (pixel i of the quad in lane
// quad_base + i, k_raster's shading phases; helper pixels run too).  SLV_JIT_DERIV_CPP selects the cpp_pixel_shader
// convention (ddx = q1 - q0, ddy = q2 - q0 for all four pixels, cpp_pixel_shader.cpp:13-19), the default is the SASL one
// (per row / per column, sasl/src/codegen/cgs_simd.cpp:275-313).  All lanes of the warp must call these together.
template <class Ctx>
SASL_FN float sasl_ddx(const Ctx& px, float v) {
  const unsigned pi = (threadIdx.x & 31u) - px.quad_base;
#ifdef SLV_JIT_DERIV_CPP
  const unsigned hi = 1, lo = 0;
  (void)pi;
#else
  const unsigned hi = pi | 1u, lo = pi & ~1u;
#endif
  const float a = __shfl_sync(0xFFFFFFFFu, v, px.quad_base + hi), b = __shfl_sync(0xFFFFFFFFu, v, px.quad_base + lo);
  return a - b;
}
template <class Ctx>
SASL_FN float sasl_ddy(const Ctx& px, float v) {
  const unsigned pi = (threadIdx.x & 31u) - px.quad_base;
#ifdef SLV_JIT_DERIV_CPP
  const unsigned hi = 2, lo = 0;
  (void)pi;
#else
  const unsigned hi = pi | 2u, lo = pi & ~2u;
#endif
  const float a = __shfl_sync(0xFFFFFFFFu, v, px.quad_base + hi), b = __shfl_sync(0xFFFFFFFFu, v, px.quad_base + lo);
  return a - b;
}
// sasl.ps.tex2d.grad -> sampler::sample_2d_grad (sampler_api.h:9-30, sampler.cpp:854-873)
template <class Ctx>
SASL_FN void sasl_tex2d_grad(const slv::RasterParams& p, const Ctx&, int slot, float u, float v, float dudx, float dvdx, float dudy,
                             float dvdy, float bias, float& r, float& g, float& b, float& a) {
  const float4 c = slv::sample_2d_grad(slot ? p.sampler1 : p.sampler0, u, v, dudx, dvdx, dudy, dvdy, bias);  // slots 0 and 1
  r = c.x; g = c.y; b = c.z; a = c.w;
}
template <class Ctx>
SASL_FN void sasl_tex2d_lod(const slv::RasterParams& p, const Ctx&, int slot, float u, float v, float lod, float& r, float& g, float& b,
                            float& a) {
  const float4 c = slv::sample_impl(slot ? p.sampler1 : p.sampler0, u, v, lod, nullptr);
  r = c.x; g = c.y; b = c.z; a = c.w;
}
#else
SASL_FN void sasl_vs_tex2d_lod(const SaslSampler&, int, float, float, float, float& r, float& g, float& b, float& a) { r = g = b = a = 0.0f; }
template <class Ctx> SASL_FN float sasl_ddx(const Ctx&, float) { return 0.0f; }
template <class Ctx> SASL_FN float sasl_ddy(const Ctx&, float) { return 0.0f; }
template <class Ctx>
SASL_FN void sasl_tex2d_grad(const slv::RasterParams&, const Ctx&, int, float, float, float, float, float, float, float, float& r, float& g,
                             float& b, float& a) { r = g = b = a = 0.0f; }
template <class Ctx>
SASL_FN void sasl_tex2d_lod(const slv::RasterParams&, const Ctx&, int, float, float, float, float& r, float& g, float& b, float& a) {
  r = g = b = a = 0.0f;
}
#endif
