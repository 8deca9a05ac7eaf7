// oracle/slv_host_shader.h - TEST INFRASTRUCTURE (never part of the product).
//
// Runtime of SASL shaders compiled for the HOST: what libsalvia_oracle.so's slv_shader_compile wraps around the code the SASL
// front end generates (salviarenderer_b200/sasl/frontend.py or host/sasl_frontend.hpp - the same text the product hands to
// NVRTC), so that the CPU checker can run SASL vertex and pixel shaders inside its pipeline and SASL scenes can be compared on
// the CPU with the cpp twins the samples ship (which are pinned to the reference).  The generated code is scalar C++ over
// sasl_rt.h's host build; this header supplies what that build stubs out:
//   * screen-space derivatives: the four pixels of a quad run as four fibers (ucontext) that meet at every ddx / ddy - the
//     pixel stores its value, yields until the other three have stored theirs, then differences two of them.  SASL convention
//     (ddx per line, ddy per column: sasl/src/codegen/cgs_simd.cpp:275-313) or, with SLV_JIT_DERIV_CPP, the cpp_pixel_shader one
//     (q1 - q0 / q2 - q0 for the whole quad, cpp_pixel_shader.cpp:13-19) - what the quad shuffles of the device build do.
//     The front end rejects derivatives under divergent control flow, so all four fibers reach the same sequence of exchanges;
//     values are double-buffered because a resumed pixel may reach the NEXT exchange before its neighbours have read this one.
//   * texture fetches: call-backs into the checker's own sampler (sample_2d_grad / sample_2d_lod).
// Order in the translation unit: sasl_rt.h, this header, the generated code, then SLV_HOST_VS_ENTRY or SLV_HOST_PS_ENTRY.
#pragma once

#include <ucontext.h>

#include <cstddef>
#include <vector>

// what the checker passes to a host shader: its sampler, bound to the draw being executed
struct SlvHostApi {
  void* ctx;
  void (*sample_grad)(void* ctx, int slot, float u, float v, float dudx, float dvdx, float dudy, float dvdy, float bias, float* rgba);
  void (*sample_lod)(void* ctx, int slot, float u, float v, float lod, float* rgba);
  void (*vs_sample_lod)(void* ctx, float u, float v, float lod, float* rgba);
};
static const SlvHostApi* slv_host_api = nullptr;  // the call in flight (a checker device draws on one thread)

struct SlvQuadState {
  ucontext_t main_ctx, fib[4];
  float xchg[2][4];
  unsigned gen[4];
  bool done[4];
};
struct SlvQuadCtx {
  float4 a[5];
  unsigned quad_base = 0;
  unsigned pix = 0;
  float4 attr(int i) const { return a[i]; }
};
static SlvQuadState slv_quad;
static SlvQuadCtx slv_quad_ctx[4];

static inline float slv_quad_exchange(const SlvQuadCtx& px, float v, unsigned hi, unsigned lo) {
  const unsigned g = slv_quad.gen[px.pix]++ & 1u;
  slv_quad.xchg[g][px.pix] = v;
  swapcontext(&slv_quad.fib[px.pix], &slv_quad.main_ctx);
  return slv_quad.xchg[g][hi] - slv_quad.xchg[g][lo];
}
#ifdef SLV_JIT_DERIV_CPP
static inline float slv_host_ddx(const SlvQuadCtx& px, float v) { return slv_quad_exchange(px, v, 1, 0); }
static inline float slv_host_ddy(const SlvQuadCtx& px, float v) { return slv_quad_exchange(px, v, 2, 0); }
#else
static inline float slv_host_ddx(const SlvQuadCtx& px, float v) { return slv_quad_exchange(px, v, px.pix | 1u, px.pix & ~1u); }
static inline float slv_host_ddy(const SlvQuadCtx& px, float v) { return slv_quad_exchange(px, v, px.pix | 2u, px.pix & ~2u); }
#endif
static inline void slv_host_tex2d_grad(const slv::RasterParams&, const SlvQuadCtx&, int slot, float u, float v, float dudx, float dvdx, float dudy,
                                       float dvdy, float bias, float& r, float& g, float& b, float& a) {
  float c[4] = {0, 0, 0, 0};
  slv_host_api->sample_grad(slv_host_api->ctx, slot, u, v, dudx, dvdx, dudy, dvdy, bias, c);
  r = c[0]; g = c[1]; b = c[2]; a = c[3];
}
static inline void slv_host_tex2d_lod(const slv::RasterParams&, const SlvQuadCtx&, int slot, float u, float v, float lod, float& r, float& g, float& b,
                                      float& a) {
  float c[4] = {0, 0, 0, 0};
  slv_host_api->sample_lod(slv_host_api->ctx, slot, u, v, lod, c);
  r = c[0]; g = c[1]; b = c[2]; a = c[3];
}
static inline void slv_host_vs_tex2d_lod(const SaslSampler&, int, float u, float v, float lod, float& r, float& g, float& b, float& a) {
  float c[4] = {0, 0, 0, 0};
  slv_host_api->vs_sample_lod(slv_host_api->ctx, u, v, lod, c);
  r = c[0]; g = c[1]; b = c[2]; a = c[3];
}
// the generated code calls sasl_rt.h's names: route them here (sasl_rt.h's host stubs stay defined and unused)
#define sasl_ddx slv_host_ddx
#define sasl_ddy slv_host_ddy
#define sasl_tex2d_grad slv_host_tex2d_grad
#define sasl_tex2d_lod slv_host_tex2d_lod
#define sasl_vs_tex2d_lod slv_host_vs_tex2d_lod

// in: 8 input registers, out: position + 5 attributes (vs_output, shader_regs.h:30-83)
#define SLV_HOST_VS_ENTRY                                                                                                        \
  extern "C" void slv_host_vs(const float* in, const unsigned char* uniforms, float* out, const SlvHostApi* api) {               \
    slv_host_api = api;                                                                                                          \
    float4 i[8], o[6];                                                                                                           \
    for (int k = 0; k < 8; ++k) i[k] = make_float4(in[4 * k], in[4 * k + 1], in[4 * k + 2], in[4 * k + 3]);                      \
    for (int k = 0; k < 6; ++k) o[k] = make_float4(0, 0, 0, 0);                                                                  \
    slv_jit_vs(i, uniforms, o, SaslSampler{});                                                                                   \
    for (int k = 0; k < 6; ++k) { out[4 * k] = o[k].x; out[4 * k + 1] = o[k].y; out[4 * k + 2] = o[k].z; out[4 * k + 3] = o[k].w; } \
  }

// attrs: 4 pixels (pixel = row * 2 + col) x 5 attributes x 4 floats; colors: 4 x 4 floats; keep: 4 flags (0 = discarded)
#define SLV_HOST_PS_ENTRY                                                                                                        \
  static slv::RasterParams slv_host_params;                                                                                      \
  static float* slv_host_colors;                                                                                                 \
  static int* slv_host_keep;                                                                                                     \
  static void slv_host_fiber(int pix) {                                                                                          \
    float4 c = make_float4(0, 0, 0, 0);                                                                                          \
    const bool keep = slv_jit_ps(slv_host_params, slv_quad_ctx[pix], c);                                                         \
    slv_host_colors[4 * pix] = c.x; slv_host_colors[4 * pix + 1] = c.y; slv_host_colors[4 * pix + 2] = c.z; slv_host_colors[4 * pix + 3] = c.w; \
    slv_host_keep[pix] = keep ? 1 : 0;                                                                                           \
    slv_quad.done[pix] = true;                                                                                                   \
    swapcontext(&slv_quad.fib[pix], &slv_quad.main_ctx);                                                                         \
  }                                                                                                                              \
  extern "C" void slv_host_ps_quad(const float* attrs, const unsigned char* uniforms, size_t n_uniform_bytes, const SlvHostApi* api, float* colors, \
                                   int* keep) {                                                                                  \
    static std::vector<char> stacks[4];                                                                                          \
    slv_host_api = api;                                                                                                          \
    slv_host_colors = colors;                                                                                                    \
    slv_host_keep = keep;                                                                                                        \
    memset(&slv_host_params, 0, sizeof(slv_host_params));                                                                        \
    memcpy(slv_host_params.ps_uniforms, uniforms, n_uniform_bytes < sizeof(slv_host_params.ps_uniforms) ? n_uniform_bytes : sizeof(slv_host_params.ps_uniforms)); \
    for (int p = 0; p < 4; ++p) {                                                                                                \
      for (int k = 0; k < 5; ++k) {                                                                                              \
        const float* s = attrs + (p * 5 + k) * 4;                                                                                \
        slv_quad_ctx[p].a[k] = make_float4(s[0], s[1], s[2], s[3]);                                                              \
      }                                                                                                                          \
      slv_quad_ctx[p].pix = (unsigned)p;                                                                                         \
      slv_quad.gen[p] = 0;                                                                                                       \
      slv_quad.done[p] = false;                                                                                                  \
      if (stacks[p].empty()) stacks[p].resize(256 * 1024);                                                                       \
      getcontext(&slv_quad.fib[p]);                                                                                              \
      slv_quad.fib[p].uc_stack.ss_sp = stacks[p].data();                                                                         \
      slv_quad.fib[p].uc_stack.ss_size = stacks[p].size();                                                                       \
      slv_quad.fib[p].uc_link = &slv_quad.main_ctx;                                                                              \
      makecontext(&slv_quad.fib[p], (void (*)())slv_host_fiber, 1, p);                                                           \
    }                                                                                                                            \
    for (bool any = true; any;) {                                                                                                \
      any = false;                                                                                                               \
      for (int p = 0; p < 4; ++p)                                                                                                \
        if (!slv_quad.done[p]) { swapcontext(&slv_quad.main_ctx, &slv_quad.fib[p]); any = true; }                                \
    }                                                                                                                            \
  }
