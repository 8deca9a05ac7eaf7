// slv_jit_unit.cu — the translation unit salviarenderer_b200/sasl/jit.py compiles at run time, one per SASL shader.
// NOT part of libsalvia_b200.so.  nvcc flags (the library's, so every float operation keeps the numerics contract):
//   -cubin -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
//   -DSLV_JIT_VS=1 -DSLV_JIT_R=<registers>        a vertex shader   -> slv_jit_k_geometry, slv_jit_k_vertex_shade
//   -DSLV_JIT_PS=1 [-DSLV_JIT_DERIV_CPP=1]        a pixel shader    -> slv_jit_k_raster_s1 / _s2 / _s4 (immediate path)
//                                                                  and slv_jit_k_shade_s1 / _s2 / _s4 (visibility-first path)
//   -DSLV_JIT_GENERATED="<path of the generated .cuh>"
// The pipeline kernels are the library's own code (slv_kernels.cuh, slv_deferred.cuh); the shader is inlined into them.
#include <cuda_runtime.h>

#include "salvia_b200.h"
#include "slv_kernels.cuh"
#ifdef SLV_JIT_PS
#include "slv_deferred.cuh"
#endif
#include "sasl_rt.h"
#include SLV_JIT_GENERATED

#ifdef SLV_JIT_VS
static_assert(SLV_JIT_R == SLV_JIT_VS_OUTPUT_ATTRS + 1, "register count does not match the shader's outputs");
extern "C" __global__ void __launch_bounds__(128, 4) slv_jit_k_geometry(const slv::GeomParams* __restrict__ draws, slv::GeomBatch hb) {
  slv::geometry_main<SLV_JIT_R>(draws, hb);
}
// two-kernel geometry: the position pass alone (k_geometry_cull of the library, with this shader's position inlined)
extern "C" __global__ void __launch_bounds__(128, 4) slv_jit_k_geometry_cull(const slv::GeomParams* __restrict__ draws, slv::GeomBatch hb) {
  slv::geometry_cull_main<SLV_JIT_R>(draws, hb);
}
// post-transform vertex cache: the shader once per referenced vertex (k_vertex_shade of the library, with this shader inlined)
extern "C" __global__ void __launch_bounds__(128) slv_jit_k_vertex_shade(const slv::GeomParams* __restrict__ draws, slv::GeomBatch hb) {
  slv::vertex_shade_main<SLV_JIT_R>(draws, hb);
}
#endif

#ifdef SLV_JIT_PS
#define SLV_JIT_RASTER(S)                                                                                              \
  extern "C" __global__ void __launch_bounds__(slv::RASTER_THREADS, slv::RASTER_CTAS_PER_SM)                           \
      slv_jit_k_raster_s##S(slv::RasterParams c, const slv::RasterParams* __restrict__ batch, uint32_t n_draws) {      \
    slv::raster_main<S, SLV_PS_JIT>(c, batch, n_draws);                                                               \
  }
SLV_JIT_RASTER(1)
SLV_JIT_RASTER(2)
SLV_JIT_RASTER(4)
// quad-granular k_shade: the four lanes of a (quad, owner) pair run the shader together, so ddx / ddy / tex2D work
#define SLV_JIT_SHADE(S)                                                                                               \
  extern "C" __global__ void __launch_bounds__(slv::DEF_THREADS, SLV_SHADE_CTAS_PER_SM)                                \
      slv_jit_k_shade_s##S(slv::RasterParams c, const slv::RasterParams* __restrict__ batch, uint32_t n_draws,         \
                           slv::DeferredBufs d) {                                                                      \
    slv::shade_quad_main<S, SLV_PS_JIT>(c, batch, n_draws, d);                                                         \
  }
SLV_JIT_SHADE(1)
SLV_JIT_SHADE(2)
SLV_JIT_SHADE(4)
#endif
