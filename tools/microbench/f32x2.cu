// Developer tool: issue throughput of packed fp32 pairs (FADD2 / FMUL2, sm_100+) against scalar FADD / FMUL and of the
// byte -> float conversions (I2F.U8 on the XU pipe against PRMT + FADD magic-number conversion), per SM per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/microbench/f32x2 tools/microbench/f32x2.cu && tools/microbench/f32x2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
               : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
               : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
constexpr int ITERS = 2048, CH = 8;

template <int MODE>
__global__ void k(float* out, float seed) {
  float2 v[CH];
  uint32_t u[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) { v[c] = make_float2(seed + c + threadIdx.x, seed * 0.5f + c); u[c] = (uint32_t)(seed * 77.0f) + c * 0x01020304u + threadIdx.x; }
  const float2 k2 = make_float2(seed, seed);
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (MODE == 0) { v[c].x = __fadd_rn(v[c].x, seed); }                                   // 1 FADD
      else if (MODE == 1) { v[c] = add2(v[c], k2); }                                         // 1 FADD2
      else if (MODE == 2) { v[c].x = __fmul_rn(v[c].x, seed); }                              // 1 FMUL
      else if (MODE == 3) { v[c] = mul2(v[c], k2); }                                         // 1 FMUL2
      else if (MODE == 4) { v[c].x = __fadd_rn(v[c].x, seed); v[c].y = __fadd_rn(v[c].y, seed); }  // 2 FADD
      else if (MODE == 5) { v[c].x += (float)((u[c] >> 8) & 0xFF); u[c] += 0x00010300u; }     // I2F.U8 (+ FADD, IADD)
      else if (MODE == 6) { v[c].x += __uint_as_float(__byte_perm(u[c], 0x4B000000u, 0x7441)) - 8388608.0f; u[c] += 0x00010300u; }  // PRMT + 2 FADD + IADD
    }
  }
  float acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += v[c].x + v[c].y + (float)u[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, double inst_per_iter_chain) {
  float* d;
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int threads = 256, blocks = sms * 8;
  cudaMalloc(&d, (size_t)threads * blocks * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(d, 1.0001f);
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k<MODE><<<blocks, threads>>>(d, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warp_inst = 5.0 * blocks * (threads / 32) * (double)ITERS * CH * inst_per_iter_chain;
  const double cycles = ms * 1e-3 * khz * 1e3;
  printf("%-44s %8.3f ms   %.3f listed warp-instructions / clk / SM (clock %d MHz as reported)\n", name, ms / 5, warp_inst / cycles / sms, khz / 1000);
  cudaFree(d);
}

int main() {
  run<0>("FADD  (1 result / lane)", 1);
  run<1>("FADD2 (2 results / lane)", 1);
  run<2>("FMUL", 1);
  run<3>("FMUL2", 1);
  run<4>("2 x FADD (same work as one FADD2)", 2);
  run<5>("I2F.U8 byte->float (+FADD +IADD)", 3);
  run<6>("PRMT+FADD byte->float (+FADD +IADD)", 4);
  return 0;
}
