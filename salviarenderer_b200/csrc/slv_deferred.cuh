// slv_deferred.cuh — the visibility-first form of phase 5 (raster + shade + merge).
//
// When every queued draw of a batch has early-Z on (no stencil), the REPLACE blend shader, a pixel shader that never
// discards and no centroid attributes, the colour of a sample at the end of the batch is the pixel-shader colour of
// the LAST fragment that passed the in-order early depth test at that sample (framebuffer.cpp:522-614 writes depth at
// test time; render_sample_quad, framebuffer.cpp:481-520, then lets the blend shader overwrite the sample).  So the
// pass is split:
//
//   k_cover<S>     in-order coverage + early-Z (+ the pipeline counters) exactly as k_raster does it, but instead of
//                  shading it records, per sample, the triangle slot of the last passing fragment ("visibility").
//                  Depth and owner live in REGISTERS (thread == pixel); nothing is shaded, no quad queue.
//   k_shade<S,PS>  runs the pixel shader once per (pixel, distinct owner): dense, order-free, no warp coupling.  The quad
//                  derivatives a cpp_pixel_shader sees (ddx = q1 - q0, ddy = q2 - q0 over the 2x2 quad evaluated with
//                  the SAME triangle, cpp_pixel_shader.cpp:13-19) are recomputed by the lane itself from the triangle's
//                  plane equations with the reference's stepping order (shader.cpp:289-367), so the lanes are
//                  independent.  Pixels whose samples belong to several triangles push their extra owners to a
//                  shared-memory queue that the CTA drains densely.
//
// The results (colour, depth, counters) are bit-identical to k_raster; tests/test_gpu_parity.py runs every case
// through whichever path the batch qualifies for, and test_deferred_equals_immediate forces both.
#pragma once

#include "slv_kernels.cuh"

namespace slv {

constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;

struct CovTri {  // one surviving triangle of the current chunk, staged in shared memory (80 B, 128-bit loads)
  float4 e0;     // A0 B0 C0 A1
  float4 e1;     // B1 C1 A2 B2
  float4 e2;     // C2 v0.x v0.y v0.z
  float4 e3;     // ddx.z ddy.z as_float(slot) as_float(bits)
  float4 aa;     // per-sample depth offsets (rasterizer.cpp:678-687)
};
// CovTri bits: 0 read_depth, 1 write_depth, 4..7 compare LUT (compare_lut)

#ifndef SLV_COVER_CTAS_PER_SM
#define SLV_COVER_CTAS_PER_SM 4
#endif
#ifndef SLV_SHADE_CTAS_PER_SM
#define SLV_SHADE_CTAS_PER_SM 3
#endif

template <int S>
__global__ void __launch_bounds__(RASTER_THREADS, SLV_COVER_CTAS_PER_SM)
    k_cover(RasterParams c, const RasterParams* __restrict__ batch, uint32_t* __restrict__ vis, uint32_t vis_pitch) {
  __shared__ CovTri s_tri[RASTER_THREADS];
  __shared__ uint16_t s_wlist[RASTER_WARPS][RASTER_THREADS];
  __shared__ uint16_t s_cnt[RASTER_WARPS + 1][RASTER_WARPS];
  __shared__ uint32_t s_item;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int q = lane >> 2, pi = lane & 3;
  const int lx = wx + (q & 3) * 2 + (pi & 1), ly = wy + (q >> 2) * 2 + (pi >> 1);
  const int bx = lx >> 2, by = ly >> 2;
  const int ix = lx & 3, iy = ly & 3;
  const uint32_t fullmask = (1u << S) - 1;

  uint32_t n_ps_quads = 0;
  uint32_t n_ztest = 0, n_zwrite = 0, n_cwrite = 0;
  uint32_t n_scanned = 0, n_surv = 0, n_pairs = 0;

  const uint32_t n_items = c.active_tiles[0] * 16u;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(c.work_counter, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t tile = c.active_tiles[1 + (item >> 4)], sub = item & 15;
    const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
    const int X16 = (sub & 3) * REGION, Y16 = (sub >> 2) * REGION;
    const int gx0 = tile_x * TILE + X16, gy0 = tile_y * TILE + Y16;
    if ((float)gx0 >= (float)c.target_w || (float)gy0 >= (float)c.target_h) continue;  // rasterizer.cpp:721-724
    const int x = gx0 + lx, y = gy0 + ly;
    const bool odd_x = x & 1, odd_y = y & 1;
    const bool in_target = (uint32_t)x < c.target_w && (uint32_t)y < c.target_h;
    const float vpx = (float)(tile_x * TILE), vpy = (float)(tile_y * TILE);
    const float hx = 0.5f + (float)(uint32_t)(x & ~1), hy = 0.5f + (float)(uint32_t)(y & ~1);
    const float left_f = (float)(gx0 + bx * 4), top_f = (float)(gy0 + by * 4);

    float z[S];
    uint32_t st[S], own[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { z[s] = 0.0f; st[s] = 0u; own[s] = VIS_NONE; }
    bool fb_loaded = false, dirty = false;

    const uint32_t list_beg = c.tile_offset[tile];
    uint32_t list_end = c.tile_offset[tile + 1];
    if (list_end > c.list_capacity) list_end = c.list_capacity;
    for (uint32_t chunk = list_beg; chunk < list_end; chunk += RASTER_THREADS) {
      // ================= filter: level-16 decision for the region + level-4 decision of its 16 blocks =================
      const uint32_t ei = chunk + tid;
      bool keep = false;
      uint32_t st_bits = 0;  // 2 bits per block: 0 rejected, 1 partial, 2 full
      CovTri ent;
      if (ei < list_end) {
        ++n_scanned;
        const uint32_t e = __ldg(c.list + ei);
        const uint32_t slot = e >> 1;
        const float4* rec = c.tris + (size_t)slot * c.tri_stride;
        const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2), bb = __ldg(rec + 3);
        TriEntry te;
        te.A[0] = e0.x; te.B[0] = e0.y; te.C[0] = e0.z;
        te.A[1] = e1.x; te.B[1] = e1.y; te.C[1] = e1.z;
        te.A[2] = e2.x; te.B[2] = e2.y; te.C[2] = e2.z;
        uint32_t full16;
        const float x_min = bb.x - vpx, x_max = bb.y - vpx, y_min = bb.z - vpy, y_max = bb.w - vpy;
        if (e & 1) {  // the whole 64x64 tile is inside the triangle (rasterizer.cpp:736-743)
          keep = true;
          full16 = 1;
        } else {  // subdivide_tile at the 16-px level (rasterizer.cpp:441-602, 698-772)
          bool rej = (x_min >= (float)(X16 + REGION)) || (x_max < (float)X16) || (y_min >= (float)(Y16 + REGION)) ||
                     (y_max < (float)Y16);
          bool acc = true;
          const float ftx = (float)(X16 / REGION), fty = (float)(Y16 / REGION);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            float A = te.A[k], B = te.B[k], C = te.C[k];
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
            st_bits = 0xAAAAAAAAu;
          } else {
            const float rl = (float)gx0, rt = (float)gy0;
            for (int b = 0; b < 16; ++b) {
              const int bbx = b & 3, bby = b >> 2;
              st_bits |= (uint32_t)block_test(te, x_min, x_max, y_min, y_max, X16 + bbx * 4, Y16 + bby * 4, rl, rt, bbx, bby)
                         << (2 * b);
            }
            keep = st_bits != 0;
          }
        }
        if (keep) {
          const float4 misc = __ldg(rec + 4);
          const RasterParams& p = batch[__float_as_uint(misc.w)];
          const int R = 1 + (int)p.n_attrs;
          const float4 v0p = __ldg(rec + TRI_HEADER), gxp = __ldg(rec + TRI_HEADER + R), gyp = __ldg(rec + TRI_HEADER + 2 * R);
          const uint32_t bits = (p.read_depth ? 1u : 0u) | (p.write_depth ? 2u : 0u) |
                                ((p.depth_enable ? compare_lut(p.depth_func) : 0xFu) << 4);
          ent.e0 = make_float4(e0.x, e0.y, e0.z, e1.x);
          ent.e1 = make_float4(e1.y, e1.z, e2.x, e2.y);
          ent.e2 = make_float4(e2.z, v0p.x, v0p.y, v0p.z);
          ent.e3 = make_float4(gxp.z, gyp.z, __uint_as_float(slot), __uint_as_float(bits));
          float aa[4];
#pragma unroll
          for (int s = 0; s < 4; ++s)
            aa[s] = (s < S && S > 1) ? (SamplePattern<S>::x(s) - 0.5f) * gxp.z + (SamplePattern<S>::y(s) - 0.5f) * gyp.z : 0.0f;
          ent.aa = make_float4(aa[0], aa[1], aa[2], aa[3]);
        }
      }
      // ---- order-preserving compaction: survivors -> s_tri, and per target warp -> s_wlist[w] ----
      uint32_t hit = 0;
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
      uint32_t my_cnt = 0;
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
      if (my_cnt && !fb_loaded) {
        fb_loaded = true;
        if (in_target && c.ds.data) {
          const float2* dp = reinterpret_cast<const float2*>(c.ds.data + ((size_t)y * c.ds.w + x) * S * 8);
          if (S == 4) {
            const float4 a = *reinterpret_cast<const float4*>(dp), b = *reinterpret_cast<const float4*>(dp + 2);
            z[0] = a.x; st[0] = __float_as_uint(a.y); z[1 % S] = a.z; st[1 % S] = __float_as_uint(a.w);
            z[2 % S] = b.x; st[2 % S] = __float_as_uint(b.y); z[3 % S] = b.z; st[3 % S] = __float_as_uint(b.w);
          } else {
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float2 v = dp[s];
              z[s] = v.x;
              st[s] = __float_as_uint(v.y);
            }
          }
        }
      }
      if (lane == 0) n_pairs += my_cnt;
      for (uint32_t wi = 0; wi < my_cnt; ++wi) {
        const uint32_t we = s_wlist[warp][wi];
        const CovTri& t = s_tri[we & 0xFF];
        const int blk = (we >> (8 + 2 * (bx & 1))) & 3;  // 0 rejected, 1 partial, 2 full
        // per-sample coverage (draw_partial_tile, rasterizer.cpp:298-439)
        uint32_t pm = 0;
        if (in_target) {
          if (blk == 2) {
            pm = fullmask;
          } else if (blk == 1) {
            const float4 e0 = t.e0, e1 = t.e1;
            const float C2 = t.e2.x;
            const float A[3] = {e0.x, e0.w, e1.z}, B[3] = {e0.y, e1.x, e1.w}, Cc[3] = {e0.z, e1.y, C2};
            float ev[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) ev[k] = Cc[k] - (left_f * A[k] + top_f * B[k]);
#pragma unroll
            for (int s = 0; s < S; ++s) {
              const float fx = SamplePattern<S>::x(s) + (float)ix, fy = SamplePattern<S>::y(s) + (float)iy;
              bool rj = false;
#pragma unroll
              for (int k = 0; k < 3; ++k) rj |= (fx * A[k] + fy * B[k]) < ev[k];
              if (!rj) pm |= 1u << s;
            }
          }
        }
        if (!__any_sync(0xFFFFFFFFu, pm != 0)) continue;
        // early-Z: test and WRITE depth now (framebuffer.cpp:522-614; Appendix B #3); record the owner
        uint32_t tested = 0;
        if (pm) {
          const float4 e2 = t.e2, e3 = t.e3, aa4 = t.aa;
          const uint32_t bits = __float_as_uint(e3.w), slot = __float_as_uint(e3.z);
          const float dx = hx - e2.y, dy = hy - e2.z;
          float depth = e2.w + (e3.x * dx + e3.y * dy);
          if (odd_x) depth += e3.x;
          if (odd_y) depth += e3.y;
          const uint32_t lut = bits >> 4;
          const bool rd = bits & 1u, wr = bits & 2u;
          const float aa[4] = {aa4.x, aa4.y, aa4.z, aa4.w};
#pragma unroll
          for (int s = 0; s < S; ++s) {
            if (pm & (1u << s)) {
              const float nd = (S == 1) ? depth : aa[s] + depth;
              const float od = rd ? z[s] : 0.0f;
              if (compare_with_lut(lut, nd, od)) {
                tested |= 1u << s;
                own[s] = slot;
                if (wr) z[s] = nd;
              }
            }
          }
          const uint32_t np = __popc(pm), nt = __popc(tested);
          n_ztest += rd ? np : 0u;
          n_zwrite += wr ? nt : 0u;
          n_cwrite += vis ? nt : 0u;
          dirty |= wr && nt;
        }
        // ps_invocations: one quad per 2x2 with a live sample after early-Z (rasterizer.cpp:1274-1321)
        const uint32_t tb = __ballot_sync(0xFFFFFFFFu, tested != 0);
        if (lane == 0) n_ps_quads += __popc((tb | (tb >> 1) | (tb >> 2) | (tb >> 3)) & 0x11111111u);
      }
      // (the barrier at the top of the next chunk / item protects s_tri, s_wlist and s_cnt)
      __syncthreads();
    }

    // ---- write back: owners always (k_shade reads every pixel of a processed region), depth when modified ----
    if (in_target) {
      if (vis) {
        uint32_t* vp = vis + ((size_t)y * vis_pitch + x) * S;
        if (S == 4) *reinterpret_cast<uint4*>(vp) = make_uint4(own[0], own[1 % S], own[2 % S], own[3 % S]);
        else if (S == 2) *reinterpret_cast<uint2*>(vp) = make_uint2(own[0], own[1 % S]);
        else *vp = own[0];
      }
    }
    if (fb_loaded) {
      const bool any_ds = __any_sync(0xFFFFFFFFu, dirty);
      if (in_target && any_ds && c.ds.data) {
        uint8_t* ds_ptr = c.ds.data + ((size_t)y * c.ds.w + x) * S * 8;
        if (S == 4) {
          *reinterpret_cast<float4*>(ds_ptr) = make_float4(z[0], __uint_as_float(st[0]), z[1 % S], __uint_as_float(st[1 % S]));
          *reinterpret_cast<float4*>(ds_ptr + 16) = make_float4(z[2 % S], __uint_as_float(st[2 % S]), z[3 % S], __uint_as_float(st[3 % S]));
        } else if (S == 2) {
          *reinterpret_cast<float4*>(ds_ptr) = make_float4(z[0], __uint_as_float(st[0]), z[1 % S], __uint_as_float(st[1 % S]));
        } else {
          *reinterpret_cast<float2*>(ds_ptr) = make_float2(z[0], __uint_as_float(st[0]));
        }
      }
    }
  }

  uint32_t a = n_ps_quads;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
    n_ztest += __shfl_xor_sync(0xFFFFFFFFu, n_ztest, o);
    n_zwrite += __shfl_xor_sync(0xFFFFFFFFu, n_zwrite, o);
    n_cwrite += __shfl_xor_sync(0xFFFFFFFFu, n_cwrite, o);
    n_scanned += __shfl_xor_sync(0xFFFFFFFFu, n_scanned, o);
    n_surv += __shfl_xor_sync(0xFFFFFFFFu, n_surv, o);
    n_pairs += __shfl_xor_sync(0xFFFFFFFFu, n_pairs, o);
  }
  if (lane == 0) {
    if (a) {
      atomicAdd(&c.stats[7], (unsigned long long)a * 4ull);
      atomicAdd(&c.stats[8], (unsigned long long)a * 4ull);  // no discard: every shaded quad reaches the backend
    }
    if (n_ztest) atomicAdd(&c.stats[9], (unsigned long long)n_ztest);
    if (n_zwrite) atomicAdd(&c.stats[10], (unsigned long long)n_zwrite);
    if (n_cwrite) atomicAdd(&c.stats[11], (unsigned long long)n_cwrite);
    if (n_scanned) atomicAdd(&c.stats[13], (unsigned long long)n_scanned);
    if (n_surv) atomicAdd(&c.stats[14], (unsigned long long)n_surv);
    if (n_pairs) atomicAdd(&c.stats[15], (unsigned long long)n_pairs);
  }
}

// ---- shading ------------------------------------------------------------------------------------------------------
struct DeferredCtx {  // what a pixel shader may read when a lane shades its pixel alone
  const float4* rec;
  int R;
  const uint32_t* mods;
  float dx, dy;            // quad origin relative to v0 (shader.cpp:277-281)
  float iw00, iw01, iw10;  // 1 / pos.w of pixels 0, 1, 2 of the quad
  float inv_w;             // of this pixel
  bool odd_x, odd_y;
  __device__ __forceinline__ float4 attr(int i) const {
    return interp_attr(rec, R, 1 + i, mods[i], dx, dy, odd_x, odd_y, false, 0.0f, 0.0f, inv_w);
  }
  // .xy of attribute register i at pixels 0, 1, 2 of the quad, as the reference's quad stepping produces them
  // (step_2d_unproj_attr_quad, shader.cpp:289-367: a00 = a0 + (ddx*dx + ddy*dy), a01 = a00 + ddx, a10 = a00 + ddy,
  //  each times 1/pos.w of its pixel unless noperspective)
  __device__ __forceinline__ void quad_xy(int i, float4, float& u0, float& v0, float& u1, float& v1, float& u2, float& v2) const {
    const uint32_t mod = mods[i];
    const float4 a0 = __ldg(rec + TRI_HEADER + 1 + i);
    float x00 = a0.x, y00 = a0.y, x01 = a0.x, y01 = a0.y, x10 = a0.x, y10 = a0.y;
    if (!(mod & SLV_AM_NOINTERPOLATION)) {
      const float4 gx = __ldg(rec + TRI_HEADER + R + 1 + i), gy = __ldg(rec + TRI_HEADER + 2 * R + 1 + i);
      x00 = a0.x + (gx.x * dx + gy.x * dy);
      y00 = a0.y + (gx.y * dx + gy.y * dy);
      x01 = x00 + gx.x; y01 = y00 + gx.y;
      x10 = x00 + gy.x; y10 = y00 + gy.y;
    }
    if (!(mod & SLV_AM_NOPERSPECTIVE)) {
      x00 *= iw00; y00 *= iw00; x01 *= iw01; y01 *= iw01; x10 *= iw10; y10 *= iw10;
    }
    u0 = x00; v0 = y00; u1 = x01; v1 = y01; u2 = x10; v2 = y10;
  }
};

template <int PS>
__device__ __forceinline__ uint32_t shade_sample_owner(const RasterParams& c, const RasterParams* __restrict__ batch, uint32_t slot,
                                                       int x, int y) {
  const float4* rec = c.tris + (size_t)slot * c.tri_stride;
  const RasterParams& p = batch[__float_as_uint(__ldg(rec + 4).w)];
  const int R = 1 + (int)p.n_attrs;
  const float4 v0p = __ldg(rec + TRI_HEADER), gxp = __ldg(rec + TRI_HEADER + R), gyp = __ldg(rec + TRI_HEADER + 2 * R);
  DeferredCtx px;
  px.rec = rec; px.R = R; px.mods = p.mods;
  px.dx = 0.5f + (float)(uint32_t)(x & ~1) - v0p.x;
  px.dy = 0.5f + (float)(uint32_t)(y & ~1) - v0p.y;
  px.odd_x = x & 1; px.odd_y = y & 1;
  // step_2d_unproj_pos_quad (shader.cpp:257-287): pos00, pos01 = pos00 + ddx, pos10 = pos00 + ddy, pos11 = pos01 + ddy
  const float w00 = v0p.w + (gxp.w * px.dx + gyp.w * px.dy);
  const float w01 = w00 + gxp.w, w10 = w00 + gyp.w, w11 = w01 + gyp.w;
  px.iw00 = 1.0f / w00; px.iw01 = 1.0f / w01; px.iw10 = 1.0f / w10;
  px.inv_w = px.odd_y ? (px.odd_x ? 1.0f / w11 : px.iw10) : (px.odd_x ? px.iw01 : px.iw00);
  float4 color;
  run_ps<PS>(p, px, color);
  return pack_color(c.color0.fmt, color);
}

constexpr int SHADE_QCAP = RASTER_THREADS * 3;  // a pixel has at most S - 1 extra owners

template <int S, int PS>
__global__ void __launch_bounds__(RASTER_THREADS, SLV_SHADE_CTAS_PER_SM)
    k_shade(RasterParams c, const RasterParams* __restrict__ batch, const uint32_t* __restrict__ vis, uint32_t vis_pitch,
            uint32_t* __restrict__ work_counter) {
  __shared__ uint32_t s_color[RASTER_THREADS][S];
  __shared__ uint2 s_q[SHADE_QCAP];  // x = pixel (tid) | sample mask << 8, y = owner slot
  __shared__ uint32_t s_qn, s_item;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wx = (warp & 1) * 8, wy = (warp >> 1) * 4;
  const int q = lane >> 2, pi = lane & 3;
  const int lx = wx + (q & 3) * 2 + (pi & 1), ly = wy + (q >> 2) * 2 + (pi >> 1);  // same pixel <-> thread map as k_cover
  const uint32_t fullmask = (1u << S) - 1;

  const uint32_t n_items = c.active_tiles[0] * 16u;
  for (;;) {
    __syncthreads();
    if (tid == 0) { s_item = atomicAdd(work_counter, 1u); s_qn = 0; }
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t tile = c.active_tiles[1 + (item >> 4)], sub = item & 15;
    const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
    const int gx0 = tile_x * TILE + (sub & 3) * REGION, gy0 = tile_y * TILE + (sub >> 2) * REGION;
    if ((float)gx0 >= (float)c.target_w || (float)gy0 >= (float)c.target_h) continue;
    const int x = gx0 + lx, y = gy0 + ly;
    const bool in_target = (uint32_t)x < c.target_w && (uint32_t)y < c.target_h;

    uint32_t own[S];
#pragma unroll
    for (int s = 0; s < S; ++s) own[s] = VIS_NONE;
    if (in_target) {
      const uint32_t* vp = vis + ((size_t)y * vis_pitch + x) * S;
      if (S == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(vp);
        own[0] = v.x; own[1 % S] = v.y; own[2 % S] = v.z; own[3 % S] = v.w;
      } else if (S == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(vp);
        own[0] = v.x; own[1 % S] = v.y;
      } else {
        own[0] = *vp;
      }
    }
    uint32_t rem = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) rem |= (own[s] != VIS_NONE) ? (1u << s) : 0u;
    const uint32_t touched = rem;
    uint32_t* cptr = reinterpret_cast<uint32_t*>(c.color0.data + ((size_t)y * c.color0.w + x) * S * 4);
    if (touched && touched != fullmask) {  // some samples keep their colour: fetch it for the 128-bit store
#pragma unroll
      for (int s = 0; s < S; ++s) s_color[tid][s] = cptr[s];
    }
    uint32_t first_slot = VIS_NONE, first_mask = 0;
    if (rem) {
#pragma unroll
      for (int s = S - 1; s >= 0; --s)
        if (rem & (1u << s)) first_slot = own[s];
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (own[s] == first_slot) first_mask |= 1u << s;
      rem &= ~first_mask;
      while (rem) {  // further distinct owners of this pixel -> the CTA's queue
        uint32_t sl = VIS_NONE, m = 0;
#pragma unroll
        for (int s = S - 1; s >= 0; --s)
          if (rem & (1u << s)) sl = own[s];
#pragma unroll
        for (int s = 0; s < S; ++s)
          if ((rem & (1u << s)) && own[s] == sl) m |= 1u << s;
        rem &= ~m;
        s_q[atomicAdd(&s_qn, 1u)] = make_uint2(tid | (m << 8), sl);
      }
      const uint32_t packed = shade_sample_owner<PS>(c, batch, first_slot, x, y);
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (first_mask & (1u << s)) s_color[tid][s] = packed;
    }
    __syncthreads();
    const uint32_t qn = s_qn;
    for (uint32_t j = tid; j < qn; j += RASTER_THREADS) {
      const uint2 it = s_q[j];
      const uint32_t pt = it.x & 0xFF, m = it.x >> 8;
      const uint32_t pl = pt & 31, pw = pt >> 5, pq = pl >> 2, pp = pl & 3;
      const int px_ = gx0 + (int)((pw & 1) * 8 + (pq & 3) * 2 + (pp & 1)), py_ = gy0 + (int)((pw >> 1) * 4 + (pq >> 2) * 2 + (pp >> 1));
      const uint32_t packed = shade_sample_owner<PS>(c, batch, it.y, px_, py_);
#pragma unroll
      for (int s = 0; s < S; ++s)
        if (m & (1u << s)) s_color[pt][s] = packed;
    }
    __syncthreads();
    if (touched) {
      if (S == 4) *reinterpret_cast<uint4*>(cptr) = make_uint4(s_color[tid][0], s_color[tid][1 % S], s_color[tid][2 % S], s_color[tid][3 % S]);
      else if (S == 2) *reinterpret_cast<uint2*>(cptr) = make_uint2(s_color[tid][0], s_color[tid][1 % S]);
      else *cptr = s_color[tid][0];
    }
  }
}

}  // namespace slv
