// slv_deferred.cuh — the visibility-first form of phase 5 (raster + shade + merge).
//
// When every queued draw of a batch has early-Z on (no stencil), the REPLACE blend shader, a pixel shader that never
// discards and no centroid attributes, the colour of a sample at the end of the batch is the pixel-shader colour of
// the LAST fragment that passed the in-order early depth test at that sample (framebuffer.cpp:522-614 writes depth at
// test time; render_sample_quad, framebuffer.cpp:481-520, then lets the blend shader overwrite the sample).  So the
// pass is split into three kernels, none of which has a CTA-wide barrier:
//
//   k_region_bin   one CTA per non-empty 64x64 tile, one warp per 16x16 region: the reference's level-16 decision
//                  (subdivide_tile, rasterizer.cpp:441-602) for every entry of the tile's sorted list, compacted IN ORDER
//                  into one list per region.
//   k_cover<S>     work item = (region, 8x4-pixel warp block), one WARP per item, warps fully independent: level-4
//                  decisions of the warp's two 4x4 blocks, per-sample coverage, in-order early-Z (+ the pipeline
//                  counters) exactly as k_raster does them, but instead of shading it records, per sample, the triangle
//                  slot of the last passing fragment ("visibility").  Depth and owner live in REGISTERS (lane == pixel).
//   k_shade<S,PS>  same items: runs the pixel shader once per (pixel, distinct owner): dense, order-free.  The quad
//                  derivatives a cpp_pixel_shader sees (ddx = q1 - q0, ddy = q2 - q0 over the 2x2 quad evaluated with
//                  the SAME triangle, cpp_pixel_shader.cpp:13-19) are recomputed by the lane itself from the triangle's
//                  plane equations with the reference's stepping order (shader.cpp:289-367), so lanes are independent.
//                  Pixels whose samples belong to several triangles push their extra owners to the warp's queue,
//                  which the warp drains densely.
//
// The results (colour, depth, counters) are bit-identical to k_raster; tests/test_gpu_parity.py runs every case
// through whichever path the batch qualifies for, and test_deferred_equals_immediate forces both.
#pragma once

#include "slv_kernels.cuh"

namespace slv {

constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;
constexpr int DEF_WARPS = 4;                 // warps per CTA of k_cover / k_shade (independent of each other)
constexpr int DEF_THREADS = DEF_WARPS * 32;
constexpr uint32_t ITEMS_PER_TILE = 128;     // 16 regions x 8 warp blocks
constexpr uint32_t FETCH = 2;               // items a warp of k_cover takes from the queue at a time

struct CovTri {  // one surviving triangle of the warp's current chunk, staged in shared memory (80 B, 128-bit loads)
  float4 e0;     // A0 B0 C0 A1
  float4 e1;     // B1 C1 A2 B2
  float4 e2;     // C2 v0.x v0.y v0.z
  float4 e3;     // ddx.z ddy.z as_float(slot) as_float(bits)
  float4 aa;     // per-sample depth offsets (rasterizer.cpp:678-687)
};
// CovTri bits: 0 read_depth, 1 write_depth, 4..7 compare LUT (compare_lut), 8..11 status of the warp's two blocks
// (2 bits each: 0 rejected, 1 partial, 2 full)

struct DeferredBufs {
  // lazy clears: the whole-surface clear that preceded the batch was not executed; k_cover / k_shade start from the clear
  // value instead of loading, and write every pixel of every item of the active tiles (tiles without triangles are
  // filled by k_inactive_tiles), so depth / colour are written ONCE per frame and depth is never read
  uint32_t lazy_depth, lazy_color;
  float clear_z;
  uint32_t clear_st, clear_color;
  // work-queue budget: fetches a warp of k_cover / k_shade performs before it exits (0 = until the queue is empty, i.e. a
  // persistent grid).  With a budget the grid is sized to cover the queue and its CTAs are short-lived, so CTAs of the NEXT
  // frame's geometry / binning kernels (higher-priority stream) are scheduled in between instead of after the kernel.
  uint32_t cover_budget, shade_budget, shade_grp;
  // fused MSAA resolve: k_shade holds every pixel's final samples when it stores them, so it also writes the resolved
  // texel (surface::resolve, surface.cpp:123-140) - into this surface, which may be another rank's memory (sort-first)
  SurfaceRef resolve_dst;
  uint32_t* region_list;      // per region 8 sub-lists (one per 8x4 warp block) of capacity region_count, entries
                              // (slot << 4) | status of the warp's two 4x4 blocks, in API order
  uint2* block_desc;          // [item = (active tile index * 16 + region) * 8 + warp block]: (first entry, entries)
  uint32_t* region_mask;      // scratch, one word per tile-list entry: regions survived | regions fully inside << 16
  uint32_t* bits_pool;        // level-4 block bits (2 bits per 4x4 block) of every partially covered (entry, region) pair, allocated
  uint32_t bits_cap;          // per entry by k_region_decide from pool_cursor (reset by k_scan_tiles)
  uint32_t* pool_cursor;
  uint32_t* region_tile_cnt;  // [tile id * 16 + region]: survivors counted by k_region_decide, consumed AND re-zeroed by k_region_bin
  uint32_t region_cap;
  uint32_t* region_offset;    // [active tile index * 16 + region]
  uint32_t* region_count;
  uint32_t* cursor;           // region_list allocation cursor (reset by k_scan_tiles)
  uint32_t* overflow_flag;    // [0] flag, [1] tile-list entries needed, [2] region-arena words needed (maxima since the last check)
  uint8_t* item_flag;         // [item] 1 when k_cover recorded an owner in the warp block
  uint32_t* vis;              // owner slot per sample, layout ((y * vis_pitch + x) * S + s); nullptr = depth-only batch
  uint32_t vis_pitch;
  uint32_t* cover_counter;    // work-queue heads (reset by k_scan_tiles)
  uint32_t* shade_counter;
};

#ifndef SLV_COVER_CTAS_PER_SM
#define SLV_COVER_CTAS_PER_SM 8
#endif
#ifndef SLV_SHADE_CTAS_PER_SM
#define SLV_SHADE_CTAS_PER_SM 8
#endif

constexpr int RBIN_THREADS = 512;  // 16 warps == the 16 regions of a tile
constexpr int RMASK_STRIDE = 6;    // scratch words per tile-list entry: 2 (region masks, offset of the entry's block bits) + a pool of
                                   // 4 per entry on average for the block bits of its partially covered regions

// level-4 decision of the 16 blocks of region `reg` of tile (tile_x, tile_y): 2 bits per block, 0 rejected, 1 partial, 2 full
__device__ __forceinline__ uint32_t region_block_bits(const TriEntry& te, float x_min, float x_max, float y_min, float y_max,
                                                      uint32_t tile_x, uint32_t tile_y, int reg) {
  const int X16 = (reg & 3) * REGION, Y16 = (reg >> 2) * REGION;
  const float rl = (float)(tile_x * TILE + X16), rt = (float)(tile_y * TILE + Y16);
  uint32_t st_bits = 0;
  for (int blk = 0; blk < 16; ++blk) {
    const int bbx = blk & 3, bby = blk >> 2;
    st_bits |= (uint32_t)block_test(te, x_min, x_max, y_min, y_max, X16 + bbx * 4, Y16 + bby * 4, rl, rt, bbx, bby) << (2 * blk);
  }
  return st_bits;
}

#ifndef SLV_JIT_PS  // k_region_bin is the library's; a JIT unit only instantiates shade_quad_main
// Level-16 and level-4 decisions of EVERY tile-list entry of the batch, one thread per entry over a flat grid (the lists are
// one dense array): the arithmetic-heavy part of region binning no longer runs one CTA per tile, where the longest list was
// the kernel's critical path and the warps of a CTA idled at the barrier behind the one with the most partially covered
// regions.  Stores survive | accept << 16 and the block bits of up to RMASK_STRIDE - 1 partial regions per entry, and counts
// the survivors per (tile, region).
__global__ void __launch_bounds__(256, 4) k_region_decide(RasterParams c, DeferredBufs d) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n_tiles = c.tiles_x * c.tiles_y;
  uint32_t total = c.tile_offset[n_tiles];
  if (total > c.list_capacity) total = c.list_capacity;
  for (uint32_t base_i = blockIdx.x * blockDim.x; base_i < total; base_i += gridDim.x * blockDim.x) {
    const uint32_t i = base_i + threadIdx.x;
    const bool valid = i < total;
    // the entry's tile: last t with tile_offset[t] <= i (offsets are non-decreasing; empty tiles share an offset)
    uint32_t tile = 0xFFFFFFFFu;
    if (valid) {
      uint32_t lo = 0, hi = n_tiles;  // invariant: tile_offset[lo] <= i < tile_offset[hi]
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(c.tile_offset + mid) <= i) lo = mid; else hi = mid;
      }
      tile = lo;
    }
    const uint32_t tile_x = valid ? tile % c.tiles_x : 0u, tile_y = valid ? tile / c.tiles_x : 0u;
    const float vpx = (float)(tile_x * TILE), vpy = (float)(tile_y * TILE);
    // regions that start outside the target: "Sub tile is out of screen" (rasterizer.cpp:721-724)
    uint32_t on_screen = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (!((float)(tile_x * TILE + (k & 3) * REGION) >= (float)c.target_w || (float)(tile_y * TILE + (k >> 2) * REGION) >= (float)c.target_h))
        on_screen |= 1u << k;
    const uint32_t e = valid ? __ldg(c.list + i) : 1u;
    uint32_t survive = 0xFFFFu, accept = 0xFFFFu;  // e & 1: the whole 64x64 tile is inside (rasterizer.cpp:736-743)
    if (!(e & 1)) {
      const float4* rec = c.tris + (size_t)(e >> 1) * c.tri_stride;
      const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2), bb = __ldg(rec + 3);
      const float x_min = bb.x - vpx, x_max = bb.y - vpx, y_min = bb.z - vpy, y_max = bb.w - vpy;
      const float A3[3] = {e0.x, e1.x, e2.x}, B3[3] = {e0.y, e1.y, e2.y}, C3[3] = {e0.z, e1.z, e2.z};
      float sx[3], sy[3], r2a[3], ev1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float A = A3[k], B = B3[k], C = C3[k];
        float step_x = TILE * A, step_y = TILE * B;
        float ra = -fabsf(step_x) - fabsf(step_y);
        float part = (float)((A > 0) * TILE) * A + (float)((B > 0) * TILE) * B;
        step_x *= 0.25f; step_y *= 0.25f; ra *= 0.25f; part *= 0.25f;
        const float ev = C - part;
        sx[k] = step_x; sy[k] = step_y; r2a[k] = ra;
        ev1[k] = ev - (vpx * A + vpy * B);
      }
      survive = 0; accept = 0;
#pragma unroll
      for (int reg = 0; reg < 16; ++reg) {
        const int X16 = (reg & 3) * REGION, Y16 = (reg >> 2) * REGION;
        bool rej = (x_min >= (float)(X16 + REGION)) || (x_max < (float)X16) || (y_min >= (float)(Y16 + REGION)) || (y_max < (float)Y16);
        bool acc = true;
        const float ftx = (float)(reg & 3), fty = (float)(reg >> 2);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float step = sx[k] * ftx + sy[k] * fty;
          rej |= (step < ev1[k]);
          acc &= !((step + r2a[k]) < ev1[k]);
        }
        survive |= rej ? 0u : (1u << reg);
        accept |= (!rej && acc) ? (1u << reg) : 0u;
      }
    }
    survive &= valid ? on_screen : 0u;
    accept &= survive;
    if (valid) d.region_mask[(size_t)i * 2] = survive | (accept << 16);
    // per-(tile, region) survivor counts: one atomic per region and distinct tile among the warp's 32 entries (a warp's
    // entries are consecutive list positions, i.e. nearly always one tile)
    {
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, tile);
      const bool leader = lane == (uint32_t)__ffs(peers) - 1u;
#pragma unroll
      for (int reg = 0; reg < 16; ++reg) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (survive >> reg) & 1u) & peers;
        if (leader && valid && bal) atomicAdd(&d.region_tile_cnt[tile * 16u + reg], (uint32_t)__popc(bal));
      }
    }
    // level-4 decisions (subdivide_tile at the 4-px level) of EVERY partially covered region of the entry, 2 bits per 4x4
    // block, into a pool slice allocated with one atomic per warp: the ordered fill of k_region_bin then only copies bits
    uint32_t partial = survive & ~accept;
    const uint32_t np = (uint32_t)__popc(partial);
    uint32_t incl = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t warp_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t wbase = 0;
    if (lane == 31 && warp_total) wbase = atomicAdd(d.pool_cursor, warp_total);
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 31);
    const uint32_t off = wbase + incl - np;
    const bool fits = (uint64_t)off + np <= (uint64_t)d.bits_cap;  // else: k_region_bin evaluates the entry's blocks itself
    if (valid) d.region_mask[(size_t)i * 2 + 1] = (np && fits) ? off : 0xFFFFFFFFu;
    if (np && fits) {
      const float4* rec = c.tris + (size_t)(e >> 1) * c.tri_stride;
      const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2), bb = __ldg(rec + 3);
      TriEntry te;
      te.A[0] = e0.x; te.B[0] = e0.y; te.C[0] = e0.z;
      te.A[1] = e1.x; te.B[1] = e1.y; te.C[1] = e1.z;
      te.A[2] = e2.x; te.B[2] = e2.y; te.C[2] = e2.z;
      const float x_min = bb.x - vpx, x_max = bb.y - vpx, y_min = bb.z - vpy, y_max = bb.w - vpy;
      for (uint32_t k = 0; partial; ++k) {
        const int reg = __ffs(partial) - 1;
        partial &= partial - 1;
        d.bits_pool[off + k] = region_block_bits(te, x_min, x_max, y_min, y_max, tile_x, tile_y, reg);
      }
    }
  }
}

// One CTA per non-empty tile, one warp per 16x16 region: warp r appends, IN ORDER, the entries that survive in region r
// (decisions and counts from k_region_decide; the reference's level-16 / level-4 subdivide_tile, rasterizer.cpp:441-602,
// 698-772) to the sub-lists of the region's eight 8x4 warp blocks.  Allocation = one atomicAdd per tile.
__global__ void __launch_bounds__(RBIN_THREADS, 2) k_region_bin(RasterParams c, DeferredBufs d) {
  __shared__ uint32_t s_cnt[16], s_base[16];
  const uint32_t b = blockIdx.x;
  if (b >= c.active_tiles[0]) return;
  const uint32_t tile = c.active_tiles[1 + b];
  const uint32_t lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
  const float vpx = (float)(tile_x * TILE), vpy = (float)(tile_y * TILE);
  const uint32_t beg = c.tile_offset[tile];
  uint32_t end = c.tile_offset[tile + 1];
  if (end > c.list_capacity) end = c.list_capacity;
  if (threadIdx.x < 16) {  // survivors per region, counted by k_region_decide; re-zeroed here for the next batch of this scratch set
    s_cnt[threadIdx.x] = d.region_tile_cnt[tile * 16u + threadIdx.x];
    d.region_tile_cnt[tile * 16u + threadIdx.x] = 0;
  }
  __syncthreads();
  const uint32_t cnt = s_cnt[r];
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int k = 0; k < 16; ++k) total += s_cnt[k];
    // every region gets 8 sub-lists (one per 8x4 warp block) of capacity cnt each
    uint32_t base = total ? atomicAdd(d.cursor, total * 8u) : 0u;
    const bool fits = (uint64_t)base + (uint64_t)total * 8u <= (uint64_t)d.region_cap;
    if (!fits) *d.overflow_flag = 1;
    if (total) atomicMax(d.overflow_flag + 2, (uint32_t)min((uint64_t)base + (uint64_t)total * 8u, (uint64_t)0xFFFFFFF0u));  // words needed
    for (int k = 0; k < 16; ++k) {
      s_base[k] = base;
      d.region_offset[b * 16 + k] = base;
      d.region_count[b * 16 + k] = fits ? s_cnt[k] : 0u;
      base += s_cnt[k] * 8u;
    }
    if (!fits) s_base[0] = 0xFFFFFFFFu;
    if (total) {
      atomicAdd(&c.stats[13], (unsigned long long)(end - beg) * 16ull);  // (entry, region) decisions evaluated
      atomicAdd(&c.stats[14], (unsigned long long)total);                 // (entry, region) survivors
    }
  }
  __syncthreads();
  const bool ok = s_base[0] != 0xFFFFFFFFu;
  // Phase 3: warp r walks the tile list once more; the lane of a surviving entry evaluates the level-4 decision of
  // the region's 16 blocks (subdivide_tile at the 4-px level) and the entry is appended, IN ORDER, to the sub-list of
  // every 8x4 warp block it touches, as (slot << 4) | status of the block pair (2 bits each: 1 partial, 2 full).
  uint32_t pos[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) pos[w] = 0;
  if (ok && cnt) {
    const uint32_t rbase = s_base[r];
    const uint32_t below = (1u << lane) - 1;
    constexpr int U = 4;  // chunks of 32 entries per iteration: the loads of all U are in flight together (a long list is a
                          // chain of dependent L2 round trips otherwise - the critical path of the front half on heavy tiles)
    for (uint32_t i = beg; i < end; i += 32 * U) {
      uint32_t m[U], slot[U], st_bits[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t ei = i + u * 32 + lane;
        m[u] = ei < end ? d.region_mask[(size_t)ei * 2] : 0u;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t ei = i + u * 32 + lane;
        st_bits[u] = 0; slot[u] = 0;
        if ((m[u] >> r) & 1u) {
          slot[u] = __ldg(c.list + ei) >> 1;
          if ((m[u] >> (16 + r)) & 1u) {
            st_bits[u] = 0xAAAAAAAAu;  // region fully inside: every block full
          } else {
            const uint32_t partial = (m[u] & 0xFFFFu) & ~(m[u] >> 16);
            const uint32_t k = __popc(partial & ((1u << r) - 1));  // rank of this region among the entry's partial ones
            const uint32_t off = d.region_mask[(size_t)ei * 2 + 1];
            st_bits[u] = off != 0xFFFFFFFFu ? d.bits_pool[off + k] : 0xFFFFFFFFu;  // marker: the pool was full, evaluated below
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (i + u * 32 >= end) break;  // warp-uniform
        if (st_bits[u] == 0xFFFFFFFFu) {
          const float4* rec = c.tris + (size_t)slot[u] * c.tri_stride;
          const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2), bb = __ldg(rec + 3);
          TriEntry te;
          te.A[0] = e0.x; te.B[0] = e0.y; te.C[0] = e0.z;
          te.A[1] = e1.x; te.B[1] = e1.y; te.C[1] = e1.z;
          te.A[2] = e2.x; te.B[2] = e2.y; te.C[2] = e2.z;
          st_bits[u] = region_block_bits(te, bb.x - vpx, bb.y - vpx, bb.z - vpy, bb.w - vpy, tile_x, tile_y, (int)r);
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) {  // warp block w owns blocks (by = w >> 1, bx = (w & 1) * 2 + {0, 1})
          const uint32_t four = (st_bits[u] >> (2 * ((w >> 1) * 4 + (w & 1) * 2))) & 0xFu;
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, four != 0);
          if (four) d.region_list[rbase + (uint32_t)w * cnt + pos[w] + __popc(bal & below)] = (slot[u] << 4) | four;
          pos[w] += __popc(bal);
        }
      }
    }
  }
  if (lane < 8) {
    uint32_t mine = pos[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mine = (lane == (uint32_t)w) ? pos[w] : mine;
    d.block_desc[(b * 16 + r) * 8 + lane] = make_uint2((ok && cnt) ? s_base[r] + lane * cnt : 0u, mine);
  }
}

#endif  // SLV_JIT_PS

// work-queue fetch: lane 0 takes FETCH consecutive items; the result is consumed one fetch later (latency hidden)
__device__ __forceinline__ uint32_t fetch_items(uint32_t* counter, uint32_t lane) {
  uint32_t v = 0;
  if (lane == 0) v = atomicAdd(counter, FETCH);
  return v;
}

template <int S>
__global__ void __launch_bounds__(DEF_THREADS, SLV_COVER_CTAS_PER_SM)
    k_cover(RasterParams c, const RasterParams* __restrict__ batch, uint32_t n_draws, DeferredBufs d) {
  __shared__ CovTri s_tri_all[DEF_WARPS][32];

  const uint32_t lane = threadIdx.x & 31;
  CovTri* s_tri = s_tri_all[threadIdx.x >> 5];
  const int q = lane >> 2, pi = lane & 3;
  const int wlx = (q & 3) * 2 + (pi & 1), wly = (q >> 2) * 2 + (pi >> 1);  // pixel inside the 8x4 warp block
  const int ix = wlx & 3, iy = wly & 3;                                    // pixel inside its 4x4 block
  const int bsel = wlx >> 2;                                               // which of the warp's two blocks
  const uint32_t fullmask = (1u << S) - 1;

  uint32_t n_ps_quads = 0;
  uint32_t n_ztest = 0, n_zwrite = 0, n_cwrite = 0, n_pairs = 0;

  const uint32_t n_items = c.active_tiles[0] * ITEMS_PER_TILE;
  uint32_t next_raw = fetch_items(d.cover_counter, lane);
  uint32_t fetches = 1;
  for (;;) {
    const uint32_t base_item = __shfl_sync(0xFFFFFFFFu, next_raw, 0);
    if (base_item >= n_items) break;
    next_raw = (d.cover_budget == 0 || fetches < d.cover_budget) ? fetch_items(d.cover_counter, lane) : 0xFFFFFFFFu;
    ++fetches;
    for (uint32_t item = base_item; item < base_item + FETCH && item < n_items; ++item) {
      const uint32_t b = item >> 7, sub = (item >> 3) & 15, w = item & 7;
      const uint2 desc = d.block_desc[item];  // item == (b * 16 + sub) * 8 + w
      const uint32_t rbeg = desc.x, rcnt = desc.y;
      if (rcnt == 0 && !d.lazy_depth) {
        if (lane == 0) d.item_flag[item] = 0;
        continue;
      }
      const uint32_t tile = c.active_tiles[1 + b];
      const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
      const int X16 = (sub & 3) * REGION, Y16 = (sub >> 2) * REGION;
      const int gx0 = tile_x * TILE + X16, gy0 = tile_y * TILE + Y16;
      const int wx = (w & 1) * 8, wy = (w >> 1) * 4;
      const int x = gx0 + wx + wlx, y = gy0 + wy + wly;
      const bool odd_x = x & 1, odd_y = y & 1;
      const bool in_target = (uint32_t)x < c.target_w && (uint32_t)y < c.target_h;
      if (rcnt == 0) {  // lazy depth clear: nothing lands here, the block just receives the clear value
        if (lane == 0) d.item_flag[item] = 0;
        if (in_target) {
          float4* ds_ptr = reinterpret_cast<float4*>(c.ds.data + ((size_t)y * c.ds.w + x) * S * 8);
          const float4 cv = make_float4(d.clear_z, __uint_as_float(d.clear_st), d.clear_z, __uint_as_float(d.clear_st));
          if (S == 4) { ds_ptr[0] = cv; ds_ptr[1] = cv; }
          else if (S == 2) ds_ptr[0] = cv;
          else *reinterpret_cast<float2*>(ds_ptr) = make_float2(cv.x, cv.y);
        }
        continue;
      }
      const float hx = 0.5f + (float)(uint32_t)(x & ~1), hy = 0.5f + (float)(uint32_t)(y & ~1);
      const int bxA = wx >> 2, by = wy >> 2;  // block A of the warp inside the region; block B = bxA + 1
      const float left_f = (float)(gx0 + (bxA + bsel) * 4), top_f = (float)(gy0 + by * 4);

      float z[S];
      uint32_t st[S], own[S];
#pragma unroll
      for (int s = 0; s < S; ++s) { z[s] = d.clear_z; st[s] = d.clear_st; own[s] = VIS_NONE; }
      bool dirty = d.lazy_depth != 0;
      if (in_target && c.ds.data && !d.lazy_depth) {
        const float2* dp = reinterpret_cast<const float2*>(c.ds.data + ((size_t)y * c.ds.w + x) * S * 8);
        if (S == 4) {
          const float4 a = *reinterpret_cast<const float4*>(dp), b2 = *reinterpret_cast<const float4*>(dp + 2);
          z[0] = a.x; st[0] = __float_as_uint(a.y); z[1 % S] = a.z; st[1 % S] = __float_as_uint(a.w);
          z[2 % S] = b2.x; st[2 % S] = __float_as_uint(b2.y); z[3 % S] = b2.z; st[3 % S] = __float_as_uint(b2.w);
        } else {
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const float2 v = dp[s];
            z[s] = v.x;
            st[s] = __float_as_uint(v.y);
          }
        }
      }

      for (uint32_t chunk = 0; chunk < rcnt; chunk += 32) {
        // ---- stage the chunk's triangles in shared memory, one lane per list entry (every entry touches this warp) ----
        const uint32_t ei = chunk + lane;
        if (ei < rcnt) {
          const uint32_t e = __ldg(d.region_list + rbeg + ei);
          const uint32_t slot = e >> 4, st4 = e & 0xFu;
          const float4* rec = c.tris + (size_t)slot * c.tri_stride;
          const float4 e0 = __ldg(rec), e1 = __ldg(rec + 1), e2 = __ldg(rec + 2);
          const float4 v0p = __ldg(rec + REC_V0), gxp = __ldg(rec + REC_DDX), gyp = __ldg(rec + REC_DDY);
          const RasterParams& p = batch[min(__float_as_uint(gxp.x), n_draws - 1)];  // draw id rides in the unused d(pos.x)/dx
          const uint32_t bits = (p.read_depth ? 1u : 0u) | (p.write_depth ? 2u : 0u) |
                                ((p.depth_enable ? compare_lut(p.depth_func) : 0xFu) << 4) | (st4 << 8);
          CovTri ent;
          ent.e0 = make_float4(e0.x, e0.y, e0.z, e1.x);
          ent.e1 = make_float4(e1.y, e1.z, e2.x, e2.y);
          ent.e2 = make_float4(e2.z, v0p.x, v0p.y, v0p.z);
          ent.e3 = make_float4(gxp.z, gyp.z, __uint_as_float(slot), __uint_as_float(bits));
          float aa[4];
#pragma unroll
          for (int s = 0; s < 4; ++s)
            aa[s] = (s < S && S > 1) ? (SamplePattern<S>::x(s) - 0.5f) * gxp.z + (SamplePattern<S>::y(s) - 0.5f) * gyp.z : 0.0f;
          ent.aa = make_float4(aa[0], aa[1], aa[2], aa[3]);
          s_tri[lane] = ent;
        }
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, ei < rcnt);
        __syncwarp();  // orders the shared-memory writes above before the reads below
        if (lane == 0) n_pairs += __popc(todo);
        // ---- the chunk's surviving triangles, in API order ----
        while (todo) {
          const int j = __ffs(todo) - 1;
          todo &= todo - 1;
          const CovTri& t = s_tri[j];
          const float4 e3 = t.e3;
          const uint32_t bits = __float_as_uint(e3.w), slot = __float_as_uint(e3.z);
          const int blk = (bits >> (8 + 2 * bsel)) & 3;  // 0 rejected, 1 partial, 2 full
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
            const float4 e2 = t.e2, aa4 = t.aa;
            const float dx = hx - e2.y, dy = hy - e2.z;
            float depth = e2.w + (e3.x * dx + e3.y * dy);
            if (odd_x) depth += e3.x;
            if (odd_y) depth += e3.y;
            const uint32_t lut = (bits >> 4) & 0xFu;
            const bool rd = bits & 1u, wr = bits & 2u;
            const float aa[4] = {aa4.x, aa4.y, aa4.z, aa4.w};
            if (lut == 0x1u && rd && wr) {  // compare_less, depth read + write: the common state, one FSETP per sample
#pragma unroll
              for (int s = 0; s < S; ++s) {
                const float nd = (S == 1) ? depth : aa[s] + depth;
                if ((pm & (1u << s)) && nd < z[s]) {
                  tested |= 1u << s;
                  own[s] = slot;
                  z[s] = nd;
                }
              }
            } else {
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
            }
            const uint32_t np = __popc(pm), nt = __popc(tested);
            n_ztest += rd ? np : 0u;
            n_zwrite += wr ? nt : 0u;
            n_cwrite += d.vis ? nt : 0u;
            dirty |= wr && nt;
          }
          // ps_invocations: one quad per 2x2 with a live sample after early-Z (rasterizer.cpp:1274-1321)
          const uint32_t tb = __ballot_sync(0xFFFFFFFFu, tested != 0);
          if (lane == 0) n_ps_quads += __popc((tb | (tb >> 1) | (tb >> 2) | (tb >> 3)) & 0x11111111u);
        }
        __syncwarp();  // s_tri is overwritten by the next chunk
      }

      // ---- write back: owners when any were recorded (k_shade skips unflagged items), depth when modified ----
      bool any_own = false;
#pragma unroll
      for (int s = 0; s < S; ++s) any_own |= own[s] != VIS_NONE;
      const bool flag = __any_sync(0xFFFFFFFFu, any_own);
      if (lane == 0) d.item_flag[item] = flag ? 1 : 0;
      if (flag && in_target && d.vis) {
        uint32_t* vp = d.vis + ((size_t)y * d.vis_pitch + x) * S;
        if (S == 4) *reinterpret_cast<uint4*>(vp) = make_uint4(own[0], own[1 % S], own[2 % S], own[3 % S]);
        else if (S == 2) *reinterpret_cast<uint2*>(vp) = make_uint2(own[0], own[1 % S]);
        else *vp = own[0];
      }
      {
        const bool any_ds = __any_sync(0xFFFFFFFFu, dirty);
        if (in_target && any_ds && c.ds.data) {
          uint8_t* ds_ptr = c.ds.data + ((size_t)y * c.ds.w + x) * S * 8;
          // depth is not read again this frame: streaming (evict-first) stores keep it from pushing the visibility surface and
          // the textures, which k_shade is about to read, out of the L2
          if (S == 4) {
            st_stream(reinterpret_cast<float4*>(ds_ptr), make_float4(z[0], __uint_as_float(st[0]), z[1 % S], __uint_as_float(st[1 % S])));
            st_stream(reinterpret_cast<float4*>(ds_ptr + 16), make_float4(z[2 % S], __uint_as_float(st[2 % S]), z[3 % S], __uint_as_float(st[3 % S])));
          } else if (S == 2) {
            st_stream(reinterpret_cast<float4*>(ds_ptr), make_float4(z[0], __uint_as_float(st[0]), z[1 % S], __uint_as_float(st[1 % S])));
          } else {
            *reinterpret_cast<float2*>(ds_ptr) = make_float2(z[0], __uint_as_float(st[0]));
          }
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
  float w11;               // pos.w of pixel 3 (its reciprocal is only taken by the lanes / programs that need it)
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
    const float4 a0 = __ldg(rec + REC_V0 + 3 * (1 + i));
    float x00 = a0.x, y00 = a0.y, x01 = a0.x, y01 = a0.y, x10 = a0.x, y10 = a0.y;
    if (!(mod & SLV_AM_NOINTERPOLATION)) {
      const float4 gx = __ldg(rec + REC_DDX + 3 * (1 + i)), gy = __ldg(rec + REC_DDY + 3 * (1 + i));
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
  // ... at all four pixels (pixel 3 = (a00 + ddx) + ddy, the order in which interp_attr steps an odd / odd pixel)
  __device__ __forceinline__ void quad_xy4(int i, float4 a, float (&u)[4], float (&v)[4]) const {
    quad_xy(i, a, u[0], v[0], u[1], v[1], u[2], v[2]);
    const uint32_t mod = mods[i];
    const float4 a0 = __ldg(rec + REC_V0 + 3 * (1 + i));
    float x11 = a0.x, y11 = a0.y;
    if (!(mod & SLV_AM_NOINTERPOLATION)) {
      const float4 gx = __ldg(rec + REC_DDX + 3 * (1 + i)), gy = __ldg(rec + REC_DDY + 3 * (1 + i));
      x11 = ((a0.x + (gx.x * dx + gy.x * dy)) + gx.x) + gy.x;
      y11 = ((a0.y + (gx.y * dx + gy.y * dy)) + gx.y) + gy.y;
    }
    if (!(mod & SLV_AM_NOPERSPECTIVE)) {
      const float iw11 = 1.0f / w11;
      x11 *= iw11; y11 *= iw11;
    }
    u[3] = x11; v[3] = y11;
  }
  __device__ __forceinline__ int quad_index() const { return (odd_x ? 1 : 0) | (odd_y ? 2 : 0); }
};

template <int PS>
__device__ __forceinline__ uint32_t shade_sample_owner(const RasterParams& c, const RasterParams* __restrict__ batch, uint32_t n_draws,
                                                       uint32_t slot, int x, int y) {
  const float4* rec = c.tris + (size_t)slot * c.tri_stride;
  const float4 v0p = __ldg(rec + REC_V0), gxp = __ldg(rec + REC_DDX), gyp = __ldg(rec + REC_DDY);
  const RasterParams& p = batch[min(__float_as_uint(gxp.x), n_draws - 1)];  // draw id rides in the unused d(pos.x)/dx
  const int R = 1 + (int)p.n_attrs;
  DeferredCtx px;
  px.rec = rec; px.R = R; px.mods = p.mods;
  px.dx = 0.5f + (float)(uint32_t)(x & ~1) - v0p.x;
  px.dy = 0.5f + (float)(uint32_t)(y & ~1) - v0p.y;
  px.odd_x = x & 1; px.odd_y = y & 1;
  // step_2d_unproj_pos_quad (shader.cpp:257-287): pos00, pos01 = pos00 + ddx, pos10 = pos00 + ddy, pos11 = pos01 + ddy
  const float w00 = v0p.w + (gxp.w * px.dx + gyp.w * px.dy);
  const float w01 = w00 + gxp.w, w10 = w00 + gyp.w, w11 = w01 + gyp.w;
  px.iw00 = 1.0f / w00; px.iw01 = 1.0f / w01; px.iw10 = 1.0f / w10;
  px.w11 = w11;
  px.inv_w = px.odd_y ? (px.odd_x ? 1.0f / w11 : px.iw10) : (px.odd_x ? px.iw01 : px.iw00);
  float4 color;
  run_ps<PS>(p, px, color);
  return pack_color(c.color0.fmt, color);
}

#ifndef SLV_SHADE_GROUP
#define SLV_SHADE_GROUP 4
#endif
constexpr uint32_t SHADE_GROUP = SLV_SHADE_GROUP;   // items a warp of k_shade takes at a time; all their (pixel, owner) pairs share one pool
#ifndef SLV_SHADE_POOL
#define SLV_SHADE_POOL 256
#endif
constexpr int SHADE_POOL = SLV_SHADE_POOL;       // pool entries per warp (an item adds at most 32 * S = 128)

__device__ __forceinline__ uint32_t fetch_group(uint32_t* counter, uint32_t lane, uint32_t grp) {
  uint32_t v = 0;
  if (lane == 0) v = atomicAdd(counter, grp);
  return v;
}

// ---- pieces shared by k_shade and shade_quad_main (the pixel- and the quad-granular form) -----------------------------------
// One item (8x4 warp block) of a shading group: owners of the lane's pixel, the per-pixel flags and the colour row the
// shader results are merged into.  false: the item needs no visit (warp-uniform).
template <int S>
__device__ __forceinline__ bool shade_item_setup(const RasterParams& c, const DeferredBufs& d, uint32_t item, uint32_t k, uint32_t grp,
                                                 uint32_t n_items, uint32_t lane, int wlx, int wly, uint32_t* s_org,
                                                 uint8_t (*s_touched)[32], uint32_t (*s_color)[32][S], uint32_t (&own)[S], uint32_t& rem) {
  const uint32_t fullmask = (1u << S) - 1;
  const bool flagged = k < grp && item < n_items && d.item_flag[item];
  // lazy colour clear / fused resolve: every item of the active tiles is visited, not only those with new owners
  const bool live = flagged || ((d.lazy_color || d.resolve_dst.data) && k < grp && item < n_items);
  if (lane == 0) s_org[k] = 0xFFFFFFFFu;
  if (!live) return false;
  const uint32_t b = item >> 7, sub = (item >> 3) & 15, w = item & 7;
  const uint32_t tile = c.active_tiles[1 + b];
  const uint32_t tile_x = tile % c.tiles_x, tile_y = tile / c.tiles_x;
  const int gx0 = tile_x * TILE + (sub & 3) * REGION + (w & 1) * 8, gy0 = tile_y * TILE + (sub >> 2) * REGION + (w >> 1) * 4;
  const int x = gx0 + wlx, y = gy0 + wly;
  const bool in_target = (uint32_t)x < c.target_w && (uint32_t)y < c.target_h;
  if (lane == 0) s_org[k] = (uint32_t)gx0 | ((uint32_t)gy0 << 16);

#pragma unroll
  for (int s = 0; s < S; ++s) own[s] = VIS_NONE;
  if (in_target && flagged) {  // unflagged items hold stale owners from an earlier batch
    const uint32_t* vp = d.vis + ((size_t)y * d.vis_pitch + x) * S;
    if (S == 4) {
      const uint4 v = ld_stream(reinterpret_cast<const uint4*>(vp));  // read exactly once
      own[0] = v.x; own[1 % S] = v.y; own[2 % S] = v.z; own[3 % S] = v.w;
    } else if (S == 2) {
      const uint2 v = *reinterpret_cast<const uint2*>(vp);
      own[0] = v.x; own[1 % S] = v.y;
    } else {
      own[0] = *vp;
    }
  }
  rem = 0;
#pragma unroll
  for (int s = 0; s < S; ++s) rem |= (own[s] != VIS_NONE) ? (1u << s) : 0u;
  const uint32_t touched = rem;
  // bit 7: the pixel is written even when no sample was touched (lazy colour clear, pixels inside the target)
  // bit 6: the pixel lies inside the target (fused resolve)
  s_touched[k][lane] = (uint8_t)(touched | ((d.lazy_color && in_target) ? 0x80u : 0u) | (in_target ? 0x40u : 0u));
  if (d.lazy_color) {
    if (touched != fullmask) {
#pragma unroll
      for (int s = 0; s < S; ++s) s_color[k][lane][s] = d.clear_color;
    }
  } else if (touched != fullmask && in_target && (touched || d.resolve_dst.data)) {
    // some samples keep their colour: fetch it for the 128-bit store (and for the resolve of untouched pixels)
    const uint32_t* cptr = reinterpret_cast<const uint32_t*>(c.color0.data + ((size_t)y * c.color0.w + x) * S * 4);
#pragma unroll
    for (int s = 0; s < S; ++s) s_color[k][lane][s] = cptr[s];
  }
  return true;
}

// The end of a group: one 128-bit colour store per touched pixel, and the fused MSAA resolve.
template <int S>
__device__ __forceinline__ void shade_group_store(const RasterParams& c, const DeferredBufs& d, uint32_t lane, int wlx, int wly, uint32_t below,
                                                  const uint32_t* s_org, const uint8_t (*s_touched)[32], const uint32_t (*s_color)[32][S],
                                                  uint2* s_pool) {
  uint32_t n_slow = 0;
#pragma unroll 1
  for (uint32_t kk = 0; kk < SHADE_GROUP; ++kk) {
    const uint32_t org = s_org[kk];
    if (org == 0xFFFFFFFFu) continue;
    const uint32_t fl = s_touched[kk][lane];
    const int x = (int)(org & 0xFFFF) + wlx, y = (int)(org >> 16) + wly;
    if (fl & 0x8Fu) {
      uint32_t* cptr = reinterpret_cast<uint32_t*>(c.color0.data + ((size_t)y * c.color0.w + x) * S * 4);
      if (S == 4) st_stream(reinterpret_cast<uint4*>(cptr), make_uint4(s_color[kk][lane][0], s_color[kk][lane][1 % S], s_color[kk][lane][2 % S], s_color[kk][lane][3 % S]));
      else if (S == 2) *reinterpret_cast<uint2*>(cptr) = make_uint2(s_color[kk][lane][0], s_color[kk][lane][1 % S]);
      else *cptr = s_color[kk][lane][0];
    }
    // fused resolve.  Pixels whose S samples are equal (the great majority) resolve to that very value: for unorm8 c,
    // ((v+v)+v)+v with v = c/255 is off 4v by < 3 ulp, so * (1/S) * 255 lies within 1e-4 of c and rounds back to c.
    // The others (triangle edges) are pooled over the group's items and resolved densely below.
    bool slow = false;
    if (d.resolve_dst.data && (fl & 0x40u)) {
      bool same = d.resolve_dst.fmt == c.color0.fmt;
#pragma unroll
      for (int s = 1; s < S; ++s) same = same && s_color[kk][lane][s] == s_color[kk][lane][0];
      if (same) *reinterpret_cast<uint32_t*>(d.resolve_dst.data + ((size_t)y * d.resolve_dst.w + x) * 4) = s_color[kk][lane][0];
      slow = !same;
    }
    if (d.resolve_dst.data) {
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, slow);
      if (slow) s_pool[n_slow + __popc(bal & below)] = make_uint2(lane | (kk << 5), 0u);
      n_slow += __popc(bal);
    }
  }
  if (n_slow) {  // sum of to_rgba32f(sample) in sample order, * (1 / S), convert (RNE)  (surface.cpp:123-140)
    __syncwarp();
    for (uint32_t j = lane; j < n_slow; j += 32) {
      const uint32_t pl = s_pool[j].x & 31, k2 = s_pool[j].x >> 5;
      const uint32_t org = s_org[k2];
      const uint32_t pq = pl >> 2, pp = pl & 3;
      const int x = (int)(org & 0xFFFF) + (int)((pq & 3) * 2 + (pp & 1)), y = (int)(org >> 16) + (int)((pq >> 2) * 2 + (pp >> 1));
      float4 clr = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float4 t = unpack_color(c.color0.fmt, s_color[k2][pl][s]);
        clr.x += t.x; clr.y += t.y; clr.z += t.z; clr.w += t.w;
      }
      const float inv = 1 / (float)S;
      clr.x *= inv; clr.y *= inv; clr.z *= inv; clr.w *= inv;
      store_texel_rgba32f(d.resolve_dst.fmt, d.resolve_dst.data + ((size_t)y * d.resolve_dst.w + x) * d.resolve_dst.bpp, clr);
    }
  }
}

template <int S, int PS>
__global__ void __launch_bounds__(DEF_THREADS, SLV_SHADE_CTAS_PER_SM)
    k_shade(RasterParams c, const RasterParams* __restrict__ batch, uint32_t n_draws, DeferredBufs d) {
  // per warp: the colour rows of the group's items, the pool of (pixel, owner) pairs, per-item origin / touched masks
  __shared__ uint32_t s_color_all[DEF_WARPS][SHADE_GROUP][32][S];
  __shared__ uint2 s_pool_all[DEF_WARPS][SHADE_POOL];  // x = lane | mask << 5 | item-in-group << 9, y = owner slot
  __shared__ uint32_t s_org_all[DEF_WARPS][SHADE_GROUP];  // gx0 | gy0 << 16 of the item's 8x4 block
  __shared__ uint8_t s_touched_all[DEF_WARPS][SHADE_GROUP][32];

  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t (*s_color)[32][S] = s_color_all[wid];
  uint2* s_pool = s_pool_all[wid];
  uint32_t* s_org = s_org_all[wid];
  uint8_t (*s_touched)[32] = s_touched_all[wid];
  const int q = lane >> 2, pi = lane & 3;
  const int wlx = (q & 3) * 2 + (pi & 1), wly = (q >> 2) * 2 + (pi >> 1);  // same pixel <-> lane map as k_cover
  const uint32_t below = (1u << lane) - 1;

  uint32_t n_exec = 0;
  const uint32_t n_items = c.active_tiles[0] * ITEMS_PER_TILE;
  // items per fetch: SHADE_GROUP when there is plenty of work, fewer when the launch has only a few items per warp
  // (sort-first shards, small targets) so that the heavy items spread over all warps
  const uint32_t grp = d.shade_grp ? d.shade_grp : min(SHADE_GROUP, max(1u, n_items / (gridDim.x * DEF_WARPS * 4u)));
  uint32_t next_raw = fetch_group(d.shade_counter, lane, grp);
  uint32_t fetches = 1;
  for (;;) {
    const uint32_t base_item = __shfl_sync(0xFFFFFFFFu, next_raw, 0);
    if (base_item >= n_items) break;
    next_raw = (d.shade_budget == 0 || fetches < d.shade_budget) ? fetch_group(d.shade_counter, lane, grp) : 0xFFFFFFFFu;
    ++fetches;
    uint32_t pool_n = 0, k = 0;
    for (;;) {
      // ---- fill: analyse items while the pool has room for a whole item ----
#pragma unroll 1
      for (; k < SHADE_GROUP && pool_n + 32 * S <= (uint32_t)SHADE_POOL; ++k) {
        uint32_t own[S], rem;
        if (!shade_item_setup<S>(c, d, base_item + k, k, grp, n_items, lane, wlx, wly, s_org, s_touched, s_color, own, rem)) continue;
        // the pixel's distinct owners -> the pool, one ballot-compacted round per owner rank (order irrelevant:
        // every (pixel, sample) has exactly one writer); round 0 = everybody's first owner, dense for covered blocks
#pragma unroll
        for (int round = 0; round < S; ++round) {
          if (round > 0 && !__any_sync(0xFFFFFFFFu, rem != 0)) break;
          uint32_t sl = VIS_NONE, m = 0;
#pragma unroll
          for (int s = S - 1; s >= 0; --s)
            if (rem & (1u << s)) sl = own[s];
#pragma unroll
          for (int s = 0; s < S; ++s)
            if ((rem & (1u << s)) && own[s] == sl) m |= 1u << s;
          rem &= ~m;
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, m != 0);
          if (m) s_pool[pool_n + __popc(bal & below)] = make_uint2(lane | (m << 5) | (k << 9), sl);
          pool_n += __popc(bal);
        }
      }
      // ---- drain (the ONE shading call site): 32 (pixel, owner) pairs per round, whatever items they come from ----
      __syncwarp();
#pragma unroll 1
      for (uint32_t j = lane; j < pool_n; j += 32) {
        const uint2 it = s_pool[j];
        const uint32_t pl = it.x & 31, m = (it.x >> 5) & 0xF, kk = it.x >> 9;
        const uint32_t org = s_org[kk];
        const uint32_t pq = pl >> 2, pp = pl & 3;
        const int px_ = (int)(org & 0xFFFF) + (int)((pq & 3) * 2 + (pp & 1)), py_ = (int)(org >> 16) + (int)((pq >> 2) * 2 + (pp >> 1));
        const uint32_t packed = shade_sample_owner<PS>(c, batch, n_draws, it.y, px_, py_);
        ++n_exec;
#pragma unroll
        for (int s = 0; s < S; ++s)
          if (m & (1u << s)) s_color[kk][pl][s] = packed;
      }
      __syncwarp();
      pool_n = 0;
      if (k >= SHADE_GROUP) break;
    }
    // ---- store the group's items, one 128-bit store per touched pixel at 4x ----
    shade_group_store<S>(c, d, lane, wlx, wly, below, s_org, s_touched, s_color, s_pool);
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_exec += __shfl_xor_sync(0xFFFFFFFFu, n_exec, o);
  if (lane == 0 && n_exec) atomicAdd(&c.stats[16], (unsigned long long)n_exec);
}

// ---- quad-granular shading: run-time compiled SASL pixel shaders on the visibility-first path -------------------------
// A SASL pixel shader differences ARBITRARY expressions across the 2x2 quad (ddx / ddy, tex2D's implicit gradients:
// sasl/src/codegen/cgs_simd.cpp:275-313), so its four pixels must execute together in four consecutive lanes, helper pixels
// included, exactly as k_raster's shading phase runs them.  shade_quad_main is k_shade with the pool holding (quad,
// distinct owner) pairs instead of (pixel, distinct owner) pairs: 8 pairs per round, lane 4e + i shades pixel i of pair e.
// Every owner of any sample of the quad is shaded once for the whole quad (the reference shades the quad once per
// triangle that has a live sample in it; the visibility-first argument at the top of the file applies unchanged).
template <int PS>
__device__ __forceinline__ uint32_t shade_quad_owner(const RasterParams& c, const RasterParams* __restrict__ batch, uint32_t n_draws,
                                                     uint32_t slot, int x, int y, uint32_t quad_base) {
  const float4* rec = c.tris + (size_t)slot * c.tri_stride;
  // step_2d_unproj_pos_quad (shader.cpp:257-287), in k_raster's order of operations
  const float4 v0p = __ldg(rec + REC_V0), gxp = __ldg(rec + REC_DDX), gyp = __ldg(rec + REC_DDY);
  const RasterParams& p = batch[min(__float_as_uint(gxp.x), n_draws - 1)];  // draw id rides in the unused d(pos.x)/dx
  PixelCtx px;
  px.rec = rec; px.R = 1 + (int)p.n_attrs; px.mods = p.mods;
  px.dx = 0.5f + (float)(uint32_t)(x & ~1) - v0p.x;
  px.dy = 0.5f + (float)(uint32_t)(y & ~1) - v0p.y;
  px.odd_x = x & 1; px.odd_y = y & 1;
  float pw = v0p.w + (gxp.w * px.dx + gyp.w * px.dy);
  if (px.odd_x) pw += gxp.w;
  if (px.odd_y) pw += gyp.w;
  px.inv_w = 1.0f / pw;
  px.quad_base = quad_base;
  px.centroid_path = false;  // batches with centroid attributes never take this path
  px.pdx = 0.0f; px.pdy = 0.0f;
  float4 color;
  run_ps<PS>(p, px, color);
  return pack_color(c.color0.fmt, color);
}

template <int S, int PS>
__device__ __forceinline__ void shade_quad_main(const RasterParams& c, const RasterParams* __restrict__ batch, uint32_t n_draws,
                                                const DeferredBufs& d) {
  __shared__ uint32_t s_color_all[DEF_WARPS][SHADE_GROUP][32][S];
  __shared__ uint2 s_pool_all[DEF_WARPS][SHADE_POOL];  // x = quad | item-in-group << 3 | 4 x S-bit sample masks << 6, y = owner slot
  __shared__ uint32_t s_org_all[DEF_WARPS][SHADE_GROUP];
  __shared__ uint8_t s_touched_all[DEF_WARPS][SHADE_GROUP][32];

  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t (*s_color)[32][S] = s_color_all[wid];
  uint2* s_pool = s_pool_all[wid];
  uint32_t* s_org = s_org_all[wid];
  uint8_t (*s_touched)[32] = s_touched_all[wid];
  const uint32_t q = lane >> 2, pi = lane & 3;
  const int wlx = (q & 3) * 2 + (pi & 1), wly = (q >> 2) * 2 + (pi >> 1);  // same pixel <-> lane map as k_cover
  const uint32_t below = (1u << lane) - 1;

  uint32_t n_exec = 0;
  const uint32_t n_items = c.active_tiles[0] * ITEMS_PER_TILE;
  const uint32_t grp = d.shade_grp ? d.shade_grp : min(SHADE_GROUP, max(1u, n_items / (gridDim.x * DEF_WARPS * 4u)));
  uint32_t next_raw = fetch_group(d.shade_counter, lane, grp);
  uint32_t fetches = 1;
  for (;;) {
    const uint32_t base_item = __shfl_sync(0xFFFFFFFFu, next_raw, 0);
    if (base_item >= n_items) break;
    next_raw = (d.shade_budget == 0 || fetches < d.shade_budget) ? fetch_group(d.shade_counter, lane, grp) : 0xFFFFFFFFu;
    ++fetches;
    uint32_t pool_n = 0, k = 0;
    for (;;) {
      // ---- fill: an item adds at most 8 quads x min(4 S, 16) owners = 32 S pairs ----
#pragma unroll 1
      for (; k < SHADE_GROUP && pool_n + 32 * S <= (uint32_t)SHADE_POOL; ++k) {
        uint32_t own[S], rem;
        if (!shade_item_setup<S>(c, d, base_item + k, k, grp, n_items, lane, wlx, wly, s_org, s_touched, s_color, own, rem)) continue;
        // the quad's distinct owners -> the pool.  Per round every quad with samples left elects the owner of its lowest
        // such lane's lowest remaining sample; all four lanes hand in the samples they hold of that owner.
        for (;;) {
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, rem != 0);
          if (!bal) break;
          const uint32_t qbits = (bal >> (q * 4)) & 0xFu;
          const uint32_t leader = q * 4 + (qbits ? (uint32_t)__ffs(qbits) - 1u : 0u);
          uint32_t sl = VIS_NONE;
#pragma unroll
          for (int s = S - 1; s >= 0; --s)
            if (rem & (1u << s)) sl = own[s];
          const uint32_t qsl = __shfl_sync(0xFFFFFFFFu, sl, leader);
          uint32_t m = 0;
#pragma unroll
          for (int s = 0; s < S; ++s)
            if (qbits && (rem & (1u << s)) && own[s] == qsl) m |= 1u << s;
          rem &= ~m;
          uint32_t m16 = m << (4 * pi);
          m16 |= __shfl_xor_sync(0xFFFFFFFFu, m16, 1);
          m16 |= __shfl_xor_sync(0xFFFFFFFFu, m16, 2);
          const bool writer = qbits && lane == leader;
          const uint32_t wb = __ballot_sync(0xFFFFFFFFu, writer);
          if (writer) s_pool[pool_n + __popc(wb & below)] = make_uint2(q | (k << 3) | (m16 << 6), qsl);
          pool_n += __popc(wb);
        }
      }
      // ---- drain (the ONE shading call site): 8 (quad, owner) pairs per round, all 32 lanes converged ----
      __syncwarp();
#pragma unroll 1
      for (uint32_t j0 = 0; j0 < pool_n; j0 += 8) {
        const uint32_t e = j0 + q;
        const bool valid = e < pool_n;
        const uint2 it = s_pool[valid ? e : pool_n - 1];
        const uint32_t eq = it.x & 7, kk = (it.x >> 3) & 7;
        const uint32_t m = valid ? (it.x >> (6 + 4 * pi)) & 0xFu : 0u;
        const uint32_t org = s_org[kk];
        const int px_ = (int)(org & 0xFFFF) + (int)((eq & 3) * 2 + (pi & 1)), py_ = (int)(org >> 16) + (int)((eq >> 2) * 2 + (pi >> 1));
        const uint32_t packed = shade_quad_owner<PS>(c, batch, n_draws, it.y, px_, py_, lane & ~3u);
        if (valid) ++n_exec;
        const uint32_t pl = eq * 4 + pi;
#pragma unroll
        for (int s = 0; s < S; ++s)
          if (m & (1u << s)) s_color[kk][pl][s] = packed;
      }
      __syncwarp();
      pool_n = 0;
      if (k >= SHADE_GROUP) break;
    }
    shade_group_store<S>(c, d, lane, wlx, wly, below, s_org, s_touched, s_color, s_pool);
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n_exec += __shfl_xor_sync(0xFFFFFFFFu, n_exec, o);
  if (lane == 0 && n_exec) atomicAdd(&c.stats[16], (unsigned long long)n_exec);
}

}  // namespace slv
