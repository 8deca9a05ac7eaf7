// salvia_b200.cu — the C ABI of include/salvia_b200.h over the sm_100a kernels in slv_kernels.cuh.
//
// Host side of the product: resource table (buffers / textures / samplers in HBM), per-draw resolution of
// the reference's render_state (render_state.h:46-94) into POD kernel parameter blocks, and the kernel
// graph of one draw:   k_geometry -> k_scan_tiles -> k_bin_fill -> k_sort_lists -> k_raster.
// There is NO CPU implementation behind these entry points: without a usable CUDA device
// slv_device_create fails and nothing else can be called.
//
// Build (see __graft_entry__.build): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
//        -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -shared -Xcompiler -fPIC

#include <cuda.h>  // types only: the driver entry points are bound with dlsym (no link-time dependency on libcuda)
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

#include "salvia_b200.h"
#include "slv_kernels.cuh"
#include "slv_deferred.cuh"

using namespace slv;

namespace {

struct Resource {
  enum Kind { NONE, BUFFER, TEXTURE, SAMPLER, MODULE } kind = NONE;
  uint8_t* dptr = nullptr;  // buffers
  size_t bytes = 0;
  TextureRef tex{};         // textures: every level is its own allocation
  uint32_t fmt = 0, samples = 1;
  slv_sampler_desc sd{};    // samplers
  slv_handle sampler_tex = 0;
  uint8_t* resolve_peer = nullptr;  // textures: slv_resolve into this texture writes the owned tiles here instead (peer memory)
  // textures: a whole-surface clear that has been recorded but not executed (lazy clear).  The next visibility-first batch
  // that renders to the surface starts from this value without reading or pre-filling it; anything else materialises it.
  bool clear_pending = false;
  uint4 clear_pattern{};
  // textures: an asynchronous readback (slv_texture_readback_async) still reading this texture on the copy stream; the next
  // writer of the texture waits for it on the device
  cudaEvent_t rb_event = nullptr;
  bool rb_pending = false;
  // ... or a sort-first assembly wait (slv_assembly_wait) on the copy stream: READERS on the main stream wait for it as well
  bool asm_pending = false;
  // shader modules (SASL shaders compiled at run time, salviarenderer_b200/sasl): the pipeline kernels with the shader inlined
  CUmodule module = nullptr;
  uint32_t module_stage = 0;        // SLV_STAGE_VS / SLV_STAGE_PS
  uint32_t module_attrs = 0;        // VS: output attributes
  CUfunction fn_geometry = nullptr; // VS: slv_jit_k_geometry
  CUfunction fn_vertex_shade = nullptr; // VS: slv_jit_k_vertex_shade (post-transform vertex cache; absent in older cubins)
  CUfunction fn_geometry_cull = nullptr; // VS: slv_jit_k_geometry_cull (two-kernel geometry; absent in older cubins)
  CUfunction fn_raster[3] = {};     // PS: slv_jit_k_raster_s1 / _s2 / _s4
  CUfunction fn_shade[3] = {};      // PS: slv_jit_k_shade_s1 / _s2 / _s4 (visibility-first path; absent in older cubins)
};

// the four driver-API entry points run-time modules need, bound on first use
struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  bool ok = false;
};
DriverApi& driver_api() {
  static DriverApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.ModuleLoadData = (decltype(api.ModuleLoadData))dlsym(h, "cuModuleLoadData");
      api.ModuleGetFunction = (decltype(api.ModuleGetFunction))dlsym(h, "cuModuleGetFunction");
      api.ModuleUnload = (decltype(api.ModuleUnload))dlsym(h, "cuModuleUnload");
      api.LaunchKernel = (decltype(api.LaunchKernel))dlsym(h, "cuLaunchKernel");
      api.ok = api.ModuleLoadData && api.ModuleGetFunction && api.ModuleUnload && api.LaunchKernel;
    }
  }
  return api;
}

uint32_t bpp_of(uint32_t fmt) {
  switch (fmt) {
  case SLV_PF_RGBA32F: return 16;
  case SLV_PF_RG32F: return 8;
  case SLV_PF_RGBA8:
  case SLV_PF_BGRA8: return 4;
  }
  return 0;
}

uint8_t host_unorm8(float x) {  // colors.h:182-192
  float m = x * 255.0f;
  m = (m > 0.0f) ? m : 0.0f;
  m = (m < 255.0f) ? m : 255.0f;
  return (uint8_t)lrintf(m);
}

}  // namespace

struct slv_device_t {
  int ordinal = 0;
  cudaStream_t stream = nullptr;      // the stream every kernel / copy is enqueued on
  cudaStream_t own_stream = nullptr;  // created with the device; `stream` may point at a caller-owned one
  std::vector<Resource> res;
  // scratch written by the front half of a batch (geometry, binning, region lists) and read by its back half (coverage,
  // shading).  N_SETS sets: the front halves of batches k+1 .. run on the two front streams while the back half of batch k is
  // still running on `stream` (frame pipelining); a set is reused N_SETS batches later, once its ev_back_done has fired.
  struct Scratch {
    float4* tris = nullptr;           // triangle records (grown on demand, never shrunk)
    uint32_t* valid_slots = nullptr;  // one entry per slot of the tris arena
    uint32_t* valid_count = nullptr;
    uint32_t* big_slots = nullptr;    // triangles whose tile range k_big_tiles counts with a warp (one entry per slot of the tris arena)
    uint32_t* big_count = nullptr;
    uint32_t* surv = nullptr;         // two-kernel geometry: ids of the primitives k_geometry_cull hands to k_geometry (one range per draw)
    uint32_t* surv_count = nullptr;   // ... and their number per queued draw
    uint32_t *tile_count = nullptr, *tile_offset = nullptr, *tile_cursor = nullptr, *active_tiles = nullptr, *large_tiles = nullptr;
    uint32_t* work_counter = nullptr;  // [0] k_raster / k_cover queue head, [1] k_shade queue head, [2] region-list cursor, [3] long lists, [4] block-bits pool cursor
    uint32_t *region_list = nullptr, *region_offset = nullptr, *region_count = nullptr;  // deferred path: per-region lists
    uint32_t* region_mask = nullptr;  // one word per tile-list entry
    uint32_t* region_tile_cnt = nullptr;  // [tile * 16 + region] survivor counters (k_region_decide -> k_region_bin, self-cleaning)
    uint8_t* item_flag = nullptr;
    uint2* block_desc = nullptr;  // (first entry, entries) of each (region, warp block) sub-list
    uint32_t* list = nullptr;
    RasterParams* d_batch = nullptr;
    GeomParams* d_geom = nullptr;
    RasterParams* h_batch = nullptr;  // pinned staging of d_batch / d_geom: the uploads never synchronise a stream
    GeomParams* h_geom = nullptr;
    // post-transform vertex cache of the batch (k_vertex_mark / k_vertex_shade -> k_geometry): clip-space positions, attributes
    // and the "referenced" marks, one range per group of draws with the same vertex state; grown on demand
    float4 *vc_pos = nullptr, *vc_attr = nullptr;
    uint8_t* vc_flags = nullptr;
    size_t vc_pos_cap = 0, vc_attr_cap = 0;  // vertices / float4s
    cudaEvent_t ev_front_done = nullptr, ev_back_done = nullptr;
    bool in_flight = false;           // ev_back_done has been recorded and not yet waited for
  };
  static constexpr int N_SETS = 4;    // batches in flight: the back half of batch k, the front halves of k+1 .. k+3
  Scratch sc[N_SETS];
  int cur = 0;                        // the set the queued draws point at
  int last_flushed = -1;              // the set of the most recently flushed batch (-1: none since the last full sync)
  Scratch& S() { return sc[cur]; }
  size_t tris_cap = 0;  // float4 units (both sets)
  uint32_t region_cap = 0;
  // front halves run on these when pipelining, alternating per batch: a front half is a chain of short, latency-bound kernels
  // (scan, sort of the longest list, ordered region fill) that leaves most of the GPU idle, so TWO of them are kept in flight -
  // what a sort-first rank needs once its back half is an N-th of the frame and the front chain is the longer of the two
  cudaStream_t front_streams[2] = {nullptr, nullptr};
  cudaStream_t front_stream_of(int set) const { return front_streams[set & 1]; }
  cudaEvent_t ev_sync = nullptr;        // scratch event: orders the front streams after buffer uploads on `stream`
  bool buffers_dirty = false;           // a vertex / index buffer was written on `stream` since the last front half
  bool pipeline = true;                 // SLV_PIPELINE=0: everything on `stream`
  cudaStream_t copy_stream = nullptr;   // asynchronous readbacks (overlap the next frame's rendering)
  cudaStream_t signal_stream = nullptr; // sort-first root: frame-buffer releases (slv_peer_signal_after_consumers), created on first use
  cudaEvent_t ev_copy = nullptr;        // orders the copy stream after the producer of the texture on `stream`
  cudaEvent_t ev_upload = nullptr;      // last buffer upload enqueued on a front stream
  bool upload_on_front = false;         // ... and not yet ordered before work on `stream`
  uint32_t* peer_flags = nullptr;       // SLV_PEER_FLAGS words other ranks raise over NVLink (slv_peer_signal / slv_flags_wait)
  uint32_t* vis = nullptr;           // visibility buffer of the deferred path: owner slot per sample
  size_t vis_cap = 0;                // in uint32 units
  // post-transform vertex cache: indexed draws run the vertex shader once per referenced vertex instead of once
  // per corner.  1 (default): only where it pays - vertex shaders that sample a texture (vertex texture fetch: ~3 k instructions
  // per run, up to six runs per primitive without the cache).  For the arithmetic-only programs the per-corner recompute is
  // FASTER on this machine: it reads the 2-3 input registers of a corner where the cache would write and re-read the 4-5 output
  // registers (measured, profiles/r02_vertex_cache.txt: k_geometry 0.065 -> 0.081 ms on the Sponza-like scene, 5.66 -> 6.13 ms on
  // the 10 M-triangle mesh).  SLV_VERTEX_CACHE=0 disables, =2 caches every indexed draw whatever its size and shader (tests).
  int vertex_cache = 1;
  // Two experiments on the front half that are BUILT, parity-tested (tests/test_gpu_geometry_split.py) and OFF by default because
  // neither paid (profiles/r02_front_half_experiments.txt):
  // * two-kernel geometry (SLV_GEOMETRY_SPLIT=1): k_geometry_cull runs the position pass of every primitive at 64 registers / full
  //   occupancy and hands only the primitives that need set-up on this rank to k_geometry.  k_geometry 0.050 -> 0.059 ms on an
  //   eighth of the frame, 0.071 -> 0.085 ms on the whole frame, 5.7 -> 6.5 ms on the 10 M-triangle mesh: the pass is not
  //   occupancy-bound.
  // * big-triangle queue (SLV_BIG_TILES=1): triangles spanning more than BIG_TILE_RANGE tiles are counted by a warp of k_big_tiles
  //   instead of by the thread that set them up.  No change (0.070 -> 0.068 ms; an eighth of the frame 0.2005 -> 0.2045 ms per
  //   pipelined frame because of the extra launch): one thread's 2,040-tile loop is not the critical path either.
  int geometry_split = 0;
  bool big_tiles = false;
  bool force_immediate = false;      // SLV_FORCE_IMMEDIATE=1: always use k_raster (tests compare both paths)
  bool jit_immediate = false;        // SLV_JIT_IMMEDIATE=1: SASL pixel shaders always take k_raster
  bool front_grids = true;           // SLV_FRONT_GRIDS=0: full-size k_sort_lists / k_region_bin / k_sort_lists_large grids
  int sort_large_grid = 0;           // SLV_SORT_LARGE_GRID: CTAs of k_sort_lists_large (0 = by tile count)
  long long bits_pool_cap = -1;      // SLV_BITS_POOL_CAP: caps the block-bits pool (tests force k_region_bin's fallback evaluation)
  int cover_grid = 0, shade_grid = 0, sm_count = 0;
  int raster_grid = 0;  // persistent raster CTAs (SM count x resident CTAs per SM)
  // ---- draw batching: geometry + binning run at slv_draw time, the raster pass of all queued draws of a
  // frame runs at the next flush point (readback, clear, resolve, state that changes the targets ...), so each
  // pixel's depth/stencil/colour is loaded once and stored once per batch instead of once per draw.
  std::vector<RasterParams> pending;
  std::vector<GeomParams> pending_geom;  // geometry parameters of the queued draws (same index as `pending`)
  std::vector<slv_handle> pending_vs_module;  // run-time vertex-shader module of each queued draw (0 = built-in program)
  std::vector<uint64_t> pending_vs_runs;      // per-corner vertex-shader runs of each queued draw (counted at flush unless the draw is cached)
  slv_handle batch_ps_module = 0;             // run-time pixel-shader module of the batch (0 = built-in program)
  slv_handle batch_color = 0, batch_ds = 0;   // texture handles of the batch's colour target 0 and depth/stencil target
  bool lazy_clear = true;                     // SLV_LAZY_CLEAR=0: clears always execute immediately
  // fused resolve: slv_resolve(src, dst) of the pending batch's colour target is handed to the batch flush, whose k_shade
  // writes the resolved texels as it stores the samples (SLV_FUSE_RESOLVE=0 disables)
  // k_cover / k_shade grids: persistent (one CTA set per SM draining the work queue) or budgeted (short-lived CTAs, grid sized to
  // the queue; SLV_PERSISTENT=0).  Measured: persistent wins when a GPU has the whole frame (0.80 vs 0.87 ms); on an eighth of the
  // frame budgeted beat fully occupied persistent grids by ~3 %, but persistent grids that leave part of each SM free (back_ctas
  // below) beat both, so persistent is the default everywhere.
  int persistent = -1;
  // resident CTAs per SM of the persistent k_cover / k_shade grids (<= the 8 they are compiled for).  Leaving part of every SM
  // free lets the NEXT frame's front half (other stream) run concurrently instead of waiting for the persistent CTAs to drain.
  // Measured (ms/frame, 8 / 7 / 6 / 5 / 4 CTAs per SM): whole frame on one GPU 0.877 / 0.835 / 0.904 / 0.976 / 1.168; an eighth of
  // the frame 0.284 / 0.270 / 0.250 / 0.258 / 0.252.  0 = choose by shard count; SLV_BACK_CTAS overrides.
  int back_ctas = 0;
  bool fuse_resolve = true;
  bool resolve_requested = false, resolve_done = false;
  SurfaceRef resolve_dst{};
  size_t tris_used = 0;      // float4 units used by the queued draws
  uint64_t slots_queued = 0; // triangle slots of the queued draws (sizes the list arena)
  uint32_t batch_S = 0;
  // profiling event pool
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { size_t a, b; int stage; };
  std::vector<Span> spans;
  uint32_t tiles_cap = 0;
  uint32_t list_cap = 0;
  uint32_t* overflow_flag = nullptr;  // device: [0] flag (1 = an arena overflowed, 2 = a peer timed out), [1] list entries needed, [2] region words needed
  // Arena sizing.  The tile-list and region-list arenas are sized from the triangle count (a heuristic: a triangle can reach every
  // tile), with generous minima; kernels that run out of room drop the entry, raise the flag and record what they WOULD have
  // needed.  The next flush point reports SLV_OUT_OF_MEMORY for that frame - once, not sticky - and the arenas are grown to the
  // recorded need before the next batch, so re-issuing the frame succeeds.  SLV_ARENA_MIN (entries) shrinks the minima (tests).
  uint64_t list_need_hint = 0, region_need_hint = 0;
  uint64_t arena_min_list = 1ull << 22, arena_min_region = 1ull << 26;
  unsigned long long* d_stats = nullptr;
  uint32_t* d_level_touched = nullptr;  // slv_texture_level_tracking: one mask of sampled mip levels per texture handle
  bool level_tracking = false;
  slv_pipeline_statistics host_stats{};  // counters that are pure functions of the draw arguments
  uint32_t shard_rank = 0, shard_n = 1;
  bool failed = false;  // sticky CUDA error
  // profiling (SLV_PROFILE=1)
  bool profile = false;
  cudaEvent_t user_ev[16] = {};
  struct SlotTable { uint32_t tiles_x, tiles_y, rank, nranks, owned; uint32_t* d_slot; };
  std::vector<SlotTable> slot_tables;  // pack/unpack: dense slot of every owned tile, per (grid, rank, nranks)
  double prof_ms[6] = {0, 0, 0, 0, 0, 0};  // geometry, binning, sort, raster (k_raster or k_cover), shade, region bin
  unsigned long long n_launches = 0;

  Resource* get(slv_handle h, Resource::Kind k) {
    if (h == 0 || h >= res.size() || res[h].kind != k) return nullptr;
    return &res[h];
  }
};

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      fprintf(stderr, "[salvia_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), __FILE__,     \
              __LINE__, cudaGetErrorString(e__));                                                        \
      return SLV_FAILED;                                                                                 \
    }                                                                                                    \
  } while (0)

namespace {

constexpr uint32_t MAX_BATCH = MAX_BATCH_DRAWS;
constexpr uint32_t LEVEL_TRACK_SLOTS = 1u << 16;  // texture handles below this can be tracked  // draws whose geometry / raster passes are fused into one launch each

slv_result flush_batch(slv_device dev);
slv_result materialize_clear(slv_device dev, Resource* r);
slv_result wait_readback(slv_device dev, Resource* r);
slv_result wait_assembly(slv_device dev, Resource* r);

slv_result sync_all(slv_device dev) {  // both streams idle
  for (auto fsx : dev->front_streams) CU(cudaStreamSynchronize(fsx));
  CU(cudaStreamSynchronize(dev->stream));
  CU(cudaStreamSynchronize(dev->copy_stream));
  if (dev->signal_stream) CU(cudaStreamSynchronize(dev->signal_stream));
  for (auto& S : dev->sc) S.in_flight = false;
  dev->last_flushed = -1;
  for (auto& r : dev->res) r.rb_pending = false;
  dev->upload_on_front = false;
  return SLV_OK;
}

// Arenas are shared by the queued draws of a batch; growing one needs the batch flushed first.
slv_result ensure_scratch(slv_device dev, size_t tris_needed_total, uint32_t n_tiles, uint64_t list_needed_total) {
  if (tris_needed_total > dev->tris_cap) {
    slv_result rc = flush_batch(dev);
    if (rc != SLV_OK) return rc;
    rc = sync_all(dev);
    if (rc != SLV_OK) return rc;
    size_t need = tris_needed_total;  // after the flush only the new draw remains; callers pass used + new
    size_t cap = std::max(need, dev->tris_cap * 2);
    for (auto& S : dev->sc) {
      if (S.tris) CU(cudaFree(S.tris));
      if (S.valid_slots) CU(cudaFree(S.valid_slots));
      if (S.surv) CU(cudaFree(S.surv));
      if (S.big_slots) CU(cudaFree(S.big_slots));
      CU(cudaMalloc(&S.big_slots, (cap / (TRI_HEADER + 3 * MAX_REGS) + 1) * sizeof(uint32_t)));
      CU(cudaMalloc(&S.tris, cap * sizeof(float4)));
      CU(cudaMalloc(&S.valid_slots, (cap / (TRI_HEADER + 3 * MAX_REGS) + 1) * sizeof(uint32_t)));
      CU(cudaMalloc(&S.surv, (cap / (TRI_HEADER + 3 * MAX_REGS) / 3 + 1) * sizeof(uint32_t)));
    }
    dev->tris_cap = cap;
  }
  if (n_tiles + 1 > dev->tiles_cap) {
    slv_result rc = flush_batch(dev);
    if (rc != SLV_OK) return rc;
    rc = sync_all(dev);
    if (rc != SLV_OK) return rc;
    uint32_t cap = std::max(n_tiles + 1, 4096u);
    for (auto& S : dev->sc) {
      if (S.tile_count) {
        CU(cudaFree(S.tile_count)); CU(cudaFree(S.tile_offset)); CU(cudaFree(S.tile_cursor));
        CU(cudaFree(S.active_tiles)); CU(cudaFree(S.large_tiles));
        CU(cudaFree(S.region_offset)); CU(cudaFree(S.region_count)); CU(cudaFree(S.item_flag)); CU(cudaFree(S.block_desc));
        CU(cudaFree(S.region_tile_cnt));
      }
      CU(cudaMalloc(&S.tile_count, cap * sizeof(uint32_t)));
      CU(cudaMalloc(&S.tile_offset, cap * sizeof(uint32_t)));
      CU(cudaMalloc(&S.tile_cursor, cap * sizeof(uint32_t)));
      CU(cudaMalloc(&S.active_tiles, (cap + 1) * sizeof(uint32_t)));
      CU(cudaMalloc(&S.large_tiles, (cap + 1) * sizeof(uint32_t)));
      CU(cudaMalloc(&S.region_offset, (size_t)cap * 16 * sizeof(uint32_t)));
      CU(cudaMalloc(&S.region_count, (size_t)cap * 16 * sizeof(uint32_t)));
      CU(cudaMalloc(&S.item_flag, (size_t)cap * 128));
      CU(cudaMalloc(&S.block_desc, (size_t)cap * 128 * sizeof(uint2)));
      CU(cudaMemsetAsync(S.tile_count, 0, cap * sizeof(uint32_t), dev->stream));
      CU(cudaMemsetAsync(S.tile_cursor, 0, cap * sizeof(uint32_t), dev->stream));
      CU(cudaMalloc(&S.region_tile_cnt, (size_t)cap * 16 * sizeof(uint32_t)));
      CU(cudaMemsetAsync(S.region_tile_cnt, 0, (size_t)cap * 16 * sizeof(uint32_t), dev->stream));
    }
    CU(cudaStreamSynchronize(dev->stream));
    dev->tiles_cap = cap;
  }
  (void)list_needed_total;
  return SLV_OK;
}

cudaEvent_t next_event(slv_device dev) {
  if (dev->ev_used == dev->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    dev->ev_pool.push_back(e);
  }
  cudaEvent_t e = dev->ev_pool[dev->ev_used++];
  cudaEventRecord(e, dev->stream);
  return e;
}
size_t mark(slv_device dev) {  // records an event on the stream, returns its index
  next_event(dev);
  return dev->ev_used - 1;
}

SurfaceRef surface_of(slv_device dev, slv_handle h) {
  SurfaceRef s{};
  auto r = dev->get(h, Resource::TEXTURE);
  if (r) s = r->tex.level[0];
  return s;
}

uint32_t vs_num_attrs(const slv_shader_binding& vs) {
  switch (vs.program) {
  case SLV_VS_MVP_PASSTHROUGH: return reinterpret_cast<const slv_vs_mvp_passthrough_uniforms*>(vs.uniforms)->n_attrs;
  case SLV_VS_PLANE_XZ: return 1;
  case SLV_VS_LIGHTS3: return 4;
  case SLV_VS_SPONZA: return 4;
  case SLV_VS_TERRAIN_VTF: return 1;
  case SLV_VS_SSM_DRAW: return 5;
  }
  return 0xFFFFFFFFu;
}

template <int R>
void launch_geometry(const GeomParams* d_draws, const GeomBatch& hb, cudaStream_t st) {
  k_geometry<R><<<hb.cta_prefix[hb.n], 128, 0, st>>>(d_draws, hb);
}

template <int S>
bool launch_raster_s(const RasterParams& rp, const RasterParams* batch, uint32_t n, uint32_t blocks, cudaStream_t st) {
  switch (rp.ps_program) {
  case SLV_PS_ATTR0_COLOR: k_raster<S, SLV_PS_ATTR0_COLOR><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_LIGHTS3: k_raster<S, SLV_PS_LIGHTS3><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_TEX_ALPHA: k_raster<S, SLV_PS_TEX_ALPHA><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_SPONZA: k_raster<S, SLV_PS_SPONZA><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_TEX_GRAD_ALPHA: k_raster<S, SLV_PS_TEX_GRAD_ALPHA><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_DISCARD_ALL: k_raster<S, SLV_PS_DISCARD_ALL><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_HEIGHT_COLOR: k_raster<S, SLV_PS_HEIGHT_COLOR><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_SSM_DRAW: k_raster<S, SLV_PS_SSM_DRAW><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  case SLV_PS_SPONZA_GRAD: k_raster<S, SLV_PS_SPONZA_GRAD><<<blocks, RASTER_THREADS, 0, st>>>(rp, batch, n); return true;
  }
  return false;
}

template <int S>
bool launch_shade_s(const RasterParams& rp, const RasterParams* batch, uint32_t n_draws, const DeferredBufs& db, uint32_t blocks,
                    cudaStream_t st) {
  switch (rp.ps_program) {
  case SLV_PS_ATTR0_COLOR: k_shade<S, SLV_PS_ATTR0_COLOR><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_LIGHTS3: k_shade<S, SLV_PS_LIGHTS3><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_TEX_ALPHA: k_shade<S, SLV_PS_TEX_ALPHA><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_SPONZA: k_shade<S, SLV_PS_SPONZA><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_TEX_GRAD_ALPHA: k_shade<S, SLV_PS_TEX_GRAD_ALPHA><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_HEIGHT_COLOR: k_shade<S, SLV_PS_HEIGHT_COLOR><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_SSM_DRAW: k_shade<S, SLV_PS_SSM_DRAW><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  case SLV_PS_SPONZA_GRAD: k_shade<S, SLV_PS_SPONZA_GRAD><<<blocks, DEF_THREADS, 0, st>>>(rp, batch, n_draws, db); return true;
  }
  return false;
}

slv_result fill_surface(slv_device dev, const SurfaceRef& s, uint4 pattern) {
  if (dev->shard_n > 1) {  // sort-first: clear the owned tiles only
    const uint32_t tiles_x = (s.w + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE, tiles_y = (s.h + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE;
    k_fill_tiles<<<tiles_x * tiles_y, 256, 0, dev->stream>>>(s, pattern, tiles_x, dev->shard_rank, dev->shard_n);
    ++dev->n_launches;
    CU(cudaGetLastError());
    return SLV_OK;
  }
  size_t n_vec = s.bytes / 16, n_words = s.bytes / 4;
  if (n_vec) {
    uint32_t blocks = (uint32_t)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
    k_fill<<<blocks, 256, 0, dev->stream>>>(reinterpret_cast<uint4*>(s.data), n_vec, pattern);
    ++dev->n_launches;
  }
  if (n_words > n_vec * 4) {
    k_fill_words<<<1, 32, 0, dev->stream>>>(reinterpret_cast<uint32_t*>(s.data), n_vec * 4, n_words, pattern);
    ++dev->n_launches;
  }
  CU(cudaGetLastError());
  return SLV_OK;
}


// a texture about to be WRITTEN on the main stream: order the write after an asynchronous readback still copying it out
slv_result wait_readback(slv_device dev, Resource* r) {
  if (r && r->rb_pending) {
    CU(cudaStreamWaitEvent(dev->stream, r->rb_event, 0));
    r->rb_pending = false;
    r->asm_pending = false;
  }
  return SLV_OK;
}
// a texture about to be READ on the main stream: order the read after a pending sort-first assembly of it
slv_result wait_assembly(slv_device dev, Resource* r) {
  if (r && r->asm_pending) return wait_readback(dev, r);
  return SLV_OK;
}

// executes a recorded whole-surface clear (see Resource::clear_pending)
slv_result materialize_clear(slv_device dev, Resource* r) {
  if (!r || !r->clear_pending) return SLV_OK;
  r->clear_pending = false;
  { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
  return fill_surface(dev, r->tex.level[0], r->clear_pattern);
}
slv_result materialize_all(slv_device dev) {
  for (auto& r : dev->res)
    if (r.kind == Resource::TEXTURE && r.clear_pending) {
      slv_result rc = materialize_clear(dev, &r);
      if (rc != SLV_OK) return rc;
    }
  return SLV_OK;
}
Resource* texture_of_data(slv_device dev, const uint8_t* data) {
  if (!data) return nullptr;
  for (auto& r : dev->res)
    if (r.kind == Resource::TEXTURE && r.tex.level[0].data == data) return &r;
  return nullptr;
}

// Batch flush: geometry and binning (scan, fill, sort, region lists) over the triangles of every queued draw - the FRONT
// half, which touches no render target - then the raster pass of all of them, in submission order - the BACK half.
// When pipelining, the front half is enqueued on a front stream, so it overlaps the back half / clears / resolve of the
// previous batch still running on the main stream; the back half waits for it with an event.
slv_result flush_batch(slv_device dev) {
  if (dev->pending.empty()) return SLV_OK;
  cudaStream_t st = dev->stream;
  slv_device_t::Scratch& S = dev->S();
  const bool piped = dev->pipeline && !dev->profile;
  cudaStream_t fs = piped ? dev->front_stream_of(dev->cur) : st;
  RasterParams& first = dev->pending[0];
  const uint32_t n = (uint32_t)dev->pending.size();
  const uint32_t n_tiles = first.tiles_x * first.tiles_y;
  const uint32_t n_slots = (uint32_t)dev->slots_queued;
  // list arena: heuristic bound, checked on the device (overflow flag -> SLV_OUT_OF_MEMORY at the next flush point)
  const uint64_t list_need = std::min<uint64_t>(std::max<uint64_t>(std::max<uint64_t>(4ull * n_slots, dev->arena_min_list), dev->list_need_hint), 1ull << 31);
  const uint64_t region_need = std::min<uint64_t>(std::max<uint64_t>(std::max<uint64_t>(2ull * list_need, dev->arena_min_region), dev->region_need_hint), 0xFFFFFFF0ull);
  if (list_need > dev->list_cap || region_need > dev->region_cap) {
    slv_result rcs = sync_all(dev);
    if (rcs != SLV_OK) return rcs;
    // (region, warp block) sub-lists of the deferred path, allocated on the device inside the region arena: 8 sub-lists of
    // capacity n per region with n surviving entries; a tile-list entry survives in 1..16 regions (about 1.5 on the
    // Sponza-like scene, i.e. ~12 words per tile-list entry; a fully covered tile costs 128: see list_need_hint above)
    const uint64_t rcap = region_need;
    for (auto& T : dev->sc) {
      if (T.list) CU(cudaFree(T.list));
      CU(cudaMalloc(&T.list, (size_t)list_need * sizeof(uint32_t)));
      if (T.region_list) CU(cudaFree(T.region_list));
      CU(cudaMalloc(&T.region_list, (size_t)rcap * sizeof(uint32_t)));
      if (T.region_mask) CU(cudaFree(T.region_mask));
      CU(cudaMalloc(&T.region_mask, (size_t)list_need * RMASK_STRIDE * sizeof(uint32_t)));
    }
    dev->list_cap = (uint32_t)list_need;
    dev->region_cap = (uint32_t)rcap;
  }
  for (RasterParams& r : dev->pending) {  // the list arena may have grown since the draws were queued
    r.list = S.list;
    r.list_capacity = dev->list_cap;
  }
  if (piped) {
    // the front half reads vertex / index buffers: order it after the uploads enqueued on the main stream
    if (dev->buffers_dirty) {  // both front streams: the batch after this one reads the same buffers on the other stream
      CU(cudaEventRecord(dev->ev_sync, st));
      for (auto fsx : dev->front_streams) CU(cudaStreamWaitEvent(fsx, dev->ev_sync, 0));
    }
    // this scratch set was last read by the back half of the batch before the previous one
    if (S.in_flight) {
      CU(cudaEventSynchronize(S.ev_front_done));  // host: the pinned parameter staging of that batch has been consumed
      CU(cudaStreamWaitEvent(fs, S.ev_back_done, 0));
      S.in_flight = false;
    }
  }
  if (!piped && dev->upload_on_front) {
    CU(cudaStreamWaitEvent(st, dev->ev_upload, 0));
    dev->upload_on_front = false;
  }
  dev->buffers_dirty = false;
  // ---- post-transform vertex cache: group the indexed draws by vertex state, give each group a range of the cache arenas
  GeomBatch vc_mark{};                              // every cached draw (k_vertex_mark: CTAs over its indices)
  std::vector<uint32_t> vc_groups;                  // representative draw of each group (k_vertex_shade: CTAs over [0, vc_cap))
  {
    std::vector<int> group_of(n, -1);
    struct Group { uint32_t rep; uint64_t indices; };
    std::vector<Group> groups;
    auto n_indices = [](const GeomParams& g) -> uint64_t { return g.topology == SLV_TOPO_TRIANGLE_LIST ? 3ull * g.prim_count : (uint64_t)g.prim_count + 2; };
    auto same_vertex_state = [&](uint32_t a, uint32_t b) {
      const GeomParams &x = dev->pending_geom[a], &y = dev->pending_geom[b];
      return dev->pending_vs_module[a] == dev->pending_vs_module[b] && x.vs_program == y.vs_program && x.n_attrs == y.n_attrs &&
             x.n_elements == y.n_elements && x.fast_layout == y.fast_layout && x.vc_cap == y.vc_cap &&
             memcmp(x.streams, y.streams, sizeof(x.streams)) == 0 && memcmp(x.elements, y.elements, sizeof(x.elements)) == 0 &&
             memcmp(x.vs_uniforms, y.vs_uniforms, sizeof(x.vs_uniforms)) == 0 && memcmp(&x.sampler0, &y.sampler0, sizeof(SamplerRef)) == 0;
    };
    if (dev->vertex_cache > 0)
      for (uint32_t i = 0; i < n; ++i) {
        const GeomParams& g = dev->pending_geom[i];
        if (!g.indices || g.vc_cap == 0) continue;
        const slv_handle m = dev->pending_vs_module[i];
        if (m && !dev->res[m].fn_vertex_shade) continue;
        for (size_t k = 0; k < groups.size() && group_of[i] < 0; ++k)
          if (same_vertex_state(groups[k].rep, i)) group_of[i] = (int)k;
        if (group_of[i] < 0) { group_of[i] = (int)groups.size(); groups.push_back({i, 0}); }
        groups[group_of[i]].indices += n_indices(g);
      }
    // a group is cached when that is cheaper than re-running the shader per corner: an expensive (texture-sampling) shader,
    // enough index references for the two extra launches, a vertex range not much larger than what the draws can reference
    size_t pos_need = 0, attr_need = 0;
    std::vector<size_t> pos_off(groups.size()), attr_off(groups.size());
    std::vector<char> cached(groups.size(), 0);
    for (size_t k = 0; k < groups.size(); ++k) {
      const GeomParams& g = dev->pending_geom[groups[k].rep];
      const bool samples_texture = g.sampler0.tex.n_levels != 0;
      const bool worth = dev->vertex_cache >= 2 || (samples_texture && groups[k].indices >= 768 && (uint64_t)g.vc_cap <= 4 * groups[k].indices + 1024);
      if (!worth || (uint64_t)g.vc_cap * (1 + g.n_attrs) > (1ull << 27)) continue;  // <= 2 GB of cache per group
      cached[k] = 1;
      pos_off[k] = pos_need;
      attr_off[k] = attr_need;
      pos_need += g.vc_cap;
      attr_need += (size_t)g.vc_cap * std::max(g.n_attrs, 1u);
    }
    if (pos_need > S.vc_pos_cap || attr_need > S.vc_attr_cap) {
      slv_result rcs = sync_all(dev);  // the arenas of this set may still be read by a batch in flight
      if (rcs != SLV_OK) return rcs;
      if (pos_need > S.vc_pos_cap) {
        const size_t cap = std::max(pos_need, S.vc_pos_cap * 2);
        if (S.vc_pos) { CU(cudaFree(S.vc_pos)); CU(cudaFree(S.vc_flags)); }
        CU(cudaMalloc(&S.vc_pos, cap * sizeof(float4)));
        CU(cudaMalloc(&S.vc_flags, cap));
        CU(cudaMemsetAsync(S.vc_flags, 0, cap, fs));
        S.vc_pos_cap = cap;
      }
      if (attr_need > S.vc_attr_cap) {
        const size_t cap = std::max(attr_need, S.vc_attr_cap * 2);
        if (S.vc_attr) CU(cudaFree(S.vc_attr));
        CU(cudaMalloc(&S.vc_attr, cap * sizeof(float4)));
        S.vc_attr_cap = cap;
      }
    }
    for (uint32_t i = 0; i < n; ++i) {
      GeomParams& g = dev->pending_geom[i];
      const int k = group_of[i];
      if (k < 0 || !cached[k]) {
        g.vc_pos = nullptr; g.vc_attr = nullptr; g.vc_flags = nullptr;
        dev->host_stats.vs_invocations += dev->pending_vs_runs[i];
        continue;
      }
      g.vc_pos = S.vc_pos + pos_off[k];
      g.vc_attr = S.vc_attr + attr_off[k];
      g.vc_flags = S.vc_flags + pos_off[k];
      vc_mark.draw_of[vc_mark.n] = i;
      vc_mark.cta_prefix[vc_mark.n + 1] = vc_mark.cta_prefix[vc_mark.n] + (uint32_t)((n_indices(g) + 256 * VC_MARK_PER_THREAD - 1) / (256 * VC_MARK_PER_THREAD));
      ++vc_mark.n;
      if (groups[k].rep == i) vc_groups.push_back(i);
    }
  }
  // ---- two-kernel geometry: which draws run k_geometry_cull first
  const bool want_split = dev->geometry_split > 0;
  bool any_split = false;
  for (uint32_t i = 0; i < n; ++i) {
    GeomParams& g = dev->pending_geom[i];
    const slv_handle m = dev->pending_vs_module[i];
    const bool split = want_split && (!m || dev->res[m].fn_geometry_cull);
    g.surv = split ? S.surv : nullptr;
    g.surv_count = split ? S.surv_count : nullptr;
    any_split = any_split || split;
  }
  // parameter upload: through this set's pinned staging when pipelining (no stream synchronisation; the staging is free
  // again once ev_front_done has fired, checked above), else straight from pageable memory (the driver stages it)
  const RasterParams* src_batch = dev->pending.data();
  const GeomParams* src_geom = dev->pending_geom.data();
  if (piped) {
    memcpy(S.h_batch, src_batch, n * sizeof(RasterParams));
    memcpy(S.h_geom, src_geom, n * sizeof(GeomParams));
    src_batch = S.h_batch;
    src_geom = S.h_geom;
  }
  CU(cudaMemcpyAsync(S.d_batch, src_batch, n * sizeof(RasterParams), cudaMemcpyHostToDevice, fs));
  CU(cudaMemcpyAsync(S.d_geom, src_geom, n * sizeof(GeomParams), cudaMemcpyHostToDevice, fs));
  BinParams bp{};
  bp.tris = S.tris;
  bp.tri_stride = first.tri_stride;
  bp.n_slots = n_slots;
  bp.tiles_x = first.tiles_x;
  bp.tiles_y = first.tiles_y;
  bp.shard_rank = dev->shard_rank;
  bp.shard_n = dev->shard_n;
  bp.tile_offset = S.tile_offset;
  bp.tile_cursor = S.tile_cursor;
  bp.list = S.list;
  bp.list_capacity = dev->list_cap;
  bp.overflow_flag = dev->overflow_flag;
  bp.valid_slots = S.valid_slots;
  bp.valid_count = S.valid_count;
  // ---- geometry of every queued draw: one launch per distinct register count (normally one)
  size_t eg0 = dev->profile ? mark(dev) : 0;
  {
    std::vector<slv_handle> mods;  // distinct vertex-shader modules of the batch (0 = the built-in programs)
    for (slv_handle m : dev->pending_vs_module)
      if (std::find(mods.begin(), mods.end(), m) == mods.end()) mods.push_back(m);
    // post-transform vertex cache: mark the referenced vertices, then the vertex shader once per marked vertex
    if (vc_mark.n) {
      k_vertex_mark<<<vc_mark.cta_prefix[vc_mark.n], 256, 0, fs>>>(S.d_geom, vc_mark);
      dev->n_launches += 1;
      for (slv_handle m : mods)
        for (uint32_t R = 1; R <= (uint32_t)MAX_REGS; ++R) {
          GeomBatch hb{};
          for (uint32_t i : vc_groups) {
            if (1 + dev->pending_geom[i].n_attrs != R || dev->pending_vs_module[i] != m) continue;
            hb.draw_of[hb.n] = i;
            hb.cta_prefix[hb.n + 1] = hb.cta_prefix[hb.n] + (dev->pending_geom[i].vc_cap + 127) / 128;
            ++hb.n;
          }
          if (!hb.n) continue;
          if (m) {
            const GeomParams* d_geom = S.d_geom;
            void* args[] = {(void*)&d_geom, (void*)&hb};
            if (driver_api().LaunchKernel(dev->res[m].fn_vertex_shade, hb.cta_prefix[hb.n], 1, 1, 128, 1, 1, 0, (CUstream)fs, args, nullptr) != CUDA_SUCCESS)
              return SLV_FAILED;
          } else {
            switch (R) {
            case 1: k_vertex_shade<1><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 2: k_vertex_shade<2><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 3: k_vertex_shade<3><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 4: k_vertex_shade<4><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 5: k_vertex_shade<5><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            default: k_vertex_shade<6><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            }
          }
          dev->n_launches += 1;
        }
    }
    for (slv_handle m : mods)
      for (uint32_t R = 1; R <= (uint32_t)MAX_REGS; ++R) {
        GeomBatch hb{};
        for (uint32_t i = 0; i < n; ++i) {
          if (1 + dev->pending_geom[i].n_attrs != R || dev->pending_vs_module[i] != m) continue;
          hb.draw_of[hb.n] = i;
          hb.cta_prefix[hb.n + 1] = hb.cta_prefix[hb.n] + (dev->pending_geom[i].prim_count + 127) / 128;
          ++hb.n;
        }
        if (!hb.n) continue;
        if (any_split && dev->pending_geom[hb.draw_of[0]].surv) {  // the position pass of every primitive, survivors -> k_geometry
          if (m) {
            const GeomParams* d_geom = S.d_geom;
            void* args[] = {(void*)&d_geom, (void*)&hb};
            if (driver_api().LaunchKernel(dev->res[m].fn_geometry_cull, hb.cta_prefix[hb.n], 1, 1, 128, 1, 1, 0, (CUstream)fs, args, nullptr) != CUDA_SUCCESS)
              return SLV_FAILED;
          } else {
            switch (R) {
            case 1: k_geometry_cull<1><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 2: k_geometry_cull<2><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 3: k_geometry_cull<3><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 4: k_geometry_cull<4><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            case 5: k_geometry_cull<5><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            default: k_geometry_cull<6><<<hb.cta_prefix[hb.n], 128, 0, fs>>>(S.d_geom, hb); break;
            }
          }
          dev->n_launches += 1;
        }
        if (m) {  // SASL vertex shader: the module's own k_geometry instance
          const GeomParams* d_geom = S.d_geom;
          void* args[] = {(void*)&d_geom, (void*)&hb};
          if (driver_api().LaunchKernel(dev->res[m].fn_geometry, hb.cta_prefix[hb.n], 1, 1, 128, 1, 1, 0, (CUstream)fs, args, nullptr) != CUDA_SUCCESS)
            return SLV_FAILED;
        } else {
          switch (R) {
          case 1: launch_geometry<1>(S.d_geom, hb, fs); break;
          case 2: launch_geometry<2>(S.d_geom, hb, fs); break;
          case 3: launch_geometry<3>(S.d_geom, hb, fs); break;
          case 4: launch_geometry<4>(S.d_geom, hb, fs); break;
          case 5: launch_geometry<5>(S.d_geom, hb, fs); break;
          default: launch_geometry<6>(S.d_geom, hb, fs); break;
          }
        }
        dev->n_launches += 1;
      }
  }
  if (dev->profile) dev->spans.push_back({eg0, mark(dev), 0});
  size_t e0 = dev->profile ? mark(dev) : 0;
  if (dev->big_tiles) {  // the tile counts of the triangles k_geometry queued: a warp per triangle
    k_big_tiles<<<dev->sm_count, 128, 0, fs>>>(S.tris, first.tri_stride, first.tiles_x, dev->shard_rank, dev->shard_n, S.tile_count, S.big_slots, S.big_count);
    dev->n_launches += 1;
  }
  k_scan_tiles<<<1, 1024, 0, fs>>>(S.tile_count, S.tile_offset, S.tile_cursor, n_tiles, S.active_tiles, S.work_counter, S.large_tiles, dev->overflow_flag, S.surv_count, S.big_count);
  k_bin_fill<<<std::max(1u, std::min((n_slots + 255) / 256, (uint32_t)dev->sm_count * 8u)), 256, 0, fs>>>(bp);
  size_t e1 = dev->profile ? mark(dev) : 0;
  // k_sort_lists and k_region_bin index the COMPACTED list of non-empty tiles, which holds at most the tiles this rank owns: a
  // sort-first rank launches an N-th of the CTAs (each empty 1024-thread CTA still has to find room on an SM next to the
  // previous frame's back half).  k_sort_lists_large needs a whole SM's shared memory per CTA: a few CTAs, not one per SM,
  // so that only a few SMs have to drain before the front half can go on.  SLV_FRONT_GRIDS=0: the former full-size grids.
  uint32_t owned_tiles = n_tiles, sort_large_grid = (uint32_t)dev->sm_count;
  if (dev->front_grids) {
    if (dev->shard_n > 1) {
      owned_tiles = 0;
      for (uint32_t ty = 0; ty < first.tiles_y; ++ty)
        for (uint32_t tx = 0; tx < first.tiles_x; ++tx) owned_tiles += ((tx + 3 * ty) % dev->shard_n == dev->shard_rank) ? 1u : 0u;
      owned_tiles = std::max(owned_tiles, 1u);
    }
    sort_large_grid = std::min<uint32_t>((uint32_t)dev->sm_count, std::max<uint32_t>(4u, owned_tiles / 32u));
    if (dev->sort_large_grid > 0) sort_large_grid = (uint32_t)dev->sort_large_grid;
  }
  k_sort_lists<<<owned_tiles, SORT_THREADS, 0, fs>>>(S.tile_offset, S.list, dev->list_cap, S.active_tiles, S.work_counter + 3);
  k_sort_lists_large<<<sort_large_grid, 1024, SORT_LARGE_SMEM * sizeof(uint32_t), fs>>>(S.tile_offset, S.list, dev->list_cap, S.large_tiles, S.valid_count);
  size_t e2 = dev->profile ? mark(dev) : 0;
  // ---- phase 5: visibility-first (k_cover + k_shade) when every queued draw qualifies, else the immediate k_raster
  bool deferred = !dev->force_immediate;
  CUfunction jit_shade_fn = nullptr;  // a batch has ONE pixel-shader program (a program change is a flush point)
  if (first.ps_program == SLV_PS_JIT && !dev->jit_immediate) {
    const Resource* m = dev->get(dev->batch_ps_module, Resource::MODULE);
    if (m) jit_shade_fn = m->fn_shade[dev->batch_S == 1 ? 0 : (dev->batch_S == 2 ? 1 : 2)];
  }
  for (const RasterParams& r : dev->pending)
    deferred = deferred && r.early_z && r.bs_program == SLV_BS_REPLACE && !r.has_centroid && r.ps_program != SLV_PS_DISCARD_ALL &&
               (r.ps_program != SLV_PS_JIT || jit_shade_fn) &&  // SASL pixel shaders: the module's quad-granular k_shade
               !r.color1.data && (!r.color0.data || r.color0.bpp == 4);
  bool ok = false;
  size_t e_mid = (size_t)-1, e_rbin = (size_t)-1;
  DeferredBufs db{};
  const bool shade = deferred && first.color0.data != nullptr;
  // ---- lazy clears of the batch's targets: consumed by the visibility-first path when the tile grid covers the whole
  // surface, executed now (on the main stream, ahead of the back half) otherwise
  Resource* rcol = dev->get(dev->batch_color, Resource::TEXTURE);
  Resource* rds = dev->get(dev->batch_ds, Resource::TEXTURE);
  const bool grid_covers = first.tiles_x * SLV_TILE_SIZE >= first.target_w && first.tiles_y * SLV_TILE_SIZE >= first.target_h;
  const bool lazy_c = shade && rcol && rcol->clear_pending && grid_covers && first.color0.w == first.target_w && first.color0.h == first.target_h;
  const bool lazy_d = deferred && rds && rds->clear_pending && first.ds.data && grid_covers && first.ds.w == first.target_w &&
                      first.ds.h == first.target_h;
  { slv_result rcw = wait_readback(dev, rcol); if (rcw != SLV_OK) return rcw; }
  { slv_result rcw = wait_readback(dev, rds); if (rcw != SLV_OK) return rcw; }
  if (!lazy_c) { slv_result rcm = materialize_clear(dev, rcol); if (rcm != SLV_OK) return rcm; }
  if (!lazy_d) { slv_result rcm = materialize_clear(dev, rds); if (rcm != SLV_OK) return rcm; }
  if (deferred) {
    if (shade) {
      const size_t need = (size_t)first.color0.w * first.color0.h * dev->batch_S;
      if (need > dev->vis_cap) {
        CU(cudaStreamSynchronize(st));
        if (dev->vis) CU(cudaFree(dev->vis));
        CU(cudaMalloc(&dev->vis, need * sizeof(uint32_t)));
        dev->vis_cap = need;
      }
    }
    db.region_list = S.region_list;
    db.region_cap = dev->region_cap;
    db.region_mask = S.region_mask;
    db.region_tile_cnt = S.region_tile_cnt;
    db.bits_pool = S.region_mask + 2 * (size_t)dev->list_cap;
    db.bits_cap = (RMASK_STRIDE - 2) * dev->list_cap;
    if (dev->bits_pool_cap >= 0) db.bits_cap = std::min<uint32_t>(db.bits_cap, (uint32_t)dev->bits_pool_cap);
    db.pool_cursor = S.work_counter + 4;
    db.region_offset = S.region_offset;
    db.region_count = S.region_count;
    db.cursor = S.work_counter + 2;
    db.overflow_flag = dev->overflow_flag;
    db.item_flag = S.item_flag;
    db.block_desc = S.block_desc;
    db.vis = shade ? dev->vis : nullptr;
    db.vis_pitch = first.color0.w;
    db.cover_counter = S.work_counter;
    db.shade_counter = S.work_counter + 1;
    db.lazy_depth = lazy_d ? 1u : 0u;
    db.lazy_color = lazy_c ? 1u : 0u;
    if (lazy_d) {
      memcpy(&db.clear_z, &rds->clear_pattern.x, 4);
      db.clear_st = rds->clear_pattern.y;
    }
    if (lazy_c) db.clear_color = rcol->clear_pattern.x;
    if (dev->resolve_requested && shade && grid_covers && first.color0.w == first.target_w && first.color0.h == first.target_h) {
      db.resolve_dst = dev->resolve_dst;
      dev->resolve_done = true;
    }
    k_region_decide<<<dev->sm_count * 8, 256, 0, fs>>>(first, db);
    k_region_bin<<<owned_tiles, RBIN_THREADS, 0, fs>>>(first, db);
    dev->n_launches += 2;
    if (dev->profile) e_rbin = mark(dev);
  }
  // ---- back half, on the main stream
  if (piped) {
    CU(cudaEventRecord(S.ev_front_done, fs));
    CU(cudaStreamWaitEvent(st, S.ev_front_done, 0));
  }
  uint32_t cover_grid = (uint32_t)dev->cover_grid, shade_grid = (uint32_t)dev->shade_grid;
  {
    int k = dev->back_ctas > 0 ? dev->back_ctas : 8;  // r02 sweeps (tools/variant_sweep.py, tools/shard_sweep.py): 8 wins on a whole frame and on an eighth
    if (!(dev->pipeline && !dev->profile)) k = SLV_COVER_CTAS_PER_SM;  // nothing to overlap with on a single stream
    if (k >= 1 && k <= SLV_COVER_CTAS_PER_SM) cover_grid = (uint32_t)(dev->sm_count * k);
    if (k >= 1 && k <= SLV_SHADE_CTAS_PER_SM) shade_grid = (uint32_t)(dev->sm_count * k);
  }
  const bool persistent = dev->persistent >= 0 ? dev->persistent != 0 : true;
  if (deferred && !persistent) {
    // short-lived CTAs: every warp performs a fixed number of queue fetches, the grid covers the worst case (every owned
    // tile active).  Items per warp: 8 (k_cover, 4 fetches of FETCH) / 2 groups (k_shade).
    uint32_t owned = 0;
    for (uint32_t ty = 0; ty < first.tiles_y; ++ty)
      for (uint32_t tx = 0; tx < first.tiles_x; ++tx) owned += (dev->shard_n <= 1 || (tx + 3 * ty) % dev->shard_n == dev->shard_rank) ? 1u : 0u;
    const uint32_t items = owned * ITEMS_PER_TILE;
    db.cover_budget = 4;
    cover_grid = std::max(1u, (items + DEF_WARPS * db.cover_budget * FETCH - 1) / (DEF_WARPS * db.cover_budget * FETCH));
    db.shade_grp = std::min<uint32_t>(SHADE_GROUP, std::max(1u, items / ((uint32_t)dev->shade_grid * DEF_WARPS * 4u)));
    db.shade_budget = 2;
    shade_grid = std::max(1u, (items + DEF_WARPS * db.shade_budget * db.shade_grp - 1) / (DEF_WARPS * db.shade_budget * db.shade_grp));
  }
  if (deferred) {
    switch (dev->batch_S) {
    case 1: k_cover<1><<<cover_grid, DEF_THREADS, 0, st>>>(first, S.d_batch, n, db); ok = true; break;
    case 2: k_cover<2><<<cover_grid, DEF_THREADS, 0, st>>>(first, S.d_batch, n, db); ok = true; break;
    case 4: k_cover<4><<<cover_grid, DEF_THREADS, 0, st>>>(first, S.d_batch, n, db); ok = true; break;
    }
    if (ok && shade) {
      if (dev->profile) e_mid = mark(dev);
      if (first.ps_program == SLV_PS_JIT) {  // SASL pixel shader: the module's own quad-granular k_shade instance
        const RasterParams* d_batch = S.d_batch;
        uint32_t nd = n;
        void* args[] = {(void*)&first, (void*)&d_batch, (void*)&nd, (void*)&db};
        ok = driver_api().LaunchKernel(jit_shade_fn, shade_grid, 1, 1, DEF_THREADS, 1, 1, 0, (CUstream)st, args, nullptr) == CUDA_SUCCESS;
      } else
      switch (dev->batch_S) {
      case 1: ok = launch_shade_s<1>(first, S.d_batch, n, db, shade_grid, st); break;
      case 2: ok = launch_shade_s<2>(first, S.d_batch, n, db, shade_grid, st); break;
      case 4: ok = launch_shade_s<4>(first, S.d_batch, n, db, shade_grid, st); break;
      }
      dev->n_launches += 1;
    }
    if (ok && (lazy_c || lazy_d || db.resolve_dst.data)) {
      // the tiles without triangles (never visited by k_cover / k_shade): clear values and, with a fused resolve, their resolve
      SurfaceRef none{};
      k_inactive_tiles<<<n_tiles, 256, 0, st>>>(first.color0, lazy_c ? 1u : 0u, lazy_c ? rcol->clear_pattern : make_uint4(0, 0, 0, 0),
                                                 lazy_d ? first.ds : none, lazy_d ? rds->clear_pattern : make_uint4(0, 0, 0, 0),
                                                 db.resolve_dst, S.tile_offset, first.tiles_x, dev->shard_rank, dev->shard_n);
      dev->n_launches += 1;
      if (lazy_c) rcol->clear_pending = false;
      if (lazy_d) rds->clear_pending = false;
    }
  } else {
    const uint32_t blocks = std::min<uint32_t>(n_tiles * 16, (uint32_t)dev->raster_grid);
    if (first.ps_program == SLV_PS_JIT) {  // SASL pixel shader: the module's own k_raster instance
      const Resource& m = dev->res[dev->batch_ps_module];
      CUfunction fn = m.fn_raster[dev->batch_S == 1 ? 0 : (dev->batch_S == 2 ? 1 : 2)];
      const RasterParams* d_batch = S.d_batch;
      uint32_t nd = n;
      void* args[] = {(void*)&first, (void*)&d_batch, (void*)&nd};
      ok = driver_api().LaunchKernel(fn, blocks, 1, 1, RASTER_THREADS, 1, 1, 0, (CUstream)st, args, nullptr) == CUDA_SUCCESS;
    } else
    switch (dev->batch_S) {
    case 1: ok = launch_raster_s<1>(first, S.d_batch, n, blocks, st); break;
    case 2: ok = launch_raster_s<2>(first, S.d_batch, n, blocks, st); break;
    case 4: ok = launch_raster_s<4>(first, S.d_batch, n, blocks, st); break;
    }
  }
  if (piped) {
    CU(cudaEventRecord(S.ev_back_done, st));
    S.in_flight = true;
    dev->last_flushed = dev->cur;
    dev->cur = (dev->cur + 1) % slv_device_t::N_SETS;  // the next batch is built in the next set
  }
  dev->n_launches += 5;
  if (dev->profile) {
    size_t e3 = mark(dev);
    dev->spans.push_back({e0, e1, 1});
    dev->spans.push_back({e1, e2, 2});
    if (e_rbin != (size_t)-1 && e_mid != (size_t)-1) {
      dev->spans.push_back({e2, e_rbin, 5});
      dev->spans.push_back({e_rbin, e_mid, 3});
      dev->spans.push_back({e_mid, e3, 4});
    } else if (e_rbin != (size_t)-1) {
      dev->spans.push_back({e2, e_rbin, 5});
      dev->spans.push_back({e_rbin, e3, 3});
    } else {
      dev->spans.push_back({e2, e3, 3});
    }
  }
  dev->pending.clear();
  dev->pending_geom.clear();
  dev->pending_vs_runs.clear();
  dev->pending_vs_module.clear();
  dev->tris_used = 0;
  dev->slots_queued = 0;
  if (!ok) return SLV_INVALID_PARAMETER;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result check_overflow(slv_device dev) {
  uint32_t flag = 0;
  CU(cudaMemcpyAsync(&flag, dev->overflow_flag, sizeof(flag), cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  if (flag == 1 && dev->pipeline)
    for (auto fsx : dev->front_streams) CU(cudaStreamSynchronize(fsx));  // the recorded needs are final
  if (flag == 2) {
    fprintf(stderr, "[salvia_b200] slv_flags_wait timed out: a peer rank never raised its flag\n");
    dev->failed = true;
    return SLV_FAILED;
  }
  if (flag) {
    // not sticky: the frame that overflowed is incomplete (reported once), the arenas grow before the next batch
    uint32_t need[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(need, dev->overflow_flag, sizeof(need), cudaMemcpyDeviceToHost, dev->stream));
    CU(cudaMemsetAsync(dev->overflow_flag, 0, 3 * sizeof(uint32_t), dev->stream));
    CU(cudaStreamSynchronize(dev->stream));
    dev->list_need_hint = std::max<uint64_t>(dev->list_need_hint, (uint64_t)need[1] + need[1] / 2);
    dev->region_need_hint = std::max<uint64_t>(dev->region_need_hint, (uint64_t)need[2] + need[2] / 2);
    fprintf(stderr, "[salvia_b200] a batch overflowed its work-list arenas (%u tile-list entries of %u, %u region-list words of %u): that "
            "frame is incomplete (SLV_OUT_OF_MEMORY); the arenas grow before the next batch\n", need[1], dev->list_cap, need[2], dev->region_cap);
    return SLV_OUT_OF_MEMORY;
  }
  return dev->failed ? SLV_FAILED : SLV_OK;
}

}  // namespace

extern "C" {

const char* slv_backend_name(void) { return "cuda-sm100a"; }
uint32_t slv_abi_version(void) { return SLV_ABI_VERSION; }

slv_result slv_device_create(int32_t ordinal, slv_device* out) {
  if (!out) return SLV_INVALID_PARAMETER;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fprintf(stderr, "[salvia_b200] no CUDA device available (%s); this library has no CPU fallback\n",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return SLV_FAILED;
  }
  if (ordinal < 0 || ordinal >= n) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(ordinal));
  auto dev = new slv_device_t;
  dev->ordinal = ordinal;
  dev->res.resize(1);
  CU(cudaStreamCreateWithFlags(&dev->own_stream, cudaStreamNonBlocking));
  dev->stream = dev->own_stream;
  {  // the front half of the NEXT frame is on the critical path: its CTAs go first whenever SM resources free up
    int least = 0, greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    for (auto& fsx : dev->front_streams) CU(cudaStreamCreateWithPriority(&fsx, cudaStreamNonBlocking, greatest));
  }
  CU(cudaEventCreateWithFlags(&dev->ev_sync, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&dev->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&dev->ev_copy, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&dev->ev_upload, cudaEventDisableTiming));
  CU(cudaMalloc(&dev->overflow_flag, 4 * sizeof(uint32_t)));
  CU(cudaMalloc(&dev->peer_flags, SLV_PEER_FLAGS * sizeof(uint32_t)));
  CU(cudaMemset(dev->peer_flags, 0, SLV_PEER_FLAGS * sizeof(uint32_t)));
  for (auto& S : dev->sc) {
    CU(cudaMalloc(&S.work_counter, 8 * sizeof(uint32_t)));
    CU(cudaMemset(S.work_counter, 0, 8 * sizeof(uint32_t)));
    CU(cudaMalloc(&S.valid_count, sizeof(uint32_t)));
    CU(cudaMalloc(&S.surv_count, MAX_BATCH * sizeof(uint32_t)));
    CU(cudaMalloc(&S.big_count, sizeof(uint32_t)));
    CU(cudaMemset(S.big_count, 0, sizeof(uint32_t)));
    CU(cudaMemset(S.surv_count, 0, MAX_BATCH * sizeof(uint32_t)));  // k_scan_tiles re-zeroes it behind every geometry pass
    CU(cudaMemset(S.valid_count, 0, sizeof(uint32_t)));  // k_sort_lists_large re-zeroes it at the end of every binning chain
    CU(cudaMalloc(&S.d_batch, MAX_BATCH * sizeof(RasterParams)));
    CU(cudaMalloc(&S.d_geom, MAX_BATCH * sizeof(GeomParams)));
    CU(cudaHostAlloc(&S.h_batch, MAX_BATCH * sizeof(RasterParams), cudaHostAllocDefault));
    CU(cudaHostAlloc(&S.h_geom, MAX_BATCH * sizeof(GeomParams), cudaHostAllocDefault));
    CU(cudaEventCreateWithFlags(&S.ev_front_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&S.ev_back_done, cudaEventDisableTiming));
  }
  {
    const char* pl = getenv("SLV_PIPELINE");
    dev->pipeline = !(pl && pl[0] == '0');
    const char* lc = getenv("SLV_LAZY_CLEAR");
    dev->lazy_clear = !(lc && lc[0] == '0');
    const char* pe = getenv("SLV_PERSISTENT");
    dev->persistent = pe ? (pe[0] == '1' ? 1 : 0) : -1;
    const char* fr = getenv("SLV_FUSE_RESOLVE");
    dev->fuse_resolve = !(fr && fr[0] == '0');
  }
  {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ordinal));
    dev->raster_grid = prop.multiProcessorCount * RASTER_CTAS_PER_SM;
    dev->sm_count = prop.multiProcessorCount;
    CU(cudaFuncSetAttribute(k_sort_lists_large, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_LARGE_SMEM * (int)sizeof(uint32_t)));
    dev->cover_grid = prop.multiProcessorCount * SLV_COVER_CTAS_PER_SM;
    dev->shade_grid = prop.multiProcessorCount * SLV_SHADE_CTAS_PER_SM;
    if (const char* bc = getenv("SLV_BACK_CTAS")) dev->back_ctas = atoi(bc);
  }
  CU(cudaMemsetAsync(dev->overflow_flag, 0, 4 * sizeof(uint32_t), dev->stream));
  if (const char* am = getenv("SLV_ARENA_MIN")) { dev->arena_min_list = std::max<uint64_t>(1024, (uint64_t)atoll(am)); dev->arena_min_region = dev->arena_min_list * 4; }
  CU(cudaMalloc(&dev->d_stats, 20 * sizeof(unsigned long long)));
  CU(cudaMemsetAsync(dev->d_stats, 0, 20 * sizeof(unsigned long long), dev->stream));
  const char* prof = getenv("SLV_PROFILE");
  dev->profile = prof && prof[0] == '1';
  const char* fi = getenv("SLV_FORCE_IMMEDIATE");
  dev->force_immediate = fi && fi[0] == '1';
  const char* fg = getenv("SLV_FRONT_GRIDS");
  dev->front_grids = !(fg && fg[0] == '0');
  if (const char* sg = getenv("SLV_SORT_LARGE_GRID")) dev->sort_large_grid = atoi(sg);
  if (const char* bp = getenv("SLV_BITS_POOL_CAP")) dev->bits_pool_cap = atoll(bp);
  const char* ji = getenv("SLV_JIT_IMMEDIATE");
  dev->jit_immediate = ji && ji[0] == '1';
  if (const char* vc = getenv("SLV_VERTEX_CACHE")) dev->vertex_cache = atoi(vc);
  if (const char* gs = getenv("SLV_GEOMETRY_SPLIT")) dev->geometry_split = atoi(gs);
  if (const char* bt = getenv("SLV_BIG_TILES")) dev->big_tiles = atoi(bt) != 0;
  for (auto& ev : dev->user_ev) CU(cudaEventCreate(&ev));

  *out = dev;
  return SLV_OK;
}

void slv_device_destroy(slv_device dev) {
  if (!dev) return;
  cudaSetDevice(dev->ordinal);
  flush_batch(dev);
  for (auto fsx : dev->front_streams) cudaStreamSynchronize(fsx);
  cudaStreamSynchronize(dev->stream);
  cudaStreamSynchronize(dev->copy_stream);
  for (auto& r : dev->res) {
    if (r.kind == Resource::BUFFER) cudaFree(r.dptr);
    if (r.kind == Resource::TEXTURE)
      for (uint32_t l = 0; l < r.tex.n_levels; ++l) cudaFree(r.tex.level[l].data);
    if (r.kind == Resource::MODULE && r.module) driver_api().ModuleUnload(r.module);
    if (r.rb_event) cudaEventDestroy(r.rb_event);
  }
  for (auto& S : dev->sc) {
    cudaFree(S.tris); cudaFree(S.valid_slots); cudaFree(S.valid_count); cudaFree(S.surv); cudaFree(S.surv_count); cudaFree(S.big_slots); cudaFree(S.big_count);
    cudaFree(S.tile_count); cudaFree(S.tile_offset); cudaFree(S.tile_cursor); cudaFree(S.active_tiles); cudaFree(S.large_tiles);
    cudaFree(S.work_counter); cudaFree(S.region_list); cudaFree(S.region_mask); cudaFree(S.region_tile_cnt); cudaFree(S.region_offset); cudaFree(S.region_count);
    cudaFree(S.item_flag); cudaFree(S.block_desc); cudaFree(S.list); cudaFree(S.d_batch); cudaFree(S.d_geom);
    cudaFreeHost(S.h_batch); cudaFreeHost(S.h_geom);
    cudaFree(S.vc_pos); cudaFree(S.vc_attr); cudaFree(S.vc_flags);
    cudaEventDestroy(S.ev_front_done); cudaEventDestroy(S.ev_back_done);
  }
  cudaFree(dev->vis);
  cudaFree(dev->overflow_flag);
  cudaFree(dev->peer_flags);
  cudaFree(dev->d_stats);
  cudaFree(dev->d_level_touched);
  for (auto& ev : dev->ev_pool) cudaEventDestroy(ev);
  for (auto& ev : dev->user_ev) cudaEventDestroy(ev);
  cudaEventDestroy(dev->ev_sync);
  cudaEventDestroy(dev->ev_copy);
  cudaEventDestroy(dev->ev_upload);
  cudaStreamDestroy(dev->copy_stream);
  if (dev->signal_stream) { cudaStreamSynchronize(dev->signal_stream); cudaStreamDestroy(dev->signal_stream); }
  for (auto fsx : dev->front_streams) cudaStreamDestroy(fsx);

  for (auto& t : dev->slot_tables) cudaFree(t.d_slot);
  cudaStreamDestroy(dev->own_stream);
  delete dev;
}

slv_result slv_buffer_create(slv_device dev, size_t bytes, slv_handle* out) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  Resource r;
  r.kind = Resource::BUFFER;
  r.bytes = bytes;
  if (cudaMalloc(&r.dptr, std::max<size_t>(bytes, 16)) != cudaSuccess) return SLV_OUT_OF_MEMORY;
  dev->res.push_back(r);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

slv_result slv_buffer_upload(slv_device dev, slv_handle h, size_t off, const void* src, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::BUFFER) : nullptr;
  if (!r || off + bytes > r->bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  if (dev->pipeline && !dev->profile) {
    // vertex / index data is only read by the front half: upload on its stream, so the copy (and the next frame's geometry
    // after it) does not queue behind the previous frame's raster work on the main stream
    // ... the stream of the NEXT batch's front half (its reader); the previous batch's front half, on the other front stream,
    // may still be reading the old contents
    cudaStream_t fs = dev->front_stream_of(dev->cur);
    if (dev->last_flushed >= 0 && dev->sc[dev->last_flushed].in_flight) CU(cudaStreamWaitEvent(fs, dev->sc[dev->last_flushed].ev_front_done, 0));
    CU(cudaMemcpyAsync(r->dptr + off, src, bytes, cudaMemcpyHostToDevice, fs));
    CU(cudaEventRecord(dev->ev_upload, fs));
    dev->upload_on_front = true;
    return SLV_OK;
  }
  CU(cudaMemcpyAsync(r->dptr + off, src, bytes, cudaMemcpyHostToDevice, dev->stream));
  dev->buffers_dirty = true;  // the next front halves (front streams) must order after this copy
  return SLV_OK;
}

slv_result slv_buffer_device_ptr(slv_device dev, slv_handle h, void** out, size_t* bytes) {
  auto r = dev ? dev->get(h, Resource::BUFFER) : nullptr;
  if (!r || !out) return SLV_INVALID_PARAMETER;
  *out = r->dptr;
  if (bytes) *bytes = r->bytes;
  return SLV_OK;
}

slv_result slv_external_write_begin(slv_device dev, void* cuda_stream, uint32_t skip_latest) {
  if (!dev || !cuda_stream) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  cudaStream_t ext = (cudaStream_t)cuda_stream;
  if (dev->pipeline && !dev->profile) {
    // buffers are read by the front halves only: the last flushed one (earlier ones precede it on the same two streams)
    for (int k = (int)skip_latest; k < (int)skip_latest + 2 && k < slv_device_t::N_SETS - 1 && dev->last_flushed >= 0; ++k) {
      const slv_device_t::Scratch& T = dev->sc[(dev->last_flushed - k + 2 * slv_device_t::N_SETS) % slv_device_t::N_SETS];
      if (T.in_flight) CU(cudaStreamWaitEvent(ext, T.ev_front_done, 0));
    }
  } else {
    CU(cudaEventRecord(dev->ev_sync, dev->stream));
    CU(cudaStreamWaitEvent(ext, dev->ev_sync, 0));
  }
  return SLV_OK;
}

slv_result slv_external_write_end(slv_device dev, void* cuda_stream) {
  if (!dev || !cuda_stream) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaEventRecord(dev->ev_sync, (cudaStream_t)cuda_stream));
  for (auto fsx : dev->front_streams) CU(cudaStreamWaitEvent(fsx, dev->ev_sync, 0));
  CU(cudaStreamWaitEvent(dev->stream, dev->ev_sync, 0));
  return SLV_OK;
}

slv_result slv_buffer_readback(slv_device dev, slv_handle h, size_t off, void* dst, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::BUFFER) : nullptr;
  if (!r || off + bytes > r->bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  if (dev->upload_on_front) {
    CU(cudaStreamWaitEvent(dev->stream, dev->ev_upload, 0));
    dev->upload_on_front = false;
  }
  CU(cudaMemcpyAsync(dst, r->dptr + off, bytes, cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  return SLV_OK;
}

static slv_result alloc_level(SurfaceRef& s, uint32_t w, uint32_t h, uint32_t samples, uint32_t fmt) {
  s.w = w; s.h = h; s.samples = samples; s.fmt = fmt; s.bpp = bpp_of(fmt);
  s.wmask = (w <= 1024 && (w & (w - 1)) == 0) ? w - 1 : 0;
  s.hmask = (h <= 1024 && (h & (h - 1)) == 0) ? h - 1 : 0;
  s.bytes = (size_t)w * h * samples * s.bpp;
  if (cudaMalloc(&s.data, std::max<size_t>(s.bytes, 16)) != cudaSuccess) return SLV_OUT_OF_MEMORY;
  return SLV_OK;
}

slv_result slv_texture_create(slv_device dev, uint32_t w, uint32_t h, uint32_t samples, uint32_t fmt, slv_handle* out) {
  if (!dev || !out || !bpp_of(fmt) || !w || !h || !samples) return SLV_INVALID_PARAMETER;
  if (w > SLV_MAX_RENDER_TARGET_SIZE || h > SLV_MAX_RENDER_TARGET_SIZE) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  Resource r;
  r.kind = Resource::TEXTURE;
  r.fmt = fmt;
  r.samples = samples;
  r.tex.n_levels = 1;
  slv_result rc = alloc_level(r.tex.level[0], w, h, samples, fmt);
  if (rc != SLV_OK) return rc;
  CU(cudaMemsetAsync(r.tex.level[0].data, 0, r.tex.level[0].bytes, dev->stream));
  dev->res.push_back(r);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

slv_result slv_texture_gen_mipmap(slv_device dev, slv_handle h, uint32_t filter) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || (filter != SLV_FILTER_POINT && filter != SLV_FILTER_LINEAR)) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  CU(cudaStreamSynchronize(dev->stream));
  for (uint32_t l = 1; l < r->tex.n_levels; ++l) CU(cudaFree(r->tex.level[l].data));
  r->tex.n_levels = 1;
  uint32_t m = std::max(r->tex.level[0].w, r->tex.level[0].h);
  uint32_t limit = 0;  // texture::calc_lod_limit (texture.h:26-35)
  while (m > 0) { m >>= 1; ++limit; }
  for (uint32_t l = 0; l + 1 < limit && l + 1 < (uint32_t)MAX_LEVELS; ++l) {
    const SurfaceRef src = r->tex.level[l];
    SurfaceRef& dst = r->tex.level[l + 1];
    slv_result rc = alloc_level(dst, (src.w + 1) / 2, (src.h + 1) / 2, src.samples, src.fmt);
    if (rc != SLV_OK) return rc;
    dim3 blk(16, 16), grd((dst.w + 15) / 16, (dst.h + 15) / 16);
    k_mipgen<<<grd, blk, 0, dev->stream>>>(src, dst, filter);
    ++dev->n_launches;
    r->tex.n_levels = l + 2;
  }
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_texture_level_count(slv_device dev, slv_handle h, uint32_t* out) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || !out) return SLV_INVALID_PARAMETER;
  *out = r->tex.n_levels;
  return SLV_OK;
}

slv_result slv_texture_level_size(slv_device dev, slv_handle h, uint32_t level, uint32_t* w, uint32_t* hh) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels) return SLV_INVALID_PARAMETER;
  *w = r->tex.level[level].w;
  *hh = r->tex.level[level].h;
  return SLV_OK;
}

slv_result slv_texture_upload(slv_device dev, slv_handle h, uint32_t level, const void* src, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels || bytes != r->tex.level[level].bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  if (level == 0) r->clear_pending = false;  // the whole level is overwritten
  { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
  CU(cudaMemcpyAsync(r->tex.level[level].data, src, bytes, cudaMemcpyHostToDevice, dev->stream));
  return SLV_OK;
}

slv_result slv_texture_readback(slv_device dev, slv_handle h, uint32_t level, void* dst, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels || bytes != r->tex.level[level].bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  { slv_result rca__ = wait_assembly(dev, r); if (rca__ != SLV_OK) return rca__; }
  CU(cudaMemcpyAsync(dst, r->tex.level[level].data, bytes, cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  return check_overflow(dev);
}

slv_result slv_texture_readback_async(slv_device dev, slv_handle h, uint32_t level, void* dst, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels || bytes != r->tex.level[level].bytes || !dst) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  if (!r->rb_event) CU(cudaEventCreateWithFlags(&r->rb_event, cudaEventDisableTiming));
  CU(cudaEventRecord(dev->ev_copy, dev->stream));            // the texture's producers are ahead of this point
  CU(cudaStreamWaitEvent(dev->copy_stream, dev->ev_copy, 0));
  CU(cudaMemcpyAsync(dst, r->tex.level[level].data, bytes, cudaMemcpyDeviceToHost, dev->copy_stream));
  CU(cudaEventRecord(r->rb_event, dev->copy_stream));
  r->rb_pending = true;                                      // the next writer of the texture waits for the copy
  return SLV_OK;
}

slv_result slv_host_register(slv_device dev, void* ptr, size_t bytes) {
  if (!dev || !ptr || !bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  return SLV_OK;
}

slv_result slv_host_unregister(slv_device dev, void* ptr) {
  if (!dev || !ptr) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaStreamSynchronize(dev->copy_stream));
  CU(cudaHostUnregister(ptr));
  return SLV_OK;
}

slv_result slv_texture_export_tiles_async(slv_device dev, slv_handle h, void* host_frame, size_t bytes) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || !host_frame || r->tex.level[0].samples != 1 || bytes != r->tex.level[0].bytes) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  void* dptr = nullptr;
  CU(cudaHostGetDevicePointer(&dptr, host_frame, 0));  // fails unless the range was registered (slv_host_register)
  const SurfaceRef& s = r->tex.level[0];
  const uint32_t tiles_x = (s.w + TILE - 1) / TILE, tiles_y = (s.h + TILE - 1) / TILE;
  if (!r->rb_event) CU(cudaEventCreateWithFlags(&r->rb_event, cudaEventDisableTiming));
  CU(cudaEventRecord(dev->ev_copy, dev->stream));            // the texture's producers are ahead of this point
  CU(cudaStreamWaitEvent(dev->copy_stream, dev->ev_copy, 0));
  {
    static const int export_ctas = getenv("SLV_EXPORT_CTAS") ? atoi(getenv("SLV_EXPORT_CTAS")) : 32;
    k_export_tiles<<<std::max(1, std::min<int>(export_ctas, (int)(tiles_x * tiles_y))), 256, 0, dev->copy_stream>>>(s, tiles_x, tiles_x * tiles_y, dev->shard_rank, dev->shard_n, (uint8_t*)dptr);
  }
  ++dev->n_launches;
  CU(cudaGetLastError());
  CU(cudaEventRecord(r->rb_event, dev->copy_stream));
  r->rb_pending = true;                                      // the next writer of the texture waits for the export
  return SLV_OK;
}

slv_result slv_readback_fence(slv_device dev, slv_handle h) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  return wait_readback(dev, r);
}

slv_result slv_readback_wait(slv_device dev) {
  if (!dev) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaStreamSynchronize(dev->copy_stream));
  return check_overflow(dev);
}

slv_result slv_sampler_create(slv_device dev, const slv_sampler_desc* d, slv_handle tex, slv_handle* out) {
  if (!dev || !d || !out || !dev->get(tex, Resource::TEXTURE)) return SLV_INVALID_PARAMETER;
  if (d->min_filter > SLV_FILTER_LINEAR || d->mag_filter > SLV_FILTER_LINEAR || d->mip_filter > SLV_FILTER_ANISOTROPIC ||
      d->addr_mode_u > SLV_ADDR_BORDER || d->addr_mode_v > SLV_ADDR_BORDER || d->mip_qual > SLV_MIP_HI_QUALITY)
    return SLV_INVALID_PARAMETER;
  Resource r;
  r.kind = Resource::SAMPLER;
  r.sd = *d;
  r.sampler_tex = tex;
  dev->res.push_back(r);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

// compile(code, profile) of the reference (salvia/include/salvia/core/renderer.h:136-147) ends in a host function pointer
// from LLVM's JIT; here it ends in a cubin (salviarenderer_b200/sasl/jit.py) that is loaded into the device's context.
slv_result slv_shader_module_load(slv_device dev, uint32_t stage, const void* image, size_t bytes, uint32_t n_vs_output_attrs,
                                  slv_handle* out) {
  if (!dev || !image || !bytes || !out || (stage != SLV_STAGE_VS && stage != SLV_STAGE_PS)) return SLV_INVALID_PARAMETER;
  if (stage == SLV_STAGE_VS && n_vs_output_attrs > SLV_MAX_VS_OUTPUT_ATTRS) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaFree(nullptr));  // the primary context exists and is current
  DriverApi& api = driver_api();
  if (!api.ok) {
    fprintf(stderr, "[salvia_b200] libcuda.so.1 is not available: cannot load run-time shader modules\n");
    return SLV_FAILED;
  }
  Resource r;
  r.kind = Resource::MODULE;
  r.module_stage = stage;
  r.module_attrs = n_vs_output_attrs;
  if (api.ModuleLoadData(&r.module, image) != CUDA_SUCCESS) return SLV_FAILED;
  bool ok = true;
  if (stage == SLV_STAGE_VS) {
    ok = api.ModuleGetFunction(&r.fn_geometry, r.module, "slv_jit_k_geometry") == CUDA_SUCCESS;
    if (api.ModuleGetFunction(&r.fn_vertex_shade, r.module, "slv_jit_k_vertex_shade") != CUDA_SUCCESS) r.fn_vertex_shade = nullptr;
    if (api.ModuleGetFunction(&r.fn_geometry_cull, r.module, "slv_jit_k_geometry_cull") != CUDA_SUCCESS) r.fn_geometry_cull = nullptr;
  } else {
    const char* names[3] = {"slv_jit_k_raster_s1", "slv_jit_k_raster_s2", "slv_jit_k_raster_s4"};
    for (int i = 0; i < 3; ++i) ok = ok && api.ModuleGetFunction(&r.fn_raster[i], r.module, names[i]) == CUDA_SUCCESS;
    const char* shade_names[3] = {"slv_jit_k_shade_s1", "slv_jit_k_shade_s2", "slv_jit_k_shade_s4"};
    for (int i = 0; i < 3; ++i)
      if (api.ModuleGetFunction(&r.fn_shade[i], r.module, shade_names[i]) != CUDA_SUCCESS) r.fn_shade[i] = nullptr;
  }
  if (!ok) {
    api.ModuleUnload(r.module);
    return SLV_INVALID_PARAMETER;
  }
  dev->res.push_back(r);
  *out = (slv_handle)(dev->res.size() - 1);
  return SLV_OK;
}

// ---- run-time shader compilation in process (NVRTC) -------------------------------------------------------------------------
namespace {
#include "slv_embedded_sources.inc"

// the NVRTC entry points, bound on first use (no link-time dependency: a machine without the toolkit still loads the library)
struct NvrtcApi {
  typedef void* Program;
  int (*CreateProgram)(Program*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(Program, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(Program, size_t*) = nullptr;
  int (*GetProgramLog)(Program, char*) = nullptr;
  int (*GetCUBINSize)(Program, size_t*) = nullptr;
  int (*GetCUBIN)(Program, char*) = nullptr;
  int (*DestroyProgram)(Program*) = nullptr;
  bool ok = false;
};
NvrtcApi& nvrtc_api() {
  static NvrtcApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = getenv("SLV_NVRTC_LIB");
    const char* names[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
    void* h = nullptr;
    for (const char* n : names)
      if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) {
      api.CreateProgram = (decltype(api.CreateProgram))dlsym(h, "nvrtcCreateProgram");
      api.CompileProgram = (decltype(api.CompileProgram))dlsym(h, "nvrtcCompileProgram");
      api.GetProgramLogSize = (decltype(api.GetProgramLogSize))dlsym(h, "nvrtcGetProgramLogSize");
      api.GetProgramLog = (decltype(api.GetProgramLog))dlsym(h, "nvrtcGetProgramLog");
      api.GetCUBINSize = (decltype(api.GetCUBINSize))dlsym(h, "nvrtcGetCUBINSize");
      api.GetCUBIN = (decltype(api.GetCUBIN))dlsym(h, "nvrtcGetCUBIN");
      api.DestroyProgram = (decltype(api.DestroyProgram))dlsym(h, "nvrtcDestroyProgram");
      api.ok = api.CreateProgram && api.CompileProgram && api.GetProgramLogSize && api.GetProgramLog && api.GetCUBINSize && api.GetCUBIN &&
               api.DestroyProgram;
    }
  }
  return api;
}

// the few host headers the embedded sources name: NVRTC has no system include path, its built-ins cover the rest
const char kShimStdint[] =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n"
    "typedef unsigned long uintptr_t; typedef long intptr_t;\n";
const char kShimEmpty[] = "#pragma once\n";

uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

// cache directory of compiled images: used only when this user owns it and nobody else can write to it (an image is code that
// runs in the process's GPU context)
std::string jit_cache_dir() {
  std::string d;
  if (const char* e = getenv("SLV_JIT_CACHE")) d = e;
  else if (const char* x = getenv("XDG_CACHE_HOME")) d = std::string(x) + "/salvia_b200_jit";
  else if (const char* hme = getenv("HOME")) { std::string c = std::string(hme) + "/.cache"; mkdir(c.c_str(), 0700); d = c + "/salvia_b200_jit"; }
  if (d.empty()) return d;
  mkdir(d.c_str(), 0700);
  struct stat st;
  if (stat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid() || (st.st_mode & 022)) return std::string();
  return d;
}
bool trusted_file(const std::string& path, struct stat* st) {
  return stat(path.c_str(), st) == 0 && S_ISREG(st->st_mode) && st->st_uid == getuid() && !(st->st_mode & 022);
}
void put_log(char* log, size_t log_bytes, const std::string& text) {
  if (!log || !log_bytes) return;
  const size_t n = std::min(text.size(), log_bytes - 1);
  memcpy(log, text.data(), n);
  log[n] = 0;
}
}  // namespace

slv_result slv_shader_compile_cubin(uint32_t stage, const char* device_code, uint32_t n_vs_output_attrs, uint32_t flags, void** image,
                                    size_t* bytes, char* log, size_t log_bytes) {
  if (log && log_bytes) log[0] = 0;
  if (!device_code || !image || !bytes || (stage != SLV_STAGE_VS && stage != SLV_STAGE_PS)) return SLV_INVALID_PARAMETER;
  if (stage == SLV_STAGE_VS && n_vs_output_attrs > SLV_MAX_VS_OUTPUT_ATTRS) return SLV_INVALID_PARAMETER;
  *image = nullptr;
  *bytes = 0;
  // ---- options: the library's own numerics flags (Makefile), the stage, the register count
  std::vector<std::string> opts = {"-arch=sm_100a", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-lineinfo",
                                   "-DSLV_JIT_GENERATED=\"slv_generated_shader.cuh\""};
  if (stage == SLV_STAGE_VS) {
    opts.push_back("-DSLV_JIT_VS=1");
    opts.push_back("-DSLV_JIT_R=" + std::to_string(n_vs_output_attrs + 1));
  } else {
    opts.push_back("-DSLV_JIT_PS=1");
    if (flags & SLV_COMPILE_DERIV_CPP) opts.push_back("-DSLV_JIT_DERIV_CPP=1");
  }
  // ---- disk cache, keyed by the generated code, the options and the embedded sources
  uint64_t h1 = 14695981039346656037ull, h2 = 0x9E3779B97F4A7C15ull;
  h1 = fnv1a(device_code, strlen(device_code), h1);
  h2 = fnv1a(device_code, strlen(device_code), h2);
  for (const std::string& o : opts) { h1 = fnv1a(o.data(), o.size() + 1, h1); h2 = fnv1a(o.data(), o.size() + 1, h2); }
  for (const EmbeddedSource& e : kEmbeddedSources) { h1 = fnv1a(e.text, strlen(e.text), h1); h2 = fnv1a(e.text, strlen(e.text), h2); }
  char key[40];
  snprintf(key, sizeof(key), "n%016llx%08llx", (unsigned long long)h1, (unsigned long long)(h2 >> 32));
  const std::string cdir = jit_cache_dir();
  const std::string cpath = cdir.empty() ? std::string() : cdir + "/" + key + ".cubin";
  struct stat st;
  if (!cpath.empty() && trusted_file(cpath, &st) && st.st_size > 0) {
    if (FILE* f = fopen(cpath.c_str(), "rb")) {
      void* buf = malloc((size_t)st.st_size);
      const bool got = buf && fread(buf, 1, (size_t)st.st_size, f) == (size_t)st.st_size;
      fclose(f);
      if (got) { *image = buf; *bytes = (size_t)st.st_size; return SLV_OK; }
      free(buf);
    }
  }
  NvrtcApi& api = nvrtc_api();
  if (!api.ok) {
    put_log(log, log_bytes, "libnvrtc.so.12 is not available (set SLV_NVRTC_LIB): run-time shader compilation needs the CUDA toolkit's NVRTC");
    return SLV_FAILED;
  }
  // ---- the program: slv_jit_unit.cu + every other embedded source, the generated shader and the header shims as named headers
  const char* main_src = nullptr;
  std::vector<const char*> h_text, h_name;
  for (const EmbeddedSource& e : kEmbeddedSources) {
    if (!strcmp(e.name, "slv_jit_unit.cu")) { main_src = e.text; continue; }
    h_text.push_back(e.text);
    h_name.push_back(e.name);
  }
  const std::pair<const char*, const char*> shims[] = {{"slv_generated_shader.cuh", device_code}, {"stdint.h", kShimStdint}, {"cstdint", kShimStdint},
                                                       {"stddef.h", kShimEmpty}, {"cuda_runtime.h", kShimEmpty}, {"cmath", kShimEmpty},
                                                       {"cstring", kShimEmpty}};
  for (auto& sh : shims) { h_name.push_back(sh.first); h_text.push_back(sh.second); }
  NvrtcApi::Program prog = nullptr;
  if (!main_src || api.CreateProgram(&prog, main_src, "slv_jit_unit.cu", (int)h_text.size(), h_text.data(), h_name.data()) != 0) return SLV_FAILED;
  std::vector<const char*> copts;
  for (const std::string& o : opts) copts.push_back(o.c_str());
  const int rc = api.CompileProgram(prog, (int)copts.size(), copts.data());
  size_t ln = 0;
  if (api.GetProgramLogSize(prog, &ln) == 0 && ln > 1) {
    std::string text(ln, '\0');
    api.GetProgramLog(prog, &text[0]);
    put_log(log, log_bytes, text);
  }
  slv_result res = SLV_FAILED;
  size_t n = 0;
  if (rc == 0 && api.GetCUBINSize(prog, &n) == 0 && n) {
    void* buf = malloc(n);
    if (buf && api.GetCUBIN(prog, (char*)buf) == 0) {
      *image = buf;
      *bytes = n;
      res = SLV_OK;
      if (!cpath.empty()) {  // unique temporary name, then an atomic rename: concurrent compiles never expose a partial file
        const std::string tmp = cpath + "." + std::to_string((long long)getpid()) + ".tmp";
        if (FILE* f = fopen(tmp.c_str(), "wb")) {
          const bool wrote = fwrite(buf, 1, n, f) == n;
          fclose(f);
          chmod(tmp.c_str(), 0600);
          if (!wrote || rename(tmp.c_str(), cpath.c_str()) != 0) unlink(tmp.c_str());
        }
      }
    } else {
      free(buf);
    }
  }
  api.DestroyProgram(&prog);
  return res;
}

slv_result slv_shader_compile(slv_device dev, uint32_t stage, const char* device_code, uint32_t n_vs_output_attrs, uint32_t flags,
                              slv_handle* out, char* log, size_t log_bytes) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  void* image = nullptr;
  size_t bytes = 0;
  slv_result rc = slv_shader_compile_cubin(stage, device_code, n_vs_output_attrs, flags, &image, &bytes, log, log_bytes);
  if (rc != SLV_OK) return rc;
  rc = slv_shader_module_load(dev, stage, image, bytes, n_vs_output_attrs, out);
  free(image);
  return rc;
}

void slv_free(void* p) { free(p); }

slv_result slv_resource_release(slv_device dev, slv_handle h) {
  if (!dev || h == 0 || h >= dev->res.size()) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  CU(cudaStreamSynchronize(dev->stream));
  Resource& r = dev->res[h];
  if (r.kind == Resource::BUFFER) CU(cudaFree(r.dptr));
  if (r.kind == Resource::TEXTURE)
    for (uint32_t l = 0; l < r.tex.n_levels; ++l) CU(cudaFree(r.tex.level[l].data));
  if (r.kind == Resource::MODULE && r.module) driver_api().ModuleUnload(r.module);
  if (r.rb_event) { CU(cudaStreamSynchronize(dev->copy_stream)); cudaEventDestroy(r.rb_event); }
  r = Resource();
  return SLV_OK;
}

static bool fill_sampler(slv_device dev, slv_handle h, SamplerRef& out) {
  auto r = dev->get(h, Resource::SAMPLER);
  if (!r) return false;
  auto t = dev->get(r->sampler_tex, Resource::TEXTURE);
  if (!t) return false;
  out.d = r->sd;
  out.tex = t->tex;
  bool fast = r->sd.min_filter == SLV_FILTER_LINEAR && r->sd.mag_filter == SLV_FILTER_LINEAR &&
              r->sd.addr_mode_u == SLV_ADDR_WRAP && r->sd.addr_mode_v == SLV_ADDR_WRAP && t->fmt == SLV_PF_RGBA8;
  for (uint32_t l = 0; l < t->tex.n_levels; ++l) {  // size 1 has mask 0 and index 0 for every coordinate: also exact
    const SurfaceRef& lv = t->tex.level[l];
    fast = fast && lv.w <= 1024 && lv.h <= 1024 && (lv.w & (lv.w - 1)) == 0 && (lv.h & (lv.h - 1)) == 0;
  }
  out.fast_wrap_rgba8 = fast ? 1u : 0u;
  out.touched = (dev->level_tracking && r->sampler_tex < LEVEL_TRACK_SLOTS) ? dev->d_level_touched + r->sampler_tex : nullptr;
  return true;
}

slv_result slv_draw(slv_device dev, const slv_draw_desc* d) {
  if (!dev || !d) return SLV_INVALID_PARAMETER;
  if (dev->failed) return SLV_FAILED;
  CU(cudaSetDevice(dev->ordinal));
  // ---- validation that the reference performs (renderer_impl.cpp:49-54,71-78,145-152,159-238)
  if (d->topology != SLV_TOPO_TRIANGLE_LIST && d->topology != SLV_TOPO_TRIANGLE_STRIP) return SLV_FAILED;
  if (d->index_buffer && d->index_format != SLV_INDEX_R16_UINT && d->index_format != SLV_INDEX_R32_UINT) return SLV_FAILED;
  const slv_viewport& vp = d->viewport;
  if (vp.x < 0 || vp.y < 0 || vp.w >= SLV_MAX_RENDER_TARGET_SIZE || vp.h >= SLV_MAX_RENDER_TARGET_SIZE) return SLV_FAILED;
  if (d->n_color_targets >= SLV_MAX_RENDER_TARGETS) return SLV_FAILED;
  // the device programs write colour target 0 (and the coverage probe target 1): further targets would be accepted and never written
  for (uint32_t i = 2; i < d->n_color_targets; ++i)
    if (d->color_targets[i]) return SLV_INVALID_PARAMETER;
  if (d->n_streams > 8 || d->n_elements > SLV_MAX_VS_INPUT_ATTRS) return SLV_INVALID_PARAMETER;
  // SASL shaders compiled at run time: SLV_PROGRAM_JIT(module) names a loaded shader module
  slv_handle vs_module = 0, ps_module = 0;
  uint32_t vs_program = d->vs.program, ps_program = d->ps.program;
  if (vs_program & 0x80000000u) {
    vs_module = vs_program & 0x7FFFFFFFu;
    auto m = dev->get(vs_module, Resource::MODULE);
    if (!m || m->module_stage != SLV_STAGE_VS) return SLV_INVALID_PARAMETER;
    vs_program = SLV_VS_JIT;
  }
  if (ps_program & 0x80000000u) {
    ps_module = ps_program & 0x7FFFFFFFu;
    auto m = dev->get(ps_module, Resource::MODULE);
    if (!m || m->module_stage != SLV_STAGE_PS) return SLV_INVALID_PARAMETER;
    ps_program = SLV_PS_JIT;
  }
  uint32_t n_attrs = vs_module ? dev->res[vs_module].module_attrs : vs_num_attrs(d->vs);
  if (n_attrs > SLV_MAX_VS_OUTPUT_ATTRS) return SLV_INVALID_PARAMETER;

  RasterParams rp{};
  float tw = FLT_MAX, th = FLT_MAX;
  uint32_t S = 0;
  for (uint32_t i = 0; i < d->n_color_targets; ++i) {
    SurfaceRef s = surface_of(dev, d->color_targets[i]);
    if (i == 0) rp.color0 = s;
    if (i == 1) rp.color1 = s;
    if (s.data) {
      tw = std::min((float)s.w, tw);
      th = std::min((float)s.h, th);
      if (S == 0) S = s.samples;
      else if (S != s.samples) return SLV_FAILED;
    }
  }
  if (d->ds_target) {
    auto r = dev->get(d->ds_target, Resource::TEXTURE);
    if (!r || r->fmt != SLV_PF_RG32F) return SLV_FAILED;
    rp.ds = r->tex.level[0];
    if (d->n_color_targets == 0) { S = rp.ds.samples; tw = (float)rp.ds.w; th = (float)rp.ds.h; }
    if ((float)rp.ds.w < tw || (float)rp.ds.h < th || rp.ds.samples != S) return SLV_FAILED;
  }
  if ((d->n_color_targets == 0 || !rp.color0.data) && !rp.ds.data) return SLV_FAILED;
  if (S != 1 && S != 2 && S != 4) return SLV_INVALID_PARAMETER;
  if (rp.color1.data && (rp.color1.w != rp.color0.w || rp.color1.h != rp.color0.h)) return SLV_INVALID_PARAMETER;
  rp.target_w = (uint32_t)tw;
  rp.target_h = (uint32_t)th;

  // ---- geometry parameters
  GeomParams gp{};
  for (uint32_t i = 0; i < d->n_streams; ++i) {
    auto r = dev->get(d->streams[i].buffer, Resource::BUFFER);
    if (!r) return SLV_INVALID_PARAMETER;
    gp.streams[i].data = r->dptr;
    gp.streams[i].stride = d->streams[i].stride;
    gp.streams[i].offset = d->streams[i].offset;
  }
  gp.n_elements = d->n_elements;
  for (uint32_t i = 0; i < d->n_elements; ++i) {
    gp.elements[i] = d->elements[i];
    if (d->elements[i].slot >= d->n_streams || d->elements[i].reg >= SLV_MAX_VS_INPUT_ATTRS) return SLV_INVALID_PARAMETER;
  }
  gp.fast_layout = 1;
  for (uint32_t i = 0; i < d->n_elements; ++i) {
    const slv_input_element& e = d->elements[i];
    const slv_vertex_stream& vs = d->streams[e.slot];
    if (e.reg != i || e.format != SLV_FMT_R32G32B32A32_FLOAT || ((vs.offset + e.aligned_byte_offset) & 15) || (vs.stride & 15)) gp.fast_layout = 0;
  }
  if (d->index_buffer) {
    auto r = dev->get(d->index_buffer, Resource::BUFFER);
    if (!r) return SLV_INVALID_PARAMETER;
    gp.indices = r->dptr;
    gp.index_stride = d->index_format == SLV_INDEX_R16_UINT ? 2 : 4;
    // vertices the bound streams hold for this layout: the post-transform vertex cache covers [0, vc_cap) (a hint here; the
    // flush decides whether the draw is cached and points vc_pos / vc_attr / vc_flags at its group's range)
    uint64_t cap = d->n_elements ? 0xFFFFFFFFull : 0;
    for (uint32_t i = 0; i < d->n_elements; ++i) {
      const slv_input_element& e = d->elements[i];
      const slv_vertex_stream& vs = d->streams[e.slot];
      const uint64_t bytes = dev->get(vs.buffer, Resource::BUFFER)->bytes;
      const uint64_t sz = gp.fast_layout ? 16 : (e.format == SLV_FMT_R32_FLOAT ? 4 : e.format == SLV_FMT_R32G32_FLOAT ? 8 : e.format == SLV_FMT_R32G32B32_FLOAT ? 12 : 16);
      const uint64_t first = (uint64_t)vs.offset + e.aligned_byte_offset + sz;
      uint64_t c = 0;
      if (first <= bytes) c = vs.stride ? (bytes - first) / vs.stride + 1 : 0xFFFFFFFFull;
      cap = std::min(cap, c);
    }
    gp.vc_cap = (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFFFull);
  }
  gp.topology = d->topology;
  gp.start = d->start;
  gp.prim_count = d->prim_count;
  gp.base_vertex = d->base_vertex;
  gp.vs_program = vs_program;
  // vertex texture fetch: vs.samplers[0] (set_vs_sampler, renderer.h:80)
  const bool vs_needs_sampler = vs_program == SLV_VS_TERRAIN_VTF || (vs_program == SLV_VS_JIT && d->vs.samplers[0] != 0);
  if (vs_needs_sampler && !fill_sampler(dev, d->vs.samplers[0], gp.sampler0)) return SLV_INVALID_PARAMETER;
  if (vs_program == SLV_VS_TERRAIN_VTF && d->n_elements < 2) return SLV_INVALID_PARAMETER;
  memcpy(gp.vs_uniforms, d->vs.uniforms, sizeof(gp.vs_uniforms));
  gp.n_attrs = n_attrs;
  rp.has_centroid = 0;
  for (uint32_t i = 0; i < SLV_MAX_VS_OUTPUT_ATTRS; ++i) {
    gp.mods[i] = rp.mods[i] = d->vs_attr_modifiers[i] ? d->vs_attr_modifiers[i] : (uint32_t)SLV_AM_LINEAR;
    if (i < n_attrs && (gp.mods[i] & SLV_AM_CENTROID)) rp.has_centroid = 1;
  }
  gp.cull_mode = d->raster.cull_mode;
  gp.front_ccw = d->raster.front_ccw;
  gp.vp = vp;
  // tile grid (rasterizer.cpp:1106-1108)
  gp.tiles_x = (uint32_t)(static_cast<size_t>(vp.w + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE);
  gp.tiles_y = (uint32_t)(static_cast<size_t>(vp.h + SLV_TILE_SIZE - 1) / SLV_TILE_SIZE);
  const uint32_t n_tiles = gp.tiles_x * gp.tiles_y;
  gp.shard_rank = dev->shard_rank;
  gp.shard_n = dev->shard_n;

  // counters that are pure functions of the arguments (default_vertex_cache.cpp:354-364,
  // geom_setup_engine.cpp:103, rasterizer.cpp:1137)
  // ... applied once the draw has passed every validation below (a rejected draw must not move the statistics)
  auto count_draw = [&]() {
    dev->host_stats.ia_vertices += 3ull * d->prim_count;
    dev->host_stats.ia_primitives += d->prim_count;
    // vs_invocations: per corner for draws without a post-transform cache (added at the flush, which decides), else the number
    // of vertices k_vertex_shade ran (device counter)
    dev->host_stats.cinvocations += d->prim_count;
  };
  if (d->bs.program < SLV_BS_REPLACE || d->bs.program > SLV_BS_REPLACE_AND_COUNT) return SLV_INVALID_PARAMETER;
  if (d->prim_count == 0 || n_tiles == 0) { count_draw(); return SLV_OK; }

  const uint32_t R = 1 + n_attrs;
  const uint32_t tri_stride = TRI_HEADER + 3 * MAX_REGS;  // uniform across the batch: slot -> record address
  const uint64_t n_slots64 = 3ull * d->prim_count;
  if (n_slots64 >= (1ull << 28)) return SLV_INVALID_PARAMETER;  // sub-list entries are (slot << 4) | block status
  const uint32_t n_slots = (uint32_t)n_slots64;
  // ---- can this draw join the queued batch? (same targets, sample count, pixel-shader program, tile grid)
  if (!dev->pending.empty()) {
    const RasterParams& f = dev->pending[0];
    bool same = f.color0.data == rp.color0.data && f.color1.data == rp.color1.data && f.ds.data == rp.ds.data &&
                f.ps_program == ps_program && dev->batch_ps_module == ps_module && dev->batch_S == S && f.tiles_x == gp.tiles_x && f.tiles_y == gp.tiles_y &&
                f.target_w == rp.target_w && f.target_h == rp.target_h && dev->pending.size() < MAX_BATCH;
    if (!same) {
      slv_result rcf = flush_batch(dev);
      if (rcf != SLV_OK) return rcf;
    }
  }
  const size_t tris_need = (size_t)n_slots * tri_stride;
  const uint64_t list_need = std::max<uint64_t>(8ull * (dev->slots_queued + n_slots), 1u << 22);
  slv_result rc = ensure_scratch(dev, dev->tris_used + tris_need, n_tiles, list_need);
  if (rc != SLV_OK) return rc;
  if (dev->tris_used + tris_need > dev->tris_cap) return SLV_OUT_OF_MEMORY;
  if (dev->slots_queued + n_slots >= (1ull << 28)) {  // 28-bit slots: sub-list entries are (slot << 4) | block status
    slv_result rcf = flush_batch(dev);
    if (rcf != SLV_OK) return rcf;
  }
  gp.tri_stride = tri_stride;
  gp.stats = dev->d_stats;

  // ---- depth/stencil function selection (framebuffer.cpp:325-425)
  const slv_depth_stencil_desc& ds = d->ds;
  rp.depth_enable = ds.depth_enable != 0;
  rp.depth_func = ds.depth_func;
  rp.stencil_enable = ds.stencil_enable != 0;
  rp.read_depth = rp.write_depth = 0;
  if (rp.ds.data && ds.depth_enable) {
    if (ds.depth_func != SLV_CMP_NEVER && ds.depth_func != SLV_CMP_ALWAYS) rp.read_depth = 1;
    if (ds.depth_write_mask && ds.depth_func != SLV_CMP_NEVER) rp.write_depth = 1;
  }
  if (!rp.ds.data) rp.stencil_enable = 0;
  rp.early_z = !ds.stencil_enable;
  rp.read_mask = ds.stencil_read_mask & 0xFF;
  rp.write_mask = ds.stencil_write_mask & 0xFF;
  rp.stencil_ref = ds.stencil_enable ? ((uint32_t)d->stencil_ref & rp.read_mask) : 0;
  rp.front_face = ds.front_face;
  rp.back_face = ds.back_face;
  rp.ps_program = ps_program;
  rp.bs_program = d->bs.program;
  if (rp.bs_program < SLV_BS_REPLACE || rp.bs_program > SLV_BS_REPLACE_AND_COUNT) return SLV_INVALID_PARAMETER;
  memcpy(rp.ps_uniforms, d->ps.uniforms, sizeof(rp.ps_uniforms));
  bool needs_sampler = (rp.ps_program == SLV_PS_JIT && d->ps.samplers[0] != 0) ||
                       rp.ps_program == SLV_PS_TEX_ALPHA || rp.ps_program == SLV_PS_TEX_GRAD_ALPHA ||
                       ((rp.ps_program == SLV_PS_SPONZA || rp.ps_program == SLV_PS_SPONZA_GRAD) &&
                        reinterpret_cast<const slv_ps_sponza_uniforms*>(d->ps.uniforms)->has_sampler);
  bool needs_sampler1 = rp.ps_program == SLV_PS_JIT && d->ps.samplers[1] != 0;  // a SASL pixel shader's second sampler
  if (rp.ps_program == SLV_PS_SSM_DRAW) {
    auto u = reinterpret_cast<const slv_ps_ssm_draw_uniforms*>(d->ps.uniforms);
    if (d->ps.uniform_bytes < sizeof(slv_ps_ssm_draw_uniforms) || n_attrs < 5) return SLV_INVALID_PARAMETER;
    needs_sampler = u->has_tex_sampler != 0;
    needs_sampler1 = u->has_depth_sampler != 0;
  }
  if (needs_sampler && !fill_sampler(dev, d->ps.samplers[0], rp.sampler0)) return SLV_INVALID_PARAMETER;
  if (needs_sampler1 && !fill_sampler(dev, d->ps.samplers[1], rp.sampler1)) return SLV_INVALID_PARAMETER;
  if (rp.ps_program == SLV_PS_TEX_ALPHA || rp.ps_program == SLV_PS_TEX_GRAD_ALPHA) {
    if (reinterpret_cast<const slv_ps_tex_alpha_uniforms*>(d->ps.uniforms)->reg >= n_attrs) return SLV_INVALID_PARAMETER;
  }
  if ((rp.ps_program == SLV_PS_LIGHTS3 || rp.ps_program == SLV_PS_SPONZA || rp.ps_program == SLV_PS_SPONZA_GRAD) && n_attrs < 4) return SLV_INVALID_PARAMETER;
  if ((rp.ps_program == SLV_PS_ATTR0_COLOR || rp.ps_program == SLV_PS_DISCARD_ALL || rp.ps_program == SLV_PS_HEIGHT_COLOR) && n_attrs < 1) return SLV_INVALID_PARAMETER;
  rp.tri_stride = tri_stride;
  rp.tiles_x = gp.tiles_x;
  rp.tiles_y = gp.tiles_y;
  rp.shard_rank = dev->shard_rank;
  rp.shard_n = dev->shard_n;
  rp.n_attrs = n_attrs;
  rp.stats = dev->d_stats;

  // sampling a texture that is a target of the queued batch: the earlier draws must land first
  if (needs_sampler && !dev->pending.empty()) {
    const uint8_t* t0 = rp.sampler0.tex.level[0].data;
    if (t0 == rp.color0.data || t0 == rp.color1.data || t0 == rp.ds.data) {
      slv_result rcf = flush_batch(dev);
      if (rcf != SLV_OK) return rcf;
    }
  }
  if (needs_sampler1 && !dev->pending.empty()) {
    const uint8_t* t1 = rp.sampler1.tex.level[0].data;
    if (t1 == rp.color0.data || t1 == rp.color1.data || t1 == rp.ds.data) {
      slv_result rcf = flush_batch(dev);
      if (rcf != SLV_OK) return rcf;
    }
  }
  if (vs_needs_sampler) {
    // the vertex stage runs on the front stream: a height map written on the main stream (upload, an earlier render pass) has to
    // be complete first - flush what is queued and let the front half order after the main stream
    const uint8_t* t0 = gp.sampler0.tex.level[0].data;
    if (!dev->pending.empty() && (t0 == dev->pending[0].color0.data || t0 == dev->pending[0].color1.data || t0 == dev->pending[0].ds.data)) {
      slv_result rcf = flush_batch(dev);
      if (rcf != SLV_OK) return rcf;
    }
    { slv_result rcm__ = materialize_clear(dev, texture_of_data(dev, t0)); if (rcm__ != SLV_OK) return rcm__; }
    dev->buffers_dirty = true;
  }

  // lazy clears: a texture this draw SAMPLES, and the second colour target, must hold their cleared contents for real
  if (needs_sampler) {
    Resource* rt = texture_of_data(dev, rp.sampler0.tex.level[0].data);
    { slv_result rcm__ = materialize_clear(dev, rt); if (rcm__ != SLV_OK) return rcm__; }
    { slv_result rca__ = wait_assembly(dev, rt); if (rca__ != SLV_OK) return rca__; }
  }
  if (needs_sampler1) {
    Resource* rt = texture_of_data(dev, rp.sampler1.tex.level[0].data);
    { slv_result rcm__ = materialize_clear(dev, rt); if (rcm__ != SLV_OK) return rcm__; }
    { slv_result rca__ = wait_assembly(dev, rt); if (rca__ != SLV_OK) return rca__; }
  }
  if (rp.color1.data) { slv_result rcm__ = materialize_clear(dev, texture_of_data(dev, rp.color1.data)); if (rcm__ != SLV_OK) return rcm__; }
  dev->batch_color = d->n_color_targets ? d->color_targets[0] : 0;
  dev->batch_ds = d->ds_target;

  // ---- bind the scratch set the batch is being built in (every flush above may have switched sets)
  {
    slv_device_t::Scratch& S = dev->S();
    gp.tris = S.tris;
    rp.tris = S.tris;
    gp.slot_base = rp.slot_base = (uint32_t)dev->slots_queued;
    gp.draw_id = (uint32_t)dev->pending.size();
    gp.tile_count = S.tile_count;
    gp.valid_slots = S.valid_slots;
    gp.valid_count = S.valid_count;
    gp.big_slots = dev->big_tiles ? S.big_slots : nullptr;
    gp.big_count = S.big_count;
    rp.tile_offset = S.tile_offset;
    rp.active_tiles = S.active_tiles;
    rp.work_counter = S.work_counter;
    rp.list = S.list;
    rp.list_capacity = dev->list_cap;
  }

  // ---- queue the draw: geometry, binning and the raster pass all run at the next flush point
  count_draw();
  dev->pending.push_back(rp);
  dev->pending_geom.push_back(gp);
  dev->pending_vs_module.push_back(vs_module);
  dev->pending_vs_runs.push_back(3ull * d->prim_count);
  dev->batch_ps_module = ps_module;
  dev->batch_S = S;
  dev->tris_used += tris_need;
  dev->slots_queued += n_slots;
  return SLV_OK;
}

slv_result slv_clear_color(slv_device dev, slv_handle h, const float rgba[4]) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || !rgba) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  const SurfaceRef& s = r->tex.level[0];
  uint32_t w[4];
  switch (s.fmt) {  // from_rgba32 conversion done once (surface.cpp:170-173)
  case SLV_PF_RGBA32F: memcpy(w, rgba, 16); break;
  case SLV_PF_RG32F: memcpy(w, rgba, 8); w[2] = w[0]; w[3] = w[1]; break;
  case SLV_PF_RGBA8: {
    uint32_t p = host_unorm8(rgba[0]) | (host_unorm8(rgba[1]) << 8) | (host_unorm8(rgba[2]) << 16) | ((uint32_t)host_unorm8(rgba[3]) << 24);
    w[0] = w[1] = w[2] = w[3] = p;
  } break;
  default: {
    uint32_t p = host_unorm8(rgba[2]) | (host_unorm8(rgba[1]) << 8) | (host_unorm8(rgba[0]) << 16) | ((uint32_t)host_unorm8(rgba[3]) << 24);
    w[0] = w[1] = w[2] = w[3] = p;
  } break;
  }
  if (dev->lazy_clear) {
    r->clear_pending = true;
    r->clear_pattern = make_uint4(w[0], w[1], w[2], w[3]);
    return SLV_OK;
  }
  { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
  return fill_surface(dev, s, make_uint4(w[0], w[1], w[2], w[3]));
}

slv_result slv_clear_depth_stencil(slv_device dev, slv_handle h, uint32_t flags, float depth, uint32_t stencil) {
  auto r = dev ? dev->get(h, Resource::TEXTURE) : nullptr;
  if (!r || r->fmt != SLV_PF_RG32F) return SLV_INVALID_PARAMETER;
  if (!(flags & (SLV_CLEAR_DEPTH | SLV_CLEAR_STENCIL))) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  const SurfaceRef& s = r->tex.level[0];
  if ((flags & 3) == 3) {
    uint32_t dbits;
    memcpy(&dbits, &depth, 4);
    if (dev->lazy_clear) {
      r->clear_pending = true;
      r->clear_pattern = make_uint4(dbits, stencil, dbits, stencil);
      return SLV_OK;
    }
    { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
    return fill_surface(dev, s, make_uint4(dbits, stencil, dbits, stencil));
  }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
  size_t n = s.bytes / 8;
  uint32_t blocks = (uint32_t)std::min<size_t>((n + 255) / 256, 148 * 16);
  k_clear_ds_partial<<<blocks, 256, 0, dev->stream>>>(reinterpret_cast<float2*>(s.data), n, flags, depth, stencil);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_resolve(slv_device dev, slv_handle src, slv_handle dst) {
  auto rs = dev ? dev->get(src, Resource::TEXTURE) : nullptr;
  auto rd = dev ? dev->get(dst, Resource::TEXTURE) : nullptr;
  if (!rs || !rd) return SLV_INVALID_PARAMETER;
  const SurfaceRef& s = rs->tex.level[0];
  const SurfaceRef& t = rd->tex.level[0];
  if (t.samples != 1 || t.w < s.w || t.h < s.h) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  SurfaceRef t2 = t;
  if (rd->resolve_peer && dev->shard_n > 1) t2.data = rd->resolve_peer;  // owned tiles go straight to the root's surface
  { slv_result rcw = wait_readback(dev, rd); if (rcw != SLV_OK) return rcw; }
  // the pending batch renders to `src`: its k_shade can write the resolved texels itself (fused resolve)
  // ... unless a queued draw READS the destination (samples it - temporal feedback - or has it bound as a second target or as
  // depth/stencil): the reference finishes every draw before it resolves, k_shade would overwrite texels other warps still read
  bool dst_in_use = false;
  for (size_t i = 0; i < dev->pending.size(); ++i) {
    const RasterParams& q = dev->pending[i];
    const GeomParams& g = dev->pending_geom[i];
    const uint8_t* used[] = {q.sampler0.tex.n_levels ? q.sampler0.tex.level[0].data : nullptr, q.sampler1.tex.n_levels ? q.sampler1.tex.level[0].data : nullptr,
                             g.sampler0.tex.n_levels ? g.sampler0.tex.level[0].data : nullptr, q.color1.data, q.ds.data};
    for (const uint8_t* u : used) dst_in_use = dst_in_use || (u && (u == t.data || u == t2.data));
  }
  if (dev->fuse_resolve && !dev->pending.empty() && dev->pending[0].color0.data == s.data && rd != rs && !dst_in_use) {
    { slv_result rcm__ = materialize_clear(dev, rd); if (rcm__ != SLV_OK) return rcm__; }
    dev->resolve_requested = true;
    dev->resolve_dst = t2;
  }
  dev->resolve_done = false;
  slv_result rcf = flush_batch(dev);
  dev->resolve_requested = false;
  if (rcf != SLV_OK) return rcf;
  if (dev->resolve_done) {
    dev->resolve_done = false;
    CU(cudaGetLastError());
    return SLV_OK;
  }
  { slv_result rcm__ = materialize_clear(dev, rs); if (rcm__ != SLV_OK) return rcm__; }
  { slv_result rcm__ = materialize_clear(dev, rd); if (rcm__ != SLV_OK) return rcm__; }
  dim3 blk(32, 8), grd((s.w + 31) / 32, (s.h + 7) / 8);
  k_resolve<<<grd, blk, 0, dev->stream>>>(s, t2, dev->shard_rank, dev->shard_n);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_flush(slv_device dev) {
  if (!dev) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_all(dev); if (rcm__ != SLV_OK) return rcm__; }
  CU(cudaStreamSynchronize(dev->stream));
  CU(cudaStreamSynchronize(dev->copy_stream));  // asynchronous readbacks have landed too
  return check_overflow(dev);
}

slv_result slv_query_begin(slv_device dev) {
  if (!dev) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  dev->host_stats = slv_pipeline_statistics{};
  CU(cudaMemsetAsync(dev->d_stats, 0, 20 * sizeof(unsigned long long), dev->stream));
  for (auto& m : dev->prof_ms) m = 0;
  if (!dev->spans.empty()) CU(cudaStreamSynchronize(dev->stream));
  dev->spans.clear();
  dev->ev_used = 0;
  dev->n_launches = 0;
  return SLV_OK;
}

slv_result slv_query_get(slv_device dev, slv_pipeline_statistics* out) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  unsigned long long h[20];
  CU(cudaMemcpyAsync(h, dev->d_stats, sizeof(h), cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  *out = dev->host_stats;
  out->vs_invocations += h[17];  // vertices shaded into the post-transform cache
  out->cprimitives = h[6];
  out->ps_invocations = h[7];
  out->backend_input_pixels = h[8];
  return check_overflow(dev);
}

slv_result slv_profile_get(slv_device dev, slv_pipeline_profiles* out) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  CU(cudaStreamSynchronize(dev->stream));
  for (auto& sp : dev->spans) {  // fold the recorded event spans into the per-stage sums
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, dev->ev_pool[sp.a], dev->ev_pool[sp.b]));
    dev->prof_ms[sp.stage] += ms;
  }
  dev->spans.clear();
  dev->ev_used = 0;
  memset(out, 0, sizeof(*out));
  out->clipping = (uint64_t)(dev->prof_ms[0] * 1e6);      // VS + clip + viewport + setup are one kernel
  out->tri_dispatch = (uint64_t)((dev->prof_ms[1] + dev->prof_ms[2] + dev->prof_ms[5]) * 1e6);
  out->ras = (uint64_t)((dev->prof_ms[3] + dev->prof_ms[4]) * 1e6);
  return SLV_OK;
}

// Developer aid (not part of the reference surface): copies internal work-list sizes of the LAST flushed batch to the
// host.  which = 0: active tile ids ([0] = count), 1: tile_offset[tiles + 1], 2: block_desc[active tiles * 128] as (first entry, entries) pairs.
slv_result slv_debug_read(slv_device dev, uint32_t which, void* dst, size_t bytes) {
  if (!dev || !dst) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  const slv_device_t::Scratch& L = dev->sc[(dev->pipeline && !dev->profile && dev->last_flushed >= 0) ? dev->last_flushed : dev->cur];  // the last flushed batch's set
  const void* src = which == 0 ? (const void*)L.active_tiles : which == 1 ? (const void*)L.tile_offset : (const void*)L.block_desc;
  const size_t cap = which == 0 ? ((size_t)dev->tiles_cap + 1) * 4 : which == 1 ? (size_t)dev->tiles_cap * 4 : (size_t)dev->tiles_cap * 128 * 8;
  if (!src || bytes > cap) return SLV_INVALID_PARAMETER;
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  return SLV_OK;
}

slv_result slv_profile_get_stages(slv_device dev, double* ms, uint32_t n) {
  if (!dev || !ms || n < 6) return SLV_INVALID_PARAMETER;
  slv_pipeline_profiles tmp;
  slv_result rc = slv_profile_get(dev, &tmp);  // folds the pending event spans
  if (rc != SLV_OK) return rc;
  for (uint32_t i = 0; i < 6; ++i) ms[i] = dev->prof_ms[i];
  return SLV_OK;
}

slv_result slv_traffic_get(slv_device dev, slv_traffic_counters* out) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  unsigned long long h[20];
  CU(cudaMemcpyAsync(h, dev->d_stats, sizeof(h), cudaMemcpyDeviceToHost, dev->stream));
  CU(cudaStreamSynchronize(dev->stream));
  out->ps_executed = h[16];
  out->list_entries_scanned = h[13];
  out->region_survivors = h[14];
  out->warp_pairs = h[15];
  out->quads_shaded = h[7] / 4;
  out->z_tested = h[9];
  out->z_written = h[10];
  out->c_written = h[11];
  out->c_read = h[12];
  return SLV_OK;
}

slv_result slv_texture_level_tracking(slv_device dev, uint32_t on) {
  if (!dev) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }  // queued draws keep the setting they were issued under
  if (on) {
    { slv_result rcs = sync_all(dev); if (rcs != SLV_OK) return rcs; }
    if (!dev->d_level_touched) CU(cudaMalloc(&dev->d_level_touched, LEVEL_TRACK_SLOTS * sizeof(uint32_t)));
    CU(cudaMemsetAsync(dev->d_level_touched, 0, LEVEL_TRACK_SLOTS * sizeof(uint32_t), dev->stream));
    CU(cudaStreamSynchronize(dev->stream));
  }
  dev->level_tracking = on != 0;
  return SLV_OK;
}

slv_result slv_texture_levels_touched(slv_device dev, slv_handle tex, uint32_t* mask) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || !mask) return SLV_INVALID_PARAMETER;
  *mask = 0;
  if (!dev->d_level_touched || tex >= LEVEL_TRACK_SLOTS) return SLV_OK;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcs = sync_all(dev); if (rcs != SLV_OK) return rcs; }
  CU(cudaMemcpy(mask, dev->d_level_touched + tex, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return SLV_OK;
}

slv_result slv_kernel_launch_count(slv_device dev, uint64_t* out) {
  if (!dev || !out) return SLV_INVALID_PARAMETER;
  *out = dev->n_launches;
  return SLV_OK;
}

slv_result slv_event_record(slv_device dev, uint32_t slot) {
  if (!dev || slot >= 16) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  CU(cudaEventRecord(dev->user_ev[slot], dev->stream));
  return SLV_OK;
}

slv_result slv_event_elapsed_ms(slv_device dev, uint32_t a, uint32_t b, float* ms) {
  if (!dev || a >= 16 || b >= 16 || !ms) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  CU(cudaEventSynchronize(dev->user_ev[b]));
  CU(cudaEventElapsedTime(ms, dev->user_ev[a], dev->user_ev[b]));
  return SLV_OK;
}

slv_result slv_set_stream(slv_device dev, void* cuda_stream) {
  if (!dev) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_all(dev); if (rcm__ != SLV_OK) return rcm__; }
  { slv_result rcs__ = sync_all(dev); if (rcs__ != SLV_OK) return rcs__; }
  dev->stream = cuda_stream ? (cudaStream_t)cuda_stream : dev->own_stream;
  return SLV_OK;
}

slv_result slv_profile_enable(slv_device dev, uint32_t on) {
  if (!dev) return SLV_INVALID_PARAMETER;
  dev->profile = on != 0;
  return SLV_OK;
}

slv_result slv_texture_device_ptr(slv_device dev, slv_handle tex, uint32_t level, void** out, size_t* bytes) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels || !out) return SLV_INVALID_PARAMETER;
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }  // the caller will touch the memory itself
  *out = r->tex.level[level].data;
  if (bytes) *bytes = r->tex.level[level].bytes;
  return SLV_OK;
}

static slv_result pack_common(slv_device dev, slv_handle tex, uint32_t rank, uint32_t nranks, void* staging, size_t* bytes,
                              int unpack) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || nranks == 0 || rank >= nranks) return SLV_INVALID_PARAMETER;
  const SurfaceRef& s = r->tex.level[0];
  if (s.samples != 1) return SLV_INVALID_PARAMETER;
  uint32_t tiles_x = (s.w + TILE - 1) / TILE, tiles_y = (s.h + TILE - 1) / TILE, n_tiles = tiles_x * tiles_y;
  // dense slot of every owned tile: built once per (grid, rank, nranks) and kept on the device, so the per-frame
  // pack / unpack calls enqueue one kernel and never synchronise the host
  slv_device_t::SlotTable* tab = nullptr;
  for (auto& t : dev->slot_tables)
    if (t.tiles_x == tiles_x && t.tiles_y == tiles_y && t.rank == rank && t.nranks == nranks) tab = &t;
  if (!tab) {
    std::vector<uint32_t> slot(n_tiles, 0);
    uint32_t owned = 0;
    for (uint32_t t = 0; t < n_tiles; ++t)
      if (nranks <= 1 || ((t % tiles_x) + 3 * (t / tiles_x)) % nranks == rank) slot[t] = owned++;
    slv_device_t::SlotTable nt{tiles_x, tiles_y, rank, nranks, owned, nullptr};
    CU(cudaSetDevice(dev->ordinal));
    CU(cudaMalloc(&nt.d_slot, n_tiles * sizeof(uint32_t)));
    CU(cudaMemcpy(nt.d_slot, slot.data(), n_tiles * sizeof(uint32_t), cudaMemcpyHostToDevice));
    dev->slot_tables.push_back(nt);
    tab = &dev->slot_tables.back();
  }
  if (bytes) *bytes = (size_t)tab->owned * TILE * TILE * s.bpp;
  if (!staging) return SLV_OK;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  if (unpack) { slv_result rcw = wait_readback(dev, r); if (rcw != SLV_OK) return rcw; }
  k_pack_tiles<<<n_tiles, 256, 0, dev->stream>>>(s, tiles_x, tiles_y, rank, nranks, (uint8_t*)staging, tab->d_slot, unpack);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_pack_tiles(slv_device dev, slv_handle tex, uint32_t rank, uint32_t nranks, void* staging, size_t* bytes) {
  return pack_common(dev, tex, rank, nranks, staging, bytes, 0);
}

slv_result slv_unpack_tiles(slv_device dev, slv_handle tex, uint32_t rank, uint32_t nranks, const void* staging) {
  if (!staging) return SLV_INVALID_PARAMETER;
  return pack_common(dev, tex, rank, nranks, const_cast<void*>(staging), nullptr, 1);
}

// ---- sort-first frame assembly over peer memory (CUDA IPC between the one-process-per-GPU ranks) -------------------
slv_result slv_peer_export_texture(slv_device dev, slv_handle tex, uint32_t level, uint8_t handle_out[SLV_PEER_HANDLE_BYTES]) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || level >= r->tex.n_levels || !handle_out) return SLV_INVALID_PARAMETER;
  static_assert(sizeof(cudaIpcMemHandle_t) <= SLV_PEER_HANDLE_BYTES, "IPC handle does not fit");
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcm__ = materialize_clear(dev, r); if (rcm__ != SLV_OK) return rcm__; }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, r->tex.level[level].data));
  memset(handle_out, 0, SLV_PEER_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  return SLV_OK;
}

slv_result slv_peer_export_flags(slv_device dev, uint8_t handle_out[SLV_PEER_HANDLE_BYTES]) {
  if (!dev || !handle_out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, dev->peer_flags));
  memset(handle_out, 0, SLV_PEER_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  return SLV_OK;
}

slv_result slv_peer_open(slv_device dev, const uint8_t handle[SLV_PEER_HANDLE_BYTES], void** dptr_out) {
  if (!dev || !handle || !dptr_out) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  CU(cudaIpcOpenMemHandle(dptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return SLV_OK;
}

slv_result slv_peer_close(slv_device dev, void* dptr) {
  if (!dev || !dptr) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcs__ = sync_all(dev); if (rcs__ != SLV_OK) return rcs__; }
  for (auto& r : dev->res)
    if (r.kind == Resource::TEXTURE && r.resolve_peer == dptr) r.resolve_peer = nullptr;
  CU(cudaIpcCloseMemHandle(dptr));
  return SLV_OK;
}

slv_result slv_resolve_target_peer(slv_device dev, slv_handle dst, void* peer_surface) {
  auto r = dev ? dev->get(dst, Resource::TEXTURE) : nullptr;
  if (!r) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  r->resolve_peer = (uint8_t*)peer_surface;
  return SLV_OK;
}

slv_result slv_peer_signal(slv_device dev, void* peer_flags, uint32_t index, uint32_t value) {
  if (!dev || index >= SLV_PEER_FLAGS) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  uint32_t* base = peer_flags ? (uint32_t*)peer_flags : dev->peer_flags;
  k_peer_signal<<<1, 1, 0, dev->stream>>>(base + index, value);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_flags_wait(slv_device dev, const void* flags, uint32_t first, uint32_t count, uint32_t value) {
  if (!dev || first + count > SLV_PEER_FLAGS) return SLV_INVALID_PARAMETER;
  if (count == 0) return SLV_OK;
  CU(cudaSetDevice(dev->ordinal));
  // the queued batch is NOT flushed: its front half touches no render target and its back half is enqueued on the main stream
  // behind this wait anyway - so a following slv_resolve can still be fused into the batch's k_shade
  k_flags_wait<<<1, 32, 0, dev->stream>>>(flags ? (const uint32_t*)flags : dev->peer_flags, first, count, value, dev->overflow_flag);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_assembly_wait(slv_device dev, slv_handle tex, const void* flags, uint32_t first, uint32_t count, uint32_t value) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || first + count > SLV_PEER_FLAGS) return SLV_INVALID_PARAMETER;
  if (count == 0) return SLV_OK;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }  // this device's own tiles of the frame are enqueued
  if (!r->rb_event) CU(cudaEventCreateWithFlags(&r->rb_event, cudaEventDisableTiming));
  CU(cudaEventRecord(dev->ev_copy, dev->stream));
  CU(cudaStreamWaitEvent(dev->copy_stream, dev->ev_copy, 0));
  k_flags_wait<<<1, 32, 0, dev->copy_stream>>>(flags ? (const uint32_t*)flags : dev->peer_flags, first, count, value, dev->overflow_flag);
  ++dev->n_launches;
  CU(cudaGetLastError());
  CU(cudaEventRecord(r->rb_event, dev->copy_stream));
  r->rb_pending = true;   // the next writer of the texture waits for the assembly (and whatever consumes it on the copy stream)
  r->asm_pending = true;  // ... and so do readers on the main stream
  return SLV_OK;
}

slv_result slv_peer_signal_after_consumers(slv_device dev, slv_handle tex, void* peer_flags, uint32_t index, uint32_t value) {
  auto r = dev ? dev->get(tex, Resource::TEXTURE) : nullptr;
  if (!r || index >= SLV_PEER_FLAGS) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  if (!dev->signal_stream) CU(cudaStreamCreateWithFlags(&dev->signal_stream, cudaStreamNonBlocking));
  if (r->rb_event && r->rb_pending) CU(cudaStreamWaitEvent(dev->signal_stream, r->rb_event, 0));
  uint32_t* base = peer_flags ? (uint32_t*)peer_flags : dev->peer_flags;
  k_peer_signal<<<1, 1, 0, dev->signal_stream>>>(base + index, value);
  ++dev->n_launches;
  CU(cudaGetLastError());
  return SLV_OK;
}

slv_result slv_set_tile_shard(slv_device dev, uint32_t rank, uint32_t nranks) {
  if (!dev || nranks == 0 || rank >= nranks) return SLV_INVALID_PARAMETER;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  dev->shard_rank = rank;
  dev->shard_n = nranks;
  return SLV_OK;
}

slv_result slv_sampler_probe(slv_device dev, slv_handle sh, uint32_t n, const float* coords, const float* ddx,
                             const float* ddy, const float* lod, uint32_t use_lod, float* out) {
  if (!dev || !coords || !out) return SLV_INVALID_PARAMETER;
  SamplerRef sm{};
  if (!fill_sampler(dev, sh, sm)) return SLV_INVALID_PARAMETER;
  if (n == 0) return SLV_OK;
  CU(cudaSetDevice(dev->ordinal));
  { slv_result rcf__ = flush_batch(dev); if (rcf__ != SLV_OK) return rcf__; }
  { slv_result rcm__ = materialize_clear(dev, texture_of_data(dev, sm.tex.level[0].data)); if (rcm__ != SLV_OK) return rcm__; }
  float *d_c = nullptr, *d_dx = nullptr, *d_dy = nullptr, *d_l = nullptr;
  float4* d_o = nullptr;
  CU(cudaMalloc(&d_c, n * 8));
  CU(cudaMalloc(&d_dx, n * 8));
  CU(cudaMalloc(&d_dy, n * 8));
  CU(cudaMalloc(&d_l, n * 4));
  CU(cudaMalloc(&d_o, n * 16));
  cudaStream_t st = dev->stream;
  CU(cudaMemcpyAsync(d_c, coords, n * 8, cudaMemcpyHostToDevice, st));
  if (ddx) CU(cudaMemcpyAsync(d_dx, ddx, n * 8, cudaMemcpyHostToDevice, st));
  if (ddy) CU(cudaMemcpyAsync(d_dy, ddy, n * 8, cudaMemcpyHostToDevice, st));
  if (lod) CU(cudaMemcpyAsync(d_l, lod, n * 4, cudaMemcpyHostToDevice, st));
  k_sampler_probe<<<(n + 127) / 128, 128, 0, st>>>(sm, n, d_c, d_dx, d_dy, d_l, use_lod, d_o);
  ++dev->n_launches;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, d_o, n * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  cudaFree(d_c); cudaFree(d_dx); cudaFree(d_dy); cudaFree(d_l); cudaFree(d_o);
  return SLV_OK;
}

}  // extern "C"
