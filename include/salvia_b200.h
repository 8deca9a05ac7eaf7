/* salvia_b200.h — the C-ABI drop-in boundary of the B200-native SALVIA draw pipeline.
 *
 * One flat `extern "C"` surface with opaque handles, POD descriptors and int32 status codes equal to
 * `salvia::result` (reference: salvia/include/salvia/common/constants.h:9).  It is what a third
 * subclass of the reference's `renderer_impl` (next to sync_renderer / async_renderer) binds in its
 * `commit_state_and_command()` override (reference: salvia/include/salvia/core/renderer_impl.h:111,
 * salvia/src/core/sync_renderer.cpp:26-29): every piece of state the reference snapshots into
 * `render_state` (salvia/include/salvia/core/render_state.h:46-94) for one draw/clear command is a
 * field of `slv_draw_desc` / an argument of `slv_clear_*`.  See INTEGRATION.md for the binding stub.
 *
 * Three shared libraries export this same table:
 *   libsalvia_b200.so   (salviarenderer_b200/csrc)  the product: hand-written sm_100a CUDA, no CPU path
 *   libsalvia_oracle.so (oracle/)                   CPU restatement — test infrastructure only
 *   libsalvia_ref.so    (oracle/_ref/)              the unmodified reference behind this ABI — ditto
 * so the parity tests drive all three with identical calls.
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef SALVIA_B200_H
#define SALVIA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLV_ABI_VERSION 1

/* ---- status codes == salvia::result (constants.h:9) -------------------------------------------- */
typedef int32_t slv_result;
enum { SLV_OK = 0, SLV_FAILED = 1, SLV_OUT_OF_MEMORY = 2, SLV_INVALID_PARAMETER = 3 };

/* ---- capacities (renderer_capacity.h:7-15) ------------------------------------------------------ */
enum {
  SLV_MAX_VS_INPUT_ATTRS = 8,
  SLV_MAX_VS_OUTPUT_ATTRS = 5, /* vs_output_ops is only populated for N = 0..5 (shader.cpp:45-52) */
  SLV_MAX_RENDER_TARGETS = 8,
  SLV_MAX_INPUT_SLOTS = 32,
  SLV_MAX_RENDER_TARGET_SIZE = 8192,
  SLV_MAX_SAMPLE_COUNT = 4,    /* sample patterns exist for 1, 2, 4 (rasterizer.cpp:1087-1103) */
  SLV_MAX_SAMPLERS = 4,
  SLV_MAX_UNIFORM_BYTES = 256,
  SLV_TILE_SIZE = 64           /* rasterizer.cpp:32 */
};

/* ---- enums with the reference's numeric values (constants.h) ------------------------------------ */
enum { SLV_TOPO_TRIANGLE_LIST = 2, SLV_TOPO_TRIANGLE_STRIP = 4 };          /* primitive_topology   */
enum { SLV_CULL_NONE = 0, SLV_CULL_FRONT = 1, SLV_CULL_BACK = 2 };         /* cull_mode            */
enum { SLV_ADDR_WRAP = 0, SLV_ADDR_MIRROR = 1, SLV_ADDR_CLAMP = 2, SLV_ADDR_BORDER = 3 };
enum { SLV_FILTER_POINT = 0, SLV_FILTER_LINEAR = 1, SLV_FILTER_ANISOTROPIC = 2 };
enum { SLV_MIP_LO_QUALITY = 0, SLV_MIP_MI_QUALITY = 1, SLV_MIP_HI_QUALITY = 2 };
enum {                                                                     /* compare_function     */
  SLV_CMP_NEVER = 0, SLV_CMP_LESS = 1, SLV_CMP_EQUAL = 2, SLV_CMP_LESS_EQUAL = 3,
  SLV_CMP_GREATER = 4, SLV_CMP_NOT_EQUAL = 5, SLV_CMP_GREATER_EQUAL = 6, SLV_CMP_ALWAYS = 7
};
enum {                                                                     /* stencil_op           */
  SLV_SOP_KEEP = 1, SLV_SOP_ZERO = 2, SLV_SOP_REPLACE = 3, SLV_SOP_INCR_SAT = 4,
  SLV_SOP_DECR_SAT = 5, SLV_SOP_INVERT = 6, SLV_SOP_INCR_WRAP = 7, SLV_SOP_DECR_WRAP = 8
};
enum { SLV_CLEAR_DEPTH = 1, SLV_CLEAR_STENCIL = 2 };                       /* clear_flag           */
enum { SLV_INDEX_NONE = 0, SLV_INDEX_R16_UINT = 57, SLV_INDEX_R32_UINT = 42 }; /* format.h         */
enum {                                                                     /* vertex element format */
  SLV_FMT_R32_FLOAT = 41, SLV_FMT_R32G32_FLOAT = 16, SLV_FMT_R32G32B32_FLOAT = 6,
  SLV_FMT_R32G32B32A32_FLOAT = 2
};
/* pixel formats of surfaces/textures, numeric values of the reference's pixel_format codes
 * (salvia/include/salvia/common/colors_convertors.h:40-47) */
enum { SLV_PF_RGBA32F = 0, SLV_PF_BGRA8 = 2, SLV_PF_RGBA8 = 3, SLV_PF_RG32F = 5 };
/* vs_output::attrib_modifier_type (shader_regs.h:53-59) */
enum { SLV_AM_LINEAR = 1, SLV_AM_CENTROID = 2, SLV_AM_NOINTERPOLATION = 4, SLV_AM_NOPERSPECTIVE = 8 };

/* ---- device shader programs ----------------------------------------------------------------------
 * cpp_vertex_shader / cpp_pixel_shader / cpp_blend_shader subclasses cannot run on the GPU, so a
 * host shader object names a registered __device__ program and carries its uniforms as a POD block
 * (layout documented per program; all floats, mat44 row-major as eflib::mat44, vec4 = 4 floats). */
enum {
  /* pos = in[0]·wvp; attr[i] = in[src[i]].   uniforms: mat44 wvp; u32 n_attrs; u32 src[5]          */
  SLV_VS_MVP_PASSTHROUGH = 1,
  /* pos = in[0]·wvp; attr0 = (in0.x, in0.z, 0, 0)  (TextureAndBlending.cpp:121-143) uniforms: wvp  */
  SLV_VS_PLANE_XZ = 2,
  /* pos = in[0]·wvp; attr0 = in[1]; attr1..3 = lightPos_k − in[0]
   * (ColorizedTriangle.cpp:29-53).  uniforms: mat44 wvp; vec4 lightPos[3]                          */
  SLV_VS_LIGHTS3 = 3,
  /* pos = in[0]·wvp; attr0 = in[1] (uv); attr1 = in[2] (normal); attr2 = lightPos − in[0];
   * attr3 = eyePos − in[0]  (Sponza.cpp:64-97).  uniforms: mat44 wvp; vec4 lightPos; vec4 eyePos   */
  SLV_VS_SPONZA = 4,
  /* vertex texture fetch (samples/VertexTextureFetch/VertexTextureFetch.cpp:38-61): uv = offset + in[1].xy * scale;
   * d = tex2Dlod(samplers[0], float4(uv, 0, 0)).x = sampler::sample_2d_lod(uv, 0); pos = float4(in[0].xyz + (0, d*20, 0), 1)·wvp;
   * attr0 = (d, 0, 0, 0).   uniforms: mat44 wvp; vec2 terrainOffset; vec2 terrainScale; vs.samplers[0] = the height map  */
  SLV_VS_TERRAIN_VTF = 5,
  /* the colour pass of samples/StandardShadowMap (resources/ssm/Draw.savs): in[0] = POSITION, in[1] = NORMAL, in[2] = TEXCOORD0;
   * pos = in[0]·cameraWvp; attr0 = in[2]; attr1 = in[1]; attr2 = lightPos − in[0]; attr3 = cameraPos − in[0];
   * attr4 = in[0]·lightWvp (the light-space position).  uniforms: slv_vs_ssm_draw_uniforms                              */
  SLV_VS_SSM_DRAW = 6,
  SLV_VS_JIT = 255          /* a SASL vertex shader compiled at run time; select it with SLV_PROGRAM_JIT(module) */
};
enum {
  SLV_PS_ATTR0_COLOR = 1,   /* color0 = attr0                                    no uniforms        */
  SLV_PS_LIGHTS3 = 2,       /* ColorizedTriangle.cpp:55-92                       no uniforms        */
  /* color0 = tex2d(sampler0, attr[reg]); color0.a = alpha (TextureAndBlending.cpp:96-166)
   * uniforms: u32 reg; f32 alpha                                                                   */
  SLV_PS_TEX_ALPHA = 3,
  /* diffuse texture × clamp(N·L) (Sponza.cpp:99-141).  uniforms: u32 has_sampler                   */
  SLV_PS_SPONZA = 4,
  /* color0 = sample_2d_grad(sampler0, attr[reg].xy, ddx, ddy, 0); a = alpha — the SASL tex2D path,
   * required for anisotropic filtering.  uniforms: u32 reg; f32 alpha; u32 sasl_derivatives */
  SLV_PS_TEX_GRAD_ALPHA = 5,
  SLV_PS_DISCARD_ALL = 6,   /* returns false for every pixel (early-Z quirk probe, App. B #3)       */
  SLV_PS_HEIGHT_COLOR = 7,  /* colour ramp over attr0.x (VertexTextureFetch.cpp:70-113)          no uniforms */
  /* draw_cpp_ps of samples/StandardShadowMap/StandardShadowMap.cpp:62-142 — TWO samplers: samplers[0] = TexSampler (diffuse,
   * tex2d of attr0, optional), samplers[1] = DepthSampler (the shadow map, nine tex2dlod taps at +-1/512 around the light-space
   * position attr4.xyz / attr4.w, exponential shadow map with esm_constant 25000 and the sample's Gaussian weights);
   * colour = tex · (ambient + (diffuse · clamp(L·N) + specular · pow(clamp(−reflect(L, N)·E), shininess)) · occlusion), a = 1.
   * exp / log evaluate in float (expf / logf), pow in double (pow(float, int) promotes).  uniforms: slv_ps_ssm_draw_uniforms */
  SLV_PS_SSM_DRAW = 8,
  /* SLV_PS_SPONZA with the diffuse texture fetched the way a SASL pixel shader's tex2D fetches it: sample_2d_grad with the
   * quad's derivatives of attr0.xy (sasl/src/codegen/cg_impl.cpp:902-909 -> sampler_api.cpp:13-18) - the only path on which the
   * reference filters anisotropically.  sasl_derivatives = 1: the SASL convention, ddx per quad ROW and
   * ddy per quad COLUMN (cgs_simd.cpp:275-313); 0: the cpp_pixel_shader one, q1 - q0 / q2 - q0 for the whole quad
   * (cpp_pixel_shader.cpp:13-19).  The twin of the SASL Sponza pixel shader bench.py compiles at run time.
   * uniforms: slv_ps_sponza_grad_uniforms */
  SLV_PS_SPONZA_GRAD = 9,
  SLV_PS_JIT = 255          /* a SASL pixel shader compiled at run time; select it with SLV_PROGRAM_JIT(module) */
};
enum {
  SLV_BS_REPLACE = 1,        /* inout.color(0,s) = src              (ColorizedTriangle.cpp:94-106)  */
  SLV_BS_LERP_SRC_ALPHA = 2, /* dst + (src − dst)·src.a             (TextureAndBlending.cpp:168-180)*/
  SLV_BS_REPLACE_AND_COUNT = 3 /* REPLACE on target 0, and target1(s) = target1(s) + 1 (coverage probe,
                                  SURVEY §8c "Extracting coverage from the oracle")                 */
};

/* uniform blocks of the programs above (mat44 = eflib::mat44, row-major data_[row][col]) */
typedef struct slv_vs_mvp_passthrough_uniforms { float wvp[16]; uint32_t n_attrs; uint32_t src[5]; } slv_vs_mvp_passthrough_uniforms;
typedef struct slv_vs_plane_xz_uniforms { float wvp[16]; } slv_vs_plane_xz_uniforms;
typedef struct slv_vs_lights3_uniforms { float wvp[16]; float light_pos[3][4]; } slv_vs_lights3_uniforms;
typedef struct slv_vs_sponza_uniforms { float wvp[16]; float light_pos[4]; float eye_pos[4]; } slv_vs_sponza_uniforms;
typedef struct slv_vs_terrain_vtf_uniforms { float wvp[16]; float offset[2]; float scale[2]; } slv_vs_terrain_vtf_uniforms;
/* sasl_derivatives (SLV_PS_TEX_GRAD_ALPHA only): 1 = ddx per quad row / ddy per quad column, the SASL convention
 * (sasl/src/codegen/cgs_simd.cpp:275-313); 0 = q1 - q0 / q2 - q0 for the whole quad (cpp_pixel_shader.cpp:13-19) */
typedef struct slv_ps_tex_alpha_uniforms { uint32_t reg; float alpha; uint32_t sasl_derivatives; } slv_ps_tex_alpha_uniforms;
typedef struct slv_ps_sponza_uniforms { uint32_t has_sampler; } slv_ps_sponza_uniforms;
typedef struct slv_ps_sponza_grad_uniforms { uint32_t has_sampler; uint32_t sasl_derivatives; } slv_ps_sponza_grad_uniforms;
typedef struct slv_vs_ssm_draw_uniforms { float camera_wvp[16]; float light_wvp[16]; float light_pos[4]; float camera_pos[4]; } slv_vs_ssm_draw_uniforms;
typedef struct slv_ps_ssm_draw_uniforms {
  float ambient[4], diffuse[4], specular[4];
  int32_t shininess;
  uint32_t has_tex_sampler;    /* samplers[0] bound (else the texture colour is white)    */
  uint32_t has_depth_sampler;  /* samplers[1] bound (else occlusion = 0: ambient only)    */
} slv_ps_ssm_draw_uniforms;

/* ---- handles ---------------------------------------------------------------------------------- */
typedef struct slv_device_t* slv_device;
typedef uint32_t slv_handle;  /* 0 = null. buffers, textures and samplers share one id space        */

/* ---- descriptors ------------------------------------------------------------------------------ */
typedef struct slv_sampler_desc {      /* sampler_desc (sampler.h:16-45) */
  uint32_t min_filter, mag_filter, mip_filter;
  uint32_t mip_qual;
  uint32_t addr_mode_u, addr_mode_v, addr_mode_w;
  float mip_lod_bias;
  uint32_t max_anisotropy;
  uint32_t comparison_func;
  float border_color[4];
  float min_lod, max_lod;
} slv_sampler_desc;

typedef struct slv_viewport {          /* viewport (viewport.h:5-12) */
  float x, y, w, h, minz, maxz;
} slv_viewport;

typedef struct slv_stencil_op_desc {   /* depth_stencil_op_desc (framebuffer.h:20-32) */
  uint32_t stencil_fail_op, stencil_depth_fail_op, stencil_pass_op, stencil_func;
} slv_stencil_op_desc;

typedef struct slv_depth_stencil_desc { /* depth_stencil_desc (framebuffer.h:34-50) */
  uint32_t depth_enable, depth_write_mask, depth_func;
  uint32_t stencil_enable, stencil_read_mask, stencil_write_mask;
  slv_stencil_op_desc front_face, back_face;
} slv_depth_stencil_desc;

typedef struct slv_raster_desc {       /* the fields of raster_desc that are read (raster_state.h:17-40) */
  uint32_t cull_mode;
  uint32_t front_ccw;
} slv_raster_desc;

typedef struct slv_vertex_stream {     /* set_vertex_buffers(slot, buf, stride, offset) */
  slv_handle buffer;
  uint32_t stride;
  uint32_t offset;
} slv_vertex_stream;

typedef struct slv_input_element {     /* input_element_desc resolved against the VS register map
                                          (stream_assembler.cpp:52-86): register <- slot/offset/fmt */
  uint32_t reg;                        /* vs_input register index 0..7                              */
  uint32_t format;                     /* SLV_FMT_*                                                 */
  uint32_t slot;                       /* index into slv_draw_desc.streams                          */
  uint32_t aligned_byte_offset;
  float default_w;                     /* 1 for POSITION semantics, else 0 (shader/constants.h:119) */
} slv_input_element;

typedef struct slv_shader_binding {
  uint32_t program;                    /* SLV_VS_* / SLV_PS_* / SLV_BS_*                            */
  uint32_t uniform_bytes;
  uint8_t uniforms[SLV_MAX_UNIFORM_BYTES];
  slv_handle samplers[SLV_MAX_SAMPLERS];
} slv_shader_binding;

typedef struct slv_draw_desc {
  /* input assembler */
  uint32_t n_streams;
  slv_vertex_stream streams[8];
  uint32_t n_elements;
  slv_input_element elements[SLV_MAX_VS_INPUT_ATTRS];
  slv_handle index_buffer;             /* 0 for draw(), else draw_index()                           */
  uint32_t index_format;               /* SLV_INDEX_*                                               */
  uint32_t topology;                   /* SLV_TOPO_*                                                */
  uint32_t start;                      /* start index (draw_index) / start vertex (draw)            */
  uint32_t prim_count;
  int32_t base_vertex;
  /* shaders */
  slv_shader_binding vs, ps, bs;
  uint32_t vs_attr_modifiers[SLV_MAX_VS_OUTPUT_ATTRS]; /* output_attribute_modifiers(i); 0 => linear */
  /* fixed-function state */
  slv_raster_desc raster;
  slv_depth_stencil_desc ds;
  int32_t stencil_ref;
  slv_viewport viewport;
  /* output merger targets: surfaces = mip 0 of textures */
  uint32_t n_color_targets;
  slv_handle color_targets[SLV_MAX_RENDER_TARGETS];
  slv_handle ds_target;
} slv_draw_desc;

typedef struct slv_pipeline_statistics { /* pipeline_statistics + internal_statistics (async_object.h:90-175) */
  uint64_t ia_vertices, ia_primitives, vs_invocations, gs_invocations, gs_primitives;
  uint64_t cinvocations, cprimitives, ps_invocations;
  uint64_t backend_input_pixels;
} slv_pipeline_statistics;

typedef struct slv_pipeline_profiles {   /* pipeline_profiles, nanoseconds (async_object.h:177-233) */
  uint64_t gather_vtx, vtx_proc, clipping, compact_clip, vp_trans, tri_dispatch, ras;
} slv_pipeline_profiles;

/* ---- entry points ----------------------------------------------------------------------------- */
/* (not visible to run-time device compilation: NVRTC translation units include this header for the POD types only) */
#ifndef __CUDACC_RTC__
/* replaces create_software_renderer()/create_benchmark_renderer() (renderer.h:133-134).
 * `ordinal` = CUDA device index (ignored by the CPU checkers).  Product: fails with SLV_FAILED when
 * no CUDA device is usable — there is no CPU fallback. */
slv_result slv_device_create(int32_t ordinal, slv_device* out);
void slv_device_destroy(slv_device dev);
/* name of the implementation: "cuda-sm100a" | "oracle" | "reference" */
const char* slv_backend_name(void);
uint32_t slv_abi_version(void);

/* renderer::create_buffer + map(write)/unmap (renderer.h:45,60-62; buffer.h:16) */
slv_result slv_buffer_create(slv_device dev, size_t bytes, slv_handle* out);
slv_result slv_buffer_upload(slv_device dev, slv_handle buf, size_t offset, const void* src, size_t bytes);
slv_result slv_buffer_readback(slv_device dev, slv_handle buf, size_t offset, void* dst, size_t bytes);

/* renderer::create_tex2d (renderer.h:46-47); level 0 is also the render-target `surface`.
 * Storage layout of a level is the reference's linear one: ((y*W + x)*S + s)*bpp (surface.cpp:277-295). */
slv_result slv_texture_create(slv_device dev, uint32_t width, uint32_t height, uint32_t samples,
                              uint32_t pixel_format, slv_handle* out);
/* texture_2d::gen_mipmap(filter, auto_gen=true) (texture2d.cpp:25-36, surface.cpp:53-92) */
slv_result slv_texture_gen_mipmap(slv_device dev, slv_handle tex, uint32_t filter);
slv_result slv_texture_level_count(slv_device dev, slv_handle tex, uint32_t* out_levels);
slv_result slv_texture_level_size(slv_device dev, slv_handle tex, uint32_t level, uint32_t* w, uint32_t* h);
/* map(surface, map_write) + memcpy / map(surface, map_read) (resource_manager.cpp:8-54) */
slv_result slv_texture_upload(slv_device dev, slv_handle tex, uint32_t level, const void* src, size_t bytes);
slv_result slv_texture_readback(slv_device dev, slv_handle tex, uint32_t level, void* dst, size_t bytes);
/* the same copy without blocking the caller (the async_renderer keeps the application running while a frame is in flight,
 * async_renderer.cpp:42-74): enqueued on a copy stream behind everything submitted so far; `dst` (pinned host memory for a
 * truly asynchronous copy) is valid after slv_readback_wait or slv_flush.  The next command that WRITES the texture waits for
 * the copy on the device, so alternating two textures overlaps the readback of frame k with the rendering of frame k+1.
 * The CPU checkers copy synchronously. */
slv_result slv_texture_readback_async(slv_device dev, slv_handle tex, uint32_t level, void* dst, size_t bytes);
slv_result slv_readback_wait(slv_device dev);
/* device-side: orders everything submitted after this call behind the pending asynchronous readback of `tex` (no host wait).
 * The sort-first root uses it before it tells the other ranks that a frame buffer may be overwritten. */
slv_result slv_readback_fence(slv_device dev, slv_handle tex);
/* External writers of a buffer on a stream of their own - e.g. the ranks of a sort-first run uploading a SLICE of the frame's vertex /
 * index data each and assembling the whole buffer with an all-gather over NVLink, so that every input byte crosses a host link
 * once instead of once per GPU.  slv_buffer_device_ptr returns the allocation.  slv_external_write_begin flushes the queued draws
 * and makes `cuda_stream` (a cudaStream_t) wait until the library's readers of buffers enqueued so far - the geometry passes of
 * the flushed batches - are done, except those of the `skip_latest` most recent batches (a writer that alternates between two
 * buffer sets passes 1: the latest batch reads the other set); slv_external_write_end makes the library's next geometry passes
 * (and its render stream) wait for everything enqueued on `cuda_stream` so far.  The CPU checkers return the host pointer and
 * ignore the ordering calls. */
slv_result slv_buffer_device_ptr(slv_device dev, slv_handle buf, void** out, size_t* bytes);
slv_result slv_external_write_begin(slv_device dev, void* cuda_stream, uint32_t skip_latest);
slv_result slv_external_write_end(slv_device dev, void* cuda_stream);
/* Sort-first frame assembly on the HOST (multi-GPU end to end): every rank writes the 64x64 tiles it owns (slv_set_tile_shard) of
 * its single-sampled resolved surface `tex` straight into ONE host frame of the texture's linear layout, shared by the ranks
 * (e.g. POSIX shared memory mapped by every process), over ITS OWN host link - so the device -> host traffic of a frame spreads
 * over N links instead of crossing rank 0's.  `host_frame` must have been registered with slv_host_register by this process
 * (page-locked + mapped: the kernel stores whole 256-byte tile rows over PCIe).  Asynchronous like slv_texture_readback_async:
 * enqueued on the copy stream behind everything submitted so far, complete after slv_readback_wait / slv_flush; the next writer
 * of `tex` waits for it on the device.  The CPU checkers copy synchronously. */
slv_result slv_host_register(slv_device dev, void* ptr, size_t bytes);
slv_result slv_host_unregister(slv_device dev, void* ptr);
slv_result slv_texture_export_tiles_async(slv_device dev, slv_handle tex, void* host_frame, size_t bytes);
/* renderer::create_sampler (renderer.h:50) */
slv_result slv_sampler_create(slv_device dev, const slv_sampler_desc* desc, slv_handle tex, slv_handle* out);
/* SASL shaders compiled at run time.  The reference's compile(code, profile) + set_vertex_shader_code /
 * set_pixel_shader_code (salvia/include/salvia/core/renderer.h:75-82,136-147) JIT the shader to host code with LLVM; here
 * salviarenderer_b200/sasl lowers it to device code and the CUDA toolchain (NVVM's NVPTX backend) produces a cubin for
 * sm_100a holding the pipeline kernel with the shader inlined (k_geometry for a vertex shader, k_raster for a pixel
 * shader).  slv_shader_module_load hands that image to the device; a draw selects it with
 * vs.program / ps.program = SLV_PROGRAM_JIT(module) and passes the shader's globals, packed as the reflection says, in
 * `uniforms` (<= SLV_MAX_UNIFORM_BYTES).  n_vs_output_attrs: the vertex shader's outputs besides SV_Position.
 * Released with slv_resource_release.  The CPU checkers return SLV_FAILED. */
#define SLV_STAGE_VS 0u
#define SLV_STAGE_PS 1u
#define SLV_PROGRAM_JIT(module) (0x80000000u | (uint32_t)(module))
slv_result slv_shader_module_load(slv_device dev, uint32_t stage, const void* image, size_t bytes, uint32_t n_vs_output_attrs,
                                  slv_handle* out);
/* compile(code, profile) (renderer.h:136-147) IN PROCESS: `device_code` is what the SASL front end generates for one shader
 * (salviarenderer_b200/sasl: the scalarised body of slv_jit_vs / slv_jit_ps); the library compiles it with NVRTC together with
 * its own pipeline-kernel sources (embedded in the library at build time) for sm_100a under the library's numerics flags, so the
 * shader is inlined into k_geometry / k_vertex_shade or k_raster / k_shade exactly like a built-in program.  No nvcc, no
 * subprocess, no GPU needed to compile.  slv_shader_compile_cubin returns the image (malloc'ed: slv_free) - the same bytes
 * slv_shader_module_load takes; slv_shader_compile = compile + load.  flags: SLV_COMPILE_DERIV_CPP selects the cpp-shader
 * derivative convention (ddx = q1 - q0, ddy = q2 - q0 for the whole quad) instead of SASL's per row / per column.  `log`
 * (optional) receives the compiler's diagnostics, NUL-terminated.  Compiled images are cached on disk by content hash in
 * $SLV_JIT_CACHE (else $XDG_CACHE_HOME/salvia_b200_jit, else ~/.cache/salvia_b200_jit; only a directory the user owns and nobody
 * else can write).  SLV_FAILED when libnvrtc is unavailable or the code does not compile.  CPU checkers: SLV_FAILED - except
 * slv_shader_compile of the restatement (oracle/), which builds the same code for the HOST and runs it inside its pipeline, so
 * that tests compare SASL frames on the CPU (oracle/slv_host_shader.h; test infrastructure). */
#define SLV_COMPILE_DERIV_CPP 1u
slv_result slv_shader_compile_cubin(uint32_t stage, const char* device_code, uint32_t n_vs_output_attrs, uint32_t flags,
                                    void** image, size_t* bytes, char* log, size_t log_bytes);
slv_result slv_shader_compile(slv_device dev, uint32_t stage, const char* device_code, uint32_t n_vs_output_attrs, uint32_t flags,
                              slv_handle* out, char* log, size_t log_bytes);
void slv_free(void* p);
/* The first half of compile(code, profile) (renderer.h:136-147; the reference's sasl library, sasl/src/{parser,semantic,codegen}):
 * the SASL front end, host only - no device, no interpreter.  `source` is SASL text, `entry` the entry function (NULL: the function
 * whose parameters or return value carry semantics).  On success `*unit` (malloc'ed: slv_free; NUL-terminated, `*unit_bytes`
 * without the NUL) holds the compiled unit as text: the line "SLVSASL 1", then one line each of
 *   stage vs|ps, n_vs_output_attrs N, uniform_bytes N, uses_derivatives 0|1,
 *   uniform NAME TYPE OFFSET SIZE   (layout of the uniform block a draw passes; TYPE "float4x4[]" = address of an array's buffer),
 *   sampler SLOT NAME, input SEMANTIC INDEX K (VS: input register K; PS: attribute K), output SEMANTIC INDEX K (VS: attribute K),
 * then "code NBYTES" and NBYTES of device code - what slv_shader_compile / slv_shader_compile_cubin take as `device_code`.
 * SLV_FAILED when the source does not compile; `log` (optional) receives the diagnostic ("line N: ...").  Implemented by
 * salviarenderer_b200/host/sasl_frontend.hpp, which C++ hosts may also include directly.  CPU checkers: SLV_FAILED. */
slv_result slv_sasl_translate(uint32_t stage, const char* source, const char* entry, char** unit, size_t* unit_bytes, char* log,
                              size_t log_bytes);
slv_result slv_resource_release(slv_device dev, slv_handle h);

/* renderer::draw / draw_index -> commit_state_and_command() (renderer_impl.cpp:337-353) */
slv_result slv_draw(slv_device dev, const slv_draw_desc* desc);
/* renderer::clear_color (render_core.cpp:96-99) */
slv_result slv_clear_color(slv_device dev, slv_handle surface_tex, const float rgba[4]);
/* renderer::clear_depth_stencil (render_core.cpp:101-111, framebuffer.cpp:616-644) */
slv_result slv_clear_depth_stencil(slv_device dev, slv_handle surface_tex, uint32_t flags, float depth,
                                   uint32_t stencil);
/* surface::resolve (surface.cpp:123-140) — src multi-sampled, dst single-sampled */
slv_result slv_resolve(slv_device dev, slv_handle src_tex, slv_handle dst_tex);
/* renderer::flush (renderer.h:128): returns when every queued command has finished */
slv_result slv_flush(slv_device dev);

/* queries: begin/end/get_data of pipeline_statistics + internal_statistics (renderer.h:116-119).
 * begin zeroes the counters; get flushes and copies them. */
slv_result slv_query_begin(slv_device dev);
slv_result slv_query_get(slv_device dev, slv_pipeline_statistics* out);

/* per-stage time accumulated since slv_query_begin (pipeline_profiles query, rasterizer.cpp:1128-1195).
 * Product: CUDA-event time of the kernels of each stage; only collected while SLV_PROFILE=1 is set in
 * the environment at device creation (it serialises the stream), otherwise all zeros. */
slv_result slv_profile_get(slv_device dev, slv_pipeline_profiles* out);

/* sort-first multi-GPU split (new; the reference is single-process): this device renders only the
 * 64x64 screen tiles t=(tx,ty) with (tx + 3*ty) % nranks == rank; geometry stages run for all
 * primitives.  (0,1) = everything (default).  Not supported by the reference backend. */
slv_result slv_set_tile_shard(slv_device dev, uint32_t rank, uint32_t nranks);

/* ---- measurement and multi-GPU plumbing (no reference twin; used by bench.py and the sharded path) ------ */
/* Algorithmic framebuffer traffic counters since slv_query_begin (B_frag): samples depth-tested,
 * samples whose depth/stencil was written, samples blended into colour, samples whose destination colour
 * was read by the blend shader. Exact integers; the oracle produces the same numbers. */
typedef struct slv_traffic_counters {
  uint64_t z_tested, z_written, c_written, c_read;
  /* work counters of the product's raster kernel (0 for the CPU checkers): tile-list entries examined by the
   * region filters, (triangle, 16x16 region) survivors, (triangle, warp) pairs walked, quads shaded */
  uint64_t list_entries_scanned, region_survivors, warp_pairs, quads_shaded;
  /* pixel-shader executions the product actually performed (the visibility-first path shades a pixel once per
   * distinct final owner, so this is <= ps_invocations, which still counts what the reference would shade) */
  uint64_t ps_executed;
} slv_traffic_counters;
slv_result slv_traffic_get(slv_device dev, slv_traffic_counters* out);
/* B_tex of SURVEY 8d counts the texture bytes a frame TOUCHES: with tracking on, every sampler call of the draws issued
 * afterwards records the mip level(s) it reads; slv_texture_levels_touched returns the bit mask (bit l = level l) accumulated
 * since tracking was last switched on (switching it on again resets every mask).  Measurement aid: tracking adds one atomic per
 * sampler call, so it is off in timed regions.  CPU checkers: SLV_OK and an empty mask. */
slv_result slv_texture_level_tracking(slv_device dev, uint32_t on);
slv_result slv_texture_levels_touched(slv_device dev, slv_handle tex, uint32_t* mask);
/* number of kernels this library launched since slv_query_begin (0 for the CPU checkers) */
slv_result slv_kernel_launch_count(slv_device dev, uint64_t* out);
/* CUDA events on the stream the kernels are launched on: record event `slot` (0..15) now; elapsed
 * milliseconds between two recorded slots (synchronises on `b`).  CPU checkers use a steady clock. */
slv_result slv_event_record(slv_device dev, uint32_t slot);
slv_result slv_event_elapsed_ms(slv_device dev, uint32_t a, uint32_t b, float* ms);
/* make the library enqueue all its work on a caller-owned CUDA stream (e.g. torch's current stream) so that
 * host plumbing such as NCCL collectives orders against the kernels on the device, without host syncs.
 * `cuda_stream` is a cudaStream_t; NULL restores the device's own stream.  CPU checkers ignore it. */
slv_result slv_set_stream(slv_device dev, void* cuda_stream);
/* per-kernel-stage event timing on/off at run time (same data as SLV_PROFILE=1; read with slv_profile_get) */
slv_result slv_profile_enable(slv_device dev, uint32_t on);
/* the same event times split by kernel stage of the product, milliseconds since slv_query_begin:
 * ms[0] k_geometry, [1] k_scan_tiles + k_bin_fill, [2] k_sort_lists, [3] k_raster or k_cover, [4] k_shade, [5] k_region_bin.  n >= 6.
 * CPU checkers return zeros. */
slv_result slv_profile_get_stages(slv_device dev, double* ms, uint32_t n);
/* raw device address of a texture level (product only; CPU checkers return their host address) so the
 * host plumbing (torch.distributed / NCCL) can exchange it without a copy */
slv_result slv_texture_device_ptr(slv_device dev, slv_handle tex, uint32_t level, void** out, size_t* bytes);
/* sort-first gather helpers: copy the 64x64 tiles owned by (rank, nranks) — see slv_set_tile_shard — of a
 * single-sampled surface into / out of a dense staging buffer (tiles in row-major tile order, each tile
 * 64 rows x 64 texels, clipped tiles padded).  `staging` is a DEVICE pointer for the product, a host pointer
 * for the CPU checkers; *bytes receives the packed size (call with staging == NULL to query it). */
slv_result slv_pack_tiles(slv_device dev, slv_handle tex, uint32_t rank, uint32_t nranks, void* staging, size_t* bytes);
slv_result slv_unpack_tiles(slv_device dev, slv_handle tex, uint32_t rank, uint32_t nranks, const void* staging);

/* sort-first frame assembly over peer memory (NVLink / NVSwitch; new, the reference is single-process).  One process
 * per GPU: the root rank exports its resolved surface and its flag block as CUDA IPC handles, every other rank opens
 * them and redirects slv_resolve into the root's surface (only its own tiles are written, see slv_set_tile_shard), so
 * the MSAA resolve and the gather of the finished tiles are ONE kernel writing over NVLink.  Flags order the ranks'
 * streams on the device: slv_peer_signal raises flags[index] = value in `peer_flags` (NULL = this device's own block)
 * after all earlier work of this device's stream; slv_flags_wait blocks this device's stream until
 * flags[first .. first+count) of `flags` (NULL = this device's own block, else an opened peer block polled over
 * NVLink) have all reached `value` (wrap-safe comparison; gives up after ~10 s and makes the next flush point fail).  The CPU checkers return SLV_FAILED from all of these. */
#define SLV_PEER_HANDLE_BYTES 64
#define SLV_PEER_FLAGS 64
slv_result slv_peer_export_texture(slv_device dev, slv_handle tex, uint32_t level, uint8_t handle_out[SLV_PEER_HANDLE_BYTES]);
slv_result slv_peer_export_flags(slv_device dev, uint8_t handle_out[SLV_PEER_HANDLE_BYTES]);
slv_result slv_peer_open(slv_device dev, const uint8_t handle[SLV_PEER_HANDLE_BYTES], void** dptr_out);
slv_result slv_peer_close(slv_device dev, void* dptr);
slv_result slv_resolve_target_peer(slv_device dev, slv_handle dst, void* peer_surface);
slv_result slv_peer_signal(slv_device dev, void* peer_flags, uint32_t index, uint32_t value);
slv_result slv_flags_wait(slv_device dev, const void* flags, uint32_t first, uint32_t count, uint32_t value);
/* The ROOT's side of the same protocol without stalling its render stream.  slv_assembly_wait: the wait for the other ranks'
 * tiles of the frame being assembled in `tex` is enqueued on the copy stream (behind this device's own work on `tex`); what
 * consumes the assembled frame orders after it - slv_texture_readback_async / slv_texture_export_tiles_async by stream order, a
 * draw that samples `tex`, a synchronous readback and the next WRITER of `tex` through the texture's event - while the next
 * frame's kernels, which touch another buffer, start at once.  slv_peer_signal_after_consumers: raises the flag once the
 * copy-stream work enqueued so far ON `tex` (its assembly wait, its readback) is done - on a stream of its own, so the release of
 * one frame buffer never waits for the assembly of another. */
slv_result slv_assembly_wait(slv_device dev, slv_handle tex, const void* flags, uint32_t first, uint32_t count, uint32_t value);
slv_result slv_peer_signal_after_consumers(slv_device dev, slv_handle tex, void* peer_flags, uint32_t index, uint32_t value);

/* sampler probe used by the sampler parity tests: evaluates sampler::sample_2d_grad
 * (sampler.cpp:854-873) [use_lod = 0] or sample_2d_lod (:850-852) [use_lod = 1] for n coordinates.
 * coords: n×2 floats; ddx, ddy: n×2 floats (ignored for lod); lod: n floats; out: n×4 floats.
 * All pointers are HOST pointers. */
slv_result slv_sampler_probe(slv_device dev, slv_handle sampler, uint32_t n, const float* coords,
                             const float* ddx, const float* ddy, const float* lod, uint32_t use_lod,
                             float* out_rgba);

#endif /* !__CUDACC_RTC__ */
#ifdef __cplusplus
}
#endif
#endif /* SALVIA_B200_H */
