/*
 * vkrt_b200.h -- C ABI of the B200-native wavefront path tracer that replaces the
 * render back-ends of nvpro-samples/vk_raytrace.
 *
 * Every entry point names the reference interface it stands in for
 * (paths relative to the reference repository).  The library is a drop-in for
 * the hot path behind `class Renderer` (src/renderer.h:30-48): a host
 * application keeps its own scene loader, hands the library the *same* buffers
 * the reference uploads to Vulkan (shaders/host_device.h layouts, restated in
 * vkrt_host_device.h) and calls vkrt_render() where it used to call
 * Renderer::run().
 *
 * Conventions: plain pointers + sizes, no C++/torch types; every function
 * returns 0 on success or a VKRT_ERR_* code (vkrt_last_error() gives the text);
 * input host memory is copied during the call and stays owned by the caller;
 * one context per GPU, calls into one context must be externally serialised
 * (same rule as the reference's renderers, src/sample_example.cpp:119-162).
 */
#ifndef VKRT_B200_H
#define VKRT_B200_H

#include <stddef.h>
#include <stdint.h>
#include "vkrt_host_device.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VKRT_OK 0
#define VKRT_ERR_INVALID 1 /* bad argument */
#define VKRT_ERR_CUDA 2    /* CUDA runtime failure */
#define VKRT_ERR_STATE 3   /* call out of order (e.g. render before commit) */
#define VKRT_ERR_NOMEM 4

typedef struct vkrt_ctx vkrt_ctx;

/* RNG seeding flavour of the two reference back-ends. */
#define VKRT_SEED_RAYQUERY 0 /* shaders/pathtrace.comp:97  tea(size.x*y+x, frame*maxSamples) */
#define VKRT_SEED_RTX 1      /* shaders/pathtrace.rgen:72  initRandom(size, xy, frame)        */

/* BVH builder selection (the reference delegates this to the Vulkan driver,
 * src/accelstruct.cpp:110-162). */
#define VKRT_BUILD_LBVH 0       /* GPU Morton/Karras LBVH (default) */
#define VKRT_BUILD_LBVH_COLLAPSED 1 /* LBVH + bottom-up SAH leaf collapse */
#define VKRT_BUILD_SAH_HOST 2       /* GPU flatten + host binned-SAH BVH2 (slow build, fast trace) */
#define VKRT_BUILD_PLOC 3           /* GPU: Morton sort + parallel locally-ordered clustering + SAH leaf collapse (SAH-like quality) */
#define VKRT_BUILD_SAH_GPU 4        /* GPU: breadth-first binned SAH (32 bins, triangle-parallel) as path-code keys + radix tree + SAH leaf collapse:
                                       the PREFER_FAST_TRACE builder (src/accelstruct.cpp:125), a few ms for 260 k triangles */
#define VKRT_BUILD_MODE_MASK 0xff
/* OR-ed into build_mode: which tree the traversal kernels walk for this scene, overriding the size rule
 * (binary while the BVH fits the L2, 4-wide above 1.5 M triangles) and the VKRT_BVH_WIDTH environment switch. */
#define VKRT_BUILD_FORCE_BVH2 0x100
#define VKRT_BUILD_FORCE_BVH4 0x200

typedef struct vkrt_render_opts
{
  int32_t seed_mode;        /* VKRT_SEED_*                                             */
  int32_t frames_in_flight; /* frames batched into one wavefront (0 = auto)            */
  int32_t tile_size;        /* edge of the sharding tiles in pixels (0 = 64)           */
  int32_t shard_rank;       /* this context renders tiles t with t % shard_count == shard_rank */
  int32_t shard_count;      /* 0 or 1 = whole image                                    */
  int32_t count_traversal;  /* 1 = run the instrumented traversal kernels (visit counters) */
  int32_t sort_rays;        /* 1 = order each bounce's compacted ray queue by direction octant + origin Morton bucket before tracing it
                               (hand-written device-side bucket sort, no host synchronisation; same image).  Default 0: measured
                               neutral to slightly slower on the bench scenes, see profiles/README.md */
  int32_t profile;          /* 1 = time every kernel family with CUDA events (vkrt_stats.ms_*) */
  int32_t precise_shading;  /* 0 (default): the shade kernel uses approximate division / sqrt and SFU sin/cos/exp/log/pow, as a GLSL
                               compiler does; 1: IEEE-exact arithmetic throughout (what the oracle comparisons at 1e-4 use).
                               Ray generation, traversal and intersection are exact in both. */
  int32_t tail_from;        /* bounce from which one fused kernel (trace + shade + shadow per path) finishes the few paths still alive:
                               0 = automatic (bounce 3 for small wavefronts: interactive frames, thin shards), -1 = never, n > 0 = from
                               bounce n.  Same image in every setting. */
  int32_t reserved[6];
} vkrt_render_opts;

typedef struct vkrt_stats
{
  uint64_t closest_rays;    /* ClosestHit() calls (shaders/traceray_rq.glsl:108)  */
  uint64_t shadow_rays;     /* AnyHit() calls     (shaders/traceray_rq.glsl:153)  */
  uint64_t nodes_visited;   /* BVH2 nodes fetched by both trace kernels (instrumented kernels only) */
  uint64_t tris_tested;     /* triangles fetched by both trace kernels (instrumented kernels only)  */
  uint64_t frames_rendered; /* frames since the last vkrt_reset_stats             */
  uint64_t kernel_launches; /* kernels of this library launched                   */
  double   ms_raygen;       /* device time per kernel family, CUDA events, summed */
  double   ms_trace_closest;
  double   ms_trace_shadow;
  double   ms_shade;
  double   ms_sort;
  double   ms_accumulate;
  double   ms_tonemap;
  double   ms_build;        /* last vkrt_scene_commit: the acceleration structure only (flatten, sort, hierarchy, refit, emit, node formats; host SAH included when selected) */
  uint32_t num_triangles;   /* world-space triangles in the BVH                   */
  uint32_t num_bvh_nodes;   /* BVH2 nodes (node_bytes each as read by the trace kernels) */
  uint32_t num_instances;   /* TLAS-equivalent nodes (src/accelstruct.cpp:132)    */
  uint32_t has_non_opaque;  /* any instance that is not FORCE_OPAQUE              */
  uint64_t shadow_nodes_visited; /* the any-hit kernel's share of nodes_visited   */
  uint64_t shadow_tris_tested;   /* the any-hit kernel's share of tris_tested     */
  uint32_t bvh_max_depth;   /* depth of the BVH2 (traversal stack bound)          */
  uint32_t node_bytes;      /* bytes per node fetched by the trace kernels: 32 (16-bit grid boxes, default) or 64 (fp32) */
  uint32_t bvh_width;       /* children per node of the active traversal kernel: 2 (default) or 4 */
  uint32_t _pad;
  /* bounce classes (opts.profile = 1): device time of the depth-0 launches alone.  Primary rays are coherent, everything
   * after the first hit is not; closest_rays at depth 0 = pixels x frames, shadow rays at depth 0 are counted here. */
  double   ms_trace_closest_primary;
  double   ms_trace_shadow_primary;
  uint64_t shadow_rays_primary;
  double   ms_tail;         /* fused tail kernel: the late bounces of small wavefronts in one launch (trace + shade + shadow) */
} vkrt_stats;

/* ---- context ---------------------------------------------------------- */

/* Renderer::setup (src/renderer.h:33-36): binds the renderer to a device. */
int vkrt_create(int cuda_device, vkrt_ctx** out_ctx);
/* Renderer::destroy (src/renderer.h:37). */
int vkrt_destroy(vkrt_ctx* ctx);
/* Renderer::name (src/renderer.h:43). */
const char* vkrt_name(void);
const char* vkrt_last_error(vkrt_ctx* ctx);

/* ---- scene upload: replaces Scene::create*Buffer + AccelStructure::create --- */

/* Scene::load start (src/scene.cpp:56-58): drops the previous scene. */
int vkrt_scene_begin(vkrt_ctx* ctx);
/* Scene::createVertexBuffer, one call per primMesh (src/scene.cpp:190-274) and the
 * matching InstanceData entry (src/scene.cpp:161-176).  `indices` are uint32 triples. */
int vkrt_scene_add_primmesh(vkrt_ctx* ctx, const VertexAttributes* vertices, uint32_t num_vertices,
                            const uint32_t* indices, uint32_t num_indices, int32_t material_index);
/* One TLAS instance (src/accelstruct.cpp:132-157): column-major 4x4 world matrix
 * (glm layout) and the primMesh it instantiates. */
int vkrt_scene_add_node(vkrt_ctx* ctx, const float world_matrix[16], uint32_t prim_mesh);
/* Scene::createMaterialBuffer (src/scene.cpp:339-382). */
int vkrt_scene_set_materials(vkrt_ctx* ctx, const GltfShadeMaterial* materials, uint32_t count);
/* Scene::createLightBuffer (src/scene.cpp:304-333).  count may be 0. */
int vkrt_scene_set_lights(vkrt_ctx* ctx, const Light* lights, uint32_t count);
/* Scene::createTextureImages (src/scene.cpp:488-580): RGBA8 image + glTF sampler codes
 * (9728 NEAREST, 9729 LINEAR, 998x mip variants; 33071 CLAMP, 33648 MIRROR, 10497 REPEAT),
 * mapped like gltfSamplerToVulkan (src/scene.cpp:447-482).  A texture without a glTF
 * sampler is LINEAR/REPEAT (src/scene.cpp:562-565): pass 9729/10497. */
int vkrt_scene_add_texture(vkrt_ctx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height,
                           int32_t mag_filter, int32_t min_filter, int32_t wrap_s, int32_t wrap_t);
/* AccelStructure::create (src/accelstruct.cpp:55-65): flattens instances to world
 * space on the GPU and builds the BVH2. */
int vkrt_scene_commit(vkrt_ctx* ctx, int32_t build_mode);

/* ---- environment: replaces HdrSampling (src/hdr_sampling.cpp) ------------- */

/* HdrSampling::createEnvironmentAccel + buildAliasmap (src/hdr_sampling.cpp:107-248),
 * host-side, pure function.  out_accel holds width*height entries. */
int vkrt_build_env_accel(const float* rgba32f, uint32_t width, uint32_t height, EnvAccel* out_accel,
                         float* out_integral, float* out_average);
/* HdrSampling::loadEnvironment upload part (src/hdr_sampling.cpp:56-99). */
int vkrt_set_environment(vkrt_ctx* ctx, const float* rgba32f, uint32_t width, uint32_t height, const EnvAccel* accel);
/* SampleExample::updateSunAndSky UBO (src/sample_example.hpp:176-193). */
int vkrt_set_sunsky(vkrt_ctx* ctx, const SunAndSky* sun_and_sky);
/* Scene::updateCamera UBO upload (src/scene.cpp:629-668). */
int vkrt_set_camera(vkrt_ctx* ctx, const SceneCamera* camera);

/* ---- output image: replaces RenderOutput (src/render_output.cpp) --------- */

/* RenderOutput::createOffscreenRender (src/render_output.cpp:78-113): allocates the
 * RGBA32F accumulation image, or adopts a caller-owned device buffer of
 * width*height*4 floats when device_rgba32f != NULL. */
int vkrt_output_create(vkrt_ctx* ctx, uint32_t width, uint32_t height, void* device_rgba32f);
/* Device pointer of the accumulation image (for NCCL reduce over NVLink). */
int vkrt_output_device_ptr(vkrt_ctx* ctx, void** out_device_ptr);
/* Zero the accumulation image (sharded renders sum disjoint tiles). */
int vkrt_output_clear(vkrt_ctx* ctx);
/* Read back the RGBA32F accumulation image (width*height*4 floats). */
int vkrt_output_read(vkrt_ctx* ctx, float* host_rgba32f);
/* RenderOutput::run + post.frag (src/render_output.cpp:174-182, shaders/post.frag:98-147):
 * tonemap the accumulation image to RGBA8 (host buffer, width*height*4 bytes). */
int vkrt_tonemap(vkrt_ctx* ctx, const Tonemapper* tonemapper, uint8_t* host_rgba8);
/* Same, result left on the device as float RGBA (for tests) or RGBA8. */
int vkrt_tonemap_float(vkrt_ctx* ctx, const Tonemapper* tonemapper, float* host_rgba32f);
/* The display path of an interactive host (src/sample_example.cpp:417-426 renders frame f+1 while frame f is presented):
 * tonemap the image as it is after the frames enqueued so far and copy the RGBA8 result to `host_rgba8` WITHOUT waiting --
 * the copy overlaps the next vkrt_render.  `slot` (0..3) names one of four device staging buffers; the host buffer
 * must stay valid (and should be page-locked: vkrt_host_alloc) until vkrt_tonemap_wait(ctx, slot) returns. */
int vkrt_tonemap_async(vkrt_ctx* ctx, const Tonemapper* tonemapper, uint8_t* host_rgba8, int32_t slot);
int vkrt_tonemap_wait(vkrt_ctx* ctx, int32_t slot);
/* Page-locked host memory for the asynchronous read-back (cudaHostAlloc / cudaFreeHost). */
int vkrt_host_alloc(size_t bytes, void** out_ptr);
int vkrt_host_free(void* ptr);

/* ---- tile shards (one process per GPU): compact exchange instead of a full-image reduce -------- */

/* Number of pixels of the shard (tile_size, shard_rank, shard_count) as vkrt_render_opts names it: tiles are dealt
 * round-robin, t % shard_count == shard_rank. */
int vkrt_output_shard_size(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, uint32_t* out_pixels);
/* image -> compact: the shard's pixels (RGBA32F, shard order) into a device buffer of shard_size float4.  What a rank sends. */
int vkrt_output_pack_shard(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, void* device_compact);
/* compact -> image: scatters a shard received from rank `shard_rank` into this context's image.  What the root does with
 * the gathered buffers; the result is bit-identical to summing zero-padded full images. */
int vkrt_output_unpack_shard(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, const void* device_compact);

/* ---- peer image: the exchange fused into the accumulate kernel (processes on one NVLink / NVSwitch node) --------
 * Instead of pack -> gather -> unpack, every non-root rank's accumulate kernel ALSO stores each pixel it owns into the root's
 * accumulation image, mapped into its address space over CUDA IPC: the shards arrive while the kernel runs, and what is left of
 * the exchange is one small collective as a cross-process fence (vkrt_output_fence on both sides of it).  A rank's local image
 * stays valid (the running mean is read from it), so nothing crosses the link twice.  The reference has no counterpart: it
 * renders on one device (src/sample_example.cpp:417-426). */
#define VKRT_IPC_HANDLE_BYTES 64
/* root: CUDA IPC handle of the library-owned accumulation image (vkrt_output_create with device_rgba32f == NULL) */
int vkrt_output_export(vkrt_ctx* ctx, uint8_t handle[VKRT_IPC_HANDLE_BYTES]);
/* other ranks: map an exported image of ANOTHER process (any GPU of the node with peer access) / unmap it */
int vkrt_peer_open(vkrt_ctx* ctx, const uint8_t handle[VKRT_IPC_HANDLE_BYTES], void** out_device_ptr);
int vkrt_peer_close(vkrt_ctx* ctx, void* device_ptr);
/* every pixel vkrt_render accumulates from now on is also stored to the same pixel of `device_rgba32f` (an image of this
 * context's dimensions: a vkrt_peer_open mapping, or any device buffer); NULL switches it off */
int vkrt_output_set_mirror(vkrt_ctx* ctx, void* device_rgba32f);
/* Puts vkrt_stream(ctx) into the order of image operations at this point: work enqueued on that stream after the call runs after
 * every image operation enqueued before it, and image operations enqueued after the call run after the work that stream held
 * before it.  Bracket a collective with two calls and it becomes a cross-process barrier between accumulates and tonemaps. */
int vkrt_output_fence(vkrt_ctx* ctx);

/* ---- render: replaces Renderer::setPushContants + Renderer::run ---------- */

/* Renders frames state->frame .. state->frame + num_frames - 1 (the reference calls
 * run() once per frame, src/sample_example.cpp:417-420; shaders/pathtrace.comp:87-134).
 * Asynchronous on the context stream; vkrt_sync() waits. */
int vkrt_render(vkrt_ctx* ctx, const RtxState* state, int32_t num_frames, const vkrt_render_opts* opts);
int vkrt_sync(vkrt_ctx* ctx);
/* The CUDA stream (cudaStream_t) the image operations of this context are ordered on: work a host enqueues there between two
 * calls (an NCCL exchange of the accumulation buffer, a Vulkan-interop copy) needs no host synchronisation. */
int vkrt_stream(vkrt_ctx* ctx, void** out_cuda_stream);

/* ---- host-side helpers of the reference's scene path (pure functions, no GPU) ------ */

/* Scene::createVertexBuffer vertex packing (src/scene.cpp:224-246): oct-compressed normal and
 * tangent (shaders/compress.glsl:111-139), tangent handedness in the LSB of texcoord.y,
 * RGBA8 colour (shaders/compress.glsl:60-72).  positions n*3, normals n*3, tangents n*4,
 * texcoords n*2, colors n*4 floats. */
int vkrt_pack_vertices(const float* positions, const float* normals, const float* tangents, const float* texcoords,
                       const float* colors, uint32_t n, VertexAttributes* out);
/* Scene::updateCamera (src/scene.cpp:629-640): lookAt view, perspectiveRH_ZO(fov, aspect, 0.001, 100000)
 * with proj[1][1] *= -1, both inverted; focalDist = |center - eye|. */
int vkrt_make_camera(const float eye[3], const float center[3], const float up[3], float fov_degrees, float aspect,
                     float aperture, int32_t nb_lights, SceneCamera* out);
/* Scene::createLightBuffer transform part (src/scene.cpp:307-309): position = M*(0,0,0,1), direction = M*(0,0,-1,0). */
int vkrt_light_from_matrix(const float world_matrix[16], Light* inout_light);

/* ---- diagnostics ------------------------------------------------------ */

/* One-ray probe of the traversal kernels (the reference's picking ray,
 * src/sample_example.cpp:469-511).  rays: n x {ox,oy,oz,tmax,dx,dy,dz,seed-bits}.
 * out_hits: n x {t,u,v,tri-id-bits} with t = 1e32 on miss; closest != 0 selects
 * ClosestHit, 0 selects AnyHit (t = 0/1e32 for hit/miss). */
int vkrt_trace_rays(vkrt_ctx* ctx, const float* rays, uint32_t n, int32_t closest, float* out_hits);
/* For a flattened triangle id returned above: {node(instance) index, primitive index}. */
int vkrt_triangle_ids(vkrt_ctx* ctx, const uint32_t* tri_ids, uint32_t n, uint32_t* out_node_prim);

/* Device-side stopwatch on the context stream (CUDA events): mark(0) .. mark(1), then
 * elapsed_ms() synchronises and returns the time between the two marks. */
int vkrt_mark(vkrt_ctx* ctx, int32_t which);
int vkrt_elapsed_ms(vkrt_ctx* ctx, float* out_ms);

int vkrt_get_stats(vkrt_ctx* ctx, vkrt_stats* out);
int vkrt_reset_stats(vkrt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* VKRT_B200_H */
