// vkrt_api.cu -- context, scene upload and the render loop behind the C ABI
// (include/vkrt_b200.h).  Host-side orchestration only: every pixel of work is done
// by the kernels in vkrt_bvh.cu / vkrt_trace.cu / vkrt_shade.cu.
//
// Reference counterparts: Renderer::setup/create/run/destroy (src/renderer.h:30-48),
// RayQuery::run (src/rayquery.cpp:96-109), Scene::load upload half (src/scene.cpp:93-109),
// AccelStructure::create (src/accelstruct.cpp:55-65), RenderOutput (src/render_output.cpp:78-182).
#include <vector>
#include <string>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include "vkrt_device.cuh"
#include "vkrt_internal.h"

using namespace vkrt;

namespace vkrt {
void launchShardCopy(cudaStream_t s, float4* image, const uint32_t* pixels, uint32_t n, float4* compact, bool toImage);
}

struct HostMesh
{
  std::vector<VertexAttributes> vtx;
  std::vector<uint32_t>         idx;
  int32_t                       material;
};
struct HostNode
{
  float    world[16];
  uint32_t mesh;
};
struct HostTexture
{
  std::vector<uint8_t> rgba;
  uint32_t             w, h;
  int32_t              linear, wrapS, wrapT;
};

enum TimerSlot { T_RAYGEN = 0, T_CLOSEST, T_SHADOW, T_SHADE, T_SORT, T_ACCUM, T_TONEMAP, T_CLOSEST0, T_SHADOW0, T_TAIL, T_COUNT };

// One wavefront in flight: its path state, its two streams (closest-hit / shade, and the any-hit kernels that overlap the next
// bounce) and the events that order them.  A context owns two: consecutive batches -- the halves of a multi-frame render, or
// consecutive one-frame calls of an interactive host -- alternate between them, so that the tail of every persistent kernel of
// one wavefront is filled by kernels of the other (round 1 measured one frame per call at half the batched rate: 16 tails).
constexpr int kMaxSlots  = 4;  // wavefront slots a context can rotate through
constexpr int kToneSlots = 4;  // staging buffers of vkrt_tonemap_async

struct WfSlot
{
  Wavefront    wf{};
  size_t       capacity = 0;  // paths
  cudaStream_t s = nullptr, s2 = nullptr;
  cudaEvent_t  evShade = nullptr, evShadow = nullptr, evJoin = nullptr, evJoin2 = nullptr;
  // ray sort (opts.sort_rays): third queue buffer (rotates with queueCur / queueNext) + keys and histogram, sized like the slot
  uint32_t* sortQueue = nullptr;
  void*     sortScratch = nullptr;
  size_t    sortCapacity = 0;
  uint32_t  counterLevels = 0;  // bounces the counter array holds
};

struct vkrt_ctx
{
  int          device = 0;
  cudaStream_t stream = nullptr;   // == slots[0].s: scene upload, image operations of the synchronous entry points, marks
  cudaStream_t post = nullptr;     // asynchronous tonemap + read-back
  WfSlot       slots[kMaxSlots];
  // Slots the renderer rotates through.  A multi-frame call is split into `numSlots` batches: 2 measured best (3: -0.2 %, 4: -0.7 %
  // -- smaller wavefronts cost more than the extra overlap gains).  One-frame calls of an interactive host take the next of
  // `numSlotsInteractive` slots, so that many frames overlap: 4 measured +3 % over 2 (profiles/README.md).  VKRT_SLOTS / VKRT_SLOTS_INTERACTIVE.
  int          numSlots = 2, numSlotsInteractive = 4;
  uint32_t     nextSlot = 0;       // batches take the slots in turn, across calls too
  cudaEvent_t  evImage = nullptr;  // last operation that touched the accumulation image (accumulate, clear, tonemap, read)
  bool         imagePending = false;
  cudaEvent_t  evFork = nullptr;
  int          numSMs = 148;
  std::string  err;

  // host staging (between scene_begin and scene_commit)
  std::vector<HostMesh>          meshes;
  std::vector<HostNode>          nodes;
  std::vector<GltfShadeMaterial> materials;
  std::vector<Light>             lights;
  std::vector<HostTexture>       textures;

  // device scene
  VertexAttributes*    dVerts = nullptr;
  uint32_t*            dIdx = nullptr;
  InstanceData*        dGeo = nullptr;
  InstanceXform*       dInst = nullptr;
  uint32_t*            dInstTriOffset = nullptr;
  GltfShadeMaterial*   dMat = nullptr;
  Light*               dLights = nullptr;
  TexDesc*             dTexDescs = nullptr;
  AlphaMat*            dAlphaMats = nullptr;
  std::vector<uchar4*> dTexData;
  float4*              dEnv = nullptr;
  EnvAccel*            dEnvAccel = nullptr;
  BvhResult            bvh;
  uint32_t             numLightsUploaded = 0;  // entries of dLights that hold caller data (the array itself never is empty)
  SceneView            view{};
  bool                 committed = false;
  bool                 hasEnv = false;

  // output image
  float4* image = nullptr;
  bool    ownsImage = false;
  float4* mirror = nullptr;        // vkrt_output_set_mirror: a second image (a peer's) the accumulate kernel also stores to
  int     W = 0, H = 0;
  float4* dToneF = nullptr;
  uchar4* dTone8 = nullptr;
  uchar4* dTone8Async[kToneSlots] = {};  // staging of vkrt_tonemap_async
  cudaEvent_t evCopy[kToneSlots] = {};
  bool    copyPending[kToneSlots] = {};
  float4* dMips = nullptr;  // auto-exposure mip pyramid (levels 1..n)
  size_t  mipCapacity = 0;

  // pixel lists of tile shards (device), keyed by {W, H, tile, rank, count}; [0] is the one the renderer uses
  struct ShardList { int key[5]; uint32_t* d; uint32_t n; };
  std::vector<ShardList> shardLists;
  uint32_t* dPixelOfLocal = nullptr;   // == the active list
  uint32_t  numLocalPixels = 0;

  std::vector<uint2> hostTriInfo;  // vkrt_triangle_ids: host copy of bvh.triInfo, dropped with the scene
  // scratch of vkrt_trace_rays (grow-only)
  void*  dProbe = nullptr;
  size_t probeCapacity = 0;

  // stats
  TraceStats* dStats = nullptr;
  vkrt_stats  stats{};
  struct Timed { cudaEvent_t a, b; int slot; };
  std::vector<Timed>       pending;
  cudaEvent_t              marks[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> eventPool;
};

#define CTX_CHECK(call)                                                                                                \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e_ = (call);                                                                                           \
    if(e_ != cudaSuccess)                                                                                              \
    {                                                                                                                  \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                                   \
      return VKRT_ERR_CUDA;                                                                                            \
    }                                                                                                                  \
  } while(0)

static int fail(vkrt_ctx* ctx, int code, const char* msg)
{
  if(ctx)
    ctx->err = msg;
  return code;
}

// every stream of the context idle (scene changes, statistics, teardown)
static cudaError_t syncAll(vkrt_ctx* c)
{
  cudaError_t first = cudaSuccess;
  for(auto& sl : c->slots)
    for(cudaStream_t st : {sl.s, sl.s2})
      if(st)
      {
        cudaError_t e = cudaStreamSynchronize(st);
        if(first == cudaSuccess) first = e;
      }
  if(c->post)
  {
    cudaError_t e = cudaStreamSynchronize(c->post);
    if(first == cudaSuccess) first = e;
  }
  return first;
}
// Image operations (accumulate, clear, tonemap, read) are enqueued in host order on different streams; each waits for the
// previous one and leaves its own completion in evImage.
static cudaError_t imageBegin(vkrt_ctx* c, cudaStream_t st)
{
  return c->imagePending ? cudaStreamWaitEvent(st, c->evImage, 0) : cudaSuccess;
}
static cudaError_t imageEnd(vkrt_ctx* c, cudaStream_t st)
{
  c->imagePending = true;
  return cudaEventRecord(c->evImage, st);
}
// stream `st` waits for everything enqueued so far on all the other streams, and they wait for it: a barrier in stream order
static cudaError_t barrierOn(vkrt_ctx* c, cudaStream_t st)
{
  cudaError_t first = cudaSuccess;
  auto ck = [&](cudaError_t e) { if(first == cudaSuccess) first = e; };
  for(auto& sl : c->slots)
  {
    if(sl.s && sl.s != st) { ck(cudaEventRecord(sl.evJoin, sl.s)); ck(cudaStreamWaitEvent(st, sl.evJoin, 0)); }
    if(sl.s2) { ck(cudaEventRecord(sl.evJoin2, sl.s2)); ck(cudaStreamWaitEvent(st, sl.evJoin2, 0)); }
  }
  return first;
}
static cudaError_t forkFrom(vkrt_ctx* c, cudaStream_t st)
{
  cudaError_t first = cudaEventRecord(c->evFork, st);
  for(auto& sl : c->slots)
    for(cudaStream_t o : {sl.s, sl.s2})
      if(o && o != st)
      {
        cudaError_t e = cudaStreamWaitEvent(o, c->evFork, 0);
        if(first == cudaSuccess) first = e;
      }
  return first;
}

static void freeScene(vkrt_ctx* c)
{
  cudaFree(c->dVerts); cudaFree(c->dIdx); cudaFree(c->dGeo); cudaFree(c->dInst); cudaFree(c->dInstTriOffset);
  cudaFree(c->dMat); cudaFree(c->dLights); cudaFree(c->dTexDescs); cudaFree(c->dAlphaMats);
  for(auto p : c->dTexData) cudaFree(p);
  c->dTexData.clear();
  cudaFree(c->bvh.nodes); cudaFree(c->bvh.qnodes); cudaFree(c->bvh.cnodes); cudaFree(c->bvh.wnodes); cudaFree(c->bvh.triInfo); cudaFree(c->bvh.triShade);
  c->dVerts = nullptr; c->dIdx = nullptr; c->dGeo = nullptr; c->dInst = nullptr; c->dInstTriOffset = nullptr;
  c->dMat = nullptr; c->dLights = nullptr; c->dTexDescs = nullptr; c->dAlphaMats = nullptr;
  c->bvh = BvhResult{};
  c->hostTriInfo.clear();
  if(c->committed)
    cudaCtxResetPersistingL2Cache();  // the BVH arena's persisting lines (device-wide set-aside) must not outlive the scene
  c->committed = false;
}

static void freeSlot(WfSlot& sl)
{
  Wavefront& w = sl.wf;
  cudaFree(w.rayO); cudaFree(w.rayD); cudaFree(w.hit); cudaFree(w.thr); cudaFree(w.rad); cudaFree(w.absb); cudaFree(w.sum);
  cudaFree(w.shO); cudaFree(w.shD); cudaFree(w.shC); cudaFree(w.queueCur); cudaFree(w.queueNext); cudaFree(w.counters);
  cudaFree(sl.sortQueue); cudaFree(sl.sortScratch);
  sl.sortQueue    = nullptr;
  sl.sortScratch  = nullptr;
  sl.sortCapacity  = 0;
  sl.counterLevels = 0;
  w                = Wavefront{};
  sl.capacity      = 0;
}
static void freeWavefront(vkrt_ctx* c)
{
  for(auto& sl : c->slots) freeSlot(sl);
}

static cudaEvent_t getEvent(vkrt_ctx* c)
{
  if(!c->eventPool.empty())
  {
    cudaEvent_t e = c->eventPool.back();
    c->eventPool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

static void drainTimers(vkrt_ctx* c)
{
  if(c->pending.empty())
    return;
  cudaStreamSynchronize(c->stream);
  double acc[T_COUNT] = {0};
  for(auto& t : c->pending)
  {
    float ms = 0.f;
    if(cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess)
      acc[t.slot] += ms;
    c->eventPool.push_back(t.a);
    c->eventPool.push_back(t.b);
  }
  c->pending.clear();
  c->stats.ms_raygen += acc[T_RAYGEN];
  c->stats.ms_trace_closest += acc[T_CLOSEST];
  c->stats.ms_trace_shadow += acc[T_SHADOW];
  c->stats.ms_shade += acc[T_SHADE];
  c->stats.ms_sort += acc[T_SORT];
  c->stats.ms_accumulate += acc[T_ACCUM];
  c->stats.ms_tonemap += acc[T_TONEMAP];
  c->stats.ms_trace_closest_primary += acc[T_CLOSEST0];
  c->stats.ms_trace_shadow_primary += acc[T_SHADOW0];
  c->stats.ms_tail += acc[T_TAIL];
}

struct ScopedTimer
{
  vkrt_ctx*   c;
  bool        on;
  cudaEvent_t a{}, b{};
  int         slot;
  ScopedTimer(vkrt_ctx* ctx, bool enable, int s) : c(ctx), on(enable), slot(s)
  {
    if(on)
    {
      a = getEvent(c);
      b = getEvent(c);
      cudaEventRecord(a, c->stream);
    }
  }
  ~ScopedTimer()
  {
    if(on)
    {
      cudaEventRecord(b, c->stream);
      c->pending.push_back({a, b, slot});
    }
  }
};

extern "C" {

const char* vkrt_name(void) { return "B200 wavefront path tracer (CUDA sm_100a)"; }

const char* vkrt_last_error(vkrt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vkrt_create(int cuda_device, vkrt_ctx** out_ctx)
{
  if(!out_ctx)
    return VKRT_ERR_INVALID;
  *out_ctx = nullptr;
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
    return VKRT_ERR_CUDA;  // no silent CPU fallback: the product path needs a GPU
  if(cuda_device < 0 || cuda_device >= n)
    return VKRT_ERR_INVALID;
  vkrt_ctx* ctx = new vkrt_ctx();
  ctx->device   = cuda_device;
  bool ok = cudaSetDevice(cuda_device) == cudaSuccess;
  for(auto& sl : ctx->slots)
  {
    ok = ok && cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&sl.s2, cudaStreamNonBlocking) == cudaSuccess;
    for(cudaEvent_t* e : {&sl.evShade, &sl.evShadow, &sl.evJoin, &sl.evJoin2})
      ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaStreamCreateWithFlags(&ctx->post, cudaStreamNonBlocking) == cudaSuccess;
  for(cudaEvent_t* e : {&ctx->evImage, &ctx->evFork})
    ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
  for(cudaEvent_t& e : ctx->evCopy)
    ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
  if(const char* e = getenv("VKRT_SLOTS"))
    ctx->numSlots = std::max(1, std::min(kMaxSlots, atoi(e)));
  if(const char* e = getenv("VKRT_SLOTS_INTERACTIVE"))
    ctx->numSlotsInteractive = std::max(1, std::min(kMaxSlots, atoi(e)));
  if(!ok)
  {
    delete ctx;  // the process is about to report the failure; the few handles created so far go with it
    return VKRT_ERR_CUDA;
  }
  ctx->stream = ctx->slots[0].s;
  {
    // the acceleration-structure build takes its temporaries from the stream-ordered pool: keep up to 1 GB of it warm between builds
    cudaMemPool_t pool = nullptr;
    if(cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess && pool)
    {
      uint64_t keep = 1ull << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, cuda_device);
  if(cudaMalloc(&ctx->dStats, sizeof(TraceStats)) != cudaSuccess)
  {
    delete ctx;
    return VKRT_ERR_NOMEM;
  }
  cudaMemset(ctx->dStats, 0, sizeof(TraceStats));
  // default sun & sky: src/sample_example.hpp:176-193 (in_use = 0)
  SunAndSky ss{};
  ss.rgb_unit_conversion[0] = ss.rgb_unit_conversion[1] = ss.rgb_unit_conversion[2] = 1.f;
  ss.multiplier = 0.0000101320f; ss.saturation = 1.0f;
  ss.ground_color[0] = ss.ground_color[1] = ss.ground_color[2] = 0.4f;
  ss.horizon_blur = 0.1f; ss.night_color[2] = 0.01f; ss.sun_disk_intensity = 0.8f;
  ss.sun_direction[1] = 0.78f; ss.sun_direction[2] = 0.62f; ss.sun_disk_scale = 5.0f; ss.sun_glow_intensity = 1.0f;
  ss.y_is_up = 1; ss.physically_scaled_sun = 1; ss.in_use = 0;
  ctx->view.sunsky = ss;
  *out_ctx = ctx;
  return VKRT_OK;
}

int vkrt_destroy(vkrt_ctx* ctx)
{
  if(!ctx)
    return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  syncAll(ctx);
  drainTimers(ctx);
  for(auto e : ctx->eventPool) cudaEventDestroy(e);
  for(auto e : ctx->marks) if(e) cudaEventDestroy(e);
  freeScene(ctx);
  freeWavefront(ctx);
  cudaFree(ctx->dEnv); cudaFree(ctx->dEnvAccel);
  for(auto& l : ctx->shardLists) cudaFree(l.d);
  if(ctx->ownsImage) cudaFree(ctx->image);
  cudaFree(ctx->dToneF); cudaFree(ctx->dTone8); cudaFree(ctx->dMips);
  for(uchar4* b : ctx->dTone8Async) cudaFree(b);
  cudaFree(ctx->dStats);
  cudaFree(ctx->dProbe);
  for(auto& sl : ctx->slots)
  {
    for(cudaEvent_t e : {sl.evShade, sl.evShadow, sl.evJoin, sl.evJoin2}) cudaEventDestroy(e);
    cudaStreamDestroy(sl.s2);
    cudaStreamDestroy(sl.s);
  }
  for(cudaEvent_t e : {ctx->evImage, ctx->evFork}) cudaEventDestroy(e);
  for(cudaEvent_t e : ctx->evCopy) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->post);
  delete ctx;
  return VKRT_OK;
}

// ---- scene -------------------------------------------------------------------------------------
int vkrt_scene_begin(vkrt_ctx* ctx)
{
  if(!ctx) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  syncAll(ctx);
  freeScene(ctx);
  ctx->meshes.clear(); ctx->nodes.clear(); ctx->materials.clear(); ctx->lights.clear(); ctx->textures.clear();
  return VKRT_OK;
}

int vkrt_scene_add_primmesh(vkrt_ctx* ctx, const VertexAttributes* vertices, uint32_t num_vertices, const uint32_t* indices,
                            uint32_t num_indices, int32_t material_index)
{
  if(!ctx) return VKRT_ERR_INVALID;
  if((!vertices && num_vertices) || (!indices && num_indices) || (num_indices % 3) != 0)
    return fail(ctx, VKRT_ERR_INVALID, "add_primmesh: bad arrays (indices must be triangle triples)");
  for(uint32_t i = 0; i < num_indices; i++)
    if(indices[i] >= num_vertices)
      return fail(ctx, VKRT_ERR_INVALID, "add_primmesh: index out of range");
  HostMesh m;
  m.vtx.assign(vertices, vertices + num_vertices);
  m.idx.assign(indices, indices + num_indices);
  m.material = material_index;
  ctx->meshes.push_back(std::move(m));
  return VKRT_OK;
}

int vkrt_scene_add_node(vkrt_ctx* ctx, const float world_matrix[16], uint32_t prim_mesh)
{
  if(!ctx || !world_matrix) return VKRT_ERR_INVALID;
  if(prim_mesh >= ctx->meshes.size())
    return fail(ctx, VKRT_ERR_INVALID, "add_node: unknown primMesh");
  HostNode n;
  std::memcpy(n.world, world_matrix, 64);
  n.mesh = prim_mesh;
  ctx->nodes.push_back(n);
  return VKRT_OK;
}

int vkrt_scene_set_materials(vkrt_ctx* ctx, const GltfShadeMaterial* materials, uint32_t count)
{
  if(!ctx || (!materials && count)) return VKRT_ERR_INVALID;
  ctx->materials.assign(materials, materials + count);
  return VKRT_OK;
}

int vkrt_scene_set_lights(vkrt_ctx* ctx, const Light* lights, uint32_t count)
{
  if(!ctx || (!lights && count)) return VKRT_ERR_INVALID;
  ctx->lights.assign(lights, lights + count);
  return VKRT_OK;
}

int vkrt_scene_add_texture(vkrt_ctx* ctx, const uint8_t* rgba8, uint32_t width, uint32_t height, int32_t mag_filter, int32_t min_filter,
                           int32_t wrap_s, int32_t wrap_t)
{
  (void)min_filter;  // LOD 0 sampling always magnifies; mip chains are never generated (src/scene.cpp:540-542)
  if(!ctx || !rgba8 || width == 0 || height == 0) return VKRT_ERR_INVALID;
  HostTexture t;
  t.rgba.assign(rgba8, rgba8 + (size_t)width * height * 4);
  t.w = width; t.h = height;
  // src/scene.cpp:451-457: only the LINEAR family maps to VK_FILTER_LINEAR; unknown codes default to NEAREST
  t.linear = (mag_filter == 9729 || mag_filter == 9985 || mag_filter == 9987) ? 1 : 0;
  auto wrap = [](int32_t m) { return m == 33071 ? 1 : (m == 33648 ? 2 : 0); };  // src/scene.cpp:467-470, default REPEAT
  t.wrapS = wrap(wrap_s);
  t.wrapT = wrap(wrap_t);
  ctx->textures.push_back(std::move(t));
  return VKRT_OK;
}

int vkrt_scene_commit(vkrt_ctx* ctx, int32_t build_mode)
{
  if(!ctx) return VKRT_ERR_INVALID;
  const int32_t widthFlags = build_mode & (VKRT_BUILD_FORCE_BVH2 | VKRT_BUILD_FORCE_BVH4);
  if((build_mode & ~(VKRT_BUILD_MODE_MASK | VKRT_BUILD_FORCE_BVH2 | VKRT_BUILD_FORCE_BVH4)) || widthFlags == (VKRT_BUILD_FORCE_BVH2 | VKRT_BUILD_FORCE_BVH4))
    return fail(ctx, VKRT_ERR_INVALID, "scene_commit: unknown build flags");
  build_mode &= VKRT_BUILD_MODE_MASK;
  if(build_mode < VKRT_BUILD_LBVH || build_mode > VKRT_BUILD_SAH_GPU)
    return fail(ctx, VKRT_ERR_INVALID, "scene_commit: unknown build mode");
  cudaSetDevice(ctx->device);
  syncAll(ctx);
  if(ctx->materials.empty())
    return fail(ctx, VKRT_ERR_STATE, "scene_commit: no materials (the importer always appends a default one)");
  for(auto& m : ctx->meshes)
    if(m.material >= (int32_t)ctx->materials.size())
      return fail(ctx, VKRT_ERR_INVALID, "scene_commit: material index out of range");
  for(auto& m : ctx->materials)
  {
    const int32_t ids[] = {m.pbrBaseColorTexture, m.pbrMetallicRoughnessTexture, m.emissiveTexture, m.normalTexture,
                           m.transmissionTexture, m.clearcoatTexture, m.clearcoatRoughnessTexture};
    for(int32_t id : ids)
      if(id >= (int32_t)ctx->textures.size())
        return fail(ctx, VKRT_ERR_INVALID, "scene_commit: texture index out of range");
  }
  freeScene(ctx);
  cudaEvent_t e0 = getEvent(ctx), e1 = getEvent(ctx);
  cudaEventRecord(e0, ctx->stream);

  // geometry: one vertex and one index arena, InstanceData per primMesh (src/scene.cpp:161-176)
  size_t nv = 0, ni = 0;
  for(auto& m : ctx->meshes) { nv += m.vtx.size(); ni += m.idx.size(); }
  CTX_CHECK(cudaMalloc(&ctx->dVerts, sizeof(VertexAttributes) * std::max<size_t>(nv, 1)));
  CTX_CHECK(cudaMalloc(&ctx->dIdx, sizeof(uint32_t) * std::max<size_t>(ni, 1)));
  std::vector<InstanceData> geo(ctx->meshes.size());
  {
    size_t ov = 0, oi = 0;
    for(size_t i = 0; i < ctx->meshes.size(); i++)
    {
      auto& m = ctx->meshes[i];
      if(!m.vtx.empty())
        CTX_CHECK(cudaMemcpyAsync(ctx->dVerts + ov, m.vtx.data(), sizeof(VertexAttributes) * m.vtx.size(), cudaMemcpyHostToDevice, ctx->stream));
      if(!m.idx.empty())
        CTX_CHECK(cudaMemcpyAsync(ctx->dIdx + oi, m.idx.data(), 4 * m.idx.size(), cudaMemcpyHostToDevice, ctx->stream));
      geo[i].vertexAddress = (uint64_t)(uintptr_t)(ctx->dVerts + ov);
      geo[i].indexAddress  = (uint64_t)(uintptr_t)(ctx->dIdx + oi);
      geo[i].materialIndex = m.material;
      geo[i]._pad          = 0;
      ov += m.vtx.size();
      oi += m.idx.size();
    }
  }
  CTX_CHECK(cudaMalloc(&ctx->dGeo, sizeof(InstanceData) * std::max<size_t>(geo.size(), 1)));
  if(!geo.empty())
    CTX_CHECK(cudaMemcpyAsync(ctx->dGeo, geo.data(), sizeof(InstanceData) * geo.size(), cudaMemcpyHostToDevice, ctx->stream));

  // instances: flags per src/accelstruct.cpp:139-149
  std::vector<InstanceXform> inst(ctx->nodes.size());
  std::vector<uint32_t>      triOff(ctx->nodes.size() + 1);
  uint64_t                   totalTris = 0;
  bool                       nonOpaque = false;
  for(size_t i = 0; i < ctx->nodes.size(); i++)
  {
    const HostNode&          n   = ctx->nodes[i];
    const HostMesh&          m   = ctx->meshes[n.mesh];
    const GltfShadeMaterial& mat = ctx->materials[std::max(0, m.material)];
    uint32_t                 fl  = 0;
    if(mat.alphaMode == 0 || (mat.pbrBaseColorFactor[3] == 1.0f && mat.pbrBaseColorTexture == -1))
      fl |= TRI_OPAQUE;
    if(getenv("VKRT_DEBUG_FORCE_OPAQUE"))
      fl |= TRI_OPAQUE;  // profiling aid only: the image changes (alpha-tested geometry becomes solid)
    if(mat.doubleSided == 1)
      fl |= TRI_CULL_DISABLE;
    const float* w = n.world;
    float det = w[0] * (w[5] * w[10] - w[9] * w[6]) - w[4] * (w[1] * w[10] - w[9] * w[2]) + w[8] * (w[1] * w[6] - w[5] * w[2]);
    if(det < 0.0f)
      fl |= TRI_FLIP_FACING;  // Vulkan decides facing in object space
    if(!(fl & TRI_OPAQUE))
      nonOpaque = true;
    float inv[16];
    if(!invertAffine(w, inv))
      return fail(ctx, VKRT_ERR_INVALID, "scene_commit: singular node matrix");
    InstanceXform& I = inst[i];
    for(int r = 0; r < 3; r++)
      for(int c = 0; c < 4; c++)
      {
        I.o2w[r * 4 + c] = w[c * 4 + r];
        I.w2o[r * 4 + c] = inv[c * 4 + r];
      }
    I.primMesh  = n.mesh;
    I.flags     = fl;
    I.triOffset = (uint32_t)totalTris;
    I.triCount  = (uint32_t)(m.idx.size() / 3);
    triOff[i]   = (uint32_t)totalTris;
    totalTris += m.idx.size() / 3;
  }
  triOff[ctx->nodes.size()] = (uint32_t)totalTris;
  if(totalTris > (uint64_t)kLeafMask)
    return fail(ctx, VKRT_ERR_INVALID, "scene_commit: more than 2^28 triangles after flattening");
  CTX_CHECK(cudaMalloc(&ctx->dInst, sizeof(InstanceXform) * std::max<size_t>(inst.size(), 1)));
  CTX_CHECK(cudaMalloc(&ctx->dInstTriOffset, 4 * triOff.size()));
  if(!inst.empty())
    CTX_CHECK(cudaMemcpyAsync(ctx->dInst, inst.data(), sizeof(InstanceXform) * inst.size(), cudaMemcpyHostToDevice, ctx->stream));
  CTX_CHECK(cudaMemcpyAsync(ctx->dInstTriOffset, triOff.data(), 4 * triOff.size(), cudaMemcpyHostToDevice, ctx->stream));

  // materials, lights (never empty: src/scene.cpp:329-330), textures
  CTX_CHECK(cudaMalloc(&ctx->dMat, sizeof(GltfShadeMaterial) * ctx->materials.size()));
  CTX_CHECK(cudaMemcpyAsync(ctx->dMat, ctx->materials.data(), sizeof(GltfShadeMaterial) * ctx->materials.size(), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<Light> lights = ctx->lights;
  if(lights.empty())
    lights.push_back(Light{});
  CTX_CHECK(cudaMalloc(&ctx->dLights, sizeof(Light) * lights.size()));
  CTX_CHECK(cudaMemcpyAsync(ctx->dLights, lights.data(), sizeof(Light) * lights.size(), cudaMemcpyHostToDevice, ctx->stream));
  ctx->numLightsUploaded = (uint32_t)ctx->lights.size();
  std::vector<TexDesc> descs(ctx->textures.size());
  for(size_t i = 0; i < ctx->textures.size(); i++)
  {
    const HostTexture& t = ctx->textures[i];
    uchar4*            d = nullptr;
    CTX_CHECK(cudaMalloc(&d, t.rgba.size()));
    ctx->dTexData.push_back(d);
    CTX_CHECK(cudaMemcpyAsync(d, t.rgba.data(), t.rgba.size(), cudaMemcpyHostToDevice, ctx->stream));
    descs[i] = TexDesc{d, (int32_t)t.w, (int32_t)t.h, t.linear, t.wrapS, t.wrapT, 0};
  }
  CTX_CHECK(cudaMalloc(&ctx->dTexDescs, sizeof(TexDesc) * std::max<size_t>(descs.size(), 1)));
  if(!descs.empty())
    CTX_CHECK(cudaMemcpyAsync(ctx->dTexDescs, descs.data(), sizeof(TexDesc) * descs.size(), cudaMemcpyHostToDevice, ctx->stream));
  // alpha records for the any-hit test (HitTest, shaders/traceray_rq.glsl:32-102): everything it reads from the material
  std::vector<AlphaMat> alphaMats(ctx->materials.size());
  for(size_t i = 0; i < ctx->materials.size(); i++)
  {
    const GltfShadeMaterial& m = ctx->materials[i];
    AlphaMat&                a = alphaMats[i];
    std::memset(&a, 0, sizeof(a));
    a.alphaFactor = m.pbrBaseColorFactor[3];
    a.alphaCutoff = m.alphaCutoff;
    a.alphaMode   = m.alphaMode;
    a.hasTexture  = (m.pbrBaseColorTexture > -1 && (size_t)m.pbrBaseColorTexture < descs.size()) ? 1 : 0;
    std::memcpy(a.uvRows, m.uvTransform, sizeof(a.uvRows));
    if(a.hasTexture)
      a.texture = descs[m.pbrBaseColorTexture];
  }
  CTX_CHECK(cudaMalloc(&ctx->dAlphaMats, sizeof(AlphaMat) * std::max<size_t>(alphaMats.size(), 1)));
  if(!alphaMats.empty())
    CTX_CHECK(cudaMemcpyAsync(ctx->dAlphaMats, alphaMats.data(), sizeof(AlphaMat) * alphaMats.size(), cudaMemcpyHostToDevice, ctx->stream));

  // acceleration structure
  std::string berr;
  uint64_t    launches = 0;
  cudaEventRecord(e0, ctx->stream);  // ms_build: the acceleration structure only (flatten, sort, hierarchy, refit, emit, node formats)
  if(!buildBvh(ctx->stream, ctx->dInst, ctx->dInstTriOffset, (uint32_t)inst.size(), ctx->dGeo, (uint32_t)totalTris, build_mode, ctx->bvh, berr, &launches))
  {
    ctx->err = "scene_commit: BVH build failed: " + berr;
    return VKRT_ERR_CUDA;
  }
  ctx->stats.kernel_launches += launches;
  if(build_mode == VKRT_BUILD_SAH_HOST && totalTris > 0)
  {
    // re-organise the GPU-flattened triangles under a host-built binned-SAH tree
    const uint32_t nt = (uint32_t)totalTris;
    std::vector<float>    htris(12ull * nt), hnodes, sorted(12ull * nt);
    std::vector<uint32_t> order;
    int                   depth = 0;
    CTX_CHECK(cudaMemcpy(htris.data(), ctx->bvh.tris, sizeof(float) * 12ull * nt, cudaMemcpyDeviceToHost));
    if(!buildSahHost(htris.data(), nt, hnodes, order, depth))
      return fail(ctx, VKRT_ERR_INVALID, "scene_commit: host SAH build failed");
    for(uint32_t i = 0; i < nt; i++)
      std::memcpy(&sorted[12ull * i], &htris[12ull * order[i]], 48);
    if(hnodes.size() / 16 > ctx->bvh.numNodes)
      return fail(ctx, VKRT_ERR_INVALID, "scene_commit: host SAH produced more nodes than the arena holds");
    ctx->bvh.numNodes = (uint32_t)(hnodes.size() / 16);  // <= n - 1: fits the arena of the GPU build
    CTX_CHECK(cudaMemcpy(ctx->bvh.nodes, hnodes.data(), hnodes.size() * sizeof(float), cudaMemcpyHostToDevice));
    CTX_CHECK(cudaMemcpy(ctx->bvh.tris, sorted.data(), sorted.size() * sizeof(float), cudaMemcpyHostToDevice));
    ctx->bvh.maxDepth = depth;
  }
  if(ctx->bvh.maxDepth > 120)
    return fail(ctx, VKRT_ERR_INVALID, "scene_commit: the BVH is deeper than the traversal stack (128 entries)");
  if(!quantizeBvh(ctx->stream, ctx->bvh, berr, &launches))
  {
    ctx->err = "scene_commit: BVH node quantisation failed: " + berr;
    return VKRT_ERR_CUDA;
  }
  if(!collapseBvhWide(ctx->stream, ctx->bvh, berr, &launches))
  {
    ctx->err = "scene_commit: BVH4 collapse failed: " + berr;
    return VKRT_ERR_CUDA;
  }
  ctx->stats.kernel_launches += 1 + (uint64_t)ctx->bvh.wideDepth;
  // a 4-wide step pushes up to three references per level: the deepest chain must fit the 128-entry traversal stack
  if(3 * ctx->bvh.wideDepth + 4 > 128)
    return fail(ctx, VKRT_ERR_INVALID, "scene_commit: the 4-wide BVH is deeper than the traversal stack (128 entries)");
  cudaEventRecord(e1, ctx->stream);
  CTX_CHECK(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ctx->eventPool.push_back(e0);
  ctx->eventPool.push_back(e1);
  ctx->stats.ms_build       = ms;
  ctx->stats.num_triangles  = (uint32_t)totalTris;
  ctx->stats.num_bvh_nodes  = ctx->bvh.numNodes;
  const int width           = widthFlags == VKRT_BUILD_FORCE_BVH2 ? 2 : (widthFlags == VKRT_BUILD_FORCE_BVH4 ? 4 : traceBvhWidth((uint32_t)totalTris));
  ctx->stats.node_bytes     = (uint32_t)traceNodeBytes(width);
  ctx->stats.bvh_width      = (uint32_t)width;
  if(width == 4)
    ctx->stats.num_bvh_nodes = ctx->bvh.numWideNodes;
  ctx->stats.num_instances  = (uint32_t)inst.size();
  ctx->stats.has_non_opaque = nonOpaque ? 1u : 0u;

  // Keep the BVH resident in L2: mark its arena persisting and everything else streaming for this stream's kernels
  // (the wavefront state streams hundreds of MB per bounce through the same L2).
  {
    const char* e = getenv("VKRT_L2_PERSIST");
    int maxWin = 0, maxPersist = 0;
    cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    if((!e || atoi(e) != 0) && maxWin > 0 && maxPersist > 0 && ctx->bvh.arenaBytes > 0)
    {
      const size_t want = std::min<size_t>(ctx->bvh.arenaBytes, (size_t)maxPersist);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.base_ptr  = ctx->bvh.nodes;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(ctx->bvh.arenaBytes, (size_t)maxWin);
      attr.accessPolicyWindow.hitRatio  = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
      attr.accessPolicyWindow.hitProp   = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp  = cudaAccessPropertyStreaming;
      for(auto& sl : ctx->slots)
        for(cudaStream_t st : {sl.s, sl.s2})
          cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  SceneView& v   = ctx->view;
  v.nodes        = ctx->bvh.nodes;
  v.qnodes       = ctx->bvh.qnodes;
  v.cnodes       = ctx->bvh.cnodes;
  v.wnodes       = ctx->bvh.wnodes;
  v.gridOrigin   = make_float3(ctx->bvh.gridOrigin[0], ctx->bvh.gridOrigin[1], ctx->bvh.gridOrigin[2]);
  v.gridInv      = make_float3(ctx->bvh.gridInv[0], ctx->bvh.gridInv[1], ctx->bvh.gridInv[2]);
  v.tris         = ctx->bvh.tris;
  v.triInfo      = ctx->bvh.triInfo;
  v.triShade     = ctx->bvh.triShade;
  v.instances    = ctx->dInst;
  v.geoInfo      = ctx->dGeo;
  v.materials    = ctx->dMat;
  v.alphaMats    = ctx->dAlphaMats;
  v.lights       = ctx->dLights;
  v.textures     = ctx->dTexDescs;
  v.numTris      = (uint32_t)totalTris;
  v.bvhWidth     = (uint32_t)width;
  v.hasNonOpaque = nonOpaque ? 1u : 0u;
  ctx->committed = true;
  return VKRT_OK;
}

// ---- environment / camera --------------------------------------------------------------------------
int vkrt_build_env_accel(const float* rgba32f, uint32_t width, uint32_t height, EnvAccel* out_accel, float* out_integral, float* out_average)
{
  if(!rgba32f || !out_accel || width == 0 || height == 0)
    return VKRT_ERR_INVALID;
  buildEnvAccelHost(rgba32f, width, height, out_accel, out_integral, out_average);
  return VKRT_OK;
}

int vkrt_set_environment(vkrt_ctx* ctx, const float* rgba32f, uint32_t width, uint32_t height, const EnvAccel* accel)
{
  if(!ctx || !rgba32f || width == 0 || height == 0) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  syncAll(ctx);
  cudaFree(ctx->dEnv); cudaFree(ctx->dEnvAccel);
  ctx->dEnv = nullptr; ctx->dEnvAccel = nullptr;
  const size_t n = (size_t)width * height;
  std::vector<EnvAccel> tmp;
  if(!accel)
  {
    tmp.resize(n);
    buildEnvAccelHost(rgba32f, width, height, tmp.data(), nullptr, nullptr);
    accel = tmp.data();
  }
  CTX_CHECK(cudaMalloc(&ctx->dEnv, n * sizeof(float4)));
  CTX_CHECK(cudaMalloc(&ctx->dEnvAccel, n * sizeof(EnvAccel)));
  CTX_CHECK(cudaMemcpyAsync(ctx->dEnv, rgba32f, n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  CTX_CHECK(cudaMemcpyAsync(ctx->dEnvAccel, accel, n * sizeof(EnvAccel), cudaMemcpyHostToDevice, ctx->stream));
  CTX_CHECK(cudaStreamSynchronize(ctx->stream));
  ctx->view.envPixels = ctx->dEnv;
  ctx->view.envAccel  = ctx->dEnvAccel;
  ctx->view.envW      = (int32_t)width;
  ctx->view.envH      = (int32_t)height;
  ctx->hasEnv         = true;
  return VKRT_OK;
}

int vkrt_set_sunsky(vkrt_ctx* ctx, const SunAndSky* s)
{
  if(!ctx || !s) return VKRT_ERR_INVALID;
  ctx->view.sunsky = *s;
  return VKRT_OK;
}

int vkrt_set_camera(vkrt_ctx* ctx, const SceneCamera* cam)
{
  if(!ctx || !cam) return VKRT_ERR_INVALID;
  ctx->view.camera = *cam;  // travels as a kernel parameter (the reference's UBO, src/scene.cpp:655)
  return VKRT_OK;
}

// ---- output ------------------------------------------------------------------------------------------
int vkrt_output_create(vkrt_ctx* ctx, uint32_t width, uint32_t height, void* device_rgba32f)
{
  if(!ctx || width == 0 || height == 0) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  syncAll(ctx);
  ctx->imagePending = false;
  for(bool& b : ctx->copyPending) b = false;
  ctx->mirror = nullptr;  // a mirror belongs to the image it was set for
  if(ctx->ownsImage) cudaFree(ctx->image);
  cudaFree(ctx->dToneF); cudaFree(ctx->dTone8);
  for(uchar4*& b : ctx->dTone8Async) { cudaFree(b); b = nullptr; }
  ctx->image = nullptr; ctx->dToneF = nullptr; ctx->dTone8 = nullptr;
  ctx->W = (int)width; ctx->H = (int)height;
  if(device_rgba32f)
  {
    ctx->image     = (float4*)device_rgba32f;
    ctx->ownsImage = false;
  }
  else
  {
    CTX_CHECK(cudaMalloc(&ctx->image, (size_t)width * height * sizeof(float4)));
    CTX_CHECK(cudaMemsetAsync(ctx->image, 0, (size_t)width * height * sizeof(float4), ctx->stream));
    CTX_CHECK(imageEnd(ctx, ctx->stream));
    ctx->ownsImage = true;
  }
  return VKRT_OK;
}

int vkrt_output_device_ptr(vkrt_ctx* ctx, void** out)
{
  if(!ctx || !out || !ctx->image) return VKRT_ERR_STATE;
  *out = ctx->image;
  return VKRT_OK;
}

int vkrt_output_clear(vkrt_ctx* ctx)
{
  if(!ctx || !ctx->image) return VKRT_ERR_STATE;
  cudaSetDevice(ctx->device);
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  CTX_CHECK(cudaMemsetAsync(ctx->image, 0, (size_t)ctx->W * ctx->H * sizeof(float4), ctx->stream));
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  return VKRT_OK;
}

int vkrt_output_read(vkrt_ctx* ctx, float* host)
{
  if(!ctx || !ctx->image || !host) return VKRT_ERR_STATE;
  cudaSetDevice(ctx->device);
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  CTX_CHECK(cudaMemcpyAsync(host, ctx->image, (size_t)ctx->W * ctx->H * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  CTX_CHECK(cudaStreamSynchronize(ctx->stream));
  return VKRT_OK;
}

// Mip pyramid + tonemap kernel on `st` (the image must already be ordered before `st`): out into dF / d8.
static int enqueueTonemap(vkrt_ctx* ctx, cudaStream_t st, const Tonemapper* tm, float4* dF, uchar4* d8, bool timed)
{
  MipView mips{};
  if(tm->autoExposure & 1)
  {
    // RenderOutput::genMipmap (src/render_output.cpp:188-193): the full chain, level 0 is the accumulation image
    size_t total = 0;
    int    w = ctx->W, h = ctx->H, levels = 1;
    while(w > 1 || h > 1) { w = std::max(1, w >> 1); h = std::max(1, h >> 1); total += (size_t)w * h; levels++; }
    if(levels > 16)
      return fail(ctx, VKRT_ERR_INVALID, "tonemap: image too large for the mip chain");
    if(total > ctx->mipCapacity)
    {
      CTX_CHECK(cudaStreamSynchronize(ctx->post));  // an asynchronous tonemap may still read the old pyramid
      cudaFree(ctx->dMips);
      ctx->dMips = nullptr;
      CTX_CHECK(cudaMalloc(&ctx->dMips, std::max<size_t>(total, 1) * sizeof(float4)));
      ctx->mipCapacity = total;
    }
    mips.level[0] = ctx->image; mips.w[0] = ctx->W; mips.h[0] = ctx->H;
    float4* dst = ctx->dMips;
    for(int l = 1; l < levels; l++)
    {
      mips.w[l] = std::max(1, mips.w[l - 1] >> 1);
      mips.h[l] = std::max(1, mips.h[l - 1] >> 1);
      launchMipDown(st, mips.level[l - 1], mips.w[l - 1], mips.h[l - 1], dst, mips.w[l], mips.h[l]);
      ctx->stats.kernel_launches++;
      mips.level[l] = dst;
      dst += (size_t)mips.w[l] * mips.h[l];
    }
    mips.count = levels;
  }
  {
    ScopedTimer t(ctx, timed, T_TONEMAP);
    launchTonemap(st, ctx->image, ctx->W, ctx->H, *tm, mips, dF, d8);
  }
  ctx->stats.kernel_launches++;
  return VKRT_OK;
}

static int tonemapCommon(vkrt_ctx* ctx, const Tonemapper* tm, float* hostF, uint8_t* host8)
{
  if(!ctx || !tm || !ctx->image) return VKRT_ERR_STATE;
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->W * ctx->H;
  if(hostF && !ctx->dToneF) CTX_CHECK(cudaMalloc(&ctx->dToneF, n * sizeof(float4)));
  if(host8 && !ctx->dTone8) CTX_CHECK(cudaMalloc(&ctx->dTone8, n * sizeof(uchar4)));
  CTX_CHECK(cudaStreamSynchronize(ctx->post));  // the mip pyramid is shared with vkrt_tonemap_async
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  int rc = enqueueTonemap(ctx, ctx->stream, tm, hostF ? ctx->dToneF : nullptr, host8 ? ctx->dTone8 : nullptr, true);
  if(rc != VKRT_OK) return rc;
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  if(hostF) CTX_CHECK(cudaMemcpyAsync(hostF, ctx->dToneF, n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  if(host8) CTX_CHECK(cudaMemcpyAsync(host8, ctx->dTone8, n * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
  CTX_CHECK(cudaStreamSynchronize(ctx->stream));
  return VKRT_OK;
}
int vkrt_tonemap(vkrt_ctx* ctx, const Tonemapper* tm, uint8_t* host_rgba8) { return tonemapCommon(ctx, tm, nullptr, host_rgba8); }
int vkrt_tonemap_float(vkrt_ctx* ctx, const Tonemapper* tm, float* host_rgba32f) { return tonemapCommon(ctx, tm, host_rgba32f, nullptr); }

int vkrt_tonemap_async(vkrt_ctx* ctx, const Tonemapper* tm, uint8_t* host_rgba8, int32_t slot)
{
  if(!ctx || !tm || !host_rgba8 || slot < 0 || slot >= kToneSlots) return VKRT_ERR_INVALID;
  if(!ctx->image) return fail(ctx, VKRT_ERR_STATE, "tonemap_async: no output image");
  cudaSetDevice(ctx->device);
  const size_t n = (size_t)ctx->W * ctx->H;
  if(!ctx->dTone8Async[slot]) CTX_CHECK(cudaMalloc(&ctx->dTone8Async[slot], n * sizeof(uchar4)));
  // the post stream runs: [wait for the frames enqueued so far] tonemap kernel -> [image free again] -> copy to the host.
  // Stream order on `post` also keeps the previous use of this staging buffer (its copy) ahead of the kernel that refills it.
  CTX_CHECK(imageBegin(ctx, ctx->post));
  int rc = enqueueTonemap(ctx, ctx->post, tm, nullptr, ctx->dTone8Async[slot], false);
  if(rc != VKRT_OK) return rc;
  CTX_CHECK(imageEnd(ctx, ctx->post));
  CTX_CHECK(cudaMemcpyAsync(host_rgba8, ctx->dTone8Async[slot], n * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->post));
  CTX_CHECK(cudaEventRecord(ctx->evCopy[slot], ctx->post));
  ctx->copyPending[slot] = true;
  return VKRT_OK;
}

int vkrt_tonemap_wait(vkrt_ctx* ctx, int32_t slot)
{
  if(!ctx || slot < 0 || slot >= kToneSlots) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if(ctx->copyPending[slot])
  {
    CTX_CHECK(cudaEventSynchronize(ctx->evCopy[slot]));
    ctx->copyPending[slot] = false;
  }
  return VKRT_OK;
}

int vkrt_host_alloc(size_t bytes, void** out_ptr)
{
  if(!out_ptr) return VKRT_ERR_INVALID;
  *out_ptr = nullptr;
  return cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? VKRT_OK : VKRT_ERR_NOMEM;
}
int vkrt_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? VKRT_OK : VKRT_ERR_CUDA; }

// ---- tile shards ---------------------------------------------------------------------------------------
// Tiles are dealt round-robin to the shards; inside a tile pixels are ordered in 8x4 blocks so that a warp of primary rays
// covers a compact screen footprint.  Lists are cached per (image size, tile, rank, count).
static int shardList(vkrt_ctx* ctx, int tile, int rank, int count, const vkrt_ctx::ShardList** out)
{
  if(tile <= 0) tile = 64;
  if(count < 1) count = 1;
  if(count == 1) rank = 0;
  if(rank < 0 || rank >= count)
    return fail(ctx, VKRT_ERR_INVALID, "shard_rank out of range");
  if(!ctx->image)
    return fail(ctx, VKRT_ERR_STATE, "no output image");
  const int key[5] = {ctx->W, ctx->H, tile, rank, count};
  for(auto& l : ctx->shardLists)
    if(std::memcmp(key, l.key, sizeof(key)) == 0)
    {
      *out = &l;
      return VKRT_OK;
    }
  std::vector<uint32_t> px;
  px.reserve((size_t)ctx->W * ctx->H / (size_t)count + 4096);
  const int tx = (ctx->W + tile - 1) / tile, ty = (ctx->H + tile - 1) / tile;
  for(int t = 0; t < tx * ty; t++)
  {
    if(count > 1 && (t % count) != rank)
      continue;
    const int x0 = (t % tx) * tile, y0 = (t / tx) * tile;
    const int x1 = std::min(x0 + tile, ctx->W), y1 = std::min(y0 + tile, ctx->H);
    for(int by = y0; by < y1; by += 4)
      for(int bx = x0; bx < x1; bx += 8)
        for(int y = by; y < std::min(by + 4, y1); y++)
          for(int x = bx; x < std::min(bx + 8, x1); x++)
            px.push_back((uint32_t)(y * ctx->W + x));
  }
  if(ctx->shardLists.size() >= 64)
  {
    // bound the cache (a host cycling through many shardings): nothing may still be using the old lists
    CTX_CHECK(syncAll(ctx));
    for(auto& l : ctx->shardLists) cudaFree(l.d);
    ctx->shardLists.clear();
    ctx->dPixelOfLocal = nullptr;
  }
  vkrt_ctx::ShardList l{};
  std::memcpy(l.key, key, sizeof(key));
  l.n = (uint32_t)px.size();
  CTX_CHECK(cudaMalloc(&l.d, 4 * std::max<size_t>(px.size(), 1)));
  if(!px.empty())
    CTX_CHECK(cudaMemcpy(l.d, px.data(), 4 * px.size(), cudaMemcpyHostToDevice));  // synchronous: visible to every stream
  ctx->shardLists.push_back(l);
  *out = &ctx->shardLists.back();
  return VKRT_OK;
}

int vkrt_output_shard_size(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, uint32_t* out_pixels)
{
  if(!ctx || !out_pixels) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const vkrt_ctx::ShardList* l = nullptr;
  int rc = shardList(ctx, tile_size, shard_rank, shard_count, &l);
  if(rc != VKRT_OK) return rc;
  *out_pixels = l->n;
  return VKRT_OK;
}

int vkrt_output_pack_shard(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, void* device_compact)
{
  if(!ctx || !device_compact) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const vkrt_ctx::ShardList* l = nullptr;
  int rc = shardList(ctx, tile_size, shard_rank, shard_count, &l);
  if(rc != VKRT_OK) return rc;
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  launchShardCopy(ctx->stream, ctx->image, l->d, l->n, (float4*)device_compact, /*toImage=*/false);
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  ctx->stats.kernel_launches++;
  CTX_CHECK(cudaGetLastError());
  return VKRT_OK;
}

int vkrt_output_unpack_shard(vkrt_ctx* ctx, int32_t tile_size, int32_t shard_rank, int32_t shard_count, const void* device_compact)
{
  if(!ctx || !device_compact) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  const vkrt_ctx::ShardList* l = nullptr;
  int rc = shardList(ctx, tile_size, shard_rank, shard_count, &l);
  if(rc != VKRT_OK) return rc;
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  launchShardCopy(ctx->stream, ctx->image, l->d, l->n, (float4*)device_compact, /*toImage=*/true);
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  ctx->stats.kernel_launches++;
  CTX_CHECK(cudaGetLastError());
  return VKRT_OK;
}

// ---- peer image: shards stored straight into the root's image by the accumulate kernel (include/vkrt_b200.h) ----------
int vkrt_output_export(vkrt_ctx* ctx, uint8_t handle[VKRT_IPC_HANDLE_BYTES])
{
  static_assert(sizeof(cudaIpcMemHandle_t) == VKRT_IPC_HANDLE_BYTES, "IPC handle size");
  if(!ctx || !handle) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if(!ctx->image || !ctx->ownsImage)
    return fail(ctx, VKRT_ERR_INVALID, "output_export: needs a library-owned image (vkrt_output_create with device_rgba32f == NULL)");
  cudaIpcMemHandle_t h;
  CTX_CHECK(cudaIpcGetMemHandle(&h, ctx->image));
  memcpy(handle, &h, sizeof(h));
  return VKRT_OK;
}

int vkrt_peer_open(vkrt_ctx* ctx, const uint8_t handle[VKRT_IPC_HANDLE_BYTES], void** out_device_ptr)
{
  if(!ctx || !handle || !out_device_ptr) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CTX_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *out_device_ptr = p;
  return VKRT_OK;
}

int vkrt_peer_close(vkrt_ctx* ctx, void* device_ptr)
{
  if(!ctx || !device_ptr) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CTX_CHECK(syncAll(ctx));  // no accumulate kernel may still store through the mapping
  if(ctx->mirror == (float4*)device_ptr)
    ctx->mirror = nullptr;
  CTX_CHECK(cudaIpcCloseMemHandle(device_ptr));
  return VKRT_OK;
}

int vkrt_output_set_mirror(vkrt_ctx* ctx, void* device_rgba32f)
{
  if(!ctx) return VKRT_ERR_INVALID;
  if(device_rgba32f && !ctx->image)
    return fail(ctx, VKRT_ERR_INVALID, "output_set_mirror: no output image yet");
  ctx->mirror = (float4*)device_rgba32f;  // read by the launches enqueued from now on (a kernel parameter)
  return VKRT_OK;
}

int vkrt_output_fence(vkrt_ctx* ctx)
{
  if(!ctx) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CTX_CHECK(imageBegin(ctx, ctx->stream));
  CTX_CHECK(imageEnd(ctx, ctx->stream));
  return VKRT_OK;
}

// ---- render --------------------------------------------------------------------------------------------
static int ensureCounters(vkrt_ctx* ctx, WfSlot& sl, uint32_t levels)
{
  if(levels <= sl.counterLevels && sl.wf.counters)
    return VKRT_OK;
  CTX_CHECK(syncAll(ctx));
  cudaFree(sl.wf.counters);
  sl.wf.counters = nullptr;
  levels         = std::max<uint32_t>(levels, 16u);
  CTX_CHECK(cudaMalloc(&sl.wf.counters, 16ull * levels));
  CTX_CHECK(cudaMemsetAsync(sl.wf.counters, 0, 16ull * levels, sl.s));
  sl.counterLevels = levels;
  return VKRT_OK;
}

static int ensureSlotSort(vkrt_ctx* ctx, WfSlot& sl)
{
  if(sl.sortCapacity >= sl.capacity)
    return VKRT_OK;
  CTX_CHECK(syncAll(ctx));
  cudaFree(sl.sortQueue); cudaFree(sl.sortScratch);
  sl.sortQueue = nullptr; sl.sortScratch = nullptr; sl.sortCapacity = 0;
  CTX_CHECK(cudaMalloc(&sl.sortQueue, 4 * sl.capacity));
  CTX_CHECK(cudaMalloc(&sl.sortScratch, sortScratchBytes((uint32_t)sl.capacity)));
  sl.sortCapacity = sl.capacity;
  return VKRT_OK;
}

static int ensureSlot(vkrt_ctx* ctx, WfSlot& sl, size_t paths)
{
  if(paths <= sl.capacity)
    return VKRT_OK;
  CTX_CHECK(syncAll(ctx));  // nothing in flight may still use the buffers about to be replaced
  freeSlot(sl);
  Wavefront& w = sl.wf;
  const size_t b = paths * sizeof(float4);
  CTX_CHECK(cudaMalloc(&w.rayO, b)); CTX_CHECK(cudaMalloc(&w.rayD, b)); CTX_CHECK(cudaMalloc(&w.hit, b));
  CTX_CHECK(cudaMalloc(&w.thr, b)); CTX_CHECK(cudaMalloc(&w.rad, b)); CTX_CHECK(cudaMalloc(&w.absb, b));
  CTX_CHECK(cudaMalloc(&w.sum, b)); CTX_CHECK(cudaMalloc(&w.shO, b)); CTX_CHECK(cudaMalloc(&w.shD, b));
  CTX_CHECK(cudaMalloc(&w.shC, b));
  CTX_CHECK(cudaMalloc(&w.queueCur, 4 * paths)); CTX_CHECK(cudaMalloc(&w.queueNext, 4 * paths));
  w.counters = nullptr;  // sized by ensureCounters (depends on maxDepth)
  sl.counterLevels = 0;
  sl.capacity = paths;
  return VKRT_OK;
}

// first CUDA error of the event plumbing inside the render loop
struct ErrLatch
{
  cudaError_t e = cudaSuccess;
  void operator()(cudaError_t x) { if(e == cudaSuccess) e = x; }
};

int vkrt_render(vkrt_ctx* ctx, const RtxState* state, int32_t num_frames, const vkrt_render_opts* opts)
{
  if(!ctx || !state || num_frames <= 0) return VKRT_ERR_INVALID;
  if(!ctx->committed) return fail(ctx, VKRT_ERR_STATE, "render: scene not committed");
  if(!ctx->hasEnv) return fail(ctx, VKRT_ERR_STATE, "render: no environment set (the reference always loads one, src/main.cpp:181)");
  if(!ctx->image) return fail(ctx, VKRT_ERR_STATE, "render: no output image");
  if(state->size[0] != ctx->W || state->size[1] != ctx->H)
    return fail(ctx, VKRT_ERR_INVALID, "render: RtxState.size differs from the output image");
  if(state->maxSamples < 1 || state->maxDepth < 0)
    return fail(ctx, VKRT_ERR_INVALID, "render: bad maxSamples/maxDepth");
  // SceneCamera.nbLights indexes the light array on the device (pathtrace.glsl:124): it must not exceed what was uploaded
  if(ctx->view.camera.nbLights < 0 || (uint32_t)ctx->view.camera.nbLights > ctx->numLightsUploaded)
    return fail(ctx, VKRT_ERR_INVALID, "render: SceneCamera.nbLights exceeds the number of lights given to vkrt_scene_set_lights");
  cudaSetDevice(ctx->device);

  vkrt_render_opts o{};
  if(opts) o = *opts;
  const int  tile     = o.tile_size > 0 ? o.tile_size : 64;
  const int  shards   = o.shard_count > 1 ? o.shard_count : 1;
  const int  rank     = shards > 1 ? o.shard_rank : 0;
  const bool heatmap  = state->debugging_mode == VKRT_DBG_HEATMAP;  // per-pixel traversal work instead of a clock (vkrt_shade.cu)
  const bool counting = o.count_traversal != 0 || heatmap;
  const bool timing   = o.profile != 0;  // per-kernel CUDA-event timing
  // per-kernel timing needs the kernels one after the other; otherwise shadow(b) overlaps closest(b+1) on a second stream and
  // consecutive batches rotate through the wavefront slots
  static const bool overlapEnv = []() { const char* e = getenv("VKRT_OVERLAP"); return !e || atoi(e) != 0; }();
  static const bool dualEnv    = []() { const char* e = getenv("VKRT_DUAL_SLOT"); return !e || atoi(e) != 0; }();
  const bool overlap  = overlapEnv && !timing && !heatmap;
  const bool dual     = overlap && dualEnv;
  // VKRT_SHADING=precise forces the IEEE-exact shade kernel for callers that cannot pass options (the C++ tools in tests)
  static const bool envPrecise = []() { const char* e = getenv("VKRT_SHADING"); return e && !strcmp(e, "precise"); }();
  if(envPrecise) o.precise_shading = 1;
  if(rank < 0 || rank >= shards)
    return fail(ctx, VKRT_ERR_INVALID, "render: shard_rank out of range");

  const vkrt_ctx::ShardList* list = nullptr;
  int rc = shardList(ctx, tile, rank, shards, &list);
  if(rc != VKRT_OK) return rc;
  const uint32_t  nLocal = list->n;
  const uint32_t* dPix   = list->d;
  if(nLocal == 0)
    return VKRT_OK;
  int fif = o.frames_in_flight;
  const bool autoFif = fif <= 0;
  if(autoFif)
  {
    // automatic: ~64 M paths in flight (11 GB of wavefront state over the slots) amortise the latency floor of thin late
    // bounces -- but never more frames than this call renders (a single interactive frame allocates for one frame), and
    // never more than half of the memory that is free right now (168 bytes of wavefront state per path, + 12 with the ray sort)
    fif = (int)std::max<size_t>(1, std::min<size_t>(64, (size_t)(64u << 20) / nLocal));
    fif = std::min(fif, num_frames);
  }
  // with n slots a batch is 1/n of the frames in flight, so that n batches exist to overlap
  const int nSlots   = dual ? (num_frames == 1 ? ctx->numSlotsInteractive : ctx->numSlots) : 1;
  int       perBatch = fif;
  if(nSlots > 1 && std::min(fif, num_frames) >= 2)
    perBatch = (std::min(fif, num_frames) + nSlots - 1) / nSlots;
  const int slotsNow = std::min(nSlots, (num_frames + perBatch - 1) / perBatch);  // slots this call fills (a one-frame call fills one; the next call the next)
  if(autoFif)
  {
    size_t have = 0;
    for(int k = 0; k < slotsNow; k++) have += ctx->slots[k].capacity;
    size_t freeB = 0, totalB = 0;
    if((size_t)nLocal * perBatch * slotsNow > have && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess)
    {
      const size_t perPath = 10 * sizeof(float4) + 8 + (o.sort_rays ? 8 : 0);
      const size_t fit     = (freeB / 2 + have * perPath) / (perPath * (size_t)nLocal * slotsNow);
      perBatch             = (int)std::max<size_t>(1, std::min<size_t>((size_t)perBatch, fit));
    }
  }
  // an explicit frames_in_flight is reserved in full (1/n per slot) even when this call renders fewer frames, so that a
  // later, longer call does not reallocate gigabytes in the middle of a render (grow-only)
  const int batchSlots = dual ? ctx->numSlots : 1;  // slots a multi-frame call uses
  const int reserve    = std::max(perBatch, autoFif ? 0 : (fif + batchSlots - 1) / batchSlots);
  if((uint64_t)nLocal * (uint64_t)reserve > 0xffffffffull)
    return fail(ctx, VKRT_ERR_INVALID, "render: pixels x frames_in_flight exceeds 2^32 paths");
  if(slotsNow > 1 || !autoFif)
    for(int k = 0; k < (autoFif ? slotsNow : batchSlots) && rc == VKRT_OK; k++)
      rc = ensureSlot(ctx, ctx->slots[k], (size_t)nLocal * reserve);
  if(rc != VKRT_OK) return rc;
  perBatch = std::min(perBatch, num_frames);

  TraceStats* dstats = ctx->dStats;
  ErrLatch    ev;
  for(int done = 0; done < num_frames;)
  {
    const int batch = std::min(perBatch, num_frames - done);
    WfSlot&   sl    = ctx->slots[nSlots > 1 ? (ctx->nextSlot++ % (uint32_t)nSlots) : 0u];
    rc = ensureSlot(ctx, sl, (size_t)nLocal * perBatch);  // no-op once the slot is large enough
    if(rc == VKRT_OK && o.sort_rays)
      rc = ensureSlotSort(ctx, sl);
    if(rc == VKRT_OK)
      rc = ensureCounters(ctx, sl, (uint32_t)state->maxDepth + 2u);
    if(rc != VKRT_OK) return rc;
    cudaStream_t st = timing ? ctx->stream : sl.s;  // ScopedTimer records on ctx->stream (== slots[0].s; timing is single slot)
    Wavefront wf      = sl.wf;
    wf.numPaths       = nLocal * (uint32_t)batch;
    wf.numLocalPixels = nLocal;
    wf.pixelOfLocal   = dPix;
    wf.clevels        = (uint32_t)state->maxDepth + 2u;
    wf.cbase          = 0;
    uint32_t* sortAlt = sl.sortQueue;  // third queue buffer of the optional ray sort (rotates with queueCur)
    // Bounce from which the fused tail kernel takes over (vkrt_render_opts.tail_from, else VKRT_TAIL_FROM: 0 never, n > 0 from bounce n; default: from bounce 3
    // for small wavefronts -- an interactive frame, a thin multi-GPU shard -- where launch gaps and kernel tails outweigh the
    // less efficient one-thread-per-path traversal; large batches keep one launch per stage).  Same image either way.
    static const int tailEnv = []() { const char* e = getenv("VKRT_TAIL_FROM"); return e ? atoi(e) : -1; }();
    static const long long tailAutoPaths = []() { const char* e = getenv("VKRT_TAIL_AUTO_PATHS"); return e ? atoll(e) : (long long)(9u << 20); }();
    int tailFrom = 1 << 30;
    if(!counting && !heatmap)
    {
      const int want = o.tail_from != 0 ? o.tail_from : (tailEnv == 0 ? -1 : (tailEnv > 0 ? tailEnv : 0));  // options, then environment
      if(want > 0)
        tailFrom = want;
      else if(want == 0 && (long long)wf.numPaths <= tailAutoPaths)
        tailFrom = 3;
    }
    const bool  fuseFinish = state->maxSamples == 1 && !heatmap;  // one sample per frame: the accumulate kernel applies the firefly clamp itself
    FrameParams fp{};
    fp.state         = *state;
    fp.state.frame   = state->frame + done;
    fp.seedMode      = o.seed_mode;
    fp.framesInBatch = batch;
    for(int smpl = 0; smpl < state->maxSamples; smpl++)
    {
      fp.sampleIndex = smpl;
      fp.depth       = 0;
      {
        ScopedTimer t(ctx, timing, T_RAYGEN);
        launchRaygen(st, ctx->view, wf, fp, ctx->numSMs);
      }
      ctx->stats.kernel_launches++;
      for(int depth = 0; depth < state->maxDepth; depth++)
      {
        fp.depth = depth;
        wf.cbase = 4u * (uint32_t)depth;
        if(depth >= tailFrom)
        {
          // fused tail: every remaining bounce of the few paths still alive in one launch (tailKernel, vkrt_shade.cu)
          if(overlap && depth > 0)
            ev(cudaStreamWaitEvent(st, sl.evShadow, 0));  // the radiance the previous bounce's shadow rays resolve into
          ScopedTimer t(ctx, timing, T_TAIL);
          if(o.precise_shading)
            launchTailPrecise(st, ctx->view, wf, fp, dstats, ctx->numSMs);
          else
            launchTailFast(st, ctx->view, wf, fp, dstats, ctx->numSMs);
          ctx->stats.kernel_launches++;
          break;
        }
        if(o.sort_rays && depth >= 1)
        {
          // order this bounce's queue by (direction octant, origin Morton bucket): three small kernels that read the live count
          // from the wavefront counters -- no host synchronisation.  Primary rays are coherent already.
          ScopedTimer t(ctx, timing, T_SORT);
          launchSortQueue(st, ctx->view, wf.rayO, wf.rayD, wf.queueCur, sortAlt, wf.counters + wf.cbase, (uint32_t)sl.capacity, sl.sortScratch, ctx->numSMs);
          std::swap(wf.queueCur, sortAlt);
          ctx->stats.kernel_launches += 3;
        }
        TraceLaunch L{};
        L.rayO = wf.rayO; L.rayD = wf.rayD; L.queue = wf.queueCur; L.count = wf.counters + wf.cbase; L.cursor = wf.counters + wf.cbase + 1;
        L.hit = wf.hit; L.stats = dstats; L.cost = heatmap ? wf.sum : nullptr;
        {
          ScopedTimer t(ctx, timing, T_CLOSEST);
          ScopedTimer t0(ctx, timing && depth == 0, T_CLOSEST0);
          launchTraceClosest(st, ctx->view, L, counting, ctx->numSMs);
        }
        if(overlap && depth > 0)
          ev(cudaStreamWaitEvent(st, sl.evShadow, 0));  // shade(b) adds to the radiance shadow(b-1) resolves into
        {
          ScopedTimer t(ctx, timing, T_SHADE);
          if(o.precise_shading)
            launchShadePrecise(st, ctx->view, wf, fp, dstats, ctx->numSMs);
          else
            launchShadeFast(st, ctx->view, wf, fp, dstats, ctx->numSMs);
        }
        TraceLaunch S{};
        S.rayO = wf.shO; S.rayD = wf.shD; S.queue = nullptr;
        S.count = wf.counters + wf.cbase + 2; S.cursor = wf.counters + wf.cbase + 3; S.primary = depth == 0 ? 1u : 0u;
        S.shC = wf.shC; S.rad = wf.rad; S.stats = dstats; S.cost = heatmap ? wf.sum : nullptr;
        if(overlap)
        {
          // the any-hit kernel of this bounce runs beside the closest-hit kernel of the next one: the tail of either
          // persistent kernel is filled by the other's CTAs
          ev(cudaEventRecord(sl.evShade, st));
          ev(cudaStreamWaitEvent(sl.s2, sl.evShade, 0));
          launchTraceShadow(sl.s2, ctx->view, S, counting, ctx->numSMs);
          ev(cudaEventRecord(sl.evShadow, sl.s2));
        }
        else
        {
          ScopedTimer t(ctx, timing, T_SHADOW);
          ScopedTimer t0(ctx, timing && depth == 0, T_SHADOW0);
          launchTraceShadow(st, ctx->view, S, counting, ctx->numSMs);
        }
        std::swap(wf.queueCur, wf.queueNext);  // the counts need no swap: every bounce has its own (Wavefront::counters)
        ctx->stats.kernel_launches += 3;
      }
      if(overlap && state->maxDepth > 0)
        ev(cudaStreamWaitEvent(st, sl.evShadow, 0));
      if(!fuseFinish)
      {
        ScopedTimer t(ctx, timing, T_ACCUM);
        launchFinishSample(st, wf, fp, ctx->numSMs);
        ctx->stats.kernel_launches++;
      }
    }
    {
      // the running mean is order dependent (pathtrace.comp:126): batches accumulate in frame order, whichever slot they ran on
      ev(imageBegin(ctx, st));
      ScopedTimer t(ctx, timing, T_ACCUM);
      launchAccumulate(st, wf, fp, ctx->image, ctx->mirror, ctx->numSMs, fuseFinish);
    }
    ev(imageEnd(ctx, st));
    ctx->stats.kernel_launches++;
    done += batch;
    ctx->stats.frames_rendered += batch;
    if(ctx->pending.size() > 4096)
      drainTimers(ctx);
  }
  if(ev.e != cudaSuccess)
  {
    ctx->err = std::string("render: stream / event ordering failed: ") + cudaGetErrorString(ev.e);
    return VKRT_ERR_CUDA;
  }
  CTX_CHECK(cudaGetLastError());
  return VKRT_OK;
}


int vkrt_sync(vkrt_ctx* ctx)
{
  if(!ctx) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CTX_CHECK(syncAll(ctx));
  return VKRT_OK;
}

int vkrt_stream(vkrt_ctx* ctx, void** out_cuda_stream)
{
  if(!ctx || !out_cuda_stream) return VKRT_ERR_INVALID;
  *out_cuda_stream = (void*)ctx->stream;
  return VKRT_OK;
}

// ---- diagnostics ------------------------------------------------------------------------------------------
int vkrt_trace_rays(vkrt_ctx* ctx, const float* rays, uint32_t n, int32_t closest, float* out_hits)
{
  if(!ctx || !rays || !out_hits) return VKRT_ERR_INVALID;
  if(!ctx->committed) return fail(ctx, VKRT_ERR_STATE, "trace_rays: scene not committed");
  if(n == 0) return VKRT_OK;
  cudaSetDevice(ctx->device);
  // de-interleave {o,tmax,d,seed} into the two float4 streams the trace kernels read
  std::vector<float4> ho(n), hd(n);
  for(uint32_t i = 0; i < n; i++)
  {
    const float* r = rays + (size_t)i * 8;
    ho[i] = make_float4(r[0], r[1], r[2], r[3]);
    hd[i] = make_float4(r[4], r[5], r[6], r[7]);
  }
  // one scratch allocation (rays, hits, counters), kept by the context and grown on demand
  const size_t need = (size_t)n * 48 + 16;
  if(need > ctx->probeCapacity)
  {
    cudaFree(ctx->dProbe);
    ctx->dProbe = nullptr;
    ctx->probeCapacity = 0;
    CTX_CHECK(cudaMalloc(&ctx->dProbe, need));
    ctx->probeCapacity = need;
  }
  float4 *dO = (float4*)ctx->dProbe, *dD = dO + n, *dH = dD + n;
  uint32_t* dCnt = (uint32_t*)(dH + n);
  const uint32_t cc[2] = {n, 0u};
  cudaMemcpyAsync(dO, ho.data(), (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dD, hd.data(), (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dCnt, cc, 8, cudaMemcpyHostToDevice, ctx->stream);
  TraceLaunch L{};
  L.rayO = dO; L.rayD = dD; L.queue = nullptr; L.count = dCnt; L.cursor = dCnt + 1; L.hit = dH;
  if(closest)
    launchTraceClosest(ctx->stream, ctx->view, L, false, ctx->numSMs);
  else
    launchTraceAnyProbe(ctx->stream, ctx->view, L, ctx->numSMs);
  ctx->stats.kernel_launches++;
  cudaMemcpyAsync(out_hits, dH, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if(e != cudaSuccess)
  {
    ctx->err = std::string("trace_rays: ") + cudaGetErrorString(e);
    return VKRT_ERR_CUDA;
  }
  return VKRT_OK;
}

int vkrt_triangle_ids(vkrt_ctx* ctx, const uint32_t* tri_ids, uint32_t n, uint32_t* out_node_prim)
{
  if(!ctx || !tri_ids || !out_node_prim) return VKRT_ERR_INVALID;
  if(!ctx->committed) return fail(ctx, VKRT_ERR_STATE, "triangle_ids: scene not committed");
  cudaSetDevice(ctx->device);
  std::vector<uint2>& info = ctx->hostTriInfo;  // fetched once per committed scene
  if(info.size() != ctx->bvh.numTris)
  {
    info.resize(ctx->bvh.numTris);
    if(!info.empty())
      CTX_CHECK(cudaMemcpy(info.data(), ctx->bvh.triInfo, sizeof(uint2) * info.size(), cudaMemcpyDeviceToHost));
  }
  for(uint32_t i = 0; i < n; i++)
  {
    if(tri_ids[i] < info.size()) { out_node_prim[2 * i] = info[tri_ids[i]].x; out_node_prim[2 * i + 1] = info[tri_ids[i]].y; }
    else { out_node_prim[2 * i] = ~0u; out_node_prim[2 * i + 1] = ~0u; }
  }
  return VKRT_OK;
}

int vkrt_mark(vkrt_ctx* ctx, int32_t which)
{
  if(!ctx || which < 0 || which > 1) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if(!ctx->marks[which]) CTX_CHECK(cudaEventCreate(&ctx->marks[which]));
  // a mark is a barrier over all streams of the context: everything enqueued before it has finished when it fires, and
  // nothing enqueued after it starts earlier
  CTX_CHECK(barrierOn(ctx, ctx->stream));
  if(ctx->imagePending) CTX_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->evImage, 0));
  CTX_CHECK(cudaEventRecord(ctx->marks[which], ctx->stream));
  CTX_CHECK(forkFrom(ctx, ctx->stream));
  return VKRT_OK;
}

int vkrt_elapsed_ms(vkrt_ctx* ctx, float* out_ms)
{
  if(!ctx || !out_ms || !ctx->marks[0] || !ctx->marks[1]) return VKRT_ERR_STATE;
  cudaSetDevice(ctx->device);
  CTX_CHECK(cudaEventSynchronize(ctx->marks[1]));
  CTX_CHECK(cudaEventElapsedTime(out_ms, ctx->marks[0], ctx->marks[1]));
  return VKRT_OK;
}

int vkrt_get_stats(vkrt_ctx* ctx, vkrt_stats* out)
{
  if(!ctx || !out) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CTX_CHECK(syncAll(ctx));
  drainTimers(ctx);
  TraceStats ts{};
  CTX_CHECK(cudaMemcpy(&ts, ctx->dStats, sizeof(ts), cudaMemcpyDeviceToHost));
  ctx->stats.closest_rays  = ts.closestRays;
  ctx->stats.shadow_rays   = ts.shadowRays;
  ctx->stats.nodes_visited = ts.nodesVisited;
  ctx->stats.tris_tested   = ts.trisTested;
  ctx->stats.shadow_nodes_visited = ts.shadowNodes;
  ctx->stats.shadow_tris_tested   = ts.shadowTris;
  ctx->stats.shadow_rays_primary  = ts.shadowRaysPrimary;
  ctx->stats.bvh_max_depth        = (uint32_t)ctx->bvh.maxDepth;
  *out = ctx->stats;
  return VKRT_OK;
}

int vkrt_reset_stats(vkrt_ctx* ctx)
{
  if(!ctx) return VKRT_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CTX_CHECK(syncAll(ctx));
  drainTimers(ctx);
  CTX_CHECK(cudaMemset(ctx->dStats, 0, sizeof(TraceStats)));
  vkrt_stats keep = ctx->stats;
  ctx->stats      = vkrt_stats{};
  ctx->stats.ms_build = keep.ms_build; ctx->stats.num_triangles = keep.num_triangles; ctx->stats.num_bvh_nodes = keep.num_bvh_nodes;
  ctx->stats.num_instances = keep.num_instances; ctx->stats.has_non_opaque = keep.has_non_opaque;
  ctx->stats.bvh_max_depth = keep.bvh_max_depth; ctx->stats.node_bytes = keep.node_bytes; ctx->stats.bvh_width = keep.bvh_width;
  return VKRT_OK;
}

}  // extern "C"
