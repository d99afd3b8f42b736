// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or imported by the product.
//
// oracle.cpp: C entry points (ctypes) of the CPU restatement of the vk_raytrace hot
// path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or runnable CPU
// path, and its Vulkan back-ends cannot be built here.  This file
// follows the reference sources line by line (citations in the o_*.h headers);
// what can be pinned independently (TEA/PCG known answers, oct round trips, alias
// table invariants) is pinned in tests/golden/.
//
// Also restated here:
//   alias table        src/hdr_sampling.cpp:107-248
//   tonemap            shaders/post.frag:98-147, shaders/tonemapping.glsl:24-65
#include <thread>
#include <atomic>
#include <numeric>
#include <mutex>
#include "o_pathtrace.h"

using namespace orc;

struct OracleCtx
{
  Scene              scene;
  int                width = 0, height = 0;
  std::vector<float> image;  // RGBA32F accumulation
  int                threads = 0;
  RenderCounters     counters;
  bool               committed = false;
};

// ---- src/hdr_sampling.cpp:107-176 -------------------------------------------------
static float buildAliasmap(const std::vector<float>& data, std::vector<EnvAccel>& accel)
{
  uint32_t size = (uint32_t)data.size();
  float    sum  = std::accumulate(data.begin(), data.end(), 0.f);
  float    fSize          = (float)size;
  float    inverseAverage = fSize / sum;
  for(uint32_t i = 0; i < size; ++i)
  {
    accel[i].q     = data[i] * inverseAverage;
    accel[i].alias = i;
  }
  std::vector<uint32_t> partitionTable(size);
  uint32_t              s = 0u, large = size;
  for(uint32_t i = 0; i < size; ++i)
  {
    if(accel[i].q < 1.f)
      partitionTable[s++] = i;
    else
      partitionTable[--large] = i;
  }
  for(s = 0; s < large && large < size; ++s)
  {
    const uint32_t smallEnergyIndex = partitionTable[s];
    const uint32_t highEnergyIndex  = partitionTable[large];
    accel[smallEnergyIndex].alias   = highEnergyIndex;
    const float differenceWithAverage = 1.f - accel[smallEnergyIndex].q;
    accel[highEnergyIndex].q -= differenceWithAverage;
    if(accel[highEnergyIndex].q < 1.0f)
      large++;
  }
  return sum;
}

// ---- src/hdr_sampling.cpp:187-248 -------------------------------------------------
static void createEnvironmentAccel(const float* pixels, uint32_t rx, uint32_t ry, EnvAccel* out, float* integral, float* average)
{
  std::vector<EnvAccel> envAccel(rx * ry);
  std::vector<float>    importanceData(rx * ry);
  float                 cosTheta0 = 1.0f;
  const float           stepPhi   = float(2.0 * M_PI) / float(rx);
  const float           stepTheta = float(M_PI) / float(ry);
  double                total     = 0;
  for(uint32_t y = 0; y < ry; ++y)
  {
    const float theta1    = float(y + 1) * stepTheta;
    const float cosTheta1 = std::cos(theta1);
    const float area      = (cosTheta0 - cosTheta1) * stepPhi;
    cosTheta0             = cosTheta1;
    for(uint32_t x = 0; x < rx; ++x)
    {
      const uint32_t idx  = y * rx + x;
      const uint32_t idx4 = idx * 4;
      float cieLuminance  = pixels[idx4] * 0.2126f + pixels[idx4 + 1] * 0.7152f + pixels[idx4 + 2] * 0.0722f;
      importanceData[idx] = area * std::max(pixels[idx4], std::max(pixels[idx4 + 1], pixels[idx4 + 2]));
      total += cieLuminance;
    }
  }
  float avg  = static_cast<float>(total) / static_cast<float>(rx * ry);
  float integ = buildAliasmap(importanceData, envAccel);
  const float invEnvIntegral = 1.0f / integ;
  for(uint32_t i = 0; i < rx * ry; ++i)
  {
    const uint32_t idx4 = i * 4;
    envAccel[i].pdf     = std::max(pixels[idx4], std::max(pixels[idx4 + 1], pixels[idx4 + 2])) * invEnvIntegral;
  }
  for(uint32_t i = 0; i < rx * ry; ++i)
    envAccel[i].aliasPdf = envAccel[envAccel[i].alias].pdf;
  std::memcpy(out, envAccel.data(), sizeof(EnvAccel) * rx * ry);
  if(integral) *integral = integ;
  if(average) *average = avg;
}

// ---- shaders/tonemapping.glsl + shaders/post.frag ------------------------------------
static vec3 linearTosRGB(vec3 c) { return pow3(c, V3(1.0f / 2.2f)); }
static vec3 sRGBToLinear(vec3 c) { return pow3(c, V3(2.2f)); }
static vec3 toneMapUncharted2Impl(vec3 color)
{
  const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
  return ((color * (A * color + C * B) + D * E) / (color * (A * color + B) + D * F)) - E / F;
}
static vec3 toneMapUncharted(vec3 color)
{
  const float W   = 11.2f;
  color           = toneMapUncharted2Impl(color * 2.0f);
  vec3 whiteScale = V3(1.0f) / toneMapUncharted2Impl(V3(W));
  return linearTosRGB(color * whiteScale);
}
static vec3 dither(vec3 linear_color, vec3 noise, float quant)
{
  vec3 c0    = floor3(linearTosRGB(linear_color) / quant) * quant;
  vec3 c1    = c0 + quant;
  vec3 discr = mix(sRGBToLinear(c0), sRGBToLinear(c1), noise);
  // mix(c0, c1, lessThan(discr, linear_color))
  return V3(discr.x < linear_color.x ? c1.x : c0.x, discr.y < linear_color.y ? c1.y : c0.y, discr.z < linear_color.z ? c1.z : c0.z);
}
// ---- mip pyramid for auto-exposure: RenderOutput::genMipmap (src/render_output.cpp:188-193) = nvvk::cmdGenerateMipmaps,
// i.e. vkCmdBlitImage(level k -> k+1, VK_FILTER_LINEAR) with extent max(1, size >> 1) per level.  A scaling blit samples the
// source at u = (x + 0.5) * srcW / dstW with linear filtering and clamp-to-edge.
struct MipChain
{
  std::vector<std::vector<float>> level;  // RGBA32F
  std::vector<int>                w, h;
};
static MipChain buildMipChain(const float* image, int W, int H)
{
  MipChain m;
  m.level.emplace_back(image, image + size_t(W) * H * 4);
  m.w.push_back(W); m.h.push_back(H);
  while(m.w.back() > 1 || m.h.back() > 1)
  {
    const int sw = m.w.back(), sh = m.h.back(), dw = std::max(1, sw >> 1), dh = std::max(1, sh >> 1);
    const std::vector<float>& src = m.level.back();
    std::vector<float> dst(size_t(dw) * dh * 4);
    for(int y = 0; y < dh; y++)
      for(int x = 0; x < dw; x++)
      {
        const float u = ((float)x + 0.5f) * ((float)sw / (float)dw) - 0.5f, v = ((float)y + 0.5f) * ((float)sh / (float)dh) - 0.5f;
        const float x0f = std::floor(u), y0f = std::floor(v), ax = u - x0f, ay = v - y0f;
        const int   x0 = std::min(sw - 1, std::max(0, (int)x0f)), x1 = std::min(sw - 1, std::max(0, (int)x0f + 1));
        const int   y0 = std::min(sh - 1, std::max(0, (int)y0f)), y1 = std::min(sh - 1, std::max(0, (int)y0f + 1));
        for(int c = 0; c < 4; c++)
        {
          const float c00 = src[(size_t(y0) * sw + x0) * 4 + c], c10 = src[(size_t(y0) * sw + x1) * 4 + c];
          const float c01 = src[(size_t(y1) * sw + x0) * 4 + c], c11 = src[(size_t(y1) * sw + x1) * 4 + c];
          const float top = c00 * (1.0f - ax) + c10 * ax, bot = c01 * (1.0f - ax) + c11 * ax;
          dst[(size_t(y) * dw + x) * 4 + c] = top * (1.0f - ay) + bot * ay;
        }
      }
    m.level.push_back(std::move(dst));
    m.w.push_back(dw); m.h.push_back(dh);
  }
  return m;
}
// NEAREST fetch of mip `lod` at uv (the post sampler is a default VkSamplerCreateInfo: NEAREST filters, NEAREST mipmap mode,
// maxLod = FLT_MAX, src/render_output.cpp:96-98); lod is clamped to the chain.
static vec3 mipFetch(const MipChain& m, vec2 uv, int lod)
{
  lod = std::min(std::max(lod, 0), (int)m.level.size() - 1);
  const int w = m.w[lod], h = m.h[lod];
  const int x = std::min(w - 1, std::max(0, (int)std::floor(uv.x * (float)w))), y = std::min(h - 1, std::max(0, (int)std::floor(uv.y * (float)h)));
  const float* p = &m.level[lod][(size_t(y) * w + x) * 4];
  return V3(p[0], p[1], p[2]);
}
static float postLuminance(vec3 c) { return dot(c, V3(0.2126f, 0.7152f, 0.0722f)); }  // post.frag:58-61
// post.frag:57 RGB2XYZ is built column-major from the row-major literature matrix, so XYZ.y uses the second *column*
static float xyzY(vec3 rgb) { return (0.3575761f * rgb.x + 0.7151522f * rgb.y) + 0.1191920f * rgb.z; }
// post.frag:64-70
static vec3 toneExposure(const Tonemapper& tm, vec3 RGB, float logAvgLum)
{
  const float XYZy = xyzY(RGB);
  const float Y    = (tm.key / logAvgLum) * XYZy;
  const float Yd   = (Y * (1.0f + Y / (tm.Ywhite * tm.Ywhite))) / (1.0f + Y);
  return RGB / XYZy * Yd;
}
// post.frag:72-95
static vec3 toneLocalExposure(const Tonemapper& tm, const MipChain& m, vec2 uv, vec3 RGB, float logAvgLum)
{
  const float XYZy = xyzY(RGB);
  const float Y    = (tm.key / logAvgLum) * XYZy;
  float       La   = 0.f;
  const float factor = tm.key / logAvgLum;
  const float epsilon = 0.05f, phi = 2.0f;
  const float scale[7] = {1, 2, 4, 8, 16, 32, 64};
  for(int i = 0; i < 7; ++i)
  {
    const float v1 = postLuminance(mipFetch(m, uv, i)) * factor;
    const float v2 = postLuminance(mipFetch(m, uv, i + 1)) * factor;
    if(std::fabs(v1 - v2) / ((tm.key * std::pow(2.0f, phi) / (scale[i] * scale[i])) + v1) > epsilon)
    {
      La = v1;
      break;
    }
    else
      La = v2;
  }
  const float Yd = Y / (1.0f + La);
  return RGB / XYZy * Yd;
}

// post.frag:98-147
static vec4 postPixel(const Tonemapper& tm, const float* image, int W, int H, int x, int y, const MipChain* mips)
{
  vec2 uvCoords = V2(((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H);
  // NEAREST sampling at uvCoords * zoom (src/render_output.cpp:96-98: default sampler)
  int sx = std::min(W - 1, std::max(0, (int)std::floor(uvCoords.x * tm.zoom * (float)W)));
  int sy = std::min(H - 1, std::max(0, (int)std::floor(uvCoords.y * tm.zoom * (float)H)));
  const float* p = image + (size_t(sy) * W + sx) * 4;
  vec3 hdr = V3(p[0], p[1], p[2]);
  if(((tm.autoExposure >> 0) & 1) == 1 && mips)
  {
    const vec3  avg     = mipFetch(*mips, V2(0.5f, 0.5f), 20);  // textureLod(inImage, vec2(0.5), 20): the 1x1 top level
    const float avgLum2 = postLuminance(avg);
    if(((tm.autoExposure >> 1) & 1) == 1)
      hdr = toneLocalExposure(tm, *mips, V2(uvCoords.x * tm.zoom, uvCoords.y * tm.zoom), hdr, avgLum2);
    else
      hdr = toneExposure(tm, hdr, avgLum2);
  }
  vec3 color = toneMapUncharted(hdr * tm.avgLum);
  if(tm.dither > 0)
  {
    uint32_t r[3] = {(uint32_t)x, (uint32_t)y, 0u};  // uvec3(gl_FragCoord.xy, 0)
    pcg3d(r);
    vec3 noise = V3(uintBitsToFloat(0x3f800000u | (r[0] >> 9)) - 1.0f, uintBitsToFloat(0x3f800000u | (r[1] >> 9)) - 1.0f,
                    uintBitsToFloat(0x3f800000u | (r[2] >> 9)) - 1.0f);
    color = dither(sRGBToLinear(color), noise, 1.f / 255.f);
  }
  color = clamp3(mix(V3(0.5f), color, tm.contrast), 0, 1);
  color = pow3(color, V3(1.0f / tm.brightness));
  vec3 i = V3(dot(color, V3(0.299f, 0.587f, 0.114f)));
  color  = mix(i, color, tm.saturation);
  vec2 uv = V2(((uvCoords.x * tm.renderingRatio[0]) - 0.5f) * 2.0f, ((uvCoords.y * tm.renderingRatio[1]) - 0.5f) * 2.0f);
  color *= 1.0f - dot(uv, uv) * tm.vignette;
  return V4(color, p[3]);
}

extern "C" {

int oracle_create(OracleCtx** out)
{
  *out = new OracleCtx();
  // shaders/host_device.h SunAndSky defaults are supplied by the caller; in_use = 0 here
  return 0;
}
int oracle_destroy(OracleCtx* c)
{
  delete c;
  return 0;
}
int oracle_set_threads(OracleCtx* c, int n)
{
  c->threads = n;
  return 0;
}
int oracle_scene_begin(OracleCtx* c)
{
  c->scene.meshes.clear(); c->scene.nodes.clear(); c->scene.materials.clear(); c->scene.lights.clear();
  c->scene.textures.clear(); c->scene.tris.clear(); c->scene.bvh.clear();
  c->committed = false;
  return 0;
}
int oracle_scene_add_primmesh(OracleCtx* c, const VertexAttributes* v, uint32_t nv, const uint32_t* idx, uint32_t ni, int32_t material)
{
  PrimMesh pm;
  pm.vtx.assign(v, v + nv);
  pm.idx.assign(idx, idx + ni);
  pm.material = material;
  for(uint32_t i = 0; i < ni; i++)
    if(idx[i] >= nv)
      return 1;
  c->scene.meshes.push_back(std::move(pm));
  return 0;
}
int oracle_scene_add_node(OracleCtx* c, const float world[16], uint32_t primMesh)
{
  if(primMesh >= c->scene.meshes.size())
    return 1;
  Node n{};
  std::memcpy(n.world, world, 64);
  n.primMesh = primMesh;
  c->scene.nodes.push_back(n);
  return 0;
}
int oracle_scene_set_materials(OracleCtx* c, const GltfShadeMaterial* m, uint32_t n)
{
  c->scene.materials.assign(m, m + n);
  return 0;
}
int oracle_scene_set_lights(OracleCtx* c, const Light* l, uint32_t n)
{
  c->scene.lights.assign(l, l + n);
  if(c->scene.lights.empty())
    c->scene.lights.push_back(Light{});  // src/scene.cpp:329-330
  return 0;
}
// src/scene.cpp:447-482 sampler mapping
int oracle_scene_add_texture(OracleCtx* c, const uint8_t* rgba, uint32_t w, uint32_t h, int32_t magFilter, int32_t minFilter, int32_t wrapS, int32_t wrapT)
{
  (void)minFilter;
  Texture t;
  t.rgba.assign(rgba, rgba + size_t(w) * h * 4);
  t.w = w; t.h = h;
  t.linear = (magFilter == 9729 || magFilter == 9985 || magFilter == 9987) ? 1 : 0;
  auto wrap = [](int32_t m) { return m == 33071 ? 1 : (m == 33648 ? 2 : 0); };
  t.wrapS = wrap(wrapS);
  t.wrapT = wrap(wrapT);
  c->scene.textures.push_back(std::move(t));
  return 0;
}
int oracle_scene_commit(OracleCtx* c, int32_t use_bvh)
{
  if(c->scene.materials.empty())
    return 3;
  if(c->scene.lights.empty())
    c->scene.lights.push_back(Light{});
  c->scene.useBvh = use_bvh != 0;
  commitScene(c->scene);
  c->committed = true;
  return 0;
}
int oracle_build_env_accel(const float* rgba, uint32_t w, uint32_t h, EnvAccel* out, float* integral, float* average)
{
  createEnvironmentAccel(rgba, w, h, out, integral, average);
  return 0;
}
int oracle_set_environment(OracleCtx* c, const float* rgba, uint32_t w, uint32_t h, const EnvAccel* accel)
{
  c->scene.envPixels.assign(rgba, rgba + size_t(w) * h * 4);
  c->scene.envW = w; c->scene.envH = h;
  c->scene.envAccel.resize(size_t(w) * h);
  if(accel)
    std::memcpy(c->scene.envAccel.data(), accel, sizeof(EnvAccel) * size_t(w) * h);
  else
    createEnvironmentAccel(rgba, w, h, c->scene.envAccel.data(), nullptr, nullptr);
  return 0;
}
int oracle_set_sunsky(OracleCtx* c, const SunAndSky* s)
{
  c->scene.sunsky = *s;
  return 0;
}
int oracle_set_camera(OracleCtx* c, const SceneCamera* cam)
{
  c->scene.camera = *cam;
  return 0;
}
int oracle_output_create(OracleCtx* c, uint32_t w, uint32_t h)
{
  c->width = (int)w; c->height = (int)h;
  c->image.assign(size_t(w) * h * 4, 0.f);
  return 0;
}
int oracle_output_read(OracleCtx* c, float* out)
{
  std::memcpy(out, c->image.data(), c->image.size() * 4);
  return 0;
}
int oracle_output_write(OracleCtx* c, const float* in)
{
  std::memcpy(c->image.data(), in, c->image.size() * 4);
  return 0;
}

// Tile-sharded variant (mirrors vkrt_render_opts.tile_size/shard_rank/shard_count): only pixels of tiles t with
// t % shard_count == shard_rank are rendered; everything else is left untouched.
int oracle_render_tiles(OracleCtx* c, const RtxState* state, int32_t num_frames, int32_t seed_mode, int32_t tile, int32_t shard_rank, int32_t shard_count)
{
  if(!c->committed || c->scene.envW == 0)
    return 3;
  const int W = state->size[0], H = state->size[1];
  if(W != c->width || H != c->height || tile <= 0 || shard_count <= 0)
    return 1;
  const int tx = (W + tile - 1) / tile;
  int nt = c->threads > 0 ? c->threads : (int)std::thread::hardware_concurrency();
  if(nt < 1) nt = 1;
  for(int f = 0; f < num_frames; f++)
  {
    RtxState st = *state;
    st.frame    = state->frame + f;
    std::atomic<int> next{0};
    std::mutex       mtx;
    auto work = [&]() {
      RenderCounters local;
      for(;;)
      {
        int y = next.fetch_add(1);
        if(y >= H)
          break;
        for(int x = 0; x < W; x++)
        {
          int t = (y / tile) * tx + (x / tile);
          if(t % shard_count == shard_rank)
            renderPixel(c->scene, st, seed_mode, x, y, c->image.data(), &local);
        }
      }
      std::lock_guard<std::mutex> g(mtx);
      c->counters.closestRays += local.closestRays;
      c->counters.shadowRays += local.shadowRays;
      c->counters.trace.nodes += local.trace.nodes;
      c->counters.trace.tris += local.trace.tris;
    };
    std::vector<std::thread> th;
    for(int t = 0; t < nt; t++) th.emplace_back(work);
    for(auto& t : th) t.join();
  }
  return 0;
}

// Renders frames state->frame .. +num_frames-1 over rows [row_begin,row_end) (0,0 = all).
int oracle_render(OracleCtx* c, const RtxState* state, int32_t num_frames, int32_t seed_mode, int32_t row_begin, int32_t row_end)
{
  if(!c->committed || c->scene.envW == 0)
    return 3;
  const int W = state->size[0], H = state->size[1];
  if(W != c->width || H != c->height)
    return 1;
  if(row_end <= row_begin) { row_begin = 0; row_end = H; }
  int nt = c->threads > 0 ? c->threads : (int)std::thread::hardware_concurrency();
  if(nt < 1) nt = 1;
  for(int f = 0; f < num_frames; f++)
  {
    RtxState st = *state;
    st.frame    = state->frame + f;
    // work item = a 128-pixel segment of a row (keeps 100+ threads balanced)
    const int        segs = (W + 127) / 128;
    std::atomic<int> next{row_begin * segs};
    std::mutex       mtx;
    auto work = [&]() {
      RenderCounters local;
      for(;;)
      {
        int item = next.fetch_add(1);
        if(item >= row_end * segs)
          break;
        const int y = item / segs, x0 = (item % segs) * 128;
        for(int x = x0; x < std::min(W, x0 + 128); x++)
          renderPixel(c->scene, st, seed_mode, x, y, c->image.data(), &local);
      }
      std::lock_guard<std::mutex> g(mtx);
      c->counters.closestRays += local.closestRays;
      c->counters.shadowRays += local.shadowRays;
      c->counters.trace.nodes += local.trace.nodes;
      c->counters.trace.tris += local.trace.tris;
    };
    if(nt == 1)
      work();
    else
    {
      std::vector<std::thread> th;
      for(int t = 0; t < nt; t++) th.emplace_back(work);
      for(auto& t : th) t.join();
    }
  }
  return 0;
}

int oracle_tonemap(OracleCtx* c, const Tonemapper* tm, float* out_rgba32f, uint8_t* out_rgba8)
{
  const int W = c->width, H = c->height;
  MipChain  mips;
  if(tm->autoExposure & 1)
    mips = buildMipChain(c->image.data(), W, H);
  for(int y = 0; y < H; y++)
    for(int x = 0; x < W; x++)
    {
      vec4   v = postPixel(*tm, c->image.data(), W, H, x, y, (tm->autoExposure & 1) ? &mips : nullptr);
      size_t o = (size_t(y) * W + x) * 4;
      if(out_rgba32f)
      {
        out_rgba32f[o] = v.x; out_rgba32f[o + 1] = v.y; out_rgba32f[o + 2] = v.z; out_rgba32f[o + 3] = v.w;
      }
      if(out_rgba8)
      {
        // UNORM8 attachment write: clamp, scale, round-to-nearest-even
        auto q = [](float f) { f = std::fmin(std::fmax(f, 0.f), 1.f); return (uint8_t)std::nearbyint(f * 255.0f); };
        out_rgba8[o] = q(v.x); out_rgba8[o + 1] = q(v.y); out_rgba8[o + 2] = q(v.z); out_rgba8[o + 3] = q(v.w);
      }
    }
  return 0;
}

// rays: n x {ox,oy,oz,tmax,dx,dy,dz,seedbits}; hits: n x {t,u,v,tri-bits}
int oracle_trace_rays(OracleCtx* c, const float* rays, uint32_t n, int32_t closest, float* out)
{
  if(!c->committed)
    return 3;
  for(uint32_t i = 0; i < n; i++)
  {
    const float* r = rays + size_t(i) * 8;
    uint32_t     seed = floatBitsToUint(r[7]);
    Hit h = traceRay(c->scene, V3(r[0], r[1], r[2]), V3(r[4], r[5], r[6]), r[3], seed, closest != 0, nullptr);
    float* o = out + size_t(i) * 4;
    if(closest)
    {
      o[0] = h.tri == ~0u ? INFINITY_RT : h.t; o[1] = h.u; o[2] = h.v; o[3] = uintBitsToFloat(h.tri);
      if(h.tri == ~0u) { o[1] = 0; o[2] = 0; }
    }
    else
    {
      o[0] = h.tri == ~0u ? INFINITY_RT : 0.0f; o[1] = 0; o[2] = 0; o[3] = uintBitsToFloat(h.tri == ~0u ? ~0u : 0u);
    }
  }
  return 0;
}
int oracle_triangle_ids(OracleCtx* c, const uint32_t* ids, uint32_t n, uint32_t* out)
{
  for(uint32_t i = 0; i < n; i++)
  {
    if(ids[i] >= c->scene.tris.size()) { out[2 * i] = ~0u; out[2 * i + 1] = ~0u; continue; }
    out[2 * i] = c->scene.tris[ids[i]].node; out[2 * i + 1] = c->scene.tris[ids[i]].prim;
  }
  return 0;
}
int oracle_get_counters(OracleCtx* c, uint64_t out[4])
{
  out[0] = c->counters.closestRays; out[1] = c->counters.shadowRays; out[2] = c->counters.trace.nodes; out[3] = c->counters.trace.tris;
  return 0;
}
int oracle_reset_counters(OracleCtx* c)
{
  c->counters = RenderCounters();
  return 0;
}
uint32_t oracle_num_triangles(OracleCtx* c) { return (uint32_t)c->scene.tris.size(); }

// ---- known-answer hooks for the integer-exact pieces ---------------------------------
uint32_t oracle_tea(uint32_t a, uint32_t b) { return tea(a, b); }
uint32_t oracle_pcg(uint32_t* state) { return pcg(*state); }
float    oracle_rand(uint32_t* state) { return rnd(*state); }
void     oracle_pcg3d(uint32_t v[3]) { pcg3d(v); }
uint32_t oracle_compress_unit_vec(float x, float y, float z) { return compress_unit_vec(V3(x, y, z)); }
void     oracle_decompress_unit_vec(uint32_t p, float out[3]) { vec3 v = decompress_unit_vec(p); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
uint32_t oracle_pack_unorm4x8(float x, float y, float z, float w) { return packUnorm4x8(V4(x, y, z, w)); }
void     oracle_offset_ray(const float p[3], const float n[3], float out[3]) { vec3 v = OffsetRay(V3(p[0], p[1], p[2]), V3(n[0], n[1], n[2])); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
float    oracle_alpha_draw(uint32_t seed, uint32_t tri) { return alphaDraw(seed, tri); }

// Batch vertex packing, src/scene.cpp:224-246 (normal/tangent oct encode, handedness LSB, RGBA8 colour).
void oracle_pack_vertices(const float* pos, const float* nrm, const float* tng, const float* uv, const float* col, uint32_t n, VertexAttributes* out)
{
  for(uint32_t i = 0; i < n; i++)
  {
    VertexAttributes v{};
    v.position[0] = pos[3 * i]; v.position[1] = pos[3 * i + 1]; v.position[2] = pos[3 * i + 2];
    v.normal      = compress_unit_vec(V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
    v.tangent     = compress_unit_vec(V3(tng[4 * i], tng[4 * i + 1], tng[4 * i + 2]));
    v.texcoord[0] = uv[2 * i];
    uint32_t value = floatBitsToUint(uv[2 * i + 1]);
    if(tng[4 * i + 3] > 0)
      value |= 1;
    else
      value &= ~1u;
    v.texcoord[1] = uintBitsToFloat(value);
    v.color       = packUnorm4x8(V4(col[4 * i], col[4 * i + 1], col[4 * i + 2], col[4 * i + 3]));
    out[i]        = v;
  }
}

// BSDF probe: evaluates both models on an explicit shading state (tests/golden/bsdf_*.json).
// in: 32 floats {V3, N3, L3, T3, B3, albedo3, metallic, roughness, transmission, ior, eta, clearcoat, clearcoatRoughness,
//                anisotropy, sheen, sheenTint3, thinwalled, subsurface, ffnormal-sign}
// out: {f.xyz, pdf}
void oracle_bsdf_eval(int32_t pbrMode, const float* in, float* out)
{
  State st{};
  vec3  V = V3(in[0], in[1], in[2]), N = V3(in[3], in[4], in[5]), L = V3(in[6], in[7], in[8]);
  st.tangent = V3(in[9], in[10], in[11]); st.bitangent = V3(in[12], in[13], in[14]);
  st.mat.albedo = V3(in[15], in[16], in[17]); st.mat.metallic = in[18]; st.mat.roughness = in[19]; st.mat.transmission = in[20];
  st.mat.ior = in[21]; st.eta = in[22]; st.mat.clearcoat = in[23]; st.mat.clearcoatRoughness = in[24]; st.mat.anisotropy = in[25];
  st.mat.sheen = in[26]; st.mat.sheenTint = V3(in[27], in[28], in[29]); st.mat.thinwalled = in[30] != 0; st.mat.subsurface = in[31];
  st.mat.specular = 0.5f; st.mat.specularTint = 1.f;
  float aspect = std::sqrt(1.0f - st.mat.anisotropy * 0.9f);
  st.mat.ax = std::fmax(0.001f, st.mat.roughness / aspect);
  st.mat.ay = std::fmax(0.001f, st.mat.roughness * aspect);
  float dsp = (st.mat.ior - 1) / (st.mat.ior + 1);
  st.mat.f0 = mix(V3(dsp * dsp), st.mat.albedo, st.mat.metallic);
  st.normal = N; st.ffnormal = N;
  float pdf = 0.f;
  vec3  f   = pbrMode == 0 ? DisneyEval(st, V, N, L, pdf) : PbrEval(st, V, N, L, pdf);
  out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = pdf;
}
// Same state, sampling: out {L.xyz, f.xyz, pdf, seed-after-bits}
void oracle_bsdf_sample(int32_t pbrMode, const float* in, uint32_t seed, float* out)
{
  State st{};
  vec3  V = V3(in[0], in[1], in[2]), N = V3(in[3], in[4], in[5]);
  st.tangent = V3(in[9], in[10], in[11]); st.bitangent = V3(in[12], in[13], in[14]);
  st.mat.albedo = V3(in[15], in[16], in[17]); st.mat.metallic = in[18]; st.mat.roughness = in[19]; st.mat.transmission = in[20];
  st.mat.ior = in[21]; st.eta = in[22]; st.mat.clearcoat = in[23]; st.mat.clearcoatRoughness = in[24]; st.mat.anisotropy = in[25];
  st.mat.sheen = in[26]; st.mat.sheenTint = V3(in[27], in[28], in[29]); st.mat.thinwalled = in[30] != 0; st.mat.subsurface = in[31];
  st.mat.specular = 0.5f; st.mat.specularTint = 1.f;
  float aspect = std::sqrt(1.0f - st.mat.anisotropy * 0.9f);
  st.mat.ax = std::fmax(0.001f, st.mat.roughness / aspect);
  st.mat.ay = std::fmax(0.001f, st.mat.roughness * aspect);
  float dsp = (st.mat.ior - 1) / (st.mat.ior + 1);
  st.mat.f0 = mix(V3(dsp * dsp), st.mat.albedo, st.mat.metallic);
  st.normal = N; st.ffnormal = N;
  float pdf = 0.f;
  vec3  L = V3(0);
  vec3  f = pbrMode == 0 ? DisneySample(st, V, N, L, pdf, seed) : PbrSample(st, V, N, L, pdf, seed);
  out[0] = L.x; out[1] = L.y; out[2] = L.z; out[3] = f.x; out[4] = f.y; out[5] = f.z; out[6] = pdf; out[7] = uintBitsToFloat(seed);
}

// Scene::updateCamera (src/scene.cpp:629-640): view = lookAt(eye, center, up) (nvh::CameraManipulator::getMatrix), proj =
// glm::perspectiveRH_ZO(radians(fov), aspect, 0.001, 100000) with proj[1][1] negated, both inverted with glm::inverse; focalDist =
// |center - eye|.  Restated in single precision like glm: matrices are column-major m[col*4+row]; the inverse is the adjugate built
// from the six 2x2 minors of the lower and of the upper row pair, divided by the determinant.
static void inverse4(const float* m, float* out)
{
  // 2x2 minors of rows 2,3 (s) and rows 0,1 (c), column pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
  auto at = [&](int col, int row) { return m[col * 4 + row]; };
  const int pr[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  float     lo[6], hi[6];
  for(int k = 0; k < 6; k++)
  {
    lo[k] = at(pr[k][0], 2) * at(pr[k][1], 3) - at(pr[k][1], 2) * at(pr[k][0], 3);
    hi[k] = at(pr[k][0], 0) * at(pr[k][1], 1) - at(pr[k][1], 0) * at(pr[k][0], 1);
  }
  const float det = hi[0] * lo[5] - hi[1] * lo[4] + hi[2] * lo[3] + hi[3] * lo[2] - hi[4] * lo[1] + hi[5] * lo[0];
  const float r   = 1.0f / det;
  float       adj[16];
  // adj(col j, row i) = cofactor(row j, col i)
  adj[0 * 4 + 0] = (at(1, 1) * lo[5] - at(2, 1) * lo[4] + at(3, 1) * lo[3]);
  adj[1 * 4 + 0] = -(at(0, 1) * lo[5] - at(2, 1) * lo[2] + at(3, 1) * lo[1]);
  adj[2 * 4 + 0] = (at(0, 1) * lo[4] - at(1, 1) * lo[2] + at(3, 1) * lo[0]);
  adj[3 * 4 + 0] = -(at(0, 1) * lo[3] - at(1, 1) * lo[1] + at(2, 1) * lo[0]);
  adj[0 * 4 + 1] = -(at(1, 0) * lo[5] - at(2, 0) * lo[4] + at(3, 0) * lo[3]);
  adj[1 * 4 + 1] = (at(0, 0) * lo[5] - at(2, 0) * lo[2] + at(3, 0) * lo[1]);
  adj[2 * 4 + 1] = -(at(0, 0) * lo[4] - at(1, 0) * lo[2] + at(3, 0) * lo[0]);
  adj[3 * 4 + 1] = (at(0, 0) * lo[3] - at(1, 0) * lo[1] + at(2, 0) * lo[0]);
  adj[0 * 4 + 2] = (at(1, 3) * hi[5] - at(2, 3) * hi[4] + at(3, 3) * hi[3]);
  adj[1 * 4 + 2] = -(at(0, 3) * hi[5] - at(2, 3) * hi[2] + at(3, 3) * hi[1]);
  adj[2 * 4 + 2] = (at(0, 3) * hi[4] - at(1, 3) * hi[2] + at(3, 3) * hi[0]);
  adj[3 * 4 + 2] = -(at(0, 3) * hi[3] - at(1, 3) * hi[1] + at(2, 3) * hi[0]);
  adj[0 * 4 + 3] = -(at(1, 2) * hi[5] - at(2, 2) * hi[4] + at(3, 2) * hi[3]);
  adj[1 * 4 + 3] = (at(0, 2) * hi[5] - at(2, 2) * hi[2] + at(3, 2) * hi[1]);
  adj[2 * 4 + 3] = -(at(0, 2) * hi[4] - at(1, 2) * hi[2] + at(3, 2) * hi[0]);
  adj[3 * 4 + 3] = (at(0, 2) * hi[3] - at(1, 2) * hi[1] + at(2, 2) * hi[0]);
  // transpose into place: out(col, row) = adj(row... ) -- adj above is indexed [col*4+row] of the cofactor matrix; the inverse is its transpose
  for(int c = 0; c < 4; c++)
    for(int rr = 0; rr < 4; rr++)
      out[c * 4 + rr] = adj[rr * 4 + c] * r;
}
void oracle_make_camera(const float eye[3], const float center[3], const float up[3], float fovDegrees, float aspect, float aperture, int32_t nbLights,
                        SceneCamera* out)
{
  const vec3 e = V3(eye[0], eye[1], eye[2]), c = V3(center[0], center[1], center[2]), u0 = V3(up[0], up[1], up[2]);
  const vec3 f = normalize(c - e);
  const vec3 s = normalize(cross(f, u0));
  const vec3 u = cross(s, f);
  // glm::lookAtRH
  const float view[16] = {s.x, u.x, -f.x, 0.f, s.y, u.y, -f.y, 0.f, s.z, u.z, -f.z, 0.f, -dot(s, e), -dot(u, e), dot(f, e), 1.f};
  // glm::perspectiveRH_ZO, then proj[1][1] *= -1
  const float zn = 0.001f, zf = 100000.0f;
  const float th = std::tan(fovDegrees * 0.01745329251994329576923690768489f / 2.0f);
  float       proj[16] = {0};
  proj[0 * 4 + 0] = 1.0f / (aspect * th);
  proj[1 * 4 + 1] = -(1.0f / th);
  proj[2 * 4 + 2] = zf / (zn - zf);
  proj[2 * 4 + 3] = -1.0f;
  proj[3 * 4 + 2] = -(zf * zn) / (zf - zn);
  SceneCamera cam{};
  inverse4(view, cam.viewInverse);
  inverse4(proj, cam.projInverse);
  cam.focalDist = length(c - e);
  cam.aperture  = aperture;
  cam.nbLights  = nbLights;
  *out          = cam;
}

void oracle_sun_and_sky(const SunAndSky* ss, const float dir[3], float out[3])
{
  vec3 v = sky::sun_and_sky(*ss, V3(dir[0], dir[1], dir[2]));
  out[0] = v.x; out[1] = v.y; out[2] = v.z;
}

}  // extern "C"
