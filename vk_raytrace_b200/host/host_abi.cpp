// host_abi.cpp -- libvkrt_host.so (include/vkrt_host.h): glTF / environment files -> an existing renderer context, through the C ABI.
#include <cstring>
#include <string>
#include <vector>
#include "../../include/vkrt_host.h"
#include "gltf_scene.hpp"

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}
}  // namespace

extern "C" {

const char* vkrt_host_last_error(void) { return g_err.c_str(); }

int vkrt_host_load_gltf(vkrt_ctx* ctx, const char* path, int32_t build_mode, vkrt_host_scene_info* out_info)
{
  if(!ctx || !path)
    return fail(VKRT_ERR_INVALID, "load_gltf: null argument");
  vkrt_host::GltfSceneData s;
  std::string              err;
  if(!vkrt_host::loadGltf(path, s, err))
    return fail(VKRT_ERR_INVALID, "load_gltf: " + err);
  auto ck = [&](int rc, const char* what) -> int {
    if(rc != VKRT_OK)
      g_err = std::string(what) + ": " + vkrt_last_error(ctx);
    return rc;
  };
  int rc = ck(vkrt_scene_begin(ctx), "vkrt_scene_begin");
  // the same calls as uploadScene() makes through the CudaRenderer adaptor (Scene::createVertexBuffer .. createTextureImages)
  std::vector<VertexAttributes> packed;
  for(size_t i = 0; i < s.primMeshes.size() && rc == VKRT_OK; i++)
  {
    const auto&    m = s.primMeshes[i];
    const uint32_t n = (uint32_t)(m.positions.size() / 3);
    packed.resize(n);
    rc = ck(vkrt_pack_vertices(m.positions.data(), m.normals.data(), m.tangents.data(), m.texcoords.data(), m.colors.data(), n, packed.data()), "vkrt_pack_vertices");
    if(rc == VKRT_OK)
      rc = ck(vkrt_scene_add_primmesh(ctx, packed.data(), n, m.indices.data(), (uint32_t)m.indices.size(), m.materialIndex), "vkrt_scene_add_primmesh");
  }
  for(size_t i = 0; i < s.nodes.size() && rc == VKRT_OK; i++)
    rc = ck(vkrt_scene_add_node(ctx, s.nodes[i].worldMatrix, s.nodes[i].primMesh), "vkrt_scene_add_node");
  if(rc == VKRT_OK) rc = ck(vkrt_scene_set_materials(ctx, s.materials.data(), (uint32_t)s.materials.size()), "vkrt_scene_set_materials");
  if(rc == VKRT_OK) rc = ck(vkrt_scene_set_lights(ctx, s.lights.empty() ? nullptr : s.lights.data(), (uint32_t)s.lights.size()), "vkrt_scene_set_lights");
  for(size_t i = 0; i < s.textures.size() && rc == VKRT_OK; i++)
  {
    const auto& t = s.textures[i];
    rc = ck(vkrt_scene_add_texture(ctx, t.image.rgba.data(), t.image.w, t.image.h, t.magFilter, t.minFilter, t.wrapS, t.wrapT), "vkrt_scene_add_texture");
  }
  if(rc == VKRT_OK) rc = ck(vkrt_scene_commit(ctx, build_mode), "vkrt_scene_commit");
  if(rc != VKRT_OK)
    return rc;
  if(out_info)
  {
    vkrt_host_scene_info& o = *out_info;
    std::memset(&o, 0, sizeof(o));
    for(int k = 0; k < 3; k++)
    {
      o.eye[k] = s.camera.eye[k]; o.center[k] = s.camera.center[k]; o.up[k] = s.camera.up[k];
      o.bounds_min[k] = s.boundsMin[k]; o.bounds_max[k] = s.boundsMax[k];
    }
    o.yfov_degrees     = s.camera.yfovDegrees;
    o.camera_from_file = s.camera.fromFile ? 1 : 0;
    o.prim_meshes = (uint32_t)s.primMeshes.size(); o.nodes = (uint32_t)s.nodes.size(); o.materials = (uint32_t)s.materials.size();
    o.textures = (uint32_t)s.textures.size(); o.lights = (uint32_t)s.lights.size();
    o.triangles = s.triangleCount;
  }
  return VKRT_OK;
}

int vkrt_host_load_environment(vkrt_ctx* ctx, const char* path, float* out_integral, uint32_t* out_width, uint32_t* out_height)
{
  if(!ctx || !path)
    return fail(VKRT_ERR_INVALID, "load_environment: null argument");
  vkrt_host::ImageRGBA32F env;
  std::string             err;
  if(!vkrt_host::loadHdrOrPfm(path, env, err))
    return fail(VKRT_ERR_INVALID, "load_environment: " + err);
  std::vector<EnvAccel> accel((size_t)env.w * env.h);
  float                 integral = 0.f, average = 0.f;
  int rc = vkrt_build_env_accel(env.rgba.data(), env.w, env.h, accel.data(), &integral, &average);
  if(rc == VKRT_OK)
    rc = vkrt_set_environment(ctx, env.rgba.data(), env.w, env.h, accel.data());
  if(rc != VKRT_OK)
    return fail(rc, std::string("load_environment: ") + vkrt_last_error(ctx));
  if(out_integral) *out_integral = integral;
  if(out_width) *out_width = env.w;
  if(out_height) *out_height = env.h;
  return VKRT_OK;
}

}  // extern "C"
