/*
 * vkrt_host.h -- C ABI of libvkrt_host.so: the host half of Scene::load for callers that do not link nvpro_core / tinygltf
 * (reference: src/scene.cpp:56-160 Scene::load, src/hdr_sampling.cpp:56-99 HdrSampling::loadEnvironment).  It reads a glTF 2.0
 * file (.gltf / .glb, sparse accessors, PNG / JPEG textures) or a Radiance .hdr / .pfm environment with the decoders of
 * vk_raytrace_b200/host/ and feeds an existing renderer context through the entry points of vkrt_b200.h -- nothing here touches
 * the GPU on its own.  Used by `bench.py --scene-file` and by Python / C hosts; C++ hosts can use gltf_scene.hpp directly.
 */
#ifndef VKRT_HOST_H
#define VKRT_HOST_H
#include "vkrt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkrt_host_scene_info
{
  float    eye[3], center[3], up[3]; /* first camera of the file, or a fit of the scene bounds (src/scene.cpp:280-297) */
  float    yfov_degrees;
  int32_t  camera_from_file;
  uint32_t prim_meshes, nodes, materials, textures, lights;
  uint64_t triangles;                /* after instancing */
  float    bounds_min[3], bounds_max[3];
} vkrt_host_scene_info;

/* Scene::load + AccelStructure::create: parse `path`, feed meshes / nodes / materials / lights / textures to `ctx`, commit with
 * `build_mode`.  Returns VKRT_OK or an error code; vkrt_host_last_error() has the message. */
int vkrt_host_load_gltf(vkrt_ctx* ctx, const char* path, int32_t build_mode, vkrt_host_scene_info* out_info);
/* HdrSampling::loadEnvironment: decode a Radiance .hdr or a .pfm, build the alias table, upload.  out_integral receives the
 * integral of the importance function (the firefly threshold is 4x that, src/sample_example.cpp:110). */
int vkrt_host_load_environment(vkrt_ctx* ctx, const char* path, float* out_integral, uint32_t* out_width, uint32_t* out_height);
const char* vkrt_host_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VKRT_HOST_H */
