// gltf_scene.hpp -- glTF 2.0 ingest for the CUDA back-end: the host half of Scene::load (src/scene.cpp:56-118).
// The reference delegates parsing to tinygltf and the conversion to nvh::GltfScene (nvpro_core, not vendored, no pinned
// version; call sites src/scene.cpp:59-73).  This is a restatement of the importer behaviour the renderer relies on
// (SURVEY.md 8c item 2): triangle primitives only, world matrices flattened per node, identical primitives shared
// between nodes become one primMesh, missing attributes generated, KHR material extensions mapped onto
// GltfShadeMaterial, KHR_lights_punctual, first camera.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/vkrt_b200.h"
#include "image_io.hpp"

class CudaRenderer;

namespace vkrt_host {

struct PrimMeshData
{
  std::vector<float>    positions, normals, tangents, texcoords, colors;  // 3,3,4,2,4 floats per vertex
  std::vector<uint32_t> indices;
  int                   materialIndex = 0;
};
struct NodeData
{
  float    worldMatrix[16];  // column-major
  uint32_t primMesh;
};
struct TextureData
{
  ImageRGBA8 image;
  int        magFilter = 9729, minFilter = 9729, wrapS = 10497, wrapT = 10497;  // no sampler: LINEAR / REPEAT (src/scene.cpp:562-565)
};
struct CameraData
{
  float eye[3] = {0, 0, 5}, center[3] = {0, 0, 0}, up[3] = {0, 1, 0};
  float yfovDegrees = 45.f;
  bool  fromFile = false;
};

struct GltfSceneData
{
  std::vector<PrimMeshData>      primMeshes;
  std::vector<NodeData>          nodes;
  std::vector<GltfShadeMaterial> materials;
  std::vector<Light>             lights;
  std::vector<TextureData>       textures;
  CameraData                     camera;
  float                          boundsMin[3] = {0, 0, 0}, boundsMax[3] = {0, 0, 0};
  size_t                         triangleCount = 0;
};

// .gltf (JSON + external / data-URI buffers) or .glb
bool loadGltf(const std::string& path, GltfSceneData& out, std::string& err);
// Scene::createVertexBuffer .. createTextureImages + AccelStructure::create, through the CudaRenderer adaptor
void uploadScene(const GltfSceneData& scene, CudaRenderer& renderer, int buildMode);

}  // namespace vkrt_host
