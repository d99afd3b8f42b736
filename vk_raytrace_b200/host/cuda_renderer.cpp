// cuda_renderer.cpp -- see cuda_renderer.hpp.  Error convention: the reference's interface is all-void
// (failures are asserts, src/rtx_pipeline.cpp:209,237); here a failed call latches ok() = false and error().
#include "cuda_renderer.hpp"

void CudaRenderer::check(int rc, const char* what)
{
  if(rc != VKRT_OK && m_ok)
  {
    m_ok    = false;
    m_error = std::string(what) + ": " + (m_ctx ? vkrt_last_error(m_ctx) : "no context");
  }
}

void CudaRenderer::setup(const VkDevice&, const VkPhysicalDevice&, uint32_t, nvvk::ResourceAllocator*)
{
  if(!m_ctx)
    check(vkrt_create(m_cudaDevice, &m_ctx), "vkrt_create");
}

void CudaRenderer::destroy()
{
  if(m_ctx)
    vkrt_destroy(m_ctx);
  m_ctx = nullptr;
}

void CudaRenderer::create(const VkExtent2D& size, const std::vector<VkDescriptorSetLayout>&, Scene*)
{
  // createRender (src/sample_example.cpp:322-337): the offscreen image is (re)created with the render size
  if(m_ctx)
    check(vkrt_output_create(m_ctx, size.width, size.height, nullptr), "vkrt_output_create");
}

void CudaRenderer::run(const VkCommandBuffer&, const VkExtent2D& size, nvvk::ProfilerVK&, const std::vector<VkDescriptorSet>&)
{
  // renderScene (src/sample_example.cpp:415-420): m_state was set through setPushContants()
  if(!m_ctx)
    return;
  RtxState s = m_state;
  s.size[0]  = (int32_t)size.width;
  s.size[1]  = (int32_t)size.height;
  vkrt_render_opts o{};
  o.seed_mode = VKRT_SEED_RAYQUERY;
  check(vkrt_render(m_ctx, &s, 1, &o), "vkrt_render");
}

void CudaRenderer::beginScene()
{
  m_numMeshes = 0;
  if(m_ctx)
    check(vkrt_scene_begin(m_ctx), "vkrt_scene_begin");
}

uint32_t CudaRenderer::feedPrimMesh(const float* positions, const float* normals, const float* tangents4, const float* texcoords,
                                    const float* colors4, uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, int materialIndex)
{
  std::vector<VertexAttributes> v(vertexCount);
  check(vkrt_pack_vertices(positions, normals, tangents4, texcoords, colors4, vertexCount, v.data()), "vkrt_pack_vertices");
  check(vkrt_scene_add_primmesh(m_ctx, v.data(), vertexCount, indices, indexCount, materialIndex), "vkrt_scene_add_primmesh");
  return m_numMeshes++;
}

void CudaRenderer::feedNode(const float worldMatrix[16], uint32_t primMesh) { check(vkrt_scene_add_node(m_ctx, worldMatrix, primMesh), "vkrt_scene_add_node"); }
void CudaRenderer::feedMaterials(const GltfShadeMaterial* m, uint32_t n) { check(vkrt_scene_set_materials(m_ctx, m, n), "vkrt_scene_set_materials"); }
void CudaRenderer::feedLights(const Light* l, uint32_t n) { check(vkrt_scene_set_lights(m_ctx, l, n), "vkrt_scene_set_lights"); }
void CudaRenderer::feedTexture(const uint8_t* rgba8, uint32_t w, uint32_t h, int magFilter, int minFilter, int wrapS, int wrapT)
{
  check(vkrt_scene_add_texture(m_ctx, rgba8, w, h, magFilter, minFilter, wrapS, wrapT), "vkrt_scene_add_texture");
}
void CudaRenderer::commitScene(int buildMode) { check(vkrt_scene_commit(m_ctx, buildMode), "vkrt_scene_commit"); }

float CudaRenderer::setEnvironment(const float* rgba32f, uint32_t w, uint32_t h)
{
  std::vector<EnvAccel> accel((size_t)w * h);
  float integral = 0.f, average = 0.f;
  check(vkrt_build_env_accel(rgba32f, w, h, accel.data(), &integral, &average), "vkrt_build_env_accel");
  check(vkrt_set_environment(m_ctx, rgba32f, w, h, accel.data()), "vkrt_set_environment");
  return integral;
}

void CudaRenderer::setSunAndSky(const SunAndSky& s) { check(vkrt_set_sunsky(m_ctx, &s), "vkrt_set_sunsky"); }

void CudaRenderer::updateCamera(const float eye[3], const float center[3], const float up[3], float fovDegrees, float aspect, float aperture, int nbLights)
{
  SceneCamera cam{};
  check(vkrt_make_camera(eye, center, up, fovDegrees, aspect, aperture, nbLights, &cam), "vkrt_make_camera");
  check(vkrt_set_camera(m_ctx, &cam), "vkrt_set_camera");
}

void CudaRenderer::tonemapTo(const Tonemapper& tm, uint8_t* hostRgba8) { check(vkrt_tonemap(m_ctx, &tm, hostRgba8), "vkrt_tonemap"); }
void CudaRenderer::readAccumulation(float* hostRgba32f) { check(vkrt_output_read(m_ctx, hostRgba32f), "vkrt_output_read"); }
void CudaRenderer::sync() { check(vkrt_sync(m_ctx), "vkrt_sync"); }
