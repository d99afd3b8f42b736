// cuda_renderer.hpp -- `CudaRenderer : Renderer`: the third back-end next to RtxPipeline and RayQuery
// (src/sample_example.hpp:86-91, src/sample_example.cpp:77-82).  It ignores the Vulkan descriptor sets and drives the
// sm_100a wavefront path tracer through the C ABI.  Because the reference keeps vertex data only in Vulkan buffers
// after load (src/scene.cpp:112-115), the scene is fed at Scene::create*Buffer time through the feed*() methods
// (INTEGRATION.md shows the five call sites).
#pragma once
#include "renderer_iface.hpp"

class CudaRenderer : public Renderer
{
public:
  explicit CudaRenderer(int cudaDevice = 0) : m_cudaDevice(cudaDevice) {}
  // The reference's Renderer has no virtual destructor (src/renderer.h:30-48: back-ends are torn down through destroy(),
  // src/sample_example.cpp:283-287), so this one is not an override; own the object by its concrete type.
  ~CudaRenderer() { destroy(); }

  // ---- Renderer (src/renderer.h:33-43) -------------------------------------------------------------------
  void setup(const VkDevice&, const VkPhysicalDevice&, uint32_t, nvvk::ResourceAllocator*) override;
  void destroy() override;
  void create(const VkExtent2D& size, const std::vector<VkDescriptorSetLayout>&, Scene* = nullptr) override;
  void run(const VkCommandBuffer&, const VkExtent2D& size, nvvk::ProfilerVK&, const std::vector<VkDescriptorSet>&) override;
  const std::string name() override { return std::string("CUDA wavefront (B200)"); }

  // ---- scene feed: mirrors Scene::load (src/scene.cpp:56-118) ---------------------------------------------------
  void beginScene();
  // Scene::createVertexBuffer (src/scene.cpp:190-274): raw importer attributes of one primMesh; packed like the reference
  uint32_t feedPrimMesh(const float* positions, const float* normals, const float* tangents4, const float* texcoords, const float* colors4,
                        uint32_t vertexCount, const uint32_t* indices, uint32_t indexCount, int materialIndex);
  void     feedNode(const float worldMatrix[16], uint32_t primMesh);                      // src/accelstruct.cpp:132-157
  void     feedMaterials(const GltfShadeMaterial* m, uint32_t n);                        // src/scene.cpp:339-382
  void     feedLights(const Light* l, uint32_t n);                                       // src/scene.cpp:304-333
  void     feedTexture(const uint8_t* rgba8, uint32_t w, uint32_t h, int magFilter, int minFilter, int wrapS, int wrapT);  // :488-580
  void     commitScene(int buildMode = VKRT_BUILD_LBVH_COLLAPSED);                       // AccelStructure::create
  // HdrSampling::loadEnvironment (src/hdr_sampling.cpp:56-99); returns the integral (firefly threshold = 4x, sample_example.cpp:110)
  float setEnvironment(const float* rgba32f, uint32_t w, uint32_t h);
  void  setSunAndSky(const SunAndSky& s);                                                 // src/sample_example.hpp:176-193
  // Scene::updateCamera (src/scene.cpp:629-640)
  void updateCamera(const float eye[3], const float center[3], const float up[3], float fovDegrees, float aspect, float aperture, int nbLights);

  // ---- RenderOutput (src/render_output.cpp:174-182) ----------------------------------------------------------
  void tonemapTo(const Tonemapper& tm, uint8_t* hostRgba8);
  void readAccumulation(float* hostRgba32f);
  void sync();
  vkrt_ctx* context() const { return m_ctx; }
  bool      ok() const { return m_ok; }
  const std::string& error() const { return m_error; }

private:
  void check(int rc, const char* what);
  int         m_cudaDevice;
  vkrt_ctx*   m_ctx = nullptr;
  uint32_t    m_numMeshes = 0;
  bool        m_ok = true;
  std::string m_error;
};
