// headless_example.cpp -- what src/main.cpp + SampleExample do around a Renderer, minus the window:
// build a tiny scene, create the back-end, run `frames` frames through Renderer::run, tonemap, write a PPM.
//   g++ -std=c++17 headless_example.cpp cuda_renderer.cpp -I../../include -L.. -lvkrt_b200 -Wl,-rpath,'$ORIGIN/..' -o vkrt_headless
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "cuda_renderer.hpp"

int main(int argc, char** argv)
{
  const int W = 256, H = 256, frames = argc > 1 ? atoi(argv[1]) : 16;
  CudaRenderer  backend(0);
  Renderer*     r = &backend;  // the application only sees the abstract interface (std::array<Renderer*, eNone>)
  r->setup(nullptr, nullptr, 0, nullptr);
  if(!backend.ok()) { std::fprintf(stderr, "%s\n", backend.error().c_str()); return 2; }

  // one ground quad + one tetrahedron, default material, white environment
  backend.beginScene();
  const float qp[] = {-2, 0, 2, 2, 0, 2, 2, 0, -2, -2, 0, -2};
  const float qn[] = {0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0};
  const float qt[] = {1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1};
  const float quv[] = {0, 0, 1, 0, 1, 1, 0, 1};
  const float qc[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
  const uint32_t qi[] = {0, 1, 2, 0, 2, 3};
  backend.feedPrimMesh(qp, qn, qt, quv, qc, 4, qi, 6, 0);
  const float tp[] = {0, 1.2f, 0, -0.7f, 0.05f, 0.6f, 0.7f, 0.05f, 0.6f, 0, 0.05f, -0.7f};
  float tn[12];
  for(int i = 0; i < 4; i++) { float l = std::sqrt(tp[3*i]*tp[3*i] + (tp[3*i+1]-0.4f)*(tp[3*i+1]-0.4f) + tp[3*i+2]*tp[3*i+2]); tn[3*i] = tp[3*i]/l; tn[3*i+1] = (tp[3*i+1]-0.4f)/l; tn[3*i+2] = tp[3*i+2]/l; }
  const uint32_t ti[] = {0, 1, 2, 0, 2, 3, 0, 3, 1, 1, 3, 2};
  backend.feedPrimMesh(tp, tn, qt, quv, qc, 4, ti, 12, 1);
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  backend.feedNode(I, 0);
  backend.feedNode(I, 1);
  GltfShadeMaterial m[2] = {};
  for(auto& x : m)
  {
    x.pbrBaseColorFactor[0] = x.pbrBaseColorFactor[1] = x.pbrBaseColorFactor[2] = 0.7f; x.pbrBaseColorFactor[3] = 1.f;
    x.pbrBaseColorTexture = x.pbrMetallicRoughnessTexture = x.emissiveTexture = x.normalTexture = x.transmissionTexture = -1;
    x.clearcoatTexture = x.clearcoatRoughnessTexture = x.thicknessTexture = -1;
    x.pbrRoughnessFactor = 0.6f; x.ior = 1.5f; x.alphaCutoff = 0.5f; x.normalTextureScale = 1.f; x.attenuationDistance = 3.4e38f;
    x.attenuationColor[0] = x.attenuationColor[1] = x.attenuationColor[2] = 1.f;
    for(int k = 0; k < 4; k++) x.uvTransform[5 * k] = 1.f;
  }
  m[1].pbrBaseColorFactor[0] = 0.9f; m[1].pbrBaseColorFactor[1] = 0.3f; m[1].pbrMetallicFactor = 1.f; m[1].pbrRoughnessFactor = 0.25f;
  backend.feedMaterials(m, 2);
  backend.feedLights(nullptr, 0);
  backend.commitScene();
  std::vector<float> env(64 * 32 * 4, 1.0f);
  const float integral = backend.setEnvironment(env.data(), 64, 32);
  const float eye[3] = {0, 1.2f, 4}, ctr[3] = {0, 0.5f, 0}, up[3] = {0, 1, 0};
  backend.updateCamera(eye, ctr, up, 45.f, (float)W / H, 0.f, 0);
  r->create(VkExtent2D{(uint32_t)W, (uint32_t)H}, {}, nullptr);

  // main loop (src/main.cpp:201-264): updateFrame -> setPushContants -> run
  RtxState st{};  // defaults of src/sample_example.hpp:162-174
  st.maxDepth = 10; st.maxSamples = 1; st.fireflyClampThreshold = integral * 4.f; st.hdrMultiplier = 1.f; st.maxHeatmap = 65000;
  nvvk::ProfilerVK* prof = nullptr;
  for(int f = 0; f < frames; f++)
  {
    st.frame = f;
    r->setPushContants(st);
    r->run(nullptr, VkExtent2D{(uint32_t)W, (uint32_t)H}, *prof, {});
  }
  Tonemapper tm{1.f, 1.f, 1.f, 0.f, 1.f, 1.f, {1.f, 1.f}, 0, 0.5f, 0.5f, 1};
  std::vector<uint8_t> img((size_t)W * H * 4);
  backend.tonemapTo(tm, img.data());
  if(!backend.ok()) { std::fprintf(stderr, "%s\n", backend.error().c_str()); return 3; }
  FILE* fp = std::fopen("headless.ppm", "wb");
  std::fprintf(fp, "P6\n%d %d\n255\n", W, H);
  for(int i = 0; i < W * H; i++) std::fwrite(&img[4 * i], 1, 3, fp);
  std::fclose(fp);
  long sum = 0;
  for(int i = 0; i < W * H; i++) sum += img[4 * i + 1];
  std::printf("%s: %d frames, mean green %.2f -> headless.ppm\n", r->name().c_str(), frames, (double)sum / (W * H));
  r->destroy();
  return 0;
}
