// gltf_render.cpp -- headless "vk_raytrace -f scene.gltf -e env.hdr" (src/main.cpp:55-57) on the CUDA back-end:
//   vkrt_gltf_render <scene.gltf|.glb> <env.hdr|.pfm> <width> <height> <frames> <maxDepth> <out-prefix> [pbrMode] [buildMode]
// writes <out-prefix>.ppm (tonemapped, src/render_output.hpp:37-49 defaults) and <out-prefix>.f32 (raw RGBA32F accumulation).
#include <cstdio>
#include <cstdlib>
#include "cuda_renderer.hpp"
#include "gltf_scene.hpp"

int main(int argc, char** argv)
{
  if(argc < 8)
  {
    std::fprintf(stderr, "usage: %s scene.gltf env.hdr|env.pfm width height frames maxDepth out-prefix [pbrMode] [buildMode]\n", argv[0]);
    return 1;
  }
  const int W = atoi(argv[3]), H = atoi(argv[4]), frames = atoi(argv[5]), depth = atoi(argv[6]);
  const std::string prefix = argv[7];
  const int pbr = argc > 8 ? atoi(argv[8]) : 0, buildMode = argc > 9 ? atoi(argv[9]) : VKRT_BUILD_LBVH_COLLAPSED;

  vkrt_host::GltfSceneData scene;
  std::string              err;
  if(!vkrt_host::loadGltf(argv[1], scene, err)) { std::fprintf(stderr, "glTF: %s\n", err.c_str()); return 2; }
  vkrt_host::ImageRGBA32F env;
  if(!vkrt_host::loadHdrOrPfm(argv[2], env, err)) { std::fprintf(stderr, "environment: %s\n", err.c_str()); return 2; }
  std::printf("loaded %zu primMeshes, %zu nodes, %zu triangles, %zu materials, %zu textures, %zu lights; env %ux%u\n", scene.primMeshes.size(),
              scene.nodes.size(), scene.triangleCount, scene.materials.size(), scene.textures.size(), scene.lights.size(), env.w, env.h);

  CudaRenderer r(0);
  r.setup(nullptr, nullptr, 0, nullptr);
  if(!r.ok()) { std::fprintf(stderr, "%s\n", r.error().c_str()); return 3; }
  vkrt_host::uploadScene(scene, r, buildMode);
  const float integral = r.setEnvironment(env.rgba.data(), env.w, env.h);
  r.updateCamera(scene.camera.eye, scene.camera.center, scene.camera.up, scene.camera.yfovDegrees, (float)W / (float)H, 0.f, (int)scene.lights.size());
  r.create(VkExtent2D{(uint32_t)W, (uint32_t)H}, {}, nullptr);
  if(!r.ok()) { std::fprintf(stderr, "%s\n", r.error().c_str()); return 3; }

  RtxState st{};
  st.maxDepth = depth; st.maxSamples = 1; st.hdrMultiplier = 1.f; st.pbrMode = pbr; st.maxHeatmap = 65000;
  st.fireflyClampThreshold = getenv("VKRT_FIREFLY") ? (float)atof(getenv("VKRT_FIREFLY")) : integral * 4.f;  // src/sample_example.cpp:110
  st.size[0] = W; st.size[1] = H;
  // progressive loop of the application; the C ABI batches the frames
  vkrt_render_opts o{};
  if(vkrt_render(r.context(), &st, frames, &o) != VKRT_OK) { std::fprintf(stderr, "%s\n", vkrt_last_error(r.context())); return 3; }
  r.sync();

  std::vector<float> acc((size_t)W * H * 4);
  r.readAccumulation(acc.data());
  FILE* f = std::fopen((prefix + ".f32").c_str(), "wb");
  std::fwrite(acc.data(), 4, acc.size(), f);
  std::fclose(f);
  Tonemapper           tm{1.f, 1.f, 1.f, 0.f, 1.f, 1.f, {1.f, 1.f}, 0, 0.5f, 0.5f, 1};
  std::vector<uint8_t> img((size_t)W * H * 4);
  r.tonemapTo(tm, img.data());
  f = std::fopen((prefix + ".ppm").c_str(), "wb");
  std::fprintf(f, "P6\n%d %d\n255\n", W, H);
  for(int i = 0; i < W * H; i++) std::fwrite(&img[4 * (size_t)i], 1, 3, f);
  std::fclose(f);
  if(!r.ok()) { std::fprintf(stderr, "%s\n", r.error().c_str()); return 3; }
  vkrt_stats s{};
  vkrt_get_stats(r.context(), &s);
  std::printf("%d frames %dx%d: %llu closest + %llu shadow rays, BVH %u nodes (build %.2f ms) -> %s.ppm / %s.f32\n", frames, W, H,
              (unsigned long long)s.closest_rays, (unsigned long long)s.shadow_rays, s.num_bvh_nodes, s.ms_build, prefix.c_str(), prefix.c_str());
  r.destroy();
  return 0;
}
