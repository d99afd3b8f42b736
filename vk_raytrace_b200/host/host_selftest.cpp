// host_selftest.cpp -- decodes an image with the host decoders and dumps the result for the Python tests:
//   vkrt_host_selftest png|jpg <file> <out.raw>   -> "w h" on stdout, RGBA8 bytes in out.raw (decoder chosen by signature)
//   vkrt_host_selftest hdr <file> <out.raw>   -> "w h" on stdout, RGBA32F floats in out.raw
//   vkrt_host_selftest gltf <file>            -> scene summary on stdout (no GPU needed)
#include <cstdio>
#include <cstring>
#include "gltf_scene.hpp"

int main(int argc, char** argv)
{
  if(argc < 3) return 1;
  std::string err;
  if((!std::strcmp(argv[1], "png") || !std::strcmp(argv[1], "jpg")) && argc > 3)
  {
    std::vector<uint8_t> f;
    vkrt_host::ImageRGBA8 img;
    if(!vkrt_host::readFile(argv[2], f) || !vkrt_host::decodeImage(f.data(), f.size(), img, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
    FILE* o = std::fopen(argv[3], "wb");
    std::fwrite(img.rgba.data(), 1, img.rgba.size(), o);
    std::fclose(o);
    std::printf("%u %u\n", img.w, img.h);
    return 0;
  }
  if(!std::strcmp(argv[1], "hdr") && argc > 3)
  {
    vkrt_host::ImageRGBA32F img;
    if(!vkrt_host::loadHdrOrPfm(argv[2], img, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
    FILE* o = std::fopen(argv[3], "wb");
    std::fwrite(img.rgba.data(), 4, img.rgba.size(), o);
    std::fclose(o);
    std::printf("%u %u\n", img.w, img.h);
    return 0;
  }
  if(!std::strcmp(argv[1], "gltf"))
  {
    vkrt_host::GltfSceneData s;
    if(!vkrt_host::loadGltf(argv[2], s, err)) { std::fprintf(stderr, "%s\n", err.c_str()); return 2; }
    std::printf("primMeshes %zu nodes %zu triangles %zu materials %zu textures %zu lights %zu camera %d fov %.6f eye %.6f %.6f %.6f\n", s.primMeshes.size(), s.nodes.size(),
                s.triangleCount, s.materials.size(), s.textures.size(), s.lights.size(), (int)s.camera.fromFile, s.camera.yfovDegrees, s.camera.eye[0], s.camera.eye[1], s.camera.eye[2]);
    // checksums the tests can reproduce from the SceneDesc
    double sumPos = 0, sumIdx = 0;
    for(auto& m : s.primMeshes) { for(float v : m.positions) sumPos += v; for(uint32_t i : m.indices) sumIdx += i; }
    double sumMat = 0;
    for(auto& m : s.materials) { sumMat += m.pbrRoughnessFactor + m.pbrMetallicFactor + m.pbrBaseColorFactor[0] + m.alphaMode + m.doubleSided + m.clearcoatFactor + m.transmissionFactor + m.ior + m.anisotropy + (double)m.sheen; }
    double sumTex = 0;
    for(auto& t : s.textures) { for(uint8_t b : t.image.rgba) sumTex += b; sumTex += 1000003.0 * t.image.w + 7919.0 * t.image.h; }
    std::printf("sumPos %.6f sumIdx %.0f sumMat %.6f sumTex %.0f\n", sumPos, sumIdx, sumMat, sumTex);
    for(auto& l : s.lights) std::printf("light type %d pos %.6f %.6f %.6f dir %.6f %.6f %.6f I %.4f cones %.6f %.6f\n", l.type, l.position[0], l.position[1], l.position[2], l.direction[0], l.direction[1], l.direction[2], l.intensity, l.innerConeCos, l.outerConeCos);
    return 0;
  }
  return 1;
}
