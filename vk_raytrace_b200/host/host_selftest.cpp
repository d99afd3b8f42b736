// host_selftest.cpp -- decodes an image with the host decoders and dumps the result for the Python tests:
//   vkrt_host_selftest png|jpg <file> <out.raw>   -> "w h" on stdout, RGBA8 bytes in out.raw (decoder chosen by signature)
//   vkrt_host_selftest hdr <file> <out.raw>   -> "w h" on stdout, RGBA32F floats in out.raw
//   vkrt_host_selftest gltf <file>            -> scene summary on stdout (no GPU needed)
//   vkrt_host_selftest fuzz <file> <seed> <n> -> decodes n seeded mutations of the file (image by signature, .hdr / .pfm, .gltf / .glb); prints
//                                                "ok <accepted> rejected <rejected>"; the test builds this tool with -fsanitize=address,undefined
#include <cstdio>
#include <cstdlib>
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
  if(!std::strcmp(argv[1], "fuzz") && argc > 4)
  {
    std::vector<uint8_t> orig;
    if(!vkrt_host::readFile(argv[2], orig) || orig.empty()) return 2;
    const std::string path = argv[2];
    const auto        ends = [&](const char* e) { const size_t n = std::strlen(e); return path.size() >= n && path.compare(path.size() - n, n, e) == 0; };
    const bool        isGltf = ends(".gltf") || ends(".glb"), isHdr = ends(".hdr") || ends(".pfm"), text = ends(".gltf");
    uint64_t          rng = 0x9E3779B97F4A7C15ull ^ (uint64_t)std::strtoull(argv[3], nullptr, 10);
    const auto        next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
    const int         n = std::atoi(argv[4]);
    int               ok = 0, bad = 0;
    static const char alphabet[] = "0123456789-,[]{}\":.e ";
    const std::string tmp = path + (isGltf ? (text ? ".fuzz.gltf" : ".fuzz.glb") : (ends(".pfm") ? ".fuzz.pfm" : ".fuzz.hdr"));
    for(int k = 0; k < n; k++)
    {
      std::vector<uint8_t> f = orig;
      const uint32_t       kind = next() % 8;
      if(kind == 0 && f.size() > 8)
        f.resize(8 + next() % (f.size() - 8));  // truncation
      const int edits = 1 + (int)(next() % 6);
      for(int e = 0; e < edits; e++)
      {
        const size_t at = next() % f.size();
        f[at] = text ? (uint8_t)alphabet[next() % (sizeof(alphabet) - 1)] : (kind == 1 ? (uint8_t)(f[at] ^ (1u << (next() % 8))) : (uint8_t)next());
      }
      bool good;
      if(isGltf || isHdr)
      {
        FILE* o = std::fopen(tmp.c_str(), "wb");
        if(!o) return 2;
        std::fwrite(f.data(), 1, f.size(), o);
        std::fclose(o);
        if(isGltf) { vkrt_host::GltfSceneData s; good = vkrt_host::loadGltf(tmp, s, err); }
        else { vkrt_host::ImageRGBA32F img; good = vkrt_host::loadHdrOrPfm(tmp, img, err); }
      }
      else
      {
        vkrt_host::ImageRGBA8 img;
        good = vkrt_host::decodeImage(f.data(), f.size(), img, err);
        if(good && img.rgba.size() != (size_t)img.w * img.h * 4) { std::fprintf(stderr, "size mismatch\n"); return 3; }
      }
      good ? ok++ : bad++;
    }
    std::remove(tmp.c_str());
    std::printf("ok %d rejected %d\n", ok, bad);
    return 0;
  }
  return 1;
}
