// gltf_scene.cpp -- see gltf_scene.hpp.  References into the reference tree are the *call sites* of the importer
// (nvh::GltfScene itself is not vendored): src/scene.cpp:56-118,123-155,190-274,280-382,447-580.
#include "gltf_scene.hpp"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include "cuda_renderer.hpp"
#include "mini_json.hpp"

namespace vkrt_host {
namespace {

std::string dirOf(const std::string& p)
{
  size_t k = p.find_last_of("/\\");
  return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

bool decodeBase64(const std::string& in, std::vector<uint8_t>& out)
{
  static int T[256];
  static bool init = false;
  if(!init)
  {
    for(int& t : T) t = -1;
    const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for(int i = 0; i < 64; i++) T[(unsigned char)a[i]] = i;
    init = true;
  }
  uint32_t val = 0;
  int      bits = -8;
  for(unsigned char c : in)
  {
    if(c == '=') break;
    if(T[c] < 0) continue;
    val = (val << 6) | (uint32_t)T[c];
    bits += 6;
    if(bits >= 0)
    {
      out.push_back((uint8_t)((val >> bits) & 0xff));
      bits -= 8;
    }
  }
  return true;
}

// column-major 4x4 helpers (glm conventions)
void matIdentity(float* m) { std::memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.f; }
void matMul(const float* a, const float* b, float* out)  // out = a * b
{
  float r[16];
  for(int c = 0; c < 4; c++)
    for(int rr = 0; rr < 4; rr++)
      r[c * 4 + rr] = a[0 * 4 + rr] * b[c * 4 + 0] + a[1 * 4 + rr] * b[c * 4 + 1] + a[2 * 4 + rr] * b[c * 4 + 2] + a[3 * 4 + rr] * b[c * 4 + 3];
  std::memcpy(out, r, 64);
}
void matFromTRS(const float* t, const float* q, const float* s, float* m)
{
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  float R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y + z * w), 2 * (x * z - y * w), 2 * (x * y - z * w), 1 - 2 * (x * x + z * z), 2 * (y * z + x * w),
                2 * (x * z + y * w), 2 * (y * z - x * w), 1 - 2 * (x * x + y * y)};  // columns
  matIdentity(m);
  for(int c = 0; c < 3; c++)
    for(int r = 0; r < 3; r++)
      m[c * 4 + r] = R[c * 3 + r] * s[c];
  m[12] = t[0]; m[13] = t[1]; m[14] = t[2];
}

struct Doc
{
  mj::Value                          root;
  std::vector<std::vector<uint8_t>>  buffers;
  std::string                        dir;
};

// One typed element array inside a buffer view: `count` elements of `ncomp` components of `ctype`, `wantComps` floats out per element.
bool readElements(const Doc& d, int bufferView, size_t accOffset, int ctype, bool normalized, int ncomp, int wantComps, size_t count,
                  bool tightlyPacked, float* out, std::string& err)
{
  const mj::Value& bv = d.root["bufferViews"][(size_t)bufferView];
  if(!bv.isObject()) { err = "bad bufferView index"; return false; }
  const int bi = bv["buffer"].integer(0);
  if(bi < 0 || (size_t)bi >= d.buffers.size()) { err = "bad buffer index"; return false; }
  const std::vector<uint8_t>& buf = d.buffers[(size_t)bi];
  const int    csize  = (ctype == 5120 || ctype == 5121) ? 1 : (ctype == 5122 || ctype == 5123) ? 2 : 4;
  const size_t off    = (size_t)bv["byteOffset"].number(0) + accOffset;
  size_t       stride = tightlyPacked ? 0 : (size_t)bv["byteStride"].number(0);
  if(stride == 0) stride = (size_t)ncomp * csize;
  if(count && (off > buf.size() || (count - 1) > (buf.size() - off) / stride || off + (count - 1) * stride + (size_t)ncomp * csize > buf.size()))
  {
    err = "accessor out of buffer range";
    return false;
  }
  for(size_t i = 0; i < count; i++)
  {
    const uint8_t* p = buf.data() + off + i * stride;
    for(int c = 0; c < std::min(ncomp, wantComps); c++)
    {
      float v = 0.f;
      switch(ctype)
      {
        case 5126: std::memcpy(&v, p + 4 * c, 4); break;
        case 5125: { uint32_t u; std::memcpy(&u, p + 4 * c, 4); v = (float)u; break; }
        case 5123: { uint16_t u; std::memcpy(&u, p + 2 * c, 2); v = normalized ? u / 65535.0f : (float)u; break; }
        case 5121: { uint8_t u = p[c]; v = normalized ? u / 255.0f : (float)u; break; }
        case 5122: { int16_t q; std::memcpy(&q, p + 2 * c, 2); v = normalized ? std::max(q / 32767.0f, -1.0f) : (float)q; break; }
        case 5120: { int8_t q = (int8_t)p[c]; v = normalized ? std::max(q / 127.0f, -1.0f) : (float)q; break; }
        default: err = "unsupported component type"; return false;
      }
      out[i * (size_t)wantComps + c] = v;
    }
    if(wantComps == 4 && ncomp == 3)
      out[i * 4 + 3] = 1.0f;  // COLOR_0 as VEC3
  }
  return true;
}

// Accessor -> floats, `wantComps` per element.  Sparse accessors (glTF 2.0 section 3.6.2.3, resolved by tinygltf before the
// reference's importer sees the data): the base array (zeros without a bufferView) with `sparse.count` elements replaced.
bool readAccessor(const Doc& d, int idx, int wantComps, std::vector<float>& out, std::string& err, uint32_t* countOut = nullptr)
{
  const mj::Value& acc = d.root["accessors"][(size_t)idx];
  if(!acc.isObject()) { err = "bad accessor index"; return false; }
  const int    ctype = acc["componentType"].integer();
  const size_t count = (size_t)acc["count"].number();
  const std::string& type = acc["type"].string();
  const int    ncomp = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : type == "MAT4" ? 16 : 0;
  const bool   normalized = acc["normalized"].boolean(false);
  if(!ncomp) { err = "unsupported accessor type " + type; return false; }
  if(count > (size_t)1 << 31) { err = "accessor count too large"; return false; }
  if(countOut) *countOut = (uint32_t)count;
  out.assign(count * (size_t)wantComps, 0.f);
  if(wantComps == 4 && ncomp == 3)
    for(size_t i = 0; i < count; i++) out[i * 4 + 3] = 1.0f;
  if(acc.has("bufferView")
     && !readElements(d, acc["bufferView"].integer(), (size_t)acc["byteOffset"].number(0), ctype, normalized, ncomp, wantComps, count, false, out.data(), err))
    return false;
  if(acc.has("sparse"))
  {
    const mj::Value& sp = acc["sparse"];
    const size_t     n  = (size_t)sp["count"].number(0);
    if(n > count) { err = "sparse accessor: more replaced elements than elements"; return false; }
    std::vector<float> where(n), vals(n * (size_t)wantComps, 0.f);
    const int ictype = sp["indices"]["componentType"].integer();
    if(ictype != 5121 && ictype != 5123 && ictype != 5125) { err = "sparse accessor: bad index component type"; return false; }
    if(n > ((size_t)1 << 24) && ictype == 5125) { err = "sparse accessor: too many replaced elements"; return false; }  // indices pass through float
    if(!readElements(d, sp["indices"]["bufferView"].integer(-1), (size_t)sp["indices"]["byteOffset"].number(0), ictype, false, 1, 1, n, true, where.data(), err))
      return false;
    if(!readElements(d, sp["values"]["bufferView"].integer(-1), (size_t)sp["values"]["byteOffset"].number(0), ctype, normalized, ncomp, wantComps, n, true,
                     vals.data(), err))
      return false;
    for(size_t k = 0; k < n; k++)
    {
      const size_t i = (size_t)where[k];
      if(i >= count) { err = "sparse accessor: index out of range"; return false; }
      std::memcpy(&out[i * (size_t)wantComps], &vals[k * (size_t)wantComps], sizeof(float) * (size_t)wantComps);
    }
  }
  return true;
}

bool readIndices(const Doc& d, int idx, std::vector<uint32_t>& out, std::string& err)
{
  std::vector<float> tmp;  // exact for < 2^24; read integers directly instead
  const mj::Value& acc = d.root["accessors"][(size_t)idx];
  const int    ctype = acc["componentType"].integer();
  const size_t count = (size_t)acc["count"].number();
  if(!acc.isObject() || !acc.has("bufferView") || acc.has("sparse")) { err = "index accessor without a plain bufferView"; return false; }
  if(ctype != 5121 && ctype != 5123 && ctype != 5125) { err = "bad index component type"; return false; }
  const mj::Value& bv = d.root["bufferViews"][(size_t)acc["bufferView"].integer()];
  const int    bi = bv["buffer"].integer(0);
  if(bi < 0 || (size_t)bi >= d.buffers.size()) { err = "bad buffer index"; return false; }
  const std::vector<uint8_t>& buf = d.buffers[(size_t)bi];
  const size_t csize = ctype == 5121 ? 1 : ctype == 5123 ? 2 : 4;
  const size_t off = (size_t)bv["byteOffset"].number(0) + (size_t)acc["byteOffset"].number(0);
  if(off > buf.size() || count > (buf.size() - off) / csize) { err = "index accessor out of range"; return false; }
  out.resize(count);
  for(size_t i = 0; i < count; i++)
  {
    const uint8_t* p = buf.data() + off + i * csize;
    if(csize == 1) out[i] = p[0];
    else if(csize == 2) { uint16_t u; std::memcpy(&u, p, 2); out[i] = u; }
    else { uint32_t u; std::memcpy(&u, p, 4); out[i] = u; }
  }
  return true;
}

void vec3Norm(float* v)
{
  float l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if(l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; }
}

// Importer behaviour for missing attributes: normals from faces, tangents from the uv parametrisation.
void generateNormals(PrimMeshData& m)
{
  const size_t n = m.positions.size() / 3;
  m.normals.assign(n * 3, 0.f);
  for(size_t t = 0; t + 2 < m.indices.size(); t += 3)
  {
    const uint32_t i0 = m.indices[t], i1 = m.indices[t + 1], i2 = m.indices[t + 2];
    const float *a = &m.positions[3 * i0], *b = &m.positions[3 * i1], *c = &m.positions[3 * i2];
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float fn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    for(uint32_t i : {i0, i1, i2})
      for(int k = 0; k < 3; k++) m.normals[3 * i + k] += fn[k];
  }
  for(size_t i = 0; i < n; i++) vec3Norm(&m.normals[3 * i]);
}
void generateTangents(PrimMeshData& m)
{
  const size_t n = m.positions.size() / 3;
  std::vector<float> tan(n * 3, 0.f), bit(n * 3, 0.f);
  for(size_t t = 0; t + 2 < m.indices.size(); t += 3)
  {
    const uint32_t i0 = m.indices[t], i1 = m.indices[t + 1], i2 = m.indices[t + 2];
    const float *a = &m.positions[3 * i0], *b = &m.positions[3 * i1], *c = &m.positions[3 * i2];
    const float *ua = &m.texcoords[2 * i0], *ub = &m.texcoords[2 * i1], *uc = &m.texcoords[2 * i2];
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const float d1[2] = {ub[0] - ua[0], ub[1] - ua[1]}, d2[2] = {uc[0] - ua[0], uc[1] - ua[1]};
    const float det = d1[0] * d2[1] - d2[0] * d1[1];
    if(std::fabs(det) < 1e-12f)
      continue;
    const float r = 1.0f / det;
    for(int k = 0; k < 3; k++)
    {
      const float ft = (e1[k] * d2[1] - e2[k] * d1[1]) * r, fb = (e2[k] * d1[0] - e1[k] * d2[0]) * r;
      for(uint32_t i : {i0, i1, i2}) { tan[3 * i + k] += ft; bit[3 * i + k] += fb; }
    }
  }
  m.tangents.assign(n * 4, 0.f);
  for(size_t i = 0; i < n; i++)
  {
    const float* nn = &m.normals[3 * i];
    float        t[3] = {tan[3 * i], tan[3 * i + 1], tan[3 * i + 2]};
    const float  d = nn[0] * t[0] + nn[1] * t[1] + nn[2] * t[2];
    for(int k = 0; k < 3; k++) t[k] -= nn[k] * d;
    if(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] < 1e-20f)
    {
      const float alt[3] = {std::fabs(nn[0]) < 0.9f ? 1.f : 0.f, std::fabs(nn[0]) < 0.9f ? 0.f : 1.f, 0.f};
      const float da = nn[0] * alt[0] + nn[1] * alt[1] + nn[2] * alt[2];
      for(int k = 0; k < 3; k++) t[k] = alt[k] - nn[k] * da;
    }
    vec3Norm(t);
    const float cx = nn[1] * t[2] - nn[2] * t[1], cy = nn[2] * t[0] - nn[0] * t[2], cz = nn[0] * t[1] - nn[1] * t[0];
    const float h = (cx * bit[3 * i] + cy * bit[3 * i + 1] + cz * bit[3 * i + 2]) < 0.f ? -1.f : 1.f;
    m.tangents[4 * i] = t[0]; m.tangents[4 * i + 1] = t[1]; m.tangents[4 * i + 2] = t[2]; m.tangents[4 * i + 3] = h;
  }
}

int texIndex(const mj::Value& texInfo) { return texInfo.isObject() ? texInfo["index"].integer(-1) : -1; }
void readVec(const mj::Value& v, float* out, int n)
{
  if(v.isArray())
    for(int i = 0; i < n && (size_t)i < v.size(); i++) out[i] = (float)v[(size_t)i].number(out[i]);
}
uint32_t unorm8(float v) { return (uint32_t)(unsigned char)std::round(std::min(std::max(v, 0.0f), 1.0f) * 255.f); }

GltfShadeMaterial defaultMaterial()
{
  GltfShadeMaterial m{};
  m.pbrBaseColorFactor[0] = m.pbrBaseColorFactor[1] = m.pbrBaseColorFactor[2] = m.pbrBaseColorFactor[3] = 1.f;
  m.pbrBaseColorTexture = m.pbrMetallicRoughnessTexture = m.emissiveTexture = m.normalTexture = -1;
  m.transmissionTexture = m.thicknessTexture = m.clearcoatTexture = m.clearcoatRoughnessTexture = -1;
  m.pbrMetallicFactor = 1.f; m.pbrRoughnessFactor = 1.f;
  m.alphaMode = VKRT_ALPHA_OPAQUE; m.alphaCutoff = 0.5f; m.normalTextureScale = 1.f;
  for(int k = 0; k < 4; k++) m.uvTransform[5 * k] = 1.f;
  m.ior = 1.5f;
  m.anisotropyDirection[1] = 1.f;  // (sin 0, cos 0, 0), src/scene.cpp:365
  m.attenuationColor[0] = m.attenuationColor[1] = m.attenuationColor[2] = 1.f;
  m.attenuationDistance = FLT_MAX;
  return m;
}

// src/scene.cpp:339-382 fed from the glTF material object
GltfShadeMaterial importMaterial(const mj::Value& jm)
{
  GltfShadeMaterial m = defaultMaterial();
  const mj::Value&  pbr = jm["pbrMetallicRoughness"];
  readVec(pbr["baseColorFactor"], m.pbrBaseColorFactor, 4);
  m.pbrBaseColorTexture         = texIndex(pbr["baseColorTexture"]);
  m.pbrMetallicFactor           = (float)pbr["metallicFactor"].number(1.0);
  m.pbrRoughnessFactor          = (float)pbr["roughnessFactor"].number(1.0);
  m.pbrMetallicRoughnessTexture = texIndex(pbr["metallicRoughnessTexture"]);
  m.emissiveTexture             = texIndex(jm["emissiveTexture"]);
  readVec(jm["emissiveFactor"], m.emissiveFactor, 3);
  const std::string& am = jm["alphaMode"].string();
  m.alphaMode   = am == "MASK" ? VKRT_ALPHA_MASK : (am == "BLEND" ? VKRT_ALPHA_BLEND : VKRT_ALPHA_OPAQUE);
  m.alphaCutoff = (float)jm["alphaCutoff"].number(0.5);
  m.doubleSided = jm["doubleSided"].boolean(false) ? 1 : 0;
  m.normalTexture      = texIndex(jm["normalTexture"]);
  m.normalTextureScale = (float)jm["normalTexture"]["scale"].number(1.0);
  // KHR_texture_transform of the base colour texture -> 3x3 (translation * rotation * scale), widened to mat4 (src/scene.cpp:360)
  const mj::Value& tt = pbr["baseColorTexture"]["extensions"]["KHR_texture_transform"];
  if(tt.isObject())
  {
    float off[2] = {0, 0}, sc[2] = {1, 1};
    readVec(tt["offset"], off, 2);
    readVec(tt["scale"], sc, 2);
    const float r = (float)tt["rotation"].number(0.0), c = std::cos(r), s = std::sin(r);
    // column-major 3x3: T * R * S
    const float m3[9] = {c * sc[0], s * sc[0], 0.f, -s * sc[1], c * sc[1], 0.f, off[0], off[1], 1.f};
    std::memset(m.uvTransform, 0, sizeof(m.uvTransform));
    for(int col = 0; col < 3; col++)
      for(int row = 0; row < 3; row++)
        m.uvTransform[col * 4 + row] = m3[col * 3 + row];
    m.uvTransform[15] = 1.f;
  }
  const mj::Value& ext = jm["extensions"];
  m.unlit = ext.has("KHR_materials_unlit") ? 1 : 0;
  const mj::Value& tr = ext["KHR_materials_transmission"];
  m.transmissionFactor  = (float)tr["transmissionFactor"].number(0.0);
  m.transmissionTexture = texIndex(tr["transmissionTexture"]);
  m.ior                 = (float)ext["KHR_materials_ior"]["ior"].number(1.5);
  const mj::Value& an = ext["KHR_materials_anisotropy"];
  m.anisotropy = (float)an["anisotropyStrength"].number(an["anisotropy"].number(0.0));
  {
    const float rot = (float)an["anisotropyRotation"].number(0.0);
    m.anisotropyDirection[0] = std::sin(rot); m.anisotropyDirection[1] = std::cos(rot); m.anisotropyDirection[2] = 0.f;
  }
  const mj::Value& vol = ext["KHR_materials_volume"];
  readVec(vol["attenuationColor"], m.attenuationColor, 3);
  m.thicknessFactor     = (float)vol["thicknessFactor"].number(0.0);
  m.thicknessTexture    = texIndex(vol["thicknessTexture"]);
  m.attenuationDistance = vol.has("attenuationDistance") ? (float)vol["attenuationDistance"].number() : FLT_MAX;
  const mj::Value& cc = ext["KHR_materials_clearcoat"];
  m.clearcoatFactor           = (float)cc["clearcoatFactor"].number(0.0);
  m.clearcoatRoughness        = (float)cc["clearcoatRoughnessFactor"].number(0.0);
  m.clearcoatTexture          = texIndex(cc["clearcoatTexture"]);
  m.clearcoatRoughnessTexture = texIndex(cc["clearcoatRoughnessTexture"]);
  const mj::Value& sh = ext["KHR_materials_sheen"];
  float sheen[4] = {0, 0, 0, 0};
  readVec(sh["sheenColorFactor"], sheen, 3);
  sheen[3] = (float)sh["sheenRoughnessFactor"].number(0.0);
  m.sheen  = unorm8(sheen[0]) | (unorm8(sheen[1]) << 8) | (unorm8(sheen[2]) << 16) | (unorm8(sheen[3]) << 24);  // src/scene.cpp:375
  return m;
}

}  // namespace

bool loadGltf(const std::string& path, GltfSceneData& out, std::string& err)
{
  out = GltfSceneData{};
  std::vector<uint8_t> file;
  if(!readFile(path, file)) { err = "cannot read " + path; return false; }
  Doc d;
  d.dir = dirOf(path);
  std::string          json;
  std::vector<uint8_t> glbBin;
  if(file.size() >= 20 && !std::memcmp(file.data(), "glTF", 4))
  {
    size_t p = 12;
    while(p + 8 <= file.size())
    {
      uint32_t len, type;
      std::memcpy(&len, &file[p], 4);
      std::memcpy(&type, &file[p + 4], 4);
      if(p + 8 + len > file.size()) { err = "truncated GLB"; return false; }
      if(type == 0x4E4F534A) json.assign(reinterpret_cast<const char*>(&file[p + 8]), len);
      else if(type == 0x004E4942) glbBin.assign(file.begin() + (long)p + 8, file.begin() + (long)(p + 8 + len));
      p += 8 + ((len + 3) & ~3u);
    }
  }
  else
    json.assign(file.begin(), file.end());
  if(!mj::parse(json, d.root, err)) { err = "JSON: " + err; return false; }

  // buffers
  const mj::Value& jb = d.root["buffers"];
  for(size_t i = 0; i < jb.size(); i++)
  {
    std::vector<uint8_t> data;
    const std::string&   uri = jb[i]["uri"].string();
    if(uri.empty())
      data = glbBin;
    else if(uri.rfind("data:", 0) == 0)
    {
      size_t k = uri.find("base64,");
      if(k == std::string::npos) { err = "unsupported data URI"; return false; }
      decodeBase64(uri.substr(k + 7), data);
    }
    else if(!readFile(d.dir + "/" + uri, data)) { err = "cannot read buffer " + uri; return false; }
    d.buffers.push_back(std::move(data));
  }

  // materials (importMaterials): a default one is appended when the file has none
  const mj::Value& jmats = d.root["materials"];
  for(size_t i = 0; i < jmats.size(); i++) out.materials.push_back(importMaterial(jmats[i]));
  if(out.materials.empty())
    out.materials.push_back(defaultMaterial());

  // textures: tinygltf decodes every image (PNG, JPEG) to RGBA8 (src/scene.cpp:527-541); undecodable -> 1x1 white (scene.cpp:496-502)
  const mj::Value& jtex = d.root["textures"];
  const mj::Value& jimg = d.root["images"];
  const mj::Value& jsmp = d.root["samplers"];
  std::vector<ImageRGBA8> images(jimg.size());
  for(size_t i = 0; i < jimg.size(); i++)
  {
    std::vector<uint8_t> bytes;
    const std::string&   uri = jimg[i]["uri"].string();
    if(!uri.empty())
    {
      if(uri.rfind("data:", 0) == 0) { size_t k = uri.find("base64,"); if(k != std::string::npos) decodeBase64(uri.substr(k + 7), bytes); }
      else readFile(d.dir + "/" + uri, bytes);
    }
    else if(jimg[i].has("bufferView"))
    {
      const mj::Value& bv = d.root["bufferViews"][(size_t)jimg[i]["bufferView"].integer()];
      const int bi = bv["buffer"].integer(0);
      if(bi >= 0 && (size_t)bi < d.buffers.size())
      {
        const size_t off = (size_t)bv["byteOffset"].number(0), len = (size_t)bv["byteLength"].number(0);
        if(off + len <= d.buffers[(size_t)bi].size())
          bytes.assign(d.buffers[(size_t)bi].begin() + (long)off, d.buffers[(size_t)bi].begin() + (long)(off + len));
      }
    }
    std::string ierr;
    if(bytes.empty() || !decodeImage(bytes.data(), bytes.size(), images[i], ierr))
    {
      images[i].w = images[i].h = 1;
      images[i].rgba = {255, 255, 255, 255};
    }
  }
  for(size_t i = 0; i < jtex.size(); i++)
  {
    TextureData t;
    const int   src = jtex[i]["source"].integer(-1);
    if(src >= 0 && (size_t)src < images.size())
      t.image = images[(size_t)src];
    else { t.image.w = t.image.h = 1; t.image.rgba = {255, 255, 255, 255}; }
    const int smp = jtex[i]["sampler"].integer(-1);
    if(smp > -1 && (size_t)smp < jsmp.size())
    {
      // tinygltf defaults: filters -1 (unset), wrap 10497; gltfSamplerToVulkan maps unknown filter codes to NEAREST (src/scene.cpp:451-476)
      t.magFilter = jsmp[(size_t)smp]["magFilter"].integer(-1);
      t.minFilter = jsmp[(size_t)smp]["minFilter"].integer(-1);
      t.wrapS     = jsmp[(size_t)smp]["wrapS"].integer(10497);
      t.wrapT     = jsmp[(size_t)smp]["wrapT"].integer(10497);
    }
    out.textures.push_back(std::move(t));
  }

  // drawable nodes (importDrawableNodes): flatten the default scene
  struct Item { int node; float world[16]; int depth; };
  std::vector<Item> stack;
  const mj::Value&  jnodes = d.root["nodes"];
  const int         sceneIdx = d.root["scene"].integer(0);
  const mj::Value&  roots = d.root["scenes"][(size_t)std::max(0, sceneIdx)]["nodes"];
  for(size_t i = 0; i < roots.size(); i++) { Item it; it.node = roots[i].integer(); it.depth = 0; matIdentity(it.world); stack.push_back(it); }
  if(roots.size() == 0)
    for(size_t i = 0; i < jnodes.size(); i++) { Item it; it.node = (int)i; it.depth = 0; matIdentity(it.world); stack.push_back(it); }
  size_t visited = 0;
  std::map<std::string, uint32_t> meshCache;
  int                             defaultMat = -1;
  const mj::Value& jmeshes = d.root["meshes"];
  const mj::Value& jlights = d.root["extensions"]["KHR_lights_punctual"]["lights"];
  const mj::Value& jcams   = d.root["cameras"];
  float bmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  std::vector<Item> work(stack.rbegin(), stack.rend());
  while(!work.empty())
  {
    Item it = work.back();
    work.pop_back();
    // a valid node hierarchy is a forest: a path longer than the node count, or more visits than a generous multiple of it
    // (instanced sub-trees), means the graph has a cycle
    if(it.node < 0 || (size_t)it.node >= jnodes.size()) { err = "node index out of range"; return false; }
    if((size_t)it.depth > jnodes.size() || ++visited > 64 * jnodes.size() + 1024) { err = "cyclic node hierarchy"; return false; }
    const mj::Value& jn = jnodes[(size_t)it.node];
    float local[16], world[16];
    if(jn.has("matrix")) { matIdentity(local); readVec(jn["matrix"], local, 16); }
    else
    {
      float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
      readVec(jn["translation"], t, 3); readVec(jn["rotation"], q, 4); readVec(jn["scale"], s, 3);
      matFromTRS(t, q, s, local);
    }
    matMul(it.world, local, world);
    const mj::Value& ch = jn["children"];
    for(size_t i = ch.size(); i-- > 0;) { Item c; c.node = ch[i].integer(); c.depth = it.depth + 1; std::memcpy(c.world, world, 64); work.push_back(c); }

    if(jn.has("mesh"))
    {
      const mj::Value& prims = jmeshes[(size_t)jn["mesh"].integer()]["primitives"];
      for(size_t pi = 0; pi < prims.size(); pi++)
      {
        const mj::Value& pr = prims[pi];
        if(pr["mode"].integer(4) != 4)
          continue;  // triangles only
        const mj::Value& at = pr["attributes"];
        if(!at.has("POSITION"))
          continue;
        const std::string key = std::to_string(at["POSITION"].integer()) + ":" + std::to_string(at["NORMAL"].integer()) + ":" + std::to_string(at["TEXCOORD_0"].integer())
                                + ":" + std::to_string(at["TANGENT"].integer()) + ":" + std::to_string(at["COLOR_0"].integer()) + ":"
                                + std::to_string(pr["indices"].integer()) + ":" + std::to_string(pr["material"].integer());
        uint32_t meshIdx;
        auto     found = meshCache.find(key);
        if(found != meshCache.end())
          meshIdx = found->second;
        else
        {
          PrimMeshData m;
          uint32_t     nv = 0;
          if(!readAccessor(d, at["POSITION"].integer(), 3, m.positions, err, &nv)) return false;
          if(pr.has("indices")) { if(!readIndices(d, pr["indices"].integer(), m.indices, err)) return false; }
          else { m.indices.resize(nv); for(uint32_t i = 0; i < nv; i++) m.indices[i] = i; }
          m.indices.resize(m.indices.size() / 3 * 3);
          for(uint32_t i : m.indices) if(i >= nv) { err = "index out of range"; return false; }
          // every attribute accessor of a primitive must have POSITION's count (glTF 2.0 section 3.7.2.1): the packer reads nv of each
          auto attr = [&](const char* name, int comps, std::vector<float>& dst) -> bool {
            uint32_t cnt = 0;
            if(!readAccessor(d, at[name].integer(), comps, dst, err, &cnt)) return false;
            if(cnt != nv) { err = std::string("attribute ") + name + " has a different count than POSITION"; return false; }
            return true;
          };
          if(at.has("TEXCOORD_0")) { if(!attr("TEXCOORD_0", 2, m.texcoords)) return false; }
          else m.texcoords.assign((size_t)nv * 2, 0.f);
          if(at.has("NORMAL")) { if(!attr("NORMAL", 3, m.normals)) return false; }
          else generateNormals(m);
          if(at.has("TANGENT")) { if(!attr("TANGENT", 4, m.tangents)) return false; }
          else generateTangents(m);
          if(at.has("COLOR_0")) { if(!attr("COLOR_0", 4, m.colors)) return false; }
          else m.colors.assign((size_t)nv * 4, 1.f);
          int mat = pr["material"].integer(-1);
          if(mat < 0 || (size_t)mat >= jmats.size())
          {
            // primitives without a material share one default material (appended unless the file has no materials at all)
            if(defaultMat < 0)
            {
              if(jmats.size() != 0)
                out.materials.push_back(defaultMaterial());
              defaultMat = (int)out.materials.size() - 1;
            }
            mat = defaultMat;
          }
          m.materialIndex = mat;
          meshIdx = (uint32_t)out.primMeshes.size();
          meshCache[key] = meshIdx;
          out.primMeshes.push_back(std::move(m));
        }
        NodeData nd;
        std::memcpy(nd.worldMatrix, world, 64);
        nd.primMesh = meshIdx;
        out.nodes.push_back(nd);
        const PrimMeshData& pm = out.primMeshes[meshIdx];
        out.triangleCount += pm.indices.size() / 3;
        for(size_t v = 0; v + 2 < pm.positions.size(); v += 3)
          for(int k = 0; k < 3; k++)
          {
            const float w = world[0 * 4 + k] * pm.positions[v] + world[1 * 4 + k] * pm.positions[v + 1] + world[2 * 4 + k] * pm.positions[v + 2] + world[12 + k];
            bmin[k] = std::min(bmin[k], w);
            bmax[k] = std::max(bmax[k], w);
          }
      }
    }
    // KHR_lights_punctual (src/scene.cpp:304-333)
    const mj::Value& lref = jn["extensions"]["KHR_lights_punctual"]["light"];
    if(lref.isNumber() && (size_t)lref.integer() < jlights.size())
    {
      const mj::Value& jl = jlights[(size_t)lref.integer()];
      Light l{};
      vkrt_light_from_matrix(world, &l);
      l.color[0] = l.color[1] = l.color[2] = 1.f;
      readVec(jl["color"], l.color, 3);
      l.intensity    = (float)jl["intensity"].number(1.0);
      l.range        = (float)jl["range"].number(0.0);
      l.innerConeCos = (float)std::cos(jl["spot"]["innerConeAngle"].number(0.0));
      l.outerConeCos = (float)std::cos(jl["spot"]["outerConeAngle"].number(0.7853981633974483));
      const std::string& t = jl["type"].string();
      l.type = t == "point" ? VKRT_LIGHT_POINT : (t == "spot" ? VKRT_LIGHT_SPOT : VKRT_LIGHT_DIRECTIONAL);
      out.lights.push_back(l);
    }
    // first camera (src/scene.cpp:280-292)
    if(jn.has("camera") && !out.camera.fromFile && (size_t)jn["camera"].integer() < jcams.size())
    {
      const mj::Value& jc = jcams[(size_t)jn["camera"].integer()];
      out.camera.fromFile = true;
      for(int k = 0; k < 3; k++) { out.camera.eye[k] = world[12 + k]; out.camera.up[k] = world[4 + k]; out.camera.center[k] = world[12 + k] - world[8 + k]; }
      out.camera.yfovDegrees = (float)(jc["perspective"]["yfov"].number(0.7853981633974483) * 180.0 / 3.14159265358979323846);
    }
  }
  if(out.triangleCount)
    for(int k = 0; k < 3; k++) { out.boundsMin[k] = bmin[k]; out.boundsMax[k] = bmax[k]; }
  if(out.camera.fromFile && out.triangleCount)
  {
    // look-at point: along the view axis at the distance of the scene centre
    float c[3], f[3], dist = 0.f;
    for(int k = 0; k < 3; k++) { c[k] = 0.5f * (bmin[k] + bmax[k]) - out.camera.eye[k]; f[k] = out.camera.center[k] - out.camera.eye[k]; }
    vec3Norm(f);
    dist = std::max(1e-3f, c[0] * f[0] + c[1] * f[1] + c[2] * f[2]);
    for(int k = 0; k < 3; k++) out.camera.center[k] = out.camera.eye[k] + f[k] * dist;
  }
  else if(out.triangleCount)
  {
    // CameraManip.fit (src/scene.cpp:296-297): look at the box centre from +Z at a distance that frames it
    float ext = 0.f;
    for(int k = 0; k < 3; k++) { out.camera.center[k] = 0.5f * (bmin[k] + bmax[k]); ext = std::max(ext, bmax[k] - bmin[k]); }
    out.camera.eye[0] = out.camera.center[0]; out.camera.eye[1] = out.camera.center[1];
    out.camera.eye[2] = out.camera.center[2] + 0.5f * ext / std::tan(0.5f * out.camera.yfovDegrees * 3.14159265f / 180.f) + 0.5f * (bmax[2] - bmin[2]);
  }
  return true;
}

void uploadScene(const GltfSceneData& s, CudaRenderer& r, int buildMode)
{
  r.beginScene();
  for(const auto& m : s.primMeshes)
    r.feedPrimMesh(m.positions.data(), m.normals.data(), m.tangents.data(), m.texcoords.data(), m.colors.data(), (uint32_t)(m.positions.size() / 3),
                   m.indices.data(), (uint32_t)m.indices.size(), m.materialIndex);
  for(const auto& n : s.nodes)
    r.feedNode(n.worldMatrix, n.primMesh);
  r.feedMaterials(s.materials.data(), (uint32_t)s.materials.size());
  r.feedLights(s.lights.empty() ? nullptr : s.lights.data(), (uint32_t)s.lights.size());
  for(const auto& t : s.textures)
    r.feedTexture(t.image.rgba.data(), t.image.w, t.image.h, t.magFilter, t.minFilter, t.wrapS, t.wrapT);
  r.commitScene(buildMode);
}

}  // namespace vkrt_host
