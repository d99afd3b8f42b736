// image_io.cpp -- the image decoders the reference gets from stb_image through tinygltf / stbi_loadf
// (src/scene.cpp:527-541 textures -> RGBA8, src/hdr_sampling.cpp:64 environment -> RGBA32F), restated without the
// dependency: PNG (8/16-bit, all colour types, non-interlaced and Adam7) with an inflate implementation, Radiance .hdr
// (RGBE, flat and new-style RLE) and PFM (exact float, used by the tests); JPEG lives in jpeg_decode.cpp.  Anything else
// (and any file that fails to decode) falls back to the 1x1 white image like the reference's addDefaultImage (src/scene.cpp:496-502).
#include "image_io.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace vkrt_host {

bool readFile(const std::string& path, std::vector<uint8_t>& out)
{
  std::ifstream f(path, std::ios::binary);
  if(!f)
    return false;
  f.seekg(0, std::ios::end);
  std::streamoff n = f.tellg();
  f.seekg(0);
  out.resize((size_t)n);
  if(n)
    f.read(reinterpret_cast<char*>(out.data()), n);
  return (bool)f;
}

// ---- inflate (RFC 1951) -----------------------------------------------------------------------------------
namespace {
struct BitReader
{
  const uint8_t* d;
  size_t         n, pos = 0;
  uint32_t       buf = 0;
  int            cnt = 0;
  bool           bad = false;
  uint32_t bits(int k)
  {
    while(cnt < k)
    {
      if(pos >= n) { bad = true; return 0; }
      buf |= (uint32_t)d[pos++] << cnt;
      cnt += 8;
    }
    uint32_t v = buf & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
    buf >>= k;
    cnt -= k;
    return v;
  }
  void alignByte() { buf = 0; cnt = 0; }
};
struct Huff
{
  uint16_t count[16], symbol[288];
  void build(const uint8_t* lens, int n)
  {
    std::memset(count, 0, sizeof(count));
    for(int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    uint16_t offs[16];
    offs[1] = 0;
    for(int i = 1; i < 15; i++) offs[i + 1] = offs[i] + count[i];
    for(int i = 0; i < n; i++)
      if(lens[i]) symbol[offs[lens[i]]++] = (uint16_t)i;
  }
  int decode(BitReader& br) const
  {
    int code = 0, first = 0, index = 0;
    for(int len = 1; len <= 15; len++)
    {
      code |= (int)br.bits(1);
      int c = count[len];
      if(code - c < first)
        return symbol[index + (code - first)];
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
      if(br.bad) return -1;
    }
    return -1;
  }
};
const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint16_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint16_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflateRaw(const uint8_t* data, size_t n, std::vector<uint8_t>& out, size_t limit)
{
  BitReader br{data, n};
  for(;;)
  {
    const uint32_t last = br.bits(1), type = br.bits(2);
    if(br.bad) return false;
    if(type == 0)
    {
      br.alignByte();
      if(br.pos + 4 > n) return false;
      uint16_t len = (uint16_t)(data[br.pos] | (data[br.pos + 1] << 8));
      br.pos += 4;
      if(br.pos + len > n || out.size() + len > limit) return false;
      out.insert(out.end(), data + br.pos, data + br.pos + len);
      br.pos += len;
    }
    else if(type == 1 || type == 2)
    {
      Huff lit, dist;
      if(type == 1)
      {
        uint8_t l[288];
        for(int i = 0; i < 144; i++) l[i] = 8;
        for(int i = 144; i < 256; i++) l[i] = 9;
        for(int i = 256; i < 280; i++) l[i] = 7;
        for(int i = 280; i < 288; i++) l[i] = 8;
        lit.build(l, 288);
        uint8_t d5[30];
        for(int i = 0; i < 30; i++) d5[i] = 5;
        dist.build(d5, 30);
      }
      else
      {
        const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        uint8_t cl[19] = {0};
        for(int i = 0; i < hclen; i++) cl[order[i]] = (uint8_t)br.bits(3);
        Huff clh;
        clh.build(cl, 19);
        uint8_t lens[320];
        int     i = 0;
        while(i < hlit + hdist)
        {
          int sym = clh.decode(br);
          if(sym < 0 || br.bad) return false;
          if(sym < 16) lens[i++] = (uint8_t)sym;
          else
          {
            int rep = 0;
            uint8_t v = 0;
            if(sym == 16) { if(i == 0) return false; v = lens[i - 1]; rep = 3 + (int)br.bits(2); }
            else if(sym == 17) rep = 3 + (int)br.bits(3);
            else rep = 11 + (int)br.bits(7);
            if(i + rep > hlit + hdist) return false;
            while(rep--) lens[i++] = v;
          }
        }
        lit.build(lens, hlit);
        dist.build(lens + hlit, hdist);
      }
      for(;;)
      {
        int sym = lit.decode(br);
        if(sym < 0 || br.bad) return false;
        if(sym < 256) { if(out.size() >= limit) return false; out.push_back((uint8_t)sym); }
        else if(sym == 256) break;
        else
        {
          sym -= 257;
          if(sym >= 29) return false;
          int len = kLenBase[sym] + (int)br.bits(kLenExtra[sym]);
          int ds  = dist.decode(br);
          if(ds < 0 || ds >= 30) return false;
          size_t d = kDistBase[ds] + br.bits(kDistExtra[ds]);
          if(d > out.size() || out.size() + (size_t)len > limit) return false;
          size_t from = out.size() - d;
          for(int k = 0; k < len; k++) out.push_back(out[from + k]);
        }
      }
    }
    else
      return false;
    if(last) break;
  }
  return true;
}

constexpr uint32_t kMaxImageDim = 1u << 15;  // decoders refuse anything larger before allocating
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c)
{
  int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

bool decodePng(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err)
{
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if(size < 8 || std::memcmp(data, sig, 8) != 0) { err = "not a PNG"; return false; }
  uint32_t w = 0, h = 0;
  int      depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  size_t p = 8;
  while(p + 12 <= size)
  {
    const uint32_t len = be32(data + p);
    const char*    t   = reinterpret_cast<const char*>(data + p + 4);
    const uint8_t* d   = data + p + 8;
    if(p + 12 + len > size) { err = "truncated PNG"; return false; }
    if(!std::strncmp(t, "IHDR", 4))
    {
      if(len != 13) { err = "bad PNG IHDR length"; return false; }
      w = be32(d); h = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
      if(d[10] != 0 || d[11] != 0 || interlace > 1) { err = "unsupported PNG compression / filter / interlace method"; return false; }
    }
    else if(!std::strncmp(t, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
    else if(!std::strncmp(t, "PLTE", 4)) plte.assign(d, d + len);
    else if(!std::strncmp(t, "tRNS", 4)) trns.assign(d, d + len);
    else if(!std::strncmp(t, "IEND", 4)) break;
    p += 12 + len;
  }
  if(w == 0 || h == 0 || idat.size() < 6) { err = "bad PNG header"; return false; }
  if(w > kMaxImageDim || h > kMaxImageDim) { err = "PNG larger than 32768 x 32768"; return false; }
  const int chans = (ctype == 0) ? 1 : (ctype == 2) ? 3 : (ctype == 3) ? 1 : (ctype == 4) ? 2 : (ctype == 6) ? 4 : 0;
  // bit depths the PNG specification allows per colour type: grey 1/2/4/8/16, palette 1/2/4/8, everything else 8/16
  const bool depthOk = (ctype == 0) ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                       : (ctype == 3) ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                      : (depth == 8 || depth == 16);
  if(!chans || !depthOk) { err = "unsupported PNG format"; return false; }
  std::vector<uint8_t> raw;
  {
    // the filtered size the header implies (one filter byte per row, per Adam7 pass when interlaced) must be what was inflated
    auto passBytes = [&](uint32_t x0, uint32_t y0, uint32_t dx, uint32_t dy) -> size_t {
      if(x0 >= w || y0 >= h) return 0;
      const size_t pw = (w - x0 + dx - 1) / dx, ph = (h - y0 + dy - 1) / dy;
      return ph * (1 + (pw * chans * depth + 7) / 8);
    };
    size_t expect = 0;
    if(!interlace) expect = passBytes(0, 0, 1, 1);
    else
    {
      static const uint32_t a7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
      for(auto& q : a7) expect += passBytes(q[0], q[1], q[2], q[3]);
    }
    // skip the zlib header; output beyond the expected size stops the decoder (no decompression bombs)
    if(!inflateRaw(idat.data() + 2, idat.size() - 2, raw, expect)) { err = "PNG inflate failed"; return false; }
    if(raw.size() != expect) { err = "PNG data size does not match its header"; return false; }
  }
  const int bpp = std::max(1, chans * depth / 8);
  out.w = w; out.h = h;
  out.rgba.assign((size_t)w * h * 4, 255);
  auto putPixel = [&](uint32_t x, uint32_t y, const uint8_t* px, int bitIndex) {
    uint8_t* o = &out.rgba[((size_t)y * w + x) * 4];
    auto sample = [&](int c) -> int {
      if(depth == 8) return px[c];
      if(depth == 16) return px[2 * c];  // high byte
      const int shift = 8 - depth - bitIndex;
      int v = (px[0] >> shift) & ((1 << depth) - 1);
      return (ctype == 3) ? v : v * 255 / ((1 << depth) - 1);
    };
    if(ctype == 0) { int g = sample(0); o[0] = o[1] = o[2] = (uint8_t)g; if(trns.size() >= 2 && depth == 8 && g == trns[1]) o[3] = 0; }
    else if(ctype == 2) { o[0] = (uint8_t)sample(0); o[1] = (uint8_t)sample(1); o[2] = (uint8_t)sample(2); }
    else if(ctype == 3)
    {
      size_t i = (size_t)sample(0);
      if(3 * i + 2 < plte.size()) { o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; }
      if(i < trns.size()) o[3] = trns[i];
    }
    else if(ctype == 4) { int g = sample(0); o[0] = o[1] = o[2] = (uint8_t)g; o[3] = (uint8_t)sample(1); }
    else { o[0] = (uint8_t)sample(0); o[1] = (uint8_t)sample(1); o[2] = (uint8_t)sample(2); o[3] = (uint8_t)sample(3); }
  };
  size_t rp = 0;
  auto pass = [&](uint32_t x0, uint32_t y0, uint32_t dx, uint32_t dy) -> bool {
    if(x0 >= w || y0 >= h) return true;
    const uint32_t pw = (w - x0 + dx - 1) / dx, ph = (h - y0 + dy - 1) / dy;
    const size_t   stride = ((size_t)pw * chans * depth + 7) / 8;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for(uint32_t yy = 0; yy < ph; yy++)
    {
      if(rp + 1 + stride > raw.size()) return false;
      const int ft = raw[rp++];
      std::memcpy(cur.data(), &raw[rp], stride);
      rp += stride;
      for(size_t i = 0; i < stride; i++)
      {
        const int a = i >= (size_t)bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= (size_t)bpp ? prev[i - bpp] : 0;
        int v = cur[i];
        if(ft == 1) v += a; else if(ft == 2) v += b; else if(ft == 3) v += (a + b) / 2; else if(ft == 4) v += paeth(a, b, c);
        cur[i] = (uint8_t)v;
      }
      for(uint32_t xx = 0; xx < pw; xx++)
      {
        const size_t bit = (size_t)xx * chans * depth;
        putPixel(x0 + xx * dx, y0 + yy * dy, &cur[bit / 8], (int)(bit % 8));
      }
      prev.swap(cur);
    }
    return true;
  };
  bool ok = true;
  if(!interlace) ok = pass(0, 0, 1, 1);
  else
  {
    static const uint32_t a7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    for(auto& q : a7) ok = ok && pass(q[0], q[1], q[2], q[3]);
  }
  if(!ok) { err = "PNG data too short"; return false; }
  return true;
}

bool decodeImage(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err)
{
  if(size >= 3 && data[0] == 0xFF && data[1] == 0xD8 && data[2] == 0xFF)
    return decodeJpeg(data, size, out, err);
  return decodePng(data, size, out, err);
}

// ---- Radiance RGBE (.hdr) and PFM -> RGBA32F, alpha = 1 (stbi_loadf(..., STBI_rgb_alpha)) ---------------------------------
bool loadHdrOrPfm(const std::string& path, ImageRGBA32F& out, std::string& err)
{
  std::vector<uint8_t> f;
  if(!readFile(path, f)) { err = "cannot read " + path; return false; }
  size_t p = 0;
  auto line = [&]() { std::string s; while(p < f.size() && f[p] != '\n') s += (char)f[p++]; p++; return s; };
  std::string first = line();
  if(first == "PF" || first == "Pf")
  {
    const int ch = first == "PF" ? 3 : 1;
    int w = 0, h = 0;
    std::sscanf(line().c_str(), "%d %d", &w, &h);
    const double scale = std::atof(line().c_str());
    if(w <= 0 || h <= 0 || w > (int)kMaxImageDim || h > (int)kMaxImageDim || p + (size_t)w * h * ch * 4 > f.size()) { err = "bad PFM"; return false; }
    out.w = w; out.h = h;
    out.rgba.resize((size_t)w * h * 4);
    for(int y = 0; y < h; y++)
      for(int x = 0; x < w; x++)
      {
        float v[3] = {0, 0, 0};
        const uint8_t* src = &f[p + ((size_t)(h - 1 - y) * w + x) * ch * 4];  // PFM rows are bottom-up
        for(int c = 0; c < ch; c++)
        {
          uint8_t b[4] = {src[4 * c], src[4 * c + 1], src[4 * c + 2], src[4 * c + 3]};
          if(scale > 0) { std::swap(b[0], b[3]); std::swap(b[1], b[2]); }  // big endian file
          std::memcpy(&v[c], b, 4);
        }
        if(ch == 1) v[1] = v[2] = v[0];
        float* o = &out.rgba[((size_t)y * w + x) * 4];
        o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = 1.0f;
      }
    return true;
  }
  if(first.rfind("#?", 0) != 0) { err = "not a Radiance/PFM file"; return false; }
  int w = 0, h = 0;
  for(;;)
  {
    std::string s = line();
    if(p >= f.size()) { err = "bad HDR header"; return false; }
    if(s.empty()) break;
  }
  {
    std::string res = line();
    char sy, sx; char ay, ax;
    if(std::sscanf(res.c_str(), "%c%c %d %c%c %d", &sy, &ay, &h, &sx, &ax, &w) != 6 || ay != 'Y' || ax != 'X') { err = "unsupported HDR orientation"; return false; }
  }
  if(w <= 0 || h <= 0 || w > (int)kMaxImageDim || h > (int)kMaxImageDim) { err = "bad HDR dimensions"; return false; }
  out.w = w; out.h = h;
  out.rgba.resize((size_t)w * h * 4);
  std::vector<uint8_t> scan((size_t)w * 4);
  for(int y = 0; y < h; y++)
  {
    if(p + 4 > f.size()) { err = "truncated HDR"; return false; }
    if(w >= 8 && w < 32768 && f[p] == 2 && f[p + 1] == 2 && !(f[p + 2] & 0x80))
    {
      p += 4;
      for(int c = 0; c < 4; c++)
        for(int x = 0; x < w;)
        {
          if(p >= f.size()) { err = "truncated HDR"; return false; }
          int n = f[p++];
          if(n > 128) { n -= 128; if(p >= f.size() || x + n > w) { err = "bad HDR RLE"; return false; } uint8_t v = f[p++]; while(n--) scan[4 * (x++) + c] = v; }
          else { if(p + n > f.size() || x + n > w) { err = "bad HDR RLE"; return false; } while(n--) scan[4 * (x++) + c] = f[p++]; }
        }
    }
    else
    {
      if(p + (size_t)w * 4 > f.size()) { err = "truncated HDR"; return false; }
      std::memcpy(scan.data(), &f[p], (size_t)w * 4);
      p += (size_t)w * 4;
    }
    for(int x = 0; x < w; x++)
    {
      const uint8_t* q = &scan[4 * x];
      float*         o = &out.rgba[((size_t)y * w + x) * 4];
      if(q[3]) { const float s = std::ldexp(1.0f, (int)q[3] - (128 + 8)); o[0] = q[0] * s; o[1] = q[1] * s; o[2] = q[2] * s; }
      else o[0] = o[1] = o[2] = 0.f;
      o[3] = 1.0f;
    }
  }
  return true;
}

}  // namespace vkrt_host
