// jpeg_decode.cpp -- JPEG (ITU-T T.81) decoder for glTF textures: baseline and extended sequential (SOF0 / SOF1, 8 bit),
// progressive (SOF2), Huffman coding, restart intervals, 1 or 3 components with arbitrary 1..4 sampling factors.
//
// Why it exists: the reference takes whatever tinygltf / stb_image decoded (src/scene.cpp:522-541), and the real
// DamagedHelmet.glb / Khronos Sponza.gltf carry JPEG textures; without a decoder they would render with the 1x1 white fallback
// (src/scene.cpp:496-502).  There is no stb_image on this image to compare with, so the arithmetic follows the IJG reference
// decoder that Pillow (libjpeg-turbo) also implements, and tests/test_host_ingest.py asserts equality with Pillow:
//   * inverse DCT: the Loeffler-Ligtenberg-Moschytz factorisation in 13-bit fixed point with a 2-bit first-pass scale ("islow");
//   * chroma upsampling: the triangle ("fancy") filters for 2:1 horizontal and 2:1 x 2:1, box replication for other ratios;
//   * YCbCr -> RGB in 16-bit fixed point (1.402, 0.344136 / 0.714136, 1.772), JFIF / Adobe APP14 colour-transform rules.
// CMYK / YCCK (4 components), arithmetic coding, lossless and 12-bit files are rejected (the caller falls back to white).
#include <algorithm>
#include <cstring>
#include "image_io.hpp"

namespace vkrt_host {
namespace {

constexpr uint32_t kMaxDim = 1u << 15;

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff
{
  bool     set = false;
  uint8_t  vals[256];
  int32_t  maxcode[18];  // largest code of each length (-1: none), [17] = sentinel
  int32_t  valptr[17], mincode[17];
  uint16_t look[512];    // 9-bit lookahead: (length << 8) | symbol, 0 = longer code

  bool build(const uint8_t counts[16], const uint8_t* symbols, int n)
  {
    std::memcpy(vals, symbols, (size_t)n);
    int code = 0, k = 0;
    std::memset(look, 0, sizeof(look));
    for(int len = 1; len <= 16; len++)
    {
      valptr[len]  = k;
      mincode[len] = code;
      for(int i = 0; i < counts[len - 1]; i++, k++, code++)
      {
        if(code >= (1 << len))
          return false;  // over-subscribed table
        if(len <= 9)
        {
          const int first = code << (9 - len), span = 1 << (9 - len);
          for(int j = 0; j < span; j++) look[first + j] = (uint16_t)((len << 8) | vals[k]);
        }
      }
      maxcode[len] = counts[len - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    set         = true;
    return k == n;
  }
};

struct Component
{
  int id = 0, h = 1, v = 1, tq = 0;
  int td = 0, ta = 0;                // tables of the current scan
  uint32_t w = 0, hgt = 0;           // size in samples: ceil(image * factor / max factor)
  uint32_t bw = 0, bh = 0;           // allocated blocks (whole MCUs)
  std::vector<int16_t> coef;         // bw * bh * 64, natural (de-zigzagged) order
  std::vector<uint8_t> pix;          // (bw * 8) x (bh * 8) after the inverse DCT
  int dcPred = 0;
};

struct BitReader
{
  const uint8_t* d;
  size_t         n, pos;
  uint32_t       buf = 0;
  int            cnt = 0;
  bool           hitMarker = false;
  void fill()
  {
    while(cnt <= 24)
    {
      uint32_t byte = 0;
      if(!hitMarker && pos < n)
      {
        byte = d[pos];
        if(byte == 0xFF)
        {
          const uint8_t next = pos + 1 < n ? d[pos + 1] : 0xD9;
          if(next == 0x00)
            pos += 2;  // stuffed zero
          else
          {
            hitMarker = true;  // leave the marker in place; feed zeros from here on
            byte      = 0;
          }
        }
        else
          pos++;
      }
      buf |= byte << (24 - cnt);
      cnt += 8;
    }
  }
  int bits(int k)
  {
    if(k == 0)
      return 0;
    if(cnt < k)
      fill();
    const int v = (int)(buf >> (32 - k));
    buf <<= k;
    cnt -= k;
    return v;
  }
  int bit() { return bits(1); }
  int peek9()
  {
    if(cnt < 9)
      fill();
    return (int)(buf >> 23);
  }
  void skip(int k) { buf <<= k; cnt -= k; }
  void reset() { buf = 0; cnt = 0; hitMarker = false; }
};

int decodeSymbol(BitReader& br, const Huff& h)
{
  const int look = h.look[br.peek9()];
  if(look)
  {
    br.skip(look >> 8);
    return look & 0xff;
  }
  int code = br.bits(9), len = 9;
  while(code > h.maxcode[len])
  {
    if(++len > 16)
      return -1;
    code = (code << 1) | br.bit();
  }
  return h.vals[(h.valptr[len] + code - h.mincode[len]) & 0xff];
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// ---- inverse DCT (LL&M, 13-bit constants, first pass scaled by 4) ---------------------------------------------------------
constexpr int kCB = 13, kP1 = 2;
constexpr int F0298 = 2446, F0390 = 3196, F0541 = 4433, F0765 = 6270, F0899 = 7373, F1175 = 9633, F1501 = 12299, F1847 = 15137, F1961 = 16069,
              F2053 = 16819, F2562 = 20995, F3072 = 25172;
// The butterflies run in 64-bit integers: for a valid stream every intermediate fits 32 bits (the IJG code computes in INT32 and so
// gives the same values), for a corrupt one -- coefficients far outside the DCT's range -- nothing overflows and the result is merely garbage.
using wide = int64_t;
inline wide descale(wide x, int n) { return (x + ((wide)1 << (n - 1))) >> n; }
inline uint8_t clamp8(wide v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

void idct8x8(const int16_t* in, const uint16_t* q, uint8_t* out, size_t stride)
{
  wide ws[64];
  auto pass = [](wide s0, wide s1, wide s2, wide s3, wide s4, wide s5, wide s6, wide s7, wide (&even)[4], wide (&odd)[4]) {
    wide z1 = (s2 + s6) * F0541;
    const wide t2 = z1 - s6 * F1847, t3 = z1 + s2 * F0765;
    const wide t0 = (s0 + s4) * (1 << kCB), t1 = (s0 - s4) * (1 << kCB);
    even[0] = t0 + t3; even[3] = t0 - t3; even[1] = t1 + t2; even[2] = t1 - t2;
    wide o0 = s7, o1 = s5, o2 = s3, o3 = s1;
    z1 = o0 + o3;
    wide z2 = o1 + o2, z3 = o0 + o2, z4 = o1 + o3;
    const wide z5 = (z3 + z4) * F1175;
    o0 *= F0298; o1 *= F2053; o2 *= F3072; o3 *= F1501;
    z1 *= -F0899; z2 *= -F2562; z3 = z3 * -F1961 + z5; z4 = z4 * -F0390 + z5;
    odd[0] = o0 + z1 + z3; odd[1] = o1 + z2 + z4; odd[2] = o2 + z2 + z3; odd[3] = o3 + z1 + z4;
  };
  for(int c = 0; c < 8; c++)
  {
    wide e[4], o[4];
    pass((wide)in[c] * q[c], (wide)in[8 + c] * q[8 + c], (wide)in[16 + c] * q[16 + c], (wide)in[24 + c] * q[24 + c], (wide)in[32 + c] * q[32 + c],
         (wide)in[40 + c] * q[40 + c], (wide)in[48 + c] * q[48 + c], (wide)in[56 + c] * q[56 + c], e, o);
    const int sh = kCB - kP1;
    ws[c]      = descale(e[0] + o[3], sh); ws[56 + c] = descale(e[0] - o[3], sh);
    ws[8 + c]  = descale(e[1] + o[2], sh); ws[48 + c] = descale(e[1] - o[2], sh);
    ws[16 + c] = descale(e[2] + o[1], sh); ws[40 + c] = descale(e[2] - o[1], sh);
    ws[24 + c] = descale(e[3] + o[0], sh); ws[32 + c] = descale(e[3] - o[0], sh);
  }
  for(int r = 0; r < 8; r++)
  {
    const wide* w = ws + 8 * r;
    wide        e[4], o[4];
    pass(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], e, o);
    const int sh = kCB + kP1 + 3;
    uint8_t*  p  = out + (size_t)r * stride;
    p[0] = clamp8(descale(e[0] + o[3], sh) + 128); p[7] = clamp8(descale(e[0] - o[3], sh) + 128);
    p[1] = clamp8(descale(e[1] + o[2], sh) + 128); p[6] = clamp8(descale(e[1] - o[2], sh) + 128);
    p[2] = clamp8(descale(e[2] + o[1], sh) + 128); p[5] = clamp8(descale(e[2] - o[1], sh) + 128);
    p[3] = clamp8(descale(e[3] + o[0], sh) + 128); p[4] = clamp8(descale(e[3] - o[0], sh) + 128);
  }
}

struct Decoder
{
  const uint8_t* d;
  size_t         n;
  std::string&   err;
  uint32_t       W = 0, H = 0;
  bool           progressive = false, haveFrame = false;
  int            hmax = 1, vmax = 1, restartInterval = 0;
  int            adobeTransform = -1;
  bool           jfif = false;
  uint16_t       qt[4][64];
  bool           qtSet[4] = {false, false, false, false};
  Huff           dc[4], ac[4];
  std::vector<Component> comps;

  bool fail(const char* m) { err = m; return false; }
  static uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

  bool parseFrame(const uint8_t* p, size_t len)
  {
    if(len < 6 || p[0] != 8) return fail("JPEG: only 8-bit precision is supported");
    H = be16(p + 1); W = be16(p + 3);
    const int nc = p[5];
    if(W == 0 || H == 0 || W > kMaxDim || H > kMaxDim) return fail("JPEG: bad dimensions");
    if(nc != 1 && nc != 3) return fail("JPEG: only 1- and 3-component images are supported");
    if(len < 6 + 3 * (size_t)nc) return fail("JPEG: truncated frame header");
    comps.assign((size_t)nc, Component{});
    hmax = vmax = 1;
    for(int i = 0; i < nc; i++)
    {
      Component& c = comps[(size_t)i];
      c.id = p[6 + 3 * i]; c.h = p[7 + 3 * i] >> 4; c.v = p[7 + 3 * i] & 15; c.tq = p[8 + 3 * i];
      if(c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("JPEG: bad sampling factors");
      hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v);
    }
    const uint32_t mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
    for(auto& c : comps)
    {
      c.w   = (W * c.h + hmax - 1) / hmax;
      c.hgt = (H * c.v + vmax - 1) / vmax;
      c.bw  = mcux * c.h;
      c.bh  = mcuy * c.v;
      c.coef.assign((size_t)c.bw * c.bh * 64, 0);
    }
    haveFrame = true;
    return true;
  }

  bool parseDqt(const uint8_t* p, size_t len)
  {
    while(len > 0)
    {
      const int pq = p[0] >> 4, tq = p[0] & 15;
      if(tq > 3 || pq > 1) return fail("JPEG: bad quantisation table");
      const size_t need = 1 + 64 * (size_t)(pq + 1);
      if(len < need) return fail("JPEG: truncated DQT");
      for(int i = 0; i < 64; i++)
        qt[tq][kZigzag[i]] = (uint16_t)(pq ? be16(p + 1 + 2 * i) : p[1 + i]);
      qtSet[tq] = true;
      p += need; len -= need;
    }
    return true;
  }

  bool parseDht(const uint8_t* p, size_t len)
  {
    while(len > 0)
    {
      if(len < 17) return fail("JPEG: truncated DHT");
      const int tc = p[0] >> 4, th = p[0] & 15;
      if(tc > 1 || th > 3) return fail("JPEG: bad Huffman table id");
      int total = 0;
      for(int i = 0; i < 16; i++) total += p[1 + i];
      if(total > 256 || len < 17 + (size_t)total) return fail("JPEG: bad Huffman table");
      if(!(tc ? ac[th] : dc[th]).build(p + 1, p + 17, total)) return fail("JPEG: inconsistent Huffman table");
      p += 17 + total; len -= 17 + (size_t)total;
    }
    return true;
  }

  // one 8x8 block of a sequential scan
  bool blockSequential(BitReader& br, Component& c, int16_t* b)
  {
    const Huff &hd = dc[c.td], &ha = ac[c.ta];
    int s = decodeSymbol(br, hd);
    if(s < 0 || s > 15) return false;
    c.dcPred += s ? extend(br.bits(s), s) : 0;
    b[0] = (int16_t)c.dcPred;
    for(int k = 1; k < 64;)
    {
      const int rs = decodeSymbol(br, ha);
      if(rs < 0) return false;
      const int r = rs >> 4;
      s           = rs & 15;
      if(s == 0)
      {
        if(r != 15) break;  // end of block
        k += 16;
        continue;
      }
      k += r;
      if(k > 63) return false;
      b[kZigzag[k++]] = (int16_t)extend(br.bits(s), s);
    }
    return true;
  }

  // progressive scans (T.81 Annex G)
  bool blockDcFirst(BitReader& br, Component& c, int16_t* b, int al)
  {
    const int s = decodeSymbol(br, dc[c.td]);
    if(s < 0 || s > 15) return false;
    c.dcPred += s ? extend(br.bits(s), s) : 0;
    b[0] = (int16_t)(c.dcPred * (1 << al));
    return true;
  }
  static void blockDcRefine(BitReader& br, int16_t* b, int al)
  {
    if(br.bit()) b[0] = (int16_t)(b[0] | (1 << al));
  }
  bool blockAcFirst(BitReader& br, const Huff& ha, int16_t* b, int ss, int se, int al, int& eobrun)
  {
    if(eobrun > 0) { eobrun--; return true; }
    for(int k = ss; k <= se;)
    {
      const int rs = decodeSymbol(br, ha);
      if(rs < 0) return false;
      const int r = rs >> 4, s = rs & 15;
      if(s)
      {
        k += r;
        if(k > 63) return false;
        b[kZigzag[k++]] = (int16_t)(extend(br.bits(s), s) * (1 << al));
      }
      else if(r == 15)
        k += 16;
      else
      {
        eobrun = (1 << r) - 1;
        if(r) eobrun += br.bits(r);
        break;
      }
    }
    return true;
  }
  bool blockAcRefine(BitReader& br, const Huff& ha, int16_t* b, int ss, int se, int al, int& eobrun)
  {
    const int p1 = 1 << al, m1 = -(1 << al);
    int       k  = ss;
    auto correct = [&](int16_t& v) {
      if(br.bit() && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1));
    };
    if(eobrun == 0)
    {
      for(; k <= se; k++)
      {
        const int rs = decodeSymbol(br, ha);
        if(rs < 0) return false;
        int r = rs >> 4, s = rs & 15, value = 0;
        if(s)
        {
          if(s != 1) return false;
          value = br.bit() ? p1 : m1;
        }
        else if(r != 15)
        {
          eobrun = 1 << r;
          if(r) eobrun += br.bits(r);
          break;
        }
        // skip r zero-history coefficients, correcting the non-zero ones passed on the way
        while(k <= se)
        {
          int16_t& v = b[kZigzag[k]];
          if(v != 0)
            correct(v);
          else if(--r < 0)
            break;
          k++;
        }
        if(value)
        {
          if(k > 63) return false;
          b[kZigzag[k]] = (int16_t)value;
        }
      }
    }
    if(eobrun > 0)
    {
      for(; k <= se; k++)
      {
        int16_t& v = b[kZigzag[k]];
        if(v != 0) correct(v);
      }
      eobrun--;
    }
    return true;
  }

  bool decodeScan(const uint8_t* hdr, size_t len, size_t& pos)
  {
    if(!haveFrame) return fail("JPEG: scan before frame header");
    if(len < 1) return fail("JPEG: truncated SOS");
    const int ns = hdr[0];
    if(ns < 1 || ns > (int)comps.size() || len < 1 + 2 * (size_t)ns + 3) return fail("JPEG: bad scan header");
    std::vector<Component*> sc;
    for(int i = 0; i < ns; i++)
    {
      Component* c = nullptr;
      for(auto& k : comps) if(k.id == hdr[1 + 2 * i]) c = &k;
      if(!c) return fail("JPEG: scan names an unknown component");
      c->td = hdr[2 + 2 * i] >> 4; c->ta = hdr[2 + 2 * i] & 15;
      if(c->td > 3 || c->ta > 3) return fail("JPEG: bad table selector");
      sc.push_back(c);
    }
    const int ss = hdr[1 + 2 * ns], se = hdr[2 + 2 * ns], ah = hdr[3 + 2 * ns] >> 4, al = hdr[3 + 2 * ns] & 15;
    if(progressive)
    {
      if(ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1) || al > 13) return fail("JPEG: bad progressive scan parameters");
    }
    else if(ss != 0 || se != 63 || ah != 0 || al != 0)
      return fail("JPEG: bad sequential scan parameters");
    for(Component* c : sc)
    {
      if((ss == 0 && ah == 0 && !dc[c->td].set) || (se > 0 && !ac[c->ta].set))
        return fail("JPEG: scan uses an undefined Huffman table");
      c->dcPred = 0;
    }
    BitReader br{d, n, pos};
    int       eobrun = 0, toRestart = restartInterval, nextRst = 0;
    auto restart = [&]() -> bool {
      // byte-align, expect RSTn, reset predictions
      br.cnt = 0; br.buf = 0;
      size_t p = br.pos;
      while(p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7))
      {
        if(d[p] == 0xFF && d[p + 1] != 0 && d[p + 1] != 0xFF) return false;  // some other marker: corrupt
        p++;
      }
      if(p + 1 >= n || d[p + 1] != 0xD0 + nextRst) return false;
      nextRst = (nextRst + 1) & 7;
      br.pos  = p + 2;
      br.reset();
      for(Component* c : sc) c->dcPred = 0;
      eobrun    = 0;
      toRestart = restartInterval;
      return true;
    };
    auto one = [&](Component& c, uint32_t bx, uint32_t by) -> bool {
      int16_t* b = &c.coef[((size_t)by * c.bw + bx) * 64];
      if(!progressive) return blockSequential(br, c, b);
      if(ss == 0) { if(ah == 0) return blockDcFirst(br, c, b, al); blockDcRefine(br, b, al); return true; }
      return ah == 0 ? blockAcFirst(br, ac[c.ta], b, ss, se, al, eobrun) : blockAcRefine(br, ac[c.ta], b, ss, se, al, eobrun);
    };
    if(ns == 1)
    {
      Component&     c  = *sc[0];
      const uint32_t nx = (c.w + 7) / 8, ny = (c.hgt + 7) / 8;  // non-interleaved: only the blocks that hold image samples
      for(uint32_t by = 0; by < ny; by++)
        for(uint32_t bx = 0; bx < nx; bx++)
        {
          if(restartInterval && toRestart == 0 && !restart()) return fail("JPEG: missing restart marker");
          if(!one(c, bx, by)) return fail("JPEG: corrupt entropy-coded data");
          toRestart--;
        }
    }
    else
    {
      const uint32_t mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
      for(uint32_t my = 0; my < mcuy; my++)
        for(uint32_t mx = 0; mx < mcux; mx++)
        {
          if(restartInterval && toRestart == 0 && !restart()) return fail("JPEG: missing restart marker");
          for(Component* c : sc)
            for(int v = 0; v < c->v; v++)
              for(int h = 0; h < c->h; h++)
                if(!one(*c, mx * c->h + h, my * c->v + v)) return fail("JPEG: corrupt entropy-coded data");
          toRestart--;
        }
    }
    // continue parsing at the marker that ended the scan
    size_t p = br.pos;
    while(p + 1 < n && !(d[p] == 0xFF && d[p + 1] != 0x00 && d[p + 1] != 0xFF && !(d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7))) p++;
    pos = p;
    return true;
  }

  bool run(ImageRGBA8& out)
  {
    if(n < 4 || d[0] != 0xFF || d[1] != 0xD8) return fail("not a JPEG");
    size_t pos = 2;
    bool   sawEoi = false, sawScan = false;
    while(pos + 4 <= n && !sawEoi)
    {
      if(d[pos] != 0xFF) { pos++; continue; }
      const uint8_t m = d[pos + 1];
      if(m == 0xFF) { pos++; continue; }
      if(m == 0xD9) { sawEoi = true; break; }
      if(m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { pos += 2; continue; }
      const size_t len = be16(d + pos + 2);
      if(len < 2 || pos + 2 + len > n) return fail("JPEG: truncated segment");
      const uint8_t* p  = d + pos + 4;
      const size_t   pl = len - 2;
      pos += 2 + len;
      switch(m)
      {
        case 0xC0: case 0xC1: progressive = false; if(!parseFrame(p, pl)) return false; break;
        case 0xC2: progressive = true; if(!parseFrame(p, pl)) return false; break;
        case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
          return fail("JPEG: lossless / hierarchical / arithmetic-coded files are not supported");
        case 0xC4: if(!parseDht(p, pl)) return false; break;
        case 0xDB: if(!parseDqt(p, pl)) return false; break;
        case 0xDD: if(pl < 2) return fail("JPEG: bad DRI"); restartInterval = (int)be16(p); break;
        case 0xE0: if(pl >= 5 && !std::memcmp(p, "JFIF\0", 5)) jfif = true; break;
        case 0xEE: if(pl >= 12 && !std::memcmp(p, "Adobe", 5)) adobeTransform = p[11]; break;
        case 0xDA:
          if(!decodeScan(p, pl, pos)) return false;
          sawScan = true;
          break;
        default: break;  // APPn, COM, ...
      }
    }
    if(!haveFrame || !sawScan) return fail("JPEG: no image data");
    // coefficients -> samples
    for(auto& c : comps)
    {
      if(!qtSet[c.tq]) return fail("JPEG: missing quantisation table");
      c.pix.assign((size_t)c.bw * 8 * c.bh * 8, 0);
      const size_t stride = (size_t)c.bw * 8;
      for(uint32_t by = 0; by < c.bh; by++)
        for(uint32_t bx = 0; bx < c.bw; bx++)
          idct8x8(&c.coef[((size_t)by * c.bw + bx) * 64], qt[c.tq], &c.pix[(size_t)by * 8 * stride + (size_t)bx * 8], stride);
      c.coef.clear();
      c.coef.shrink_to_fit();
    }
    // full-resolution planes
    std::vector<std::vector<uint8_t>> full(comps.size());
    for(size_t i = 0; i < comps.size(); i++)
      upsample(comps[i], full[i]);
    out.w = W; out.h = H;
    out.rgba.assign((size_t)W * H * 4, 255);
    if(comps.size() == 1)
    {
      for(size_t i = 0; i < (size_t)W * H; i++)
        out.rgba[4 * i] = out.rgba[4 * i + 1] = out.rgba[4 * i + 2] = full[0][i];
      return true;
    }
    // colour space: Adobe transform flag wins; JFIF means YCbCr; otherwise ids 'R','G','B' mean RGB, anything else YCbCr
    bool ycc = true;
    if(adobeTransform == 0) ycc = false;
    else if(adobeTransform < 0 && !jfif && comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B') ycc = false;
    auto fix = [](double x) { return (int)(x * 65536.0 + 0.5); };
    const int kCrR = fix(1.40200), kCbB = fix(1.77200), kCrG = fix(0.71414), kCbG = fix(0.34414), half = 1 << 15;
    for(size_t i = 0; i < (size_t)W * H; i++)
    {
      const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
      uint8_t*  o = &out.rgba[4 * i];
      if(!ycc) { o[0] = full[0][i]; o[1] = full[1][i]; o[2] = full[2][i]; continue; }
      o[0] = clamp8(y + ((kCrR * cr + half) >> 16));
      o[1] = clamp8(y + ((-kCbG * cb - kCrG * cr + half) >> 16));
      o[2] = clamp8(y + ((kCbB * cb + half) >> 16));
    }
    return true;
  }

  // component plane (c.w x c.hgt samples inside a padded buffer) -> W x H
  void upsample(const Component& c, std::vector<uint8_t>& dst) const
  {
    const size_t stride = (size_t)c.bw * 8;
    dst.assign((size_t)W * H, 0);
    const int fx = hmax / c.h, fy = vmax / c.v;
    const bool exactX = c.h * fx == hmax, exactY = c.v * fy == vmax;
    auto in = [&](uint32_t x, uint32_t y) -> int { return c.pix[(size_t)std::min(y, c.hgt - 1) * stride + std::min(x, c.w - 1)]; };
    if(exactX && exactY && fx == 1 && fy == 1)
    {
      for(uint32_t y = 0; y < H; y++) std::memcpy(&dst[(size_t)y * W], &c.pix[(size_t)y * stride], W);
      return;
    }
    // (the triangle filters need more than two input columns; narrower components are replicated, like the IJG decoder does)
    if(exactX && exactY && fx == 2 && fy == 1 && c.w > 2)
    {
      // 2:1 horizontal triangle filter: 3/4 nearer + 1/4 further sample, rounding offsets 1 and 2 alternate
      for(uint32_t y = 0; y < H; y++)
        for(uint32_t x = 0; x < W; x++)
        {
          const uint32_t i = x >> 1;
          int v;
          if(c.w == 1) v = in(0, y);
          else if(x & 1) v = (i + 1 < c.w) ? (3 * in(i, y) + in(i + 1, y) + 2) >> 2 : in(i, y);
          else v = (i > 0) ? (3 * in(i, y) + in(i - 1, y) + 1) >> 2 : in(i, y);
          dst[(size_t)y * W + x] = (uint8_t)v;
        }
      return;
    }
    if(exactX && exactY && fx == 2 && fy == 2 && c.w > 2)
    {
      // 2:1 x 2:1 triangle filter: vertical 3:1 column sums first (nearer row weighted 3), then the horizontal 3:1 with rounding 8 / 7
      for(uint32_t y = 0; y < H; y++)
      {
        const uint32_t r     = y >> 1;
        const uint32_t other = (y & 1) ? std::min(r + 1, c.hgt - 1) : (r > 0 ? r - 1 : 0);
        auto colsum = [&](uint32_t i) { return 3 * in(i, r) + in(i, other); };
        for(uint32_t x = 0; x < W; x++)
        {
          const uint32_t i = x >> 1;
          const int      cur = colsum(i);
          int            v;
          if(c.w == 1) v = (x & 1) ? (cur * 4 + 7) >> 4 : (cur * 4 + 8) >> 4;
          else if(x & 1) v = (i + 1 < c.w) ? (cur * 3 + colsum(i + 1) + 7) >> 4 : (cur * 4 + 7) >> 4;
          else v = (i > 0) ? (cur * 3 + colsum(i - 1) + 8) >> 4 : (cur * 4 + 8) >> 4;
          dst[(size_t)y * W + x] = (uint8_t)v;
        }
      }
      return;
    }
    // any other ratio: replicate (integral factors) or nearest sample (fractional ones)
    for(uint32_t y = 0; y < H; y++)
      for(uint32_t x = 0; x < W; x++)
        dst[(size_t)y * W + x] = (uint8_t)in((uint32_t)((uint64_t)x * c.h / hmax), (uint32_t)((uint64_t)y * c.v / vmax));
  }
};

}  // namespace

bool decodeJpeg(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err)
{
  Decoder dec{data, size, err};
  return dec.run(out);
}

}  // namespace vkrt_host
