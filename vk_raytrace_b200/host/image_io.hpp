// image_io.hpp -- see image_io.cpp
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vkrt_host {

struct ImageRGBA8
{
  uint32_t             w = 0, h = 0;
  std::vector<uint8_t> rgba;
};
struct ImageRGBA32F
{
  uint32_t           w = 0, h = 0;
  std::vector<float> rgba;
};

bool readFile(const std::string& path, std::vector<uint8_t>& out);
bool decodePng(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err);
// baseline / extended / progressive Huffman JPEG, 1 or 3 components (jpeg_decode.cpp)
bool decodeJpeg(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err);
// PNG or JPEG by signature
bool decodeImage(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err);
bool loadHdrOrPfm(const std::string& path, ImageRGBA32F& out, std::string& err);

}  // namespace vkrt_host
