// Texture of the public API (reference src/texture.h, src/texture.cc:10-72).  The device samples the same pixels
// with the same filter (csrc/device/shade.cuh: TextureFetch3); the host-side fetches restate Texture::FetchFloatN
// -> BilinearFilter with clamp addressing (reference src/image-utils.cc:99-167) for callers of the C++ API.
#ifndef PBRLAB_B200_TEXTURE_H_
#define PBRLAB_B200_TEXTURE_H_
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>
namespace pbrlab {
class Texture {
public:
  Texture() : width_(0), height_(0), channels_(0) {}
  Texture(const std::vector<float>& pixels, const uint32_t width, const uint32_t height, const uint32_t channels,
          const std::string& name)
      : width_(width), height_(height), channels_(channels), pixels_(pixels), name_(name) {}
  bool Reset(const std::vector<float>& pixels, const uint32_t width, const uint32_t height, const uint32_t channels) {
    if (pixels.size() != size_t(width) * height * channels) return false;   // texture.cc:21-34
    pixels_ = pixels; width_ = width; height_ = height; channels_ = channels;
    return true;
  }
  void SetName(const std::string& name) { name_ = name; }
  uint32_t GetWidth(void) const { return width_; }
  uint32_t GetHeight(void) const { return height_; }
  uint32_t GetChannels(void) const { return channels_; }
  std::string GetName(void) const { return name_; }
  const std::vector<float>& GetPixels(void) const { return pixels_; }

  void FetchFloatN(const float u, const float v, const uint32_t n, float* dst) const {
    const float uu = std::min(std::max(u, 0.0f), 1.0f), vv = std::min(std::max(v, 0.0f), 1.0f);
    const float px = float(width_) * uu, py = float(height_) * vv;
    const int w = int(width_), h = int(height_);
    const int x0 = std::max(0, std::min(w - 1, int(px))), y0 = std::max(0, std::min(h - 1, int(py)));
    const int x1 = (x0 + 1 >= w) ? w - 1 : x0 + 1, y1 = (y0 + 1 >= h) ? h - 1 : y0 + 1;
    const float dx = px - float(x0), dy = py - float(y0);
    const float w0 = (1.0f - dx) * (1.0f - dy), w1 = (1.0f - dx) * dy, w2 = dx * (1.0f - dy), w3 = dx * dy;
    const int st = int(channels_);
    const int i00 = st * (y0 * w + x0), i01 = st * (y0 * w + x1), i10 = st * (y1 * w + x0), i11 = st * (y1 * w + x1);
    for (uint32_t c = 0; c < n; ++c) {
      if (c < channels_)
        dst[c] = pixels_[size_t(i00) + c] * w0 + pixels_[size_t(i10) + c] * w1 + pixels_[size_t(i01) + c] * w2 +
                 pixels_[size_t(i11) + c] * w3;
      else
        dst[c] = 0.f;
    }
  }
  void FetchFloat(const float u, const float v, float* dst) const { FetchFloatN(u, v, 1, dst); }
  void FetchFloat2(const float u, const float v, float* dst) const { FetchFloatN(u, v, 2, dst); }
  void FetchFloat3(const float u, const float v, float* dst) const { FetchFloatN(u, v, 3, dst); }
  void FetchFloat4(const float u, const float v, float* dst) const { FetchFloatN(u, v, 4, dst); }

private:
  uint32_t width_, height_, channels_;
  std::vector<float> pixels_;
  std::string name_;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_TEXTURE_H_
