// Texture container of the public API (reference src/texture.h, src/texture.cc:10-41).  Sampling on the device is
// a "next" row; the host class only stores pixels so loaders and AddTexture() keep working.
#ifndef PBRLAB_B200_TEXTURE_H_
#define PBRLAB_B200_TEXTURE_H_
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
  uint32_t GetWidth(void) const { return width_; }
  uint32_t GetHeight(void) const { return height_; }
  uint32_t GetChannels(void) const { return channels_; }
  std::string GetName(void) const { return name_; }
  const std::vector<float>& GetPixels(void) const { return pixels_; }
private:
  uint32_t width_, height_, channels_;
  std::vector<float> pixels_;
  std::string name_;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_TEXTURE_H_
