// Accumulation buffers of a render: SUMS, the caller divides (reference src/render-layer.h:11-26,
// src/render-layer.cc:13-30).  rgba: 4 floats per pixel (r, g, b, number of samples as float); count: u32 per pixel.
#ifndef PBRLAB_B200_RENDER_LAYER_H_
#define PBRLAB_B200_RENDER_LAYER_H_
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <mutex>
#include <vector>
namespace pbrlab {
struct RenderLayer {
  RenderLayer() : width(0), height(0) {}
  RenderLayer(const size_t w, const size_t h) : width(0), height(0) {
    Resize(w, h);
    Clear();
  }
  void Clear(void) {
    std::lock_guard<std::mutex> lock(mtx);
    std::fill(rgba.begin(), rgba.end(), 0.0f);
    std::fill(count.begin(), count.end(), uint32_t(0));
  }
  void Resize(const size_t w, const size_t h) {
    std::lock_guard<std::mutex> lock(mtx);
    width = w;
    height = h;
    rgba.resize(w * h * 4);
    count.resize(w * h);
  }
  size_t width;
  size_t height;
  std::vector<float> rgba;
  std::vector<uint32_t> count;
  mutable std::mutex mtx;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_RENDER_LAYER_H_
