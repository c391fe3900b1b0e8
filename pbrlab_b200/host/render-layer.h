// Accumulation buffers of a render: SUMS, the caller divides (reference src/render-layer.h:11-26,
// src/render-layer.cc:13-30).  rgba: 4 floats per pixel (r, g, b, number of samples as float); count: u32 per pixel.
#ifndef PBRLAB_B200_RENDER_LAYER_H_
#define PBRLAB_B200_RENDER_LAYER_H_
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
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
    // (Render() clears the layer at the start of every frame, render.cc:99-100: a 4K layer is 166 MB, so large
    // layers are zeroed by four threads)
    const size_t n = rgba.size(), parts = n >= (size_t(1) << 22) ? 4 : 1;
    if (parts == 1) {
      std::fill(rgba.begin(), rgba.end(), 0.0f);
      std::fill(count.begin(), count.end(), uint32_t(0));
      return;
    }
    std::vector<std::thread> th;
    for (size_t k = 0; k < parts; ++k)
      th.emplace_back([this, k, parts]() {
        const size_t a = rgba.size() * k / parts, b = rgba.size() * (k + 1) / parts;
        std::memset(rgba.data() + a, 0, (b - a) * sizeof(float));
        const size_t c = count.size() * k / parts, d = count.size() * (k + 1) / parts;
        std::memset(count.data() + c, 0, (d - c) * sizeof(uint32_t));
      });
    for (auto& t : th) t.join();
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
