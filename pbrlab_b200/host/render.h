// pbrlab::Render — the drop-in boundary function (reference src/render.h:14-17, src/render.cc:192-241).
// Same signature and contract: blocking; resizes and clears `layer` first; accumulates SUMS (rgba incl. alpha = number
// of samples, count) that the caller divides; honours cancel_render_flag between sample batches and then returns
// early with a partially accumulated layer; *finish_pass starts at 0 and rises monotonically to num_sample; always
// returns true unless the device backend failed: like the reference it never throws; on a device error it prints the
// message, leaves a cleared layer and returns false (LastRenderError() holds the text — there is no CPU fallback).
// The body is a call into the B200 wavefront path tracer through the C ABI (include/pbrgpu.h: pbrgpu_render).
#ifndef PBRLAB_B200_RENDER_H_
#define PBRLAB_B200_RENDER_H_
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <string>

#include "api-types.h"
#include "render-layer.h"
#include "scene.h"
#include "type.h"

namespace pbrlab {

bool Render(const Scene& scene, const uint32_t width, const uint32_t height, const uint32_t num_sample,
            const std::atomic_bool& cancel_render_flag, RenderLayer* layer, std::atomic_size_t* finish_pass);

// Seed of the per-path PCG32 streams used by Render() (the reference seeds per worker thread with a constant,
// src/render.cc:215; here path (pixel p, sample s) uses pcg32_srandom(seed + s, p)).
std::string LastRenderError(void);
void SetRenderSeed(uint64_t seed);
uint64_t GetRenderSeed(void);

}  // namespace pbrlab
#endif  // PBRLAB_B200_RENDER_H_
