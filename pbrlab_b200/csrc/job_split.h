// How one frame's samples are dealt out (reference src/render.cc:211-222 with the sample as the job
// unit): the samples  offset + j*stride < spp  of a job go round-robin first over the R ranks of a multi-process job,
// then over the D devices of a rank's context.  Plain host arithmetic, shared by the CUDA library (pbrgpu.cu:
// RenderImpl) and by the CPU tests (tests/host_emul, tests/test_multirank.py).
#pragma once
#include <cstdint>

namespace pbrjob {

struct Share {
  uint32_t offset, stride;   // this worker renders samples offset + k*stride < spp
  bool ok;                   // false: the stride does not fit 32 bits
};

inline Share ShareOf(uint32_t job_offset, uint32_t job_stride, uint32_t rank, uint32_t world, uint32_t device,
                     uint32_t num_devices) {
  Share s;
  const uint64_t stride = uint64_t(job_stride) * world * num_devices;
  const uint64_t offset = uint64_t(job_offset) + (uint64_t(rank) + uint64_t(device) * world) * job_stride;
  s.ok = stride <= 0xffffffffull && offset <= 0xffffffffull && job_stride != 0 && world != 0 && num_devices != 0;
  s.offset = uint32_t(offset);
  s.stride = uint32_t(stride);
  return s;
}

// samples of [0, spp) that fall to a worker
inline uint32_t CountOf(const Share& s, uint32_t spp) {
  return (spp > s.offset) ? (spp - s.offset + s.stride - 1) / s.stride : 0u;
}

// ---- the order in which one worker starts its camera samples (wavefront.cuh: FrameParams)
// Sample ids 0 .. npix*(probe_passes + rest_passes) map one to one onto (pixel, local sample index): the first
// probe_passes samples of every pixel in raster order, sample-major; then the remaining rest_passes samples pixel
// block by pixel block (order_block pixels, the last block may be shorter), sample-major inside a block, through the
// permutation `order` of the pixels.  order == nullptr or rest_passes == 0: raster order throughout.
#if defined(__CUDACC__)
#define PBRJOB_HD __host__ __device__ __forceinline__
#else
#define PBRJOB_HD inline
#endif
struct SampleOrder {
  const uint32_t* order;
  uint32_t npix, probe_passes, rest_passes, order_block;
};
PBRJOB_HD void SampleOfId(const SampleOrder& f, unsigned long long id, uint32_t* pixel, uint32_t* s_local) {
  const unsigned long long probe = (unsigned long long)f.probe_passes * f.npix;
  if (f.order == nullptr || f.rest_passes == 0u || id < probe) {
    const uint32_t s = uint32_t(id / f.npix);
    *s_local = s;
    *pixel = uint32_t(id - (unsigned long long)s * f.npix);
    return;
  }
  const unsigned long long r = id - probe;
  const unsigned long long per_block = (unsigned long long)f.order_block * f.rest_passes;
  const uint32_t block = uint32_t(r / per_block);
  const uint32_t within = uint32_t(r - (unsigned long long)block * per_block);
  const uint32_t first = block * f.order_block;
  const uint32_t left = f.npix - first;
  const uint32_t width = f.order_block < left ? f.order_block : left;
  const uint32_t s = within / width;
  *s_local = f.probe_passes + s;
  *pixel = f.order[first + (within - s * width)];
}

}  // namespace pbrjob
