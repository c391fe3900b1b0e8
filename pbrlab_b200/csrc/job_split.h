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

}  // namespace pbrjob
