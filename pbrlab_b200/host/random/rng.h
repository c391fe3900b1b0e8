// Host-side PCG32 wrapper with the reference's interface (reference src/random/rng.h:40-70); the generator itself
// is the one the kernels use (csrc/device/rng.cuh).
#ifndef PBRLAB_B200_RNG_H_
#define PBRLAB_B200_RNG_H_
#include "../../csrc/device/rng.cuh"
namespace pbrlab {
class RNG {
public:
  explicit RNG(uint64_t initstate, uint64_t initseq) { pbr::pcg32_srandom(&state_, initstate, initseq); }
  inline float Draw() const { return pbr::Draw(&state_); }
private:
  mutable pbr::Pcg32 state_;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_RNG_H_
