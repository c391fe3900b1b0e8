// PCG32 (XSH-RR) exactly as src/random/rng.h:17-65: the step, the seeding sequence and the
// ((x >> 9) | 0x3f800000) - 1 float mapping.  One generator per path; its 16 bytes live in the SoA path state.
#pragma once
#include "common.cuh"

namespace pbr {

struct Pcg32 {
  uint64_t state;
  uint64_t inc;
};

PBR_HD uint32_t pcg32_random(Pcg32* rng) {                      // src/random/rng.h:17-27
  const uint64_t oldstate = rng->state;
  rng->state = oldstate * 6364136223846793005ULL + rng->inc;
  const uint32_t xorshifted = uint32_t(((oldstate >> 18u) ^ oldstate) >> 27u);
  const uint32_t rot = uint32_t(oldstate >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}

PBR_HD void pcg32_srandom(Pcg32* rng, uint64_t initstate, uint64_t initseq) {  // src/random/rng.h:29-36
  rng->state = 0U;
  rng->inc = (initseq << 1U) | 1U;
  pcg32_random(rng);
  rng->state += initstate;
  pcg32_random(rng);
}

PBR_HD float Draw(Pcg32* rng) {                                   // src/random/rng.h:52-65
  const float f = u2f((pcg32_random(rng) >> 9) | 0x3f800000u);
  return f - 1.0f;
}

}  // namespace pbr
