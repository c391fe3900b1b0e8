// TEST INFRASTRUCTURE ONLY.  Ray counters for the reference, without touching its sources: the
// `libpbrlab_ref_count.so` flavour is linked with -Wl,--wrap=rtcIntersect1 -Wl,--wrap=rtcOccluded1 so every call
// the reference makes at src/raytracer/raytracer_impl.cc:275 / :285 lands here first.  Counting costs ~25 % of
// the run time, so this flavour is never used for timing.
#include <atomic>
#include <cstdint>

#include <embree4/rtcore.h>

extern "C" {
void __real_rtcIntersect1(RTCScene scene, struct RTCRayHit* rayhit, struct RTCIntersectArguments* args);
void __real_rtcOccluded1(RTCScene scene, struct RTCRay* ray, struct RTCOccludedArguments* args);

static std::atomic<uint64_t> g_intersect(0), g_occluded(0);

void __wrap_rtcIntersect1(RTCScene scene, struct RTCRayHit* rayhit, struct RTCIntersectArguments* args) {
  g_intersect.fetch_add(1, std::memory_order_relaxed);
  __real_rtcIntersect1(scene, rayhit, args);
}
void __wrap_rtcOccluded1(RTCScene scene, struct RTCRay* ray, struct RTCOccludedArguments* args) {
  g_occluded.fetch_add(1, std::memory_order_relaxed);
  __real_rtcOccluded1(scene, ray, args);
}
__attribute__((visibility("default"))) void ref_ray_counts(uint64_t* out2) {
  out2[0] = g_intersect.load();
  out2[1] = g_occluded.load();
}
__attribute__((visibility("default"))) void ref_ray_counts_reset() {
  g_intersect = 0;
  g_occluded = 0;
}
}
