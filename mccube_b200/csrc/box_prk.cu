// placeholder, replaced below
#include "common.cuh"
extern "C" size_t mccb200_box_prk_step_workspace_bytes(uint64_t, int, int, int, uint64_t) { return 0; }
extern "C" int mccb200_box_prk_step(const float*, uint64_t, int, const mccb200_drift*, float, const mccb200_cubature*,
                                    const int32_t*, int, int, const float*, const uint32_t*, uint64_t, uint64_t,
                                    float*, void*, size_t, void*) {
  return mccb::fail(MCCB200_ERR_UNSUPPORTED, "box_prk_step: not built yet");
}
