// Entry points declared in the header whose kernels have not landed yet: they fail loudly.
#include "fc_common.cuh"

extern "C" int64_t fc_match_workspace_bytes(int, int) { return 0; }
extern "C" int fc_match_knn2(const float*, int, const float*, int, double, int32_t*, float*, uint8_t*, void*, void*) {
  return FC_ERR_UNSUPPORTED;
}
