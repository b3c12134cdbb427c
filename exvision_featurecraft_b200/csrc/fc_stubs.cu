// Entry points declared in the header whose kernels have not landed yet: they fail loudly.
#include "fc_common.cuh"

extern "C" int64_t fc_match_workspace_bytes(int, int) { return 0; }
extern "C" int fc_match_knn2(const float*, int, const float*, int, double, int32_t*, float*, uint8_t*, void*, void*) {
  return FC_ERR_UNSUPPORTED;
}
extern "C" int fc_match_knn2_exact(const float*, int, const float*, int, double, int32_t*, float*, uint8_t*, void*) {
  return FC_ERR_UNSUPPORTED;
}
extern "C" int fc_gauss_octave(const void*, int, int, int, int, int, const double*, const int*, int, int, double, double,
                               void*, uint32_t*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_downsample2(const void*, int, int, int, int, int, int, void*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_convert_f64(const double*, int64_t, int, void*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_gradient_field(const void*, int, int, int, int, int, int, void*, void*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_orientation_bins(const void*, const void*, int, int, int, int, const int32_t*, int, const double*, int,
                                   uint64_t*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_orientation_scan(const uint64_t*, int, int32_t*, int32_t*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_orientation_emit(const int32_t*, const uint64_t*, const int32_t*, int, int32_t*, void*) { return FC_ERR_UNSUPPORTED; }
extern "C" int fc_descriptors(const void*, const void*, int, int, int, int, const int32_t*, int, const double*, const double*,
                              void*, int, void*) { return FC_ERR_UNSUPPORTED; }
