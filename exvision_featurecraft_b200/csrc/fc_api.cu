// Library-level entry points: version, status strings, error and launch bookkeeping.
#include <atomic>

#include "fc_common.cuh"

namespace fc {
static std::atomic<int> g_last_cuda_error{0};
static std::atomic<unsigned long long> g_launches{0};
void note_cuda_error(cudaError_t e) { g_last_cuda_error.store((int)e); }
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n); }
static std::atomic<unsigned long long> g_paths[FC_PATH_COUNT];
void note_path(int which) { g_paths[which].fetch_add(1ull); }
}  // namespace fc

extern "C" int fc_version(void) { return 100; }

extern "C" const char* fc_status_string(int status) {
  switch (status) {
    case FC_OK: return "ok";
    case FC_ERR_BAD_ARG: return "bad argument";
    case FC_ERR_CAPACITY: return "output capacity too small";
    case FC_ERR_CUDA: return "CUDA error";
    case FC_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

extern "C" int fc_last_cuda_error(void) { return fc::g_last_cuda_error.load(); }
extern "C" unsigned long long fc_launch_count(void) { return fc::g_launches.load(); }
extern "C" unsigned long long fc_path_count(int path) {
  return (path >= 0 && path < FC_PATH_COUNT) ? fc::g_paths[path].load() : 0ull;
}
