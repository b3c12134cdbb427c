// int32 exclusive scan (n+1 outputs) used by every compaction stage of the library.
// Two-level reduce-then-scan: chunk sums -> recursive scan of the sums -> per-chunk scan.
#include "fc_common.cuh"

namespace fc {

static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 8;
static constexpr int kScanChunk = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = warp_inclusive_scan(v);
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    int si = warp_inclusive_scan(s);
    if (lane < kScanThreads / 32) warp_sums[lane] = si;
  }
  __syncthreads();
  int base = warp > 0 ? warp_sums[warp - 1] : 0;
  *total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_chunk_sums_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                        int32_t* __restrict__ sums) {
  const int64_t start = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (start + i < n) s += in[start + i];
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_chunk_kernel(const int32_t* __restrict__ in, int64_t n,
                                                                   int32_t* __restrict__ out,
                                                                   const int32_t* __restrict__ chunk_base) {
  const int64_t start = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (start + i < n) ? in[start + i] : 0;
    s += v[i];
  }
  int total;
  int run = block_exclusive_scan(s, &total) + (chunk_base ? chunk_base[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (start + i < n) out[start + i] = run;
    run += v[i];
    if (start + i == n - 1) out[n] = run;
  }
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) out[0] = 0;
}

int64_t scan_scratch_elems(int64_t n) {
  if (n <= kScanChunk) return 0;
  int64_t nc = div_up64(n, kScanChunk);
  return nc + (nc + 1) + scan_scratch_elems(nc);
}

int exclusive_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* scratch, cudaStream_t stream) {
  if (n < 0) return FC_ERR_BAD_ARG;
  if (n <= kScanChunk) {
    scan_chunk_kernel<<<1, kScanThreads, 0, stream>>>(in, n, out, nullptr);
    note_launch();
    FC_RETURN_IF_LAUNCH_FAILED();
    return FC_OK;
  }
  const int64_t nc = div_up64(n, kScanChunk);
  int32_t* sums = scratch;
  int32_t* bases = scratch + nc;
  scan_chunk_sums_kernel<<<(unsigned)nc, kScanThreads, 0, stream>>>(in, n, sums);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  int rc = exclusive_scan_i32(sums, nc, bases, scratch + nc + (nc + 1), stream);
  if (rc != FC_OK) return rc;
  scan_chunk_kernel<<<(unsigned)nc, kScanThreads, 0, stream>>>(in, n, out, bases);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

}  // namespace fc


// Publish a few int32 values to PINNED host memory straight from a kernel (the pointer of cudaHostAlloc'ed memory
// is valid on the device under UVA).  Unlike a 4-byte cudaMemcpy D2H this does not queue behind bulk transfers on
// the device-to-host copy engine, so the host can learn an output size while a large result copy is in flight.
namespace fc {
__global__ void publish_i32_kernel(const int32_t* __restrict__ src, volatile int32_t* __restrict__ dst_host, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst_host[i] = src[i];
  __threadfence_system();
}
}  // namespace fc

namespace fc {
struct GatherIdx {
  long long idx[32];
};
__global__ void publish_gather_i32_kernel(const int32_t* __restrict__ src, GatherIdx g, volatile int32_t* __restrict__ dst_host, int n) {
  const int i = threadIdx.x;
  if (i < n) dst_host[i] = src[g.idx[i]];
  __threadfence_system();
}
__global__ void popcount_u64_kernel(const unsigned long long* __restrict__ m, long long n, int32_t* __restrict__ counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = __popcll(m[i]);
}
}  // namespace fc

// dst_host_pinned[i] = src[indices_host[i]], i < n <= 32, with ONE launch: the output sizes of all (octave, scale) blocks
// of a batch are the values of one concatenated scan at the block boundaries.
extern "C" int fc_publish_gather_i32(const int32_t* src, const int64_t* indices_host, int n, int32_t* dst_host_pinned, void* stream) {
  if (!src || !indices_host || !dst_host_pinned || n <= 0 || n > 32) return FC_ERR_BAD_ARG;
  fc::GatherIdx g;
  for (int i = 0; i < 32; ++i) g.idx[i] = i < n ? (long long)indices_host[i] : 0;
  fc::publish_gather_i32_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(src, g, dst_host_pinned, n);
  fc::note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

// out[i] = sum(in[0..i)), i in [0, n]: the scan behind fc_mask_scan / fc_orientation_scan on its own, so that the counts of
// several masks (or bin-mask arrays) can be concatenated and scanned with one call.
extern "C" int fc_exclusive_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* scratch, void* stream) {
  if (!in || !out || n < 0 || (n > 2048 && !scratch)) return FC_ERR_BAD_ARG;
  return fc::exclusive_scan_i32(in, n, out, scratch, (cudaStream_t)stream);
}

extern "C" int fc_popcount_u64(const uint64_t* masks, int64_t n, int32_t* counts, void* stream) {
  if (!masks || !counts || n <= 0) return FC_ERR_BAD_ARG;
  fc::popcount_u64_kernel<<<(unsigned)fc::div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)masks, n, counts);
  fc::note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_publish_i32(const int32_t* src, int32_t* dst_host_pinned, int n, void* stream) {
  if (!src || !dst_host_pinned || n <= 0) return FC_ERR_BAD_ARG;
  fc::publish_i32_kernel<<<fc::div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(src, dst_host_pinned, n);
  fc::note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
