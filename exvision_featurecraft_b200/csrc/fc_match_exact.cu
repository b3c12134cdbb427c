// Exact float32 brute-force 2-NN matcher on CUDA cores: the verification twin of the tcgen05 path and
// the re-check kernel for the rows that path flags as ambiguous.
//
// The definition it implements is the oracle's (oracle/featurecraft_oracle.py, knn_match): L2 distance
// accumulated in float32 as sum((a-b)^2) over the 128 components in order, sqrt, two smallest per query
// row, ties -> lower train index; then the ratio test m.distance < t * n.distance evaluated in double on
// the float32 distances, strict (APP/utils/keypoint_descriptor.py:338-349).  That is exact against the
// oracle; cv2.BFMatcher's own kernel sums with SIMD partial accumulators, so its distances can differ from
// these in the last float32 ulp (<= 5e-6 in the golden case) and a ratio test that close can fall either way.
#include "fc_common.cuh"

namespace fc {

static constexpr int kMQ = 64;    // queries per block
static constexpr int kMT = 64;    // train rows per tile
static constexpr int kDim = 128;
static constexpr int kPitch = kDim + 1;

struct Top2 {
  float d0, d1;
  int i0, i1;
  __device__ __forceinline__ void init() { d0 = d1 = INFINITY; i0 = i1 = -1; }
  // lexicographic (distance, index) insert; callers feed indices in increasing order per thread, the
  // cross-thread merge uses the explicit index compare
  __device__ __forceinline__ void push(float d, int i) {
    if (d < d0 || (d == d0 && i < i0)) {
      d1 = d0; i1 = i0; d0 = d; i0 = i;
    } else if (d < d1 || (d == d1 && i < i1)) {
      d1 = d; i1 = i;
    }
  }
};

// rows == nullptr: queries [q_begin, q_begin + kMQ) of the block; otherwise the block handles the query
// rows listed in rows[blockIdx.x*kMQ ...] (n_rows entries) -- used for the flagged-row re-check.
__global__ void __launch_bounds__(256)
match_exact_kernel(const float* __restrict__ query, int n_query, const float* __restrict__ train, int n_train,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows_ptr, double ratio,
                   int32_t* __restrict__ knn_idx, float* __restrict__ knn_dist, uint8_t* __restrict__ good) {
  extern __shared__ float smem[];
  float* sQ = smem;                 // kMQ x kPitch
  float* sT = smem + kMQ * kPitch;  // kMT x kPitch
  __shared__ int sRow[kMQ];
  const int n_rows = rows ? *n_rows_ptr : n_query;
  const int q_begin = blockIdx.x * kMQ;
  if (q_begin >= n_rows) return;
  for (int t = threadIdx.x; t < kMQ; t += blockDim.x) {
    const int r = q_begin + t;
    sRow[t] = r < n_rows ? (rows ? rows[r] : r) : -1;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kMQ * (kDim / 4); t += blockDim.x) {
    const int r = t / (kDim / 4), c4 = t - r * (kDim / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sRow[r] >= 0) v = reinterpret_cast<const float4*>(query + (size_t)sRow[r] * kDim)[c4];
    float* d = sQ + r * kPitch + c4 * 4;
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  const int tq = threadIdx.x >> 4, tt = threadIdx.x & 15;
  Top2 best[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) best[a].init();
  for (int t0 = 0; t0 < n_train; t0 += kMT) {
    __syncthreads();
    for (int t = threadIdx.x; t < kMT * (kDim / 4); t += blockDim.x) {
      const int r = t / (kDim / 4), c4 = t - r * (kDim / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t0 + r < n_train) v = __ldg(reinterpret_cast<const float4*>(train + (size_t)(t0 + r) * kDim) + c4);
      float* d = sT + r * kPitch + c4 * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
    const float* q = sQ + (tq * 4) * kPitch;
    const float* tr = sT + tt * kPitch;
#pragma unroll 4
    for (int k = 0; k < kDim; ++k) {
      float qa[4], tb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) qa[a] = q[a * kPitch + k];
#pragma unroll
      for (int b = 0; b < 4; ++b) tb[b] = tr[(16 * b) * kPitch + k];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float df = __fsub_rn(qa[a], tb[b]);
          acc[a][b] = __fadd_rn(acc[a][b], __fmul_rn(df, df));  // no FMA: the float32 reference rounds the square
        }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = t0 + tt + 16 * b;
      if (j < n_train) {
#pragma unroll
        for (int a = 0; a < 4; ++a) best[a].push(sqrtf(acc[a][b]), j);
      }
    }
  }
  // merge the 16 partial lists of every query through shared memory (reuse sT)
  __syncthreads();
  float* mD = sT;                                   // [kMQ][16][2]
  int* mI = reinterpret_cast<int*>(sT + kMQ * 32);  // [kMQ][16][2]
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int o = ((tq * 4 + a) * 16 + tt) * 2;
    mD[o] = best[a].d0; mD[o + 1] = best[a].d1;
    mI[o] = best[a].i0; mI[o + 1] = best[a].i1;
  }
  __syncthreads();
  if (threadIdx.x < kMQ) {
    const int row = sRow[threadIdx.x];
    if (row >= 0) {
      Top2 m;
      m.init();
      for (int c = 0; c < 32; ++c) {
        const int i = mI[threadIdx.x * 32 + c];
        if (i >= 0) m.push(mD[threadIdx.x * 32 + c], i);
      }
      knn_idx[2 * (size_t)row] = m.i0;
      knn_idx[2 * (size_t)row + 1] = m.i1;
      knn_dist[2 * (size_t)row] = m.d0;
      knn_dist[2 * (size_t)row + 1] = m.d1;
      good[row] = (m.i1 >= 0 && (double)m.d0 < __dmul_rn(ratio, (double)m.d1)) ? 1 : 0;
    }
  }
}

int launch_match_exact(const float* query, int n_query, const float* train, int n_train, const int32_t* rows,
                       const int32_t* n_rows_ptr, int max_rows, double ratio, int32_t* knn_idx, float* knn_dist,
                       uint8_t* good, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)(kMQ + kMT) * kPitch;
  FC_CUDA_TRY(cudaFuncSetAttribute(match_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = div_up(max_rows, kMQ);
  if (blocks == 0) return FC_OK;
  match_exact_kernel<<<blocks, 256, smem, st>>>(query, n_query, train, n_train, rows, n_rows_ptr, ratio, knn_idx, knn_dist, good);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

}  // namespace fc

extern "C" int fc_match_knn2_exact(const float* query, int n_query, const float* train, int n_train, double ratio,
                                   int32_t* knn_idx, float* knn_dist, uint8_t* good, void* stream) {
  if (!query || !train || !knn_idx || !knn_dist || !good || n_query <= 0 || n_train <= 0) return FC_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(train)) & 15) return FC_ERR_BAD_ARG;
  return fc::launch_match_exact(query, n_query, train, n_train, nullptr, nullptr, n_query, ratio, knn_idx, knn_dist, good,
                                (cudaStream_t)stream);
}
