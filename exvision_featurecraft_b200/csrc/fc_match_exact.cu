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

// rows == nullptr: the block's work items are the query blocks [64 rb, 64 rb + 64); otherwise the query rows listed in
// rows[64 rb ...] (n_rows entries, device-side count) -- the flagged-row re-check of the tensor path.
//
// Work items = (query block rb, train split s).  With few query blocks -- the usual re-check has a handful of rows -- one
// CTA walking all train tiles is a latency chain (11 ms for one row against 65 536 train rows), so the train range is
// cut into S = min(kExactMaxSplits, kExactItems / query blocks) pieces: every piece leaves its (distance, index) top-2
// in `part`, the last piece to arrive for a query block (counter in `done`, zero on entry and on exit) merges them.
// The lexicographic (distance, index) order makes the result independent of how the range is cut.  part == nullptr: S = 1.
static constexpr int kExactItems = 148 * 4;
static constexpr int kExactMaxSplits = 64;

__global__ void __launch_bounds__(256)
match_exact_kernel(const float* __restrict__ query, int n_query, const float* __restrict__ train, int n_train,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows_ptr, double ratio,
                   int32_t* __restrict__ knn_idx, float* __restrict__ knn_dist, uint8_t* __restrict__ good,
                   float4* __restrict__ part, unsigned int* __restrict__ done) {
  extern __shared__ float smem[];
  float* sQ = smem;                 // kMQ x kPitch
  float* sT = smem + kMQ * kPitch;  // kMT x kPitch
  __shared__ int sRow[kMQ];
  __shared__ int sLast;
  const int n_rows = rows ? *n_rows_ptr : n_query;
  const int row_blocks = (n_rows + kMQ - 1) / kMQ;
  if (row_blocks == 0) return;
  int S = 1, chunk = n_train;
  if (part && row_blocks * 2 <= kExactItems) {
    S = min(kExactMaxSplits, kExactItems / row_blocks);
    chunk = (((n_train + S - 1) / S + kMT - 1) / kMT) * kMT;
    S = (n_train + chunk - 1) / chunk;
  }
  const int items = row_blocks * S;
  const int tq = threadIdx.x >> 4, tt = threadIdx.x & 15;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
  const int rb = item / S, split = item - rb * S;
  const int q_begin = rb * kMQ;
  const int t_begin = split * chunk, t_end = min(n_train, t_begin + chunk);
  __syncthreads();  // the previous item's merge has finished with sT / sRow
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
  Top2 best[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) best[a].init();
  for (int t0 = t_begin; t0 < t_end; t0 += kMT) {
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
  Top2 m;
  m.init();
  if (threadIdx.x < kMQ) {
    for (int c = 0; c < 32; ++c) {
      const int i = mI[threadIdx.x * 32 + c];
      if (i >= 0) m.push(mD[threadIdx.x * 32 + c], i);
    }
  }
  bool finish = true;
  if (S > 1) {
    if (threadIdx.x < kMQ)
      part[(size_t)item * kMQ + threadIdx.x] = make_float4(m.d0, m.d1, __int_as_float(m.i0), __int_as_float(m.i1));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sLast = atomicAdd(done + rb, 1u) == (unsigned)(S - 1);
    __syncthreads();
    finish = sLast != 0;
    if (finish) {
      __threadfence();
      if (threadIdx.x < kMQ) {
        m.init();
        for (int s2 = 0; s2 < S; ++s2) {
          const float4 v = __ldcg(part + ((size_t)rb * S + s2) * kMQ + threadIdx.x);
          const int i0 = __float_as_int(v.z), i1 = __float_as_int(v.w);
          if (i0 >= 0) m.push(v.x, i0);
          if (i1 >= 0) m.push(v.y, i1);
        }
      }
      if (threadIdx.x == 0) done[rb] = 0u;  // ready for the next call on this workspace
    }
  }
  if (finish && threadIdx.x < kMQ) {
    const int row = sRow[threadIdx.x];
    if (row >= 0) {
      knn_idx[2 * (size_t)row] = m.i0;
      knn_idx[2 * (size_t)row + 1] = m.i1;
      knn_dist[2 * (size_t)row] = m.d0;
      knn_dist[2 * (size_t)row + 1] = m.d1;
      good[row] = (m.i1 >= 0 && (double)m.d0 < __dmul_rn(ratio, (double)m.d1)) ? 1 : 0;
    }
  }
  }  // items
}

// bytes of the split scratch: kExactItems partial lists of 64 rows, and the per-query-block arrival counters
size_t match_exact_part_bytes() { return sizeof(float4) * (size_t)(kExactItems + kExactMaxSplits) * kMQ; }
size_t match_exact_done_words() { return (size_t)kExactItems; }

int launch_match_exact(const float* query, int n_query, const float* train, int n_train, const int32_t* rows,
                       const int32_t* n_rows_ptr, int max_rows, double ratio, int32_t* knn_idx, float* knn_dist,
                       uint8_t* good, void* part, unsigned int* done, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)(kMQ + kMT) * kPitch;
  FC_CUDA_TRY(cudaFuncSetAttribute(match_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = div_up(max_rows, kMQ);
  if (blocks == 0) return FC_OK;
  if (part && blocks > kExactItems) blocks = kExactItems;  // device-side count: the CTAs stride over the work items
  match_exact_kernel<<<blocks, 256, smem, st>>>(query, n_query, train, n_train, rows, n_rows_ptr, ratio, knn_idx, knn_dist, good,
                                               reinterpret_cast<float4*>(part), done);
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
                                nullptr, nullptr, (cudaStream_t)stream);
}
