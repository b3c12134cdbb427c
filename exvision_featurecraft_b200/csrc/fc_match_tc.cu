// Tensor-core 2-NN descriptor matching for sm_100a: tcgen05.mma (fp16 in, fp32 accumulate in TMEM), operand
// tiles staged by bulk async copies (TMA unit, 1-D mode), fused per-row running top-4 of 16-row GROUP minima in
// the epilogue, exact float32 re-rank of the members of the kept groups and an exact CUDA-core re-check of the rows the bounds cannot decide.
//
// Reference: cv2.BFMatcher().knnMatch(query, train, k=2) + ratio test (APP/utils/keypoint_descriptor.py:338-349).
//
//   score[m][n] = |t_n|^2 - 2 q_m . t_n   (= d^2 - |q_m|^2: same ranking per query row)
// is ONE GEMM: the operands are padded from K = 128 to 144, A' = [-2 q | 1 1 0..], B' = [t | hi(|t|^2) lo(|t|^2) 0..],
// so the accumulator already holds the score and the epilogue is a pure running-minimum.
//
//   match_prep_kernel    fp32 -> fp16 operands written in the UMMA canonical K-major no-swizzle layout
//                        (8-row x 16-byte core matrices; row-group stride SBO = 18*128 B, K stride LBO = 128 B),
//                        so a 128-row tile is one contiguous 36 KB block and needs no tensor map.
//   match_tc2_kernel     1 CTA pair per (256 queries, train split).  warp 0: bulk-copy producer (4-stage ring of
//                        256-row train tiles, half per CTA), warp 1: MMA issuer (9 tcgen05.mma.cta_group::2 256x256x16
//                        per tile, two accumulator stages = all 512 TMEM columns), warps 2-9: epilogue (tcgen05.ld,
//                        min-tree fast path, sorted top-4 insert on the rare slow path).
//   match_rerank_kernel  one warp per query: exact float32 distances of the members of the kept groups (same operation order
//                        as the exact kernel), top-2, ratio test, and the bound test that proves the answer equals
//                        the brute-force one; undecidable rows are appended to a list ...
//   match_exact_kernel   ... and re-done exactly (fc_match_exact.cu).
#include <cuda_fp16.h>

#include "fc_common.cuh"

namespace fc {

int launch_match_exact(const float* query, int n_query, const float* train, int n_train, const int32_t* rows,
                       const int32_t* n_rows_ptr, int max_rows, double ratio, int32_t* knn_idx, float* knn_dist,
                       uint8_t* good, void* part, unsigned int* done, cudaStream_t st);
size_t match_exact_part_bytes();
size_t match_exact_done_words();

namespace tc {
constexpr int kDim = 128;
constexpr int KP = 144;                        // padded K
constexpr int KC = KP / 8;                     // core matrices along K
constexpr int kRowGroupBytes = KC * 128;       // SBO: one 8-row group
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows / 8 * kRowGroupBytes;  // 36864
constexpr int kQueriesPerCta = 256;
constexpr int kCand = 4;   // kept groups per (query row, split)
constexpr int kGroup = 8;   // train rows per group
constexpr int kMaxSplits = 4;
constexpr int kThreads = 320;  // producer warp, MMA warp, 8 epilogue warps
constexpr float kPadNorm = 60000.0f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: start >> 4 | LBO(128 B) >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(kRowGroupBytes >> 4) << 32) |
         (1ull << 46);
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));  // FMNMX3
  return r;
}

struct TopK {
  float s[kCand];
  int i[kCand];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int c = 0; c < kCand; ++c) { s[c] = INFINITY; i[c] = -1; }
  }
  // sorted ascending by score; indices arrive in increasing order, so strict < keeps the lower index on ties
  __device__ __forceinline__ void insert(float v, int idx) {
    if (v < s[3]) {
      if (v < s[2]) {
        s[3] = s[2]; i[3] = i[2];
        if (v < s[1]) {
          s[2] = s[1]; i[2] = i[1];
          if (v < s[0]) { s[1] = s[0]; i[1] = i[0]; s[0] = v; i[0] = idx; }
          else { s[1] = v; i[1] = idx; }
        } else { s[2] = v; i[2] = idx; }
      } else { s[3] = v; i[3] = idx; }
    }
  }
};

__device__ __forceinline__ float min8(const uint32_t* r) {
  const float a = min3(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]));
  const float b = min3(__uint_as_float(r[3]), __uint_as_float(r[4]), __uint_as_float(r[5]));
  return min3(min3(a, b, __uint_as_float(r[6])), __uint_as_float(r[7]), INFINITY);
}

// 32 accumulator columns of one row = four groups of 8 consecutive train rows.  Only the MINIMUM of each group
// is tracked (FMNMX3 trees, no per-element branches): the running top-4 GROUPS per query row are kept, and the
// re-rank kernel evaluates all 8 members of the kept groups exactly.  The k-th best row always lies in one of
// the k best groups, and every row outside the kept groups scores >= the 4th kept group minimum.
__device__ __forceinline__ void scan_chunk(const uint32_t (&r)[32], int group0, TopK& best) {
  const float g0 = min8(r), g1 = min8(r + 8), g2 = min8(r + 16), g3 = min8(r + 24);
  if (min3(fminf(g0, g1), g2, g3) < best.s[3]) {
    best.insert(g0, group0);
    best.insert(g1, group0 + 1);
    best.insert(g2, group0 + 2);
    best.insert(g3, group0 + 3);
  }
}

// ------------------------------------------------------------------------------------------------
// operand preparation: row r, column k -> byte offset (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2
// one warp per row; is_query: values are scaled by -2 and the pad columns hold 1, 1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
match_prep_kernel(const float* __restrict__ src, int n_rows, int n_rows_padded, int is_query, __half* __restrict__ dst,
                  unsigned int* __restrict__ stats /* [0] max |t|^2 bits, [1] overflow flag */) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows_padded) return;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n_rows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * kDim) + lane);
  float nrm = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  const float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
  const float sc = is_query ? -2.0f : 1.0f;
  char* base = reinterpret_cast<char*>(dst) + (size_t)(row >> 3) * kRowGroupBytes + (row & 7) * 16;
  // lane holds columns 4*lane .. 4*lane+3 -> core matrix lane/2, half (lane&1)
  __half2 h01 = __floats2half2_rn(sc * v.x, sc * v.y), h23 = __floats2half2_rn(sc * v.z, sc * v.w);
  uint2 packed = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
  *reinterpret_cast<uint2*>(base + (lane >> 1) * 128 + (lane & 1) * 8) = packed;
  if (lane < 2) {
    // pad core matrices 16, 17: 8 halves each
    __half h[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = __float2half_rn(0.f);
    if (lane == 0) {
      if (is_query) {
        h[0] = __float2half_rn(1.f);
        h[1] = __float2half_rn(1.f);
      } else {
        const float n_use = row < n_rows ? nrm : kPadNorm;
        const __half hi = __float2half_rn(n_use);
        h[0] = hi;
        h[1] = __float2half_rn(n_use - __half2float(hi));
      }
    }
    *reinterpret_cast<uint4*>(base + (16 + lane) * 128) = *reinterpret_cast<uint4*>(h);
  }
  if (lane == 0 && row < n_rows) {
    // NaN / overflow of the fp16 operands -> the whole call is re-done exactly
    if (!(amax <= 30000.f) || !(nrm <= 60000.f)) atomicOr(stats + 1, 1u);
    if (!is_query) atomicMax(stats, __float_as_uint(nrm));  // nrm >= 0: bit pattern order == value order
  }
}

// ------------------------------------------------------------------------------------------------
// A cluster of two CTAs issues tcgen05.mma.cta_group::2 (M = 256 = 2 x 128 query rows, N = 256 train rows per
// instruction).  Each CTA stages its own 128-row query tile and ITS HALF of every 256-row train tile, so per SM and MMA
// only 4 KB (A) + 4 KB (half of B) are read from shared memory instead of 8 + 8 KB for the same FLOPs: a 1-SM
// (cta_group::1) version of this kernel is bound by the 128 B/clk shared-memory port, this one is not.
//   leader CTA (rank 0): MMA issue; waits for its own and the peer's operand halves, commits with multicast so the
//                        smem-free and accumulator-ready barriers of BOTH CTAs are signalled;
//   peer CTA (rank 1):   warp 1 relays "my half has landed" to the leader; epilogue warps arrive remotely on the
//                        leader's accumulator-free barrier.
// ------------------------------------------------------------------------------------------------
constexpr int kStages2 = 4;
constexpr int kTileRows2 = 256;                 // train rows per stage (128 per CTA)
constexpr size_t kSmemBytes2 = (1 + kStages2) * kTileBytes + 512;
// kind::f16: D = f32 (bit 4), A = B = f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc2 = (1u << 4) | ((256u >> 3) << 17) | ((256u >> 4) << 24);  // M = 256, N = 256

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc2), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
match_tc2_kernel(const __half* __restrict__ Ah, const __half* __restrict__ Bh, int n_query_padded, int n_tiles_total,
                 int tiles_per_split, float* __restrict__ cand_score, int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem;                         // my 128 query rows
  unsigned char* sB = smem + kTileBytes;            // kStages2 x my half (128 rows) of a 256-row train tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + kStages2) * kTileBytes);
  uint64_t* full = bars;                         // [kStages2] my half landed (tx)
  uint64_t* peer_full = full + kStages2;         // [kStages2] leader only: the peer's half landed
  uint64_t* empty = peer_full + kStages2;        // [kStages2] MMAs that read the stage are done (multicast commit)
  uint64_t* a_full = empty + kStages2;           // my query tile landed (tx)
  uint64_t* peer_a_full = a_full + 1;            // leader only
  uint64_t* t_full = peer_a_full + 1;            // [2] accumulators ready (multicast commit)
  uint64_t* t_empty = t_full + 2;                // [2] leader only: 8 local + 8 remote epilogue warps drained the stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int cq = blockIdx.x >> 1, split = blockIdx.y;
  const int tile0 = split * tiles_per_split;
  const int n_tiles = min(tiles_per_split, n_tiles_total - tile0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) { mbar_init(full + s, 1); mbar_init(peer_full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(a_full, 1);
    mbar_init(peer_a_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const char* a_src = reinterpret_cast<const char*>(Ah) + ((size_t)cq * 2 + rank) * kTileBytes;
      mbar_expect_tx(a_full, kTileBytes);
      bulk_g2s(sA, a_src, kTileBytes, a_full);
      const char* b_src = reinterpret_cast<const char*>(Bh) + ((size_t)tile0 * 2 + rank) * kTileBytes;
      for (int it = 0; it < n_tiles; ++it) {
        const int st = it % kStages2;
        const uint32_t ph = (it / kStages2) & 1;
        mbar_wait(empty + st, ph ^ 1);
        mbar_expect_tx(full + st, kTileBytes);
        bulk_g2s(sB + (size_t)st * kTileBytes, b_src + (size_t)it * 2 * kTileBytes, kTileBytes, full + st);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank == 1) {
        // relay: tell the leader when my operands are in my shared memory
        mbar_wait(a_full, 0);
        mbar_arrive_remote(peer_a_full, 0);
        for (int it = 0; it < n_tiles; ++it) {
          const int st = it % kStages2;
          mbar_wait(full + st, (it / kStages2) & 1);
          mbar_arrive_remote(peer_full + st, 0);
        }
      } else {
        mbar_wait(a_full, 0);
        mbar_wait_cluster(peer_a_full, 0);
        const uint32_t a_addr = smem_u32(sA);
        for (int it = 0; it < n_tiles; ++it) {
          const int st = it % kStages2;
          const uint32_t ph = (it / kStages2) & 1;
          const int acc = it & 1;
          const uint32_t aph = (it >> 1) & 1;
          mbar_wait_cluster(t_empty + acc, aph ^ 1);
          mbar_wait(full + st, ph);
          mbar_wait_cluster(peer_full + st, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t b_addr = smem_u32(sB + (size_t)st * kTileBytes);
          const uint32_t d = tmem_base + acc * 256;
#pragma unroll
          for (int kk = 0; kk < KP / 16; ++kk) umma2_f16(d, umma_desc(a_addr + kk * 256), umma_desc(b_addr + kk * 256), kk > 0);
          umma2_commit_both(empty + st);    // both CTAs may refill this stage
          umma2_commit_both(t_full + acc);  // both CTAs' epilogues may read the accumulators
        }
      }
    }
  } else {
    // epilogue (both CTAs): my 128 query rows x 256 train columns per tile; warps 2-5 scan columns 0..127, warps 6-9
    // columns 128..255 of the same rows
    const int lg = warp & 3;
    const int h = (warp - 2) >> 2;
    const int row = lg * 32 + lane;
    TopK best;
    best.init();
    for (int it = 0; it < n_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(t_full + acc, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n0 = ((tile0 + it) * kTileRows2 + h * 128) / kGroup;  // first 8-row group of my 128 columns
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * 256 + h * 128;
      uint32_t ra[32], rb[32];
      tmem_ld32_issue(taddr, ra);
      tmem_ld_wait();
      tmem_ld32_issue(taddr + 32, rb);
      scan_chunk(ra, n0, best);
      tmem_ld_wait();
      tmem_ld32_issue(taddr + 64, ra);
      scan_chunk(rb, n0 + 4, best);
      tmem_ld_wait();
      tmem_ld32_issue(taddr + 96, rb);
      scan_chunk(ra, n0 + 8, best);
      tmem_ld_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(t_empty + acc); else mbar_arrive_remote(t_empty + acc, 0);
      }
      scan_chunk(rb, n0 + 12, best);
    }
    // merge the two column halves of every row through shared memory (all operand stages are idle by now)
    float* mS = reinterpret_cast<float*>(sB);
    int* mI = reinterpret_cast<int*>(sB) + 128 * kCand;
    if (h == 1) {
#pragma unroll
      for (int c = 0; c < kCand; ++c) { mS[row * kCand + c] = best.s[c]; mI[row * kCand + c] = best.i[c]; }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (h == 0) {
#pragma unroll
      for (int c = 0; c < kCand; ++c) best.insert(mS[row * kCand + c], mI[row * kCand + c]);  // groups of h = 1 are larger ids: ties keep h = 0
      const size_t q = ((size_t)cq * 2 + rank) * 128 + row;
      const size_t o = ((size_t)split * n_query_padded + q) * kCand;
      *reinterpret_cast<float4*>(cand_score + o) = make_float4(best.s[0], best.s[1], best.s[2], best.s[3]);
      *reinterpret_cast<int4*>(cand_idx + o) = make_int4(best.i[0], best.i[1], best.i[2], best.i[3]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves before everything is done
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------
// exact re-rank of the candidates + decision bounds (see DESIGN.md "Matching: why the tensor path is exact")
// ------------------------------------------------------------------------------------------------
static constexpr int kRerankWarps = 4;

struct Best2 {
  float d0, d1;
  int i0, i1;
  __device__ __forceinline__ void init() { d0 = d1 = INFINITY; i0 = i1 = 0x7fffffff; }
  __device__ __forceinline__ void push(float d, int i) {
    if (d < d0 || (d == d0 && i < i0)) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
    else if (d < d1 || (d == d1 && i < i1)) { d1 = d; i1 = i; }
  }
};

__global__ void __launch_bounds__(kRerankWarps * 32)
match_rerank_kernel(const float* __restrict__ query, int n_query, int n_query_padded, const float* __restrict__ train,
                    int n_train, int n_splits, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_idx,
                    const unsigned int* __restrict__ stats, double ratio, int32_t* __restrict__ knn_idx,
                    float* __restrict__ knn_dist, uint8_t* __restrict__ good, int32_t* __restrict__ amb_rows,
                    int32_t* __restrict__ amb_count) {
  __shared__ float sQ[kRerankWarps][kDim];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kRerankWarps + warp;
  if (q >= n_query) return;
  const float4 qv = __ldg(reinterpret_cast<const float4*>(query + (size_t)q * kDim) + lane);
  sQ[warp][4 * lane] = qv.x; sQ[warp][4 * lane + 1] = qv.y; sQ[warp][4 * lane + 2] = qv.z; sQ[warp][4 * lane + 3] = qv.w;
  float q2 = qv.x * qv.x + qv.y * qv.y + qv.z * qv.z + qv.w * qv.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, o);
  __syncwarp();
  // kept groups: lane l < n_splits*kCand holds (group id, group minimum)
  int grp = -1;
  float gmin = INFINITY;
  if (lane < n_splits * kCand) {
    const size_t o = ((size_t)(lane / kCand) * n_query_padded + q) * kCand + (lane % kCand);
    grp = cand_idx[o];
    gmin = cand_score[o];
  }
  // Only the kEval = 2 best groups of every split are evaluated exactly: the best row of a split lies in its best
  // group and its second best row in its two best groups.  Every row that is NOT evaluated (members of the 3rd / 4th
  // kept group and everything outside) has an approximate score >= the 3rd kept minimum of its split.
  constexpr int kEval = 2;
  float s4 = (lane < n_splits * kCand && (lane % kCand) == kEval) ? gmin : INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s4 = fminf(s4, __shfl_xor_sync(0xffffffffu, s4, o));
  // exact float32 distances (same operation order as match_exact_kernel) of every member of the evaluated groups
  Best2 b;
  b.init();
  const int n_cand = n_splits * kEval * kGroup;
  for (int c0 = 0; c0 < n_cand; c0 += 32) {  // warp-uniform trip count: the shuffle below needs all 32 lanes
    const int c = c0 + lane;
    const int gi = (c / kGroup) % (kMaxSplits * kEval);  // evaluated group number: split gi / kEval, rank gi % kEval
    const int g = __shfl_sync(0xffffffffu, grp, (gi / kEval) * kCand + (gi % kEval));
    const int row = g * kGroup + (c % kGroup);
    if (c < n_cand && g >= 0 && row < n_train) {
      const float4* t = reinterpret_cast<const float4*>(train + (size_t)row * kDim);
      float acc = 0.0f;
#pragma unroll 16
      for (int k4 = 0; k4 < kDim / 4; ++k4) {
        const float4 tv = __ldg(t + k4);
        float df = __fsub_rn(sQ[warp][4 * k4], tv.x);
        acc = __fadd_rn(acc, __fmul_rn(df, df));
        df = __fsub_rn(sQ[warp][4 * k4 + 1], tv.y);
        acc = __fadd_rn(acc, __fmul_rn(df, df));
        df = __fsub_rn(sQ[warp][4 * k4 + 2], tv.z);
        acc = __fadd_rn(acc, __fmul_rn(df, df));
        df = __fsub_rn(sQ[warp][4 * k4 + 3], tv.w);
        acc = __fadd_rn(acc, __fmul_rn(df, df));
      }
      const float d = sqrtf(acc);
      if (d == d) b.push(d, row);
    }
  }
  // warp merge of the lane-local best-2 lists: two rounds of lexicographic arg-min; the winning lane pops
  float d0, d1;
  int i0, i1;
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    float bd = b.d0;
    int bi = b.i0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    if (round == 0) { d0 = bd; i0 = bi; } else { d1 = bd; i1 = bi; }
    if (b.i0 == bi && bi != 0x7fffffff) { b.d0 = b.d1; b.i0 = b.i1; b.d1 = INFINITY; b.i1 = 0x7fffffff; }
  }
  if (i0 == 0x7fffffff) i0 = -1;
  if (i1 == 0x7fffffff) i1 = -1;
  if (lane == 0) {
    const float tmax = __uint_as_float(stats[0]);
    const bool overflow = stats[1] != 0;
    // |approx - true score| <= delta: fp16 rounding of both operands (relative 2^-11 each, absolute 2^-25 floor),
    // the two-term fp16 split of |t|^2 and fp32 accumulation
    const float qn = sqrtf(q2), tn = sqrtf(tmax);
    const float delta = 2.0f * (0.0009775f * qn * tn + 6.8e-7f * (qn + tn)) + 2e-6f * (1.0f + q2 + tmax);
    const float lb2 = s4 + q2 - delta;  // lower bound on the true squared distance of every row that was not evaluated
    const float lb = lb2 > 0.f ? sqrtf(lb2) : 0.f;
    const double r = ratio;
    const bool is_good = i1 >= 0 && (double)d0 < r * (double)d1;
    bool decided;
    if (lb > d1) {
      decided = true;  // the kept groups contain the true two nearest neighbours
    } else if (is_good) {
      decided = (d0 < lb) && ((double)d0 < r * (double)fminf(lb, d1));
    } else {
      decided = ((double)lb >= r * (double)d0);
    }
    if (overflow) decided = false;
    else if (i0 < 0) decided = true;  // NaN query row: no match (the reference would raise)
    knn_idx[2 * (size_t)q] = i0;
    knn_idx[2 * (size_t)q + 1] = i1;
    knn_dist[2 * (size_t)q] = d0;
    knn_dist[2 * (size_t)q + 1] = d1;
    good[q] = is_good ? 1 : 0;
    if (!decided) amb_rows[atomicAdd(amb_count, 1)] = q;
  }
}

__global__ void zero_u32_kernel(unsigned int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}

struct Plan {
  int nq_pad, nt_pad, n_qtiles, n_tiles, splits, tiles_per_split;  // tiles = kTileRows2-row train tiles
  size_t off_a, off_b, off_cs, off_ci, off_stats, off_rows, off_part, total;
};

static Plan make_plan(int nq, int nt) {
  Plan p;
  const int tile_rows = kTileRows2;
  p.nq_pad = div_up(nq, kQueriesPerCta) * kQueriesPerCta;
  p.nt_pad = div_up(nt, tile_rows) * tile_rows;
  p.n_qtiles = p.nq_pad / kQueriesPerCta;  // 256 queries per CTA pair
  p.n_tiles = p.nt_pad / tile_rows;
  // number of train splits: fill the 148 SMs (1 CTA each) in whole waves, keep >= 16 tiles per CTA when possible
  int want = 1;
  double best_eff = -1.0;
  // (every extra split costs a second set of candidates in the re-rank: only split when it buys >= 15 %)
  for (int s = 1; s <= kMaxSplits && s <= p.n_tiles; ++s) {
    if (s > 1 && p.n_tiles / s < 16) break;
    const int ctas = p.n_qtiles * s * 2;
    const double eff = (double)ctas / ((double)div_up(ctas, 148) * 148.0);
    if (eff > best_eff + 0.15) { best_eff = eff; want = s; }
  }
  p.tiles_per_split = div_up(p.n_tiles, want);
  p.splits = div_up(p.n_tiles, p.tiles_per_split);
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t o = 0;
  p.off_a = o; o = align(o + (size_t)p.nq_pad * KP * 2);
  p.off_b = o; o = align(o + (size_t)p.nt_pad * KP * 2);
  p.off_cs = o; o = align(o + (size_t)p.splits * p.nq_pad * kCand * 4);
  p.off_ci = o; o = align(o + (size_t)p.splits * p.nq_pad * kCand * 4);
  p.off_stats = o; o = align(o + 16 + 4 * match_exact_done_words());  // [0..3] stats, then the re-check's arrival counters
  p.off_rows = o; o = align(o + (size_t)nq * 4);
  p.off_part = o; o = align(o + match_exact_part_bytes());
  p.total = o;
  return p;
}

}  // namespace tc
}  // namespace fc

using namespace fc;

extern "C" int64_t fc_match_workspace_bytes(int n_query, int n_train) {
  if (n_query <= 0 || n_train <= 0) return 0;
  return (int64_t)tc::make_plan(n_query, n_train).total;
}

extern "C" int64_t fc_match_workspace_stats_offset(int n_query, int n_train) {
  if (n_query <= 0 || n_train <= 0) return -1;
  return (int64_t)tc::make_plan(n_query, n_train).off_stats;
}

extern "C" int fc_match_knn2(const float* query, int n_query, const float* train, int n_train, double ratio,
                             int32_t* knn_idx, float* knn_dist, uint8_t* good, void* workspace, void* stream) {
  if (!query || !train || !knn_idx || !knn_dist || !good || !workspace || n_query <= 0 || n_train <= 0) return FC_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(train)) & 15) return FC_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const tc::Plan p = tc::make_plan(n_query, n_train);
  char* ws = reinterpret_cast<char*>(workspace);
  __half* Ah = reinterpret_cast<__half*>(ws + p.off_a);
  __half* Bh = reinterpret_cast<__half*>(ws + p.off_b);
  float* cs = reinterpret_cast<float*>(ws + p.off_cs);
  int32_t* ci = reinterpret_cast<int32_t*>(ws + p.off_ci);
  unsigned int* stats = reinterpret_cast<unsigned int*>(ws + p.off_stats);  // [0] max|t|^2, [1] overflow, [2] amb count
  int32_t* rows = reinterpret_cast<int32_t*>(ws + p.off_rows);

  const int n_zero = 4 + (int)match_exact_done_words();
  tc::zero_u32_kernel<<<div_up(n_zero, 256), 256, 0, st>>>(stats, n_zero);
  note_launch();
  tc::match_prep_kernel<<<div_up(p.nq_pad * 32, 256), 256, 0, st>>>(query, n_query, p.nq_pad, 1, Ah, stats);
  note_launch();
  tc::match_prep_kernel<<<div_up(p.nt_pad * 32, 256), 256, 0, st>>>(train, n_train, p.nt_pad, 0, Bh, stats);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  {
    FC_CUDA_TRY(cudaFuncSetAttribute(tc::match_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes2));
    dim3 grid(2 * p.n_qtiles, p.splits);
    tc::match_tc2_kernel<<<grid, tc::kThreads, tc::kSmemBytes2, st>>>(Ah, Bh, p.nq_pad, p.n_tiles, p.tiles_per_split, cs, ci);
  }
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  tc::match_rerank_kernel<<<div_up(n_query, tc::kRerankWarps), tc::kRerankWarps * 32, 0, st>>>(
      query, n_query, p.nq_pad, train, n_train, p.splits, cs, ci, stats, ratio, knn_idx, knn_dist, good, rows,
      reinterpret_cast<int32_t*>(stats + 2));
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return launch_match_exact(query, n_query, train, n_train, rows, reinterpret_cast<const int32_t*>(stats + 2), n_query, ratio,
                            knn_idx, knn_dist, good, ws + p.off_part, stats + 4, st);
}
