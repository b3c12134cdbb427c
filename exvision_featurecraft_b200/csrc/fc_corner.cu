// Harris / lambda-minus corner response kernels (sm_100a).
//
//   harris5_f32_kernel     app Harris, 5x5 window (the reference default and the benchmarked config):
//                          sliding-window kernel, one warp per 128-column strip, 4 pixels per thread,
//                          exact packed-float32 sums (bytes unpacked by PRMT, no conversion-unit ops).
//   lambda5_packed_kernel  lambda-minus, same strip structure, exact integer sums in 16-bit lanes + LAPACK's dlae2 sequence.
//   corner_tiled_kernel    every other combination (any odd window <= 31, K_X gradient, lambda-minus):
//                          shared-memory tiled, exact integer sums.
//
// Reference arithmetic: APP/FeatureCraft_Backend.py:336-355 (Harris), :402-437 (lambda-minus),
// implementation_without_ui/Corner Detection/harris.ipynb:150-190 (notebook Harris).
// HBM traffic per pixel: 1 B gray in + 8 B float64 response out (+ 1 bit mask).
#include "fc_common.cuh"

namespace fc {

enum { GRAD_CENTRAL = 0, GRAD_KX = 1 };
enum { RESP_HARRIS = 0, RESP_LAMBDA_MINUS = 1 };

// ------------------------------------------------------------------------------------------------
// Shared pieces of the two strip kernels: packed image rows and exact integer -> float64 moves
// ------------------------------------------------------------------------------------------------
struct Row3 {
  uint32_t w[3];  // bytes of columns c0-4 .. c0+7 of one image row
};

__device__ __forceinline__ Row3 load_row3(const uint8_t* __restrict__ img, int row, int H, int W, int c0) {
  row = min(max(row, 0), H - 1);
  const uint32_t* p = reinterpret_cast<const uint32_t*>(img + (size_t)row * W);
  const int wi = c0 >> 2;
  const int nw = W >> 2;
  Row3 r;
  r.w[0] = (wi - 1 >= 0 && wi - 1 < nw) ? __ldg(p + wi - 1) : 0u;
  r.w[1] = (wi < nw) ? __ldg(p + wi) : 0u;
  r.w[2] = (wi + 1 < nw) ? __ldg(p + wi + 1) : 0u;
  return r;
}

// value of column c0 - 4 + k (k = 0..11) of a packed row: one PRMT
__device__ __forceinline__ int col_i(const Row3& r, int k) {
  return (int)__byte_perm(r.w[k >> 2], 0u, 0x4440u + (unsigned)(k & 3));
}

// exact int64 -> float64 for |n| < 2^51: build 2^52 + (n + 2^51) from bits, subtract 2^52 + 2^51
__device__ __forceinline__ double exact_i64_to_f64(long long n) {
  const unsigned long long u = (unsigned long long)(n + (1ll << 51));
  const double biased = __hiloint2double((int)(0x43300000u | (unsigned)(u >> 32)), (int)(unsigned)u);
  return __dsub_rn(biased, 6755399441055744.0);  // 2^52 + 2^51
}

static constexpr int kStripWarps = 4;
static constexpr int kStripCols = 128;  // per warp
static constexpr int kStripThreads = kStripWarps * 32;

// ------------------------------------------------------------------------------------------------
// Harris fast path, packed-fp32 variant (the one fc_harris_response launches)
// ------------------------------------------------------------------------------------------------
// Every quantity before the response is a multiple of 1/4 below 2^22, i.e. exactly representable in float32:
// gradients (multiples of 1/2, |g| <= 255), products (<= 65 025), 5-tap row sums (<= 325 125), 5x5 sums
// (<= 1 625 625 -> 23 significant bits with the two fraction bits).  float32 adds / multiplies of such values are
// exact, so the sums equal the reference's float64 sums bit for bit, and sm_100's packed FADD2 / FMUL2 do two of
// them per issue slot.  Bytes are unpacked straight into "2^22 + byte/2" floats by one PRMT (0x4A800000 | byte:
// the mantissa LSB of that binade is 1/2); the bias cancels in every difference, so the halved central difference
// of np.gradient costs one subtraction.  Only the response itself (det, trace^2 up to 2^44) needs float64: three
// exact float->double conversions and six float64 operations per pixel, two of which round -- the same two as in
// the reference (k * trace^2 and the final subtraction, APP/FeatureCraft_Backend.py:351-355).
// 56 instructions per pixel (an exact int32 / int64 predecessor of this kernel needed 81).
struct RowF {
  float lo;      // column c0 - 3            (k = 1 in Row3 numbering)
  float2 e[4];   // columns c0 - 2 .. c0 + 5 (k = 2 .. 9), register-pair aligned
  float hi;      // column c0 + 6            (k = 10)
  template <int K> __device__ __forceinline__ float get() const {
    static_assert(K >= 1 && K <= 10, "column out of range");
    if constexpr (K == 1) return lo;
    else if constexpr (K == 10) return hi;
    else if constexpr ((K & 1) == 0) return e[(K - 2) / 2].x;
    else return e[(K - 2) / 2].y;
  }
};

__device__ __forceinline__ float unpack_half_biased(uint32_t w, int byte) {
  return __uint_as_float(__byte_perm(w, 0x4A800000u, 0x7640u + (unsigned)byte));
}

__device__ __forceinline__ RowF unpack_row(const Row3& r) {
  RowF f;
  f.lo = unpack_half_biased(r.w[0], 1);
  f.e[0] = make_float2(unpack_half_biased(r.w[0], 2), unpack_half_biased(r.w[0], 3));
  f.e[1] = make_float2(unpack_half_biased(r.w[1], 0), unpack_half_biased(r.w[1], 1));
  f.e[2] = make_float2(unpack_half_biased(r.w[1], 2), unpack_half_biased(r.w[1], 3));
  f.e[3] = make_float2(unpack_half_biased(r.w[2], 0), unpack_half_biased(r.w[2], 1));
  f.hi = unpack_half_biased(r.w[2], 2);
  return f;
}

__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

// 5-tap sums h[j] = p[j] + .. + p[j+4], j = 0..3, of eight consecutive values given as four pairs (9 additions)
__device__ __forceinline__ float4 five_tap_sums(const float2 (&p)[4]) {
  const float t1 = p[0].y + p[1].x, t2 = p[1].y + p[2].x, t3 = p[2].y + p[3].x;
  const float c = t1 + t2, c2 = t2 + t3;
  return make_float4(c + p[0].x, c + p[2].y, c2 + p[1].x, c2 + p[3].y);
}

template <int I> __device__ __forceinline__ void border_grad(const RowF& up, const RowF& mid, const RowF& dn, float sy,
                                                             int c0, int W, float (&gx)[8], float (&gy)[8]) {
  constexpr int K = I + 2;
  const int c = c0 - 2 + I;
  const bool inside = (c >= 0) && (c < W);
  const float vx = sy * (dn.get<K>() - up.get<K>());
  float vy;
  if (c == 0) vy = 2.f * (mid.get<K + 1>() - mid.get<K>());
  else if (c == W - 1) vy = 2.f * (mid.get<K>() - mid.get<K - 1>());
  else vy = mid.get<K + 1>() - mid.get<K - 1>();
  gx[I] = inside ? vx : 0.f;
  gy[I] = inside ? vy : 0.f;
}

// h[0] = xx, h[1] = xy, h[2] = yy row sums for output columns c0 .. c0+3
template <bool INTERIOR>
__device__ __forceinline__ void harris_row_sums_f32(const RowF& up, const RowF& mid, const RowF& dn, int row, int H, int W,
                                                    int c0, float4 (&h)[3]) {
  float2 gx[4], gy[4];  // product columns c0-2 .. c0+5
  if (INTERIOR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) gx[i] = sub2(dn.e[i], up.e[i]);
    gy[0] = make_float2(mid.e[0].y - mid.lo, mid.e[1].x - mid.e[0].x);
    gy[1] = make_float2(mid.e[1].y - mid.e[0].y, mid.e[2].x - mid.e[1].x);
    gy[2] = make_float2(mid.e[2].y - mid.e[1].y, mid.e[3].x - mid.e[2].x);
    gy[3] = make_float2(mid.e[3].y - mid.e[2].y, mid.hi - mid.e[3].x);
  } else {
    // rows were clamped when loaded, so dn - up is half the one-sided difference on the first / last row
    const float sy = (row == 0 || row == H - 1) ? 2.f : 1.f;
    float ax[8], ay[8];
    border_grad<0>(up, mid, dn, sy, c0, W, ax, ay); border_grad<1>(up, mid, dn, sy, c0, W, ax, ay);
    border_grad<2>(up, mid, dn, sy, c0, W, ax, ay); border_grad<3>(up, mid, dn, sy, c0, W, ax, ay);
    border_grad<4>(up, mid, dn, sy, c0, W, ax, ay); border_grad<5>(up, mid, dn, sy, c0, W, ax, ay);
    border_grad<6>(up, mid, dn, sy, c0, W, ax, ay); border_grad<7>(up, mid, dn, sy, c0, W, ax, ay);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gx[i] = make_float2(ax[2 * i], ax[2 * i + 1]);
      gy[i] = make_float2(ay[2 * i], ay[2 * i + 1]);
    }
  }
  float2 xx[4], xy[4], yy[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xx[i] = __fmul2_rn(gx[i], gx[i]);
    xy[i] = __fmul2_rn(gx[i], gy[i]);
    yy[i] = __fmul2_rn(gy[i], gy[i]);
  }
  h[0] = five_tap_sums(xx);
  h[1] = five_tap_sums(xy);
  h[2] = five_tap_sums(yy);
}

// the reference's R from exact float32 window sums: det and trace^2 are exact in float64, two roundings remain
__device__ __forceinline__ double harris_from_f32_sums(float sxx, float sxy, float syy, double k) {
  const double a = (double)sxx, b = (double)sxy, c = (double)syy;
  const double det = __fma_rn(-b, b, __dmul_rn(a, c));
  const double tr = __dadd_rn(a, c);
  return __dsub_rn(det, __dmul_rn(k, __dmul_rn(tr, tr)));
}

struct HarrisF32State {
  float4 S[3];           // running 5x5 sums (xx, xy, yy) of the four output columns
  int slot;              // ring offset of the row that leaves the window next
  const uint32_t* in_p;  // FAST: this lane's word (column c0) of the next row to stage
  uint32_t* stage;       // FAST: this warp's staging rows in shared memory, [kHarrisStageRows][36] words
  int stage_slot;        // FAST: staging row that holds image row hr+2
  double* out_p;         // FAST: response address of (row hr-2, column c0)
  uint32_t* mask_p;      // FAST: mask word of (row hr-2, column c0)
  Row3 nxt;              // bytes of row hr+2, loaded at the end of the previous step
};

static constexpr int kHarrisStageRows = 4;  // image rows in flight per warp (three ahead of the one being unpacked)

// FAST path input staging: one 4-byte cp.async per lane per image row (lanes 0 and 1 also fetch the two halo
// words), landing in a small per-warp ring in shared memory.  Completion is tracked by cp.async groups, not by
// register scoreboards, so three rows (about three steps of run time) are in flight without holding a register,
// and every byte is fetched from L1/L2 once instead of three times.
__device__ __forceinline__ void harris_stage_row(const uint32_t* __restrict__ own_word, uint32_t* stage_row, int lane) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(stage_row + 1 + lane);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(own_word) : "memory");
  if (lane < 2) {
    // lane 0: word left of the strip -> stage word 0; lane 1: word right of the strip -> stage word 33
    const uint32_t dh = (uint32_t)__cvta_generic_to_shared(stage_row + lane * 33);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dh), "l"(own_word - 1 + lane * 32) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// One image row hr of the band: row sums of hr -> ring -> window sums -> response of row hr-2; then the bytes of
// row hr+2 replace the buffer of row hr-1 (FAST: from the staging ring, and row hr+5 is requested; otherwise from
// registers loaded at the end of the previous step).
// FAST: the whole strip and rows hr-1 .. hr+3 are inside the image (no clamping, no border gradient, every lane
// active, pointer increments instead of index arithmetic).
template <bool HAS_MASK, bool FAST>
__device__ __forceinline__ void harris_f32_step(HarrisF32State& st, RowF& up, const RowF& mid, const RowF& dn,
                                                const uint8_t* __restrict__ img, int hr, int H, int W, int c0,
                                                bool cols_interior, bool active, int y0, float4* __restrict__ ring,
                                                double k, double thr, double* __restrict__ o, uint32_t* __restrict__ m,
                                                int mask_pitch, int lane) {
  float4 h[3];
  if (FAST) {
    harris_row_sums_f32<true>(up, mid, dn, hr, H, W, c0, h);
  } else if (hr < 0 || hr >= H) {
    h[0] = h[1] = h[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (cols_interior && hr >= 1 && hr <= H - 2) {
    harris_row_sums_f32<true>(up, mid, dn, hr, H, W, c0, h);
  } else {
    harris_row_sums_f32<false>(up, mid, dn, hr, H, W, c0, h);
  }
  float4* slot = ring + st.slot;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float4 old = slot[q * kStripThreads];
    slot[q * kStripThreads] = h[q];
    const float2 d0 = sub2(make_float2(h[q].x, h[q].y), make_float2(old.x, old.y));
    const float2 d1 = sub2(make_float2(h[q].z, h[q].w), make_float2(old.z, old.w));
    const float2 s0 = __fadd2_rn(make_float2(st.S[q].x, st.S[q].y), d0);
    const float2 s1 = __fadd2_rn(make_float2(st.S[q].z, st.S[q].w), d1);
    st.S[q] = make_float4(s0.x, s0.y, s1.x, s1.y);
  }
  st.slot = (st.slot == 4 * 3 * kStripThreads) ? 0 : st.slot + 3 * kStripThreads;
  const int r = hr - 2;  // rows r-2 .. r+2 are now summed
  if (r >= y0) {
    double res[4];
    res[0] = harris_from_f32_sums(st.S[0].x, st.S[1].x, st.S[2].x, k);
    res[1] = harris_from_f32_sums(st.S[0].y, st.S[1].y, st.S[2].y, k);
    res[2] = harris_from_f32_sums(st.S[0].z, st.S[1].z, st.S[2].z, k);
    res[3] = harris_from_f32_sums(st.S[0].w, st.S[1].w, st.S[2].w, k);
    if (FAST || active) {
      double2* dst = reinterpret_cast<double2*>(FAST ? st.out_p : o + (size_t)r * W + c0);
      __stcs(dst, make_double2(res[0], res[1]));
      __stcs(dst + 1, make_double2(res[2], res[3]));
    }
    if (HAS_MASK) {
      uint32_t nib = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) nib |= (res[j] > thr) ? (1u << j) : 0u;
      uint32_t v = (FAST || active) ? (nib << (4 * (lane & 7))) : 0u;
      v |= __shfl_xor_sync(0xffffffffu, v, 1);  // (a REDUX with per-group member masks was measured 20 % slower)
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 4);
      if ((lane & 7) == 0 && (FAST || active)) *(FAST ? st.mask_p : m + (size_t)r * mask_pitch + (c0 >> 5)) = v;
    }
    if (FAST) {
      st.out_p += W;
      if (HAS_MASK) st.mask_p += mask_pitch;
    }
  }
  if (FAST) {
    // the copy of row hr+2 was committed three groups ago; the warp barrier makes every lane's words visible and
    // also orders this step's reads after the previous step's, so the slot refilled below is no longer in use
    asm volatile("cp.async.wait_group %0;" ::"n"(kHarrisStageRows - 2) : "memory");
    __syncwarp();
    const uint32_t* srow = st.stage + st.stage_slot * 36 + lane;
    Row3 cur;
    cur.w[0] = srow[0]; cur.w[1] = srow[1]; cur.w[2] = srow[2];
    up = unpack_row(cur);  // the buffer of row hr-1 becomes row hr+2
    const int refill = (st.stage_slot + kHarrisStageRows - 1) & (kHarrisStageRows - 1);  // slot of row hr+1
    harris_stage_row(st.in_p, st.stage + refill * 36, lane);                              // <- row hr+5
    st.in_p += W >> 2;
    st.stage_slot = (st.stage_slot + 1) & (kHarrisStageRows - 1);
  } else {
    up = unpack_row(st.nxt);
    st.nxt = load_row3(img, hr + 3, H, W, c0);
  }
}

// 4 CTAs per SM (<= 128 registers): at 5 the step spills, and with the staged input 16 warps per SM are enough
template <bool HAS_MASK>
__global__ void __launch_bounds__(kStripThreads, 4)
harris5_f32_kernel(const uint8_t* __restrict__ gray, int H, int W, int rows_per_band, double k, double thr,
                   double* __restrict__ out, uint32_t* __restrict__ mask, int mask_pitch) {
  // ring of the last 5 rows of horizontal sums, private to each thread: [slot][quantity][thread] float4 (conflict free)
  __shared__ float4 sRing[5 * 3 * kStripThreads];
  __shared__ uint32_t sStage[kStripWarps * kHarrisStageRows * 36];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x * kStripWarps + warp;
  const int strip_c0 = strip * kStripCols;
  if (strip_c0 >= W) return;  // warp-uniform, and the kernel has no block-wide barrier
  const int c0 = strip_c0 + lane * 4;
  const bool active = c0 < W;
  const size_t img_off = (size_t)blockIdx.z * H * W;
  const uint8_t* img = gray + img_off;
  double* o = out + img_off;
  uint32_t* m = HAS_MASK ? mask + (size_t)blockIdx.z * H * mask_pitch : nullptr;
  const int y0 = blockIdx.y * rows_per_band;
  const int y1 = min(y0 + rows_per_band, H);
  const bool cols_interior = (strip_c0 - 2 >= 1) && (strip_c0 + kStripCols + 1 <= W - 2);
  float4* ring = sRing + threadIdx.x;
  HarrisF32State st;
  st.slot = 0;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    st.S[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 5; ++s) ring[(s * 3 + q) * kStripThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  RowF A = unpack_row(load_row3(img, y0 - 3, H, W, c0));
  RowF B = unpack_row(load_row3(img, y0 - 2, H, W, c0));
  RowF C = unpack_row(load_row3(img, y0 - 1, H, W, c0));
  const int last = y1 + 1;
  int hr = y0 - 2;
#define FC_STEP(FAST, U, M, D) \
  harris_f32_step<HAS_MASK, FAST>(st, U, M, D, img, hr, H, W, c0, cols_interior, active, y0, ring, k, thr, o, m, mask_pitch, lane)
  if (cols_interior && y0 - 3 >= 0 && y1 + 6 <= H - 1) {
    // band and strip entirely inside the image.  The roles up / mid / dn rotate through A, B, C without moving a
    // register: three rows per trip.  Rows y0 .. y0+2 are staged before the first step (hr = y0-2 consumes row y0).
    st.stage = sStage + warp * (kHarrisStageRows * 36);
    st.stage_slot = 0;
    st.in_p = reinterpret_cast<const uint32_t*>(img + (size_t)y0 * W + c0);
#pragma unroll
    for (int i = 0; i < kHarrisStageRows - 1; ++i) {
      harris_stage_row(st.in_p, st.stage + i * 36, lane);
      st.in_p += W >> 2;
    }
    st.out_p = o + (size_t)y0 * W + c0;
    st.mask_p = HAS_MASK ? m + (size_t)y0 * mask_pitch + (c0 >> 5) : nullptr;
    for (;;) {
      FC_STEP(true, A, B, C); if (++hr > last) break;
      FC_STEP(true, B, C, A); if (++hr > last) break;
      FC_STEP(true, C, A, B); if (++hr > last) break;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // the last three staged rows are never consumed
  } else {
    // bands / strips that touch the image border: same step with every check, rolled (registers are moved)
    st.nxt = load_row3(img, y0, H, W, c0);
#pragma unroll 1
    for (; hr <= last; ++hr) {
      FC_STEP(false, A, B, C);
      const RowF t = A; A = B; B = C; C = t;
    }
  }
#undef FC_STEP
}

// ------------------------------------------------------------------------------------------------
// Fast path: lambda-minus (K_X gradients, 5x5 window / 25, smallest eigenvalue), same strip structure
// ------------------------------------------------------------------------------------------------
// K_X = a (x) [-1 0 0 0 1] + b (x) [0 -1 0 1 0] with a = (1,2,3,2,1), b = (2,3,5,3,2) down the rows
// (APP/FeatureCraft_Backend.py:405-413), applied as zero-padded correlation.  Pixels are >= 0, so |g| <= 255 x (sum of
// the positive weights = 24) = 6120, products <= 3.75e7, 5-tap row sums <= 1.9e8, 5x5 sums <= 9.4e8: everything is int32.
// One IEEE division by 25 per sum and LAPACK's dlae2 sequence give numpy's eigvalsh bit for bit.
// exact int32 -> float64: bits of 2^52 + (n + 2^31), minus 2^52 + 2^31
__device__ __forceinline__ double exact_i32_to_f64(int n) {
  return __dsub_rn(__hiloint2double(0x43300000, n ^ (int)0x80000000), 4503601774854144.0);
}

// uint64 -> float64 with one rounding: exact halves, hi * 2^32 + lo in one fused operation
__device__ __forceinline__ double u64_to_f64(unsigned long long n);

// exact uint32 -> float64: bits of 2^52 + n, minus 2^52
__device__ __forceinline__ double exact_u32_to_f64(uint32_t n) {
  return __dsub_rn(__hiloint2double(0x43300000, (int)n), 4503599627370496.0);
}

__device__ __forceinline__ double u64_to_f64(unsigned long long n) {
  return __fma_rn(exact_u32_to_f64((uint32_t)(n >> 32)), 4294967296.0, exact_u32_to_f64((uint32_t)n));
}

// n / 25 correctly rounded for every integer |n| < 2^33 (what `sum / 25` gives in numpy): reciprocal multiply +
// one Markstein correction (two FMAs).  fc_selftest_div25 checks all 2^33 + 1 magnitudes against __ddiv_rn on the
// device (tests/test_corners_gpu.py runs it), so this is a proven identity, not an approximation.
__device__ __forceinline__ double div25_exact(double n) {
  const double y = 0.04;  // rn(1/25)
  const double q = __dmul_rn(n, y);
  const double r = __fma_rn(-q, 25.0, n);
  return __fma_rn(r, y, q);
}

__global__ void selftest_div25_kernel(unsigned long long* __restrict__ mismatches) {
  unsigned long long n = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (; n <= (1ull << 33); n += stride) {
    const double d = (double)n;
    bad += (div25_exact(d) != __ddiv_rn(d, 25.0)) + (div25_exact(-d) != __ddiv_rn(-d, 25.0));
  }
  if (bad) atomicAdd(mismatches, bad);
}

// div_seq / sqrt_seq against the library's correctly rounded operations on pseudo-random operands of the range the
// lambda-minus kernel feeds them (magnitudes 2^-10 .. 2^30, quotients of any size in between, radicands in [1, 2] and
// across the range), 2^32 samples per call
__global__ void selftest_divsqrt_kernel(unsigned long long* __restrict__ mismatches) {
  unsigned long long n = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (; n < (1ull << 32); n += stride) {
    // splitmix64
    unsigned long long z = n * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    unsigned long long w = z * 0xD6E8FEB86659FD93ull;
    w ^= w >> 32;
    // mantissas from the hash, exponents in [-10, 30]
    const int ea = (int)(z % 41) - 10, eb = (int)((z >> 8) % 41) - 10;
    const double a = __longlong_as_double((long long)((z >> 12) | ((unsigned long long)(1023 + ea) << 52)));
    const double b = __longlong_as_double((long long)((w >> 12) | ((unsigned long long)(1023 + eb) << 52)));
    const double sa = (z & 1) ? -a : a;
    bad += (div_seq(sa, b, rcp_seq(b)) != __ddiv_rn(sa, b));
    bad += (sqrt_seq(a) != __dsqrt_rn(a));
    const double x12 = __longlong_as_double((long long)((w >> 12) | (1023ull << 52)));  // [1, 2)
    bad += (sqrt_seq(x12) != __dsqrt_rn(x12));
    // multiples of 1/25 like the kernel's operands
    const double m1 = __ddiv_rn((double)(unsigned)(z >> 33), 25.0), m2 = __ddiv_rn((double)(unsigned)((w >> 33) | 1u), 25.0);
    bad += (div_seq(m1, m2, rcp_seq(m2)) != __ddiv_rn(m1, m2));
  }
  if (bad) atomicAdd(mismatches, bad);
}

// out-of-line copy of the rare LAPACK split path (one body instead of one per pixel)
static __device__ __noinline__ double dsterf_2x2_min_call(double a, double b, double c) { return dsterf_2x2_min(a, b, c); }

// ------------------------------------------------------------------------------------------------
// lambda-minus, packed form: the integer half of the kernel in 16-bit lanes
// ------------------------------------------------------------------------------------------------
// ncu on the first version (every row unpacked to int32 columns): 945 instructions per thread and row (4 pixels), 240 of them float64 (the LAPACK
// sequence, irreducible) and ~420 the K_X gradients: per row every one of the 12 columns was unpacked from 5 rows and
// pushed through both 5-tap kernels, every gradient through a 5x2-tap sum.  Here
//   * values live as PAIRS of 16-bit lanes in one register (every intermediate fits: |.| <= 12240), encoded as
//     hi * 65536 + lo with lo signed, so plain 32-bit adds / subtracts are exact lane-wise arithmetic and one IADD3
//     adds three pairs.  Pairs at even columns are E[j] = (v[2j], v[2j+1]); the odd-aligned O[j] = (v[2j+1], v[2j+2])
//     come from one PRMT -- built from unsigned quantities only, where the encoding has no borrow;
//   * both 5-tap kernels are two box filters: (1,2,3,2,1) = box3 * box3 and (2,3,5,3,2) = 2 * (1,2,3,2,1) - box3;
//   * K_X is separable in the other direction too (K_Y = K_X^T), so  gy = A'(r+2) - A'(r-2) + B'(r+1) - B'(r-1)  with
//     A', B' the HORIZONTAL kernels of single rows: each row is filtered once, when it enters, and kept in a
//     per-thread shared-memory ring (5 rows x 32 bytes); the vertical kernels of gx keep running box sums in registers.
// A row costs ~125 integer instructions instead of ~420; sums, quotients and the eigenvalue sequence are unchanged.
struct LamPairs {
  uint32_t e[6];
};

__device__ __forceinline__ LamPairs lam_unpack(const uint32_t (&w)[3]) {
  LamPairs v;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    v.e[2 * q] = __byte_perm(w[q], 0u, 0x4140u);      // (byte 0, byte 1) as 16-bit lanes
    v.e[2 * q + 1] = __byte_perm(w[q], 0u, 0x4342u);  // (byte 2, byte 3)
  }
  return v;
}

// odd-aligned pair (x[2j+1], x[2j+2]) of two even-aligned pairs with non-negative lanes
__device__ __forceinline__ uint32_t lam_odd(uint32_t ej, uint32_t ej1) { return __byte_perm(ej, ej1, 0x5432u); }

// sign-extended low lane: PTX prmt replicates the sign of a byte when bit 3 of its selector nibble is set
// (__byte_perm only honours the low three bits)
__device__ __forceinline__ int lam_lo(uint32_t p) {
  int r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(p), "r"(0u), "r"(0x9910u));
  return r;
}
__device__ __forceinline__ int lam_hi(uint32_t p) { return (int)(p + 0x8000u) >> 16; }         // high lane incl. the borrow

// FAST: the opt-in response mode of fc_lambda_minus_response_fast (see there); false = LAPACK's arithmetic, bit for bit.
template <bool FAST>
__global__ void __launch_bounds__(kStripThreads, 4)
lambda5_packed_kernel(const uint8_t* __restrict__ gray, int H, int W, int rows_per_band, double* __restrict__ out,
                      unsigned long long* __restrict__ img_max_bits) {
  __shared__ int4 sRing[5 * 3 * kStripThreads];   // [slot][xx, xy, yy][thread]: row sums of the last 5 rows
  extern __shared__ __align__(16) uint4 sHor[];   // [row % 5][A', B'][thread]: horizontal kernels of the last 5 rows (dynamic: 20 KB)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strip = blockIdx.x * kStripWarps + warp;
  const int strip_c0 = strip * kStripCols;
  if (strip_c0 >= W) return;  // warp-uniform, no block-wide barrier in this kernel
  const int c0 = strip_c0 + lane * 4;
  const bool active = c0 < W;
  const size_t img_off = (size_t)blockIdx.z * H * W;
  const uint8_t* img = gray + img_off;
  double* o = out + img_off;
  const int y0 = blockIdx.y * rows_per_band;
  const int y1 = min(y0 + rows_per_band, H);
  const bool cols_interior = (strip_c0 - 2 >= 0) && (strip_c0 + kStripCols + 1 <= W - 1);
  int4* ring = sRing + threadIdx.x;
  uint4* hor = sHor + threadIdx.x;
  uint32_t Sxx[4], Syy[4];
  int Sxy[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { Sxx[j] = 0u; Syy[j] = 0u; Sxy[j] = 0; }
#pragma unroll
  for (int i = 0; i < 15; ++i) ring[i * kStripThreads] = make_int4(0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < 10; ++i) hor[i * kStripThreads] = make_uint4(0u, 0u, 0u, 0u);
  // the thread's three words of a row (columns c0-4 .. c0+7), zero outside the image (padding_matrix)
  const int wi = c0 >> 2, nw = W >> 2;
  const bool ok0 = wi - 1 >= 0 && wi - 1 < nw, ok1 = wi < nw, ok2 = wi + 1 < nw;
  const uint32_t* img32 = reinterpret_cast<const uint32_t*>(img) + wi;
  auto load_row = [&](int row, uint32_t (&w)[3]) {
    const bool in = (unsigned)row < (unsigned)H;  // warp-uniform
    const uint32_t* p = img32 + (size_t)(in ? row : 0) * nw;
    w[0] = (in && ok0) ? __ldg(p - 1) : 0u;
    w[1] = (in && ok1) ? __ldg(p) : 0u;
    w[2] = (in && ok2) ? __ldg(p + 1) : 0u;
  };
  // vertical state of the gx kernels: raw pairs of rows q-2, q-1 and box sums s(q-3), s(q-2) when row q enters
  LamPairs vA, vB, sP, sC;
#pragma unroll
  for (int j = 0; j < 6; ++j) vA.e[j] = vB.e[j] = sP.e[j] = sC.e[j] = 0u;
  uint32_t nxt[3];
  load_row(y0 - 4, nxt);
  double local_max = -INFINITY;
  int slot = 0;   // row-sum ring
  int hslot = 0;  // slot of the entering row in the horizontal ring (rows older by 4, 3, 1 sit at +1, +2, +4 mod 5)
#pragma unroll 1
  for (int q = y0 - 4; q <= y1 + 3; ++q) {  // row q enters; hr = q - 2 is the gradient row, r = q - 4 the output row
    uint32_t w[3] = {nxt[0], nxt[1], nxt[2]};
    load_row(q + 1, nxt);  // consumed one iteration later
    const LamPairs vN = lam_unpack(w);
    // horizontal kernels of row q at the even pairs j = 1..4 (columns c0-2 .. c0+5)
    uint32_t od[5], se[5], so[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) od[j] = lam_odd(vN.e[j], vN.e[j + 1]);
#pragma unroll
    for (int j = 0; j < 5; ++j) so[j] = vN.e[j] + od[j] + vN.e[j + 1];
#pragma unroll
    for (int j = 1; j < 5; ++j) se[j] = od[j - 1] + vN.e[j] + od[j];
    uint32_t An[4], Bn[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      An[p] = so[p] + se[p + 1] + so[p + 1];
      Bn[p] = 2u * An[p] - se[p + 1];
    }
    const int s1 = hslot + 1 >= 5 ? hslot - 4 : hslot + 1;  // row q-4
    const int s2 = hslot + 2 >= 5 ? hslot - 3 : hslot + 2;  // row q-3
    const int s4 = hslot + 4 >= 5 ? hslot - 1 : hslot + 4;  // row q-1
    const uint4 Aold = hor[(2 * s1) * kStripThreads];
    const uint4 Bup = hor[(2 * s2 + 1) * kStripThreads];
    const uint4 Bdn = hor[(2 * s4 + 1) * kStripThreads];
    hor[(2 * hslot) * kStripThreads] = make_uint4(An[0], An[1], An[2], An[3]);
    hor[(2 * hslot + 1) * kStripThreads] = make_uint4(Bn[0], Bn[1], Bn[2], Bn[3]);
    hslot = hslot == 4 ? 0 : hslot + 1;
    // vertical kernels at row hr = q-2: s(hr+1) = v(hr) + v(hr+1) + v(hr+2), A = s(hr-1) + s(hr) + s(hr+1), B = 2A - s(hr)
    LamPairs sN, A, B;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      sN.e[j] = vA.e[j] + vB.e[j] + vN.e[j];
      A.e[j] = sP.e[j] + sC.e[j] + sN.e[j];
      B.e[j] = 2u * A.e[j] - sC.e[j];
    }
    const int hr = q - 2;
    int h[12];
    if (hr < 0 || hr >= H || hr < y0 - 2) {
#pragma unroll
      for (int i = 0; i < 12; ++i) h[i] = 0;
    } else {
      uint32_t bo[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) bo[j] = lam_odd(B.e[j], B.e[j + 1]);
      const uint32_t ao[4] = {Aold.x, Aold.y, Aold.z, Aold.w}, bu[4] = {Bup.x, Bup.y, Bup.z, Bup.w},
                     bd[4] = {Bdn.x, Bdn.y, Bdn.z, Bdn.w};
      int gx[8], gy[8];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const uint32_t GX = (A.e[p + 2] - A.e[p]) + (bo[p + 1] - bo[p]);
        const uint32_t GY = (An[p] - ao[p]) + (bd[p] - bu[p]);
        gx[2 * p] = lam_lo(GX); gx[2 * p + 1] = lam_hi(GX);
        gy[2 * p] = lam_lo(GY); gy[2 * p + 1] = lam_hi(GY);
      }
      if (!cols_interior) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = c0 - 2 + i;
          if (c < 0 || c >= W) { gx[i] = 0; gy[i] = 0; }  // the window sum is zero padded
        }
      }
      int xx = gx[0] * gx[0], xy = gx[0] * gy[0], yy = gy[0] * gy[0];
#pragma unroll
      for (int i = 1; i < 5; ++i) {
        xx += gx[i] * gx[i];
        xy += gx[i] * gy[i];
        yy += gy[i] * gy[i];
      }
      h[0] = xx; h[4] = xy; h[8] = yy;
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        xx += gx[j + 4] * gx[j + 4] - gx[j - 1] * gx[j - 1];
        xy += gx[j + 4] * gy[j + 4] - gx[j - 1] * gy[j - 1];
        yy += gy[j + 4] * gy[j + 4] - gy[j - 1] * gy[j - 1];
        h[j] = xx; h[4 + j] = xy; h[8 + j] = yy;
      }
    }
    // slide the vertical state
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      vA.e[j] = vB.e[j]; vB.e[j] = vN.e[j];
      sP.e[j] = sC.e[j]; sC.e[j] = sN.e[j];
    }
    if (hr < y0 - 2) continue;  // the four rows above the band only prime the state
    int4* s = ring + slot;
    const int4 oxx = s[0], oxy = s[kStripThreads], oyy = s[2 * kStripThreads];
    s[0] = make_int4(h[0], h[1], h[2], h[3]);
    s[kStripThreads] = make_int4(h[4], h[5], h[6], h[7]);
    s[2 * kStripThreads] = make_int4(h[8], h[9], h[10], h[11]);
    slot = (slot == 4 * 3 * kStripThreads) ? 0 : slot + 3 * kStripThreads;
    const int ox[4] = {oxx.x, oxx.y, oxx.z, oxx.w}, om[4] = {oxy.x, oxy.y, oxy.z, oxy.w}, oy[4] = {oyy.x, oyy.y, oyy.z, oyy.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Sxx[j] += (uint32_t)(h[j] - ox[j]);
      Sxy[j] += h[4 + j] - om[j];
      Syy[j] += (uint32_t)(h[8 + j] - oy[j]);
    }
    const int r = hr - 2;  // rows r-2 .. r+2 are now summed
    if (r >= y0) {
      double res[4];
      if constexpr (FAST) {
        // smaller eigenvalue = det / larger eigenvalue with the determinant and the radicand of the larger one formed
        // EXACTLY in 64-bit integers (Sxx Syy - Sxy^2 < 2^60, (Sxx - Syy)^2 + 4 Sxy^2 < 2^63), so the cancellation that
        // makes the textbook formula lose digits on edges never happens: three roundings (two conversions, one square
        // root) and one division separate the result from the exact value
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long sxy2 = (long long)Sxy[j] * (long long)Sxy[j];
          const long long det = (long long)((unsigned long long)Sxx[j] * (unsigned long long)Syy[j]) - sxy2;
          const long long d = (long long)Sxx[j] - (long long)Syy[j];
          const unsigned long long rad = (unsigned long long)(d * d) + 4ull * (unsigned long long)sxy2;
          const double R = rad ? sqrt_seq(u64_to_f64(rad)) : 0.0;  // isotropic window: both eigenvalues are Sxx / 25
          const double den = __dadd_rn(exact_u32_to_f64(Sxx[j] + Syy[j]), R);  // 50 x the larger eigenvalue
          const double num = __dmul_rn(u64_to_f64((unsigned long long)det), 0.08);  // 2 det / 25
          res[j] = det > 0 ? div_seq(num, den, rcp_seq(den)) : 0.0;           // det > 0 implies den > 0
        }
      } else {
      double av[4], bv[4], cv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        av[j] = div25_exact(exact_u32_to_f64(Sxx[j]));
        bv[j] = div25_exact(exact_i32_to_f64(Sxy[j]));
        cv[j] = div25_exact(exact_u32_to_f64(Syy[j]));
        res[j] = lambda_min_psd(av[j], bv[j], cv[j]);  // straight-line: the four float64 chains interleave
      }
      bool rare = false;
#pragma unroll
      for (int j = 0; j < 4; ++j) rare |= lambda_needs_dsterf(av[j], bv[j], cv[j]);
      if (rare) {  // tiny non-zero off-diagonal: LAPACK's split test decides (practically never on image data)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (lambda_needs_dsterf(av[j], bv[j], cv[j])) res[j] = dsterf_2x2_min_call(av[j], bv[j], cv[j]);
      }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (active) local_max = res[j] > local_max ? res[j] : local_max;  // res is never NaN
      }
      if (active) {
        double2* dst = reinterpret_cast<double2*>(o + (size_t)r * W + c0);
        __stcs(dst, make_double2(res[0], res[1]));
        __stcs(dst + 1, make_double2(res[2], res[3]));
      }
    }
  }
  if (img_max_bits) {
    const double m = warp_max(local_max);
    if (lane == 0 && m > -INFINITY) atomicMax(img_max_bits + blockIdx.z, ordered_bits(m));
  }
}

// ------------------------------------------------------------------------------------------------
// General path: shared-memory tiled, any odd window <= 31, both gradients, both responses
// ------------------------------------------------------------------------------------------------
static constexpr int kTileX = 32;
static constexpr int kTileY = 16;
static constexpr int kTiledThreads = 256;

__device__ __forceinline__ int kx_a(int d) { return d == 2 ? 3 : (d == 1 || d == 3) ? 2 : 1; }  // 1 2 3 2 1
__device__ __forceinline__ int kx_b(int d) { return d == 2 ? 5 : (d == 1 || d == 3) ? 3 : 2; }  // 2 3 5 3 2

template <int GRAD>
__global__ void __launch_bounds__(kTiledThreads)
corner_tiled_kernel(const uint8_t* __restrict__ gray, int H, int W, int R, int resp, double k,
                    double* __restrict__ out, unsigned long long* __restrict__ img_max_bits) {
  constexpr int G = (GRAD == GRAD_KX) ? 2 : 1;  // gradient reach
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gw = kTileX + 2 * R, gh = kTileY + 2 * R;  // gradient tile
  const int iw = gw + 2 * G, ih = gh + 2 * G;          // intensity tile
  int* sI = reinterpret_cast<int*>(smem_raw);          // ih * iw
  int* sGx = sI + ih * iw;                             // gh * gw
  int* sGy = sGx + gh * gw;                            // gh * gw
  int* sH = sGy + gh * gw;                             // 3 * gh * kTileX
  const size_t img_off = (size_t)blockIdx.z * H * W;
  const uint8_t* img = gray + img_off;
  const int x0 = blockIdx.x * kTileX, y0 = blockIdx.y * kTileY;

  // 1. intensities.  CENTRAL: clamped coordinates (np.gradient's one-sided edges fall out of the
  //    clamp + a factor 2); KX: zeros outside the image (padding_matrix).
  for (int t = threadIdx.x; t < ih * iw; t += kTiledThreads) {
    const int ly = t / iw, lx = t - ly * iw;
    int y = y0 - R - G + ly, x = x0 - R - G + lx;
    int v;
    if (GRAD == GRAD_CENTRAL) {
      y = min(max(y, 0), H - 1);
      x = min(max(x, 0), W - 1);
      v = img[(size_t)y * W + x];
    } else {
      v = (y >= 0 && y < H && x >= 0 && x < W) ? (int)img[(size_t)y * W + x] : 0;
    }
    sI[t] = v;
  }
  __syncthreads();
  // 2. gradients on the (gh x gw) tile; zero outside the image (the window sum is zero padded)
  for (int t = threadIdx.x; t < gh * gw; t += kTiledThreads) {
    const int ly = t / gw, lx = t - ly * gw;
    const int y = y0 - R + ly, x = x0 - R + lx;
    int gx = 0, gy = 0;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const int* c = sI + (ly + G) * iw + (lx + G);
      if (GRAD == GRAD_CENTRAL) {
        // 2 * np.gradient: axis 0 first (named Ix in the reference), then axis 1
        gx = (c[iw] - c[-iw]) * ((y == 0 || y == H - 1) ? 2 : 1);
        gy = (c[1] - c[-1]) * ((x == 0 || x == W - 1) ? 2 : 1);
      } else {
#pragma unroll
        for (int d = 0; d < 5; ++d) {
          const int* rr = c + (d - 2) * iw;
          gx += kx_a(d) * (rr[2] - rr[-2]) + kx_b(d) * (rr[1] - rr[-1]);          // K_X
          gy += kx_a(d) * (c[2 * iw + d - 2] - c[-2 * iw + d - 2]) +
                kx_b(d) * (c[iw + d - 2] - c[-iw + d - 2]);                        // K_X^T
        }
      }
    }
    sGx[t] = gx;
    sGy[t] = gy;
  }
  __syncthreads();
  // 3. horizontal window sums of the products (int32: (2R+1) * 6120^2 < 2^31 for R <= 15)
  for (int t = threadIdx.x; t < gh * kTileX; t += kTiledThreads) {
    const int ly = t / kTileX, lx = t - ly * kTileX;
    const int* px = sGx + ly * gw + lx;
    const int* py = sGy + ly * gw + lx;
    int xx = 0, xy = 0, yy = 0;
    for (int d = 0; d <= 2 * R; ++d) {
      const int a = px[d], b = py[d];
      xx += a * a;
      xy += a * b;
      yy += b * b;
    }
    sH[t] = xx;
    sH[gh * kTileX + t] = xy;
    sH[2 * gh * kTileX + t] = yy;
  }
  __syncthreads();
  // 4. vertical sums (int64) + response
  double local_max = -INFINITY;
  for (int t = threadIdx.x; t < kTileY * kTileX; t += kTiledThreads) {
    const int ly = t / kTileX, lx = t - ly * kTileX;
    const int y = y0 + ly, x = x0 + lx;
    if (y >= H || x >= W) continue;
    long long a = 0, b = 0, c = 0;
    for (int d = 0; d <= 2 * R; ++d) {
      const int idx = (ly + d) * kTileX + lx;
      a += sH[idx];
      b += sH[gh * kTileX + idx];
      c += sH[2 * gh * kTileX + idx];
    }
    // CENTRAL sums are in quarter units (exact scaling); KX sums are the reference's integers
    const double scale = (GRAD == GRAD_CENTRAL) ? 0.25 : 1.0;
    double sxx = (double)a * scale, sxy = (double)b * scale, syy = (double)c * scale;
    double r;
    if (resp == RESP_HARRIS) {
      r = harris_from_sums(sxx, sxy, syy, k);
    } else {
      r = dsterf_2x2_min(__ddiv_rn(sxx, 25.0), __ddiv_rn(sxy, 25.0), __ddiv_rn(syy, 25.0));
      local_max = fmax(local_max, r);
    }
    out[img_off + (size_t)y * W + x] = r;
  }
  if (resp == RESP_LAMBDA_MINUS && img_max_bits) {
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max > -INFINITY)
      atomicMax(img_max_bits + blockIdx.z, ordered_bits(local_max));
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, int n, unsigned long long v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void decode_max_kernel(unsigned long long* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<double*>(p)[i] = from_ordered_bits(p[i]);
}

static size_t tiled_smem_bytes(int R, int G) {
  const int gw = kTileX + 2 * R, gh = kTileY + 2 * R;
  const int iw = gw + 2 * G, ih = gh + 2 * G;
  return sizeof(int) * ((size_t)ih * iw + 2 * (size_t)gh * gw + 3 * (size_t)gh * kTileX);
}

template <int GRAD>
static int launch_tiled(const uint8_t* gray, int B, int H, int W, int R, int resp, double k, double* out,
                        unsigned long long* max_bits, cudaStream_t st) {
  const size_t smem = tiled_smem_bytes(R, GRAD == GRAD_KX ? 2 : 1);
  FC_CUDA_TRY(cudaFuncSetAttribute(corner_tiled_kernel<GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(div_up(W, kTileX), div_up(H, kTileY), B);
  corner_tiled_kernel<GRAD><<<grid, kTiledThreads, smem, st>>>(gray, H, W, R, resp, k, out, max_bits);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

// ------------------------------------------------------------------------------------------------
// gray conversion, thresholds, compaction
// ------------------------------------------------------------------------------------------------
__global__ void rgb_to_gray_kernel(const uint8_t* __restrict__ rgb, int64_t n, uint8_t* __restrict__ gray) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
  // np.dot over the 3 channels, left to right, then the truncating astype(uint8)
  const double v = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2989), __dmul_rn(g, 0.5870)), __dmul_rn(b, 0.1140));
  gray[i] = (uint8_t)(int)v;
}

// one warp per (image row, 128 columns): 4 pixels per lane (two 16-byte loads in flight), nibbles merged by shuffles
__global__ void __launch_bounds__(256)
threshold_mask_kernel(const double* __restrict__ map, int B, int H, int W, double thr_abs, double fraction,
                      const double* __restrict__ image_max, int border, uint32_t* __restrict__ mask, int pitch) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = (pitch + 3) >> 2;  // 128-column groups per row
  const int64_t total = (int64_t)B * H * groups;
  if (warp >= total) return;
  const int grp = (int)(warp % groups);
  const int64_t row_g = warp / groups;
  const int y = (int)(row_g % H);
  const int b = (int)(row_g / H);
  const double thr = image_max ? __dmul_rn(fraction, image_max[b]) : thr_abs;
  const int x0 = grp * 128 + lane * 4;
  const double* src = map + row_g * W;
  const bool row_ok = (y >= border && y < H - border);
  uint32_t nib = 0;
  if (row_ok && x0 < W) {
    double v[4];
    if (((W & 1) == 0) && x0 + 3 < W && ((reinterpret_cast<uintptr_t>(src + x0) & 15) == 0)) {
      const double2 p = __ldcs(reinterpret_cast<const double2*>(src + x0));
      const double2 q = __ldcs(reinterpret_cast<const double2*>(src + x0) + 1);
      v[0] = p.x; v[1] = p.y; v[2] = q.x; v[3] = q.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (x0 + j < W) ? src[x0 + j] : -INFINITY;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      if (x < W && x >= border && x < W - border && v[j] > thr) nib |= 1u << j;
    }
  }
  uint32_t v32 = nib << (4 * (lane & 7));
  v32 |= __shfl_xor_sync(0xffffffffu, v32, 1);
  v32 |= __shfl_xor_sync(0xffffffffu, v32, 2);
  v32 |= __shfl_xor_sync(0xffffffffu, v32, 4);
  const int word = grp * 4 + (lane >> 3);
  if ((lane & 7) == 0 && word < pitch) mask[row_g * pitch + word] = v32;
}

// one thread per mask row: popcount of the row
__global__ void mask_row_count_kernel(const uint32_t* __restrict__ mask, int64_t rows, int pitch,
                                      int32_t* __restrict__ counts) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int c = 0;
  for (int w = lane; w < pitch; w += 32) c += __popc(mask[warp * pitch + w]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[warp] = c;
}

// one warp per mask row: ordered emission of (row, col) and the gathered map values
__global__ void mask_emit_kernel(const uint32_t* __restrict__ mask, const int32_t* __restrict__ row_offsets,
                                 int64_t rows, int H, int W, int pitch, const double* __restrict__ map,
                                 int32_t* __restrict__ coords, double* __restrict__ values, int triples) {
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  int base = row_offsets[warp];
  if (row_offsets[warp + 1] == base) return;
  const int y = (int)(warp % H);
  for (int w0 = 0; w0 < pitch; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = (w < pitch) ? mask[warp * pitch + w] : 0u;
    const int cnt = __popc(bits);
    const int inc = warp_inclusive_scan(cnt);
    int pos = base + inc - cnt;
    while (bits) {
      const int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      const int x = w * 32 + bit;
      if (triples) {  // (image, row, col) keypoint records
        coords[3 * (size_t)pos] = (int)(warp / H);
        coords[3 * (size_t)pos + 1] = y;
        coords[3 * (size_t)pos + 2] = x;
      } else {
        coords[2 * (size_t)pos] = y;
        coords[2 * (size_t)pos + 1] = x;
      }
      if (values) values[pos] = map[warp * W + x];
      ++pos;
    }
    base += __shfl_sync(0xffffffffu, inc, 31);
  }
}

}  // namespace fc

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace fc;

static bool bad_image_args(const void* p, int B, int H, int W) { return !p || B <= 0 || H <= 0 || W <= 0; }

extern "C" int fc_rgb_to_gray_u8(const uint8_t* rgb, int batch, int height, int width, uint8_t* gray, void* stream) {
  if (bad_image_args(rgb, batch, height, width) || !gray) return FC_ERR_BAD_ARG;
  const int64_t n = (int64_t)batch * height * width;
  rgb_to_gray_kernel<<<(unsigned)div_up64(n, 256), 256, 0, (cudaStream_t)stream>>>(rgb, n, gray);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_harris_response(const uint8_t* gray, int batch, int height, int width, int window_size, double k,
                                  double threshold, double* response, uint32_t* mask, void* stream) {
  if (bad_image_args(gray, batch, height, width) || !response) return FC_ERR_BAD_ARG;
  if (window_size < 1 || window_size > 31 || (window_size & 1) == 0) return FC_ERR_BAD_ARG;
  if (height < 2 || width < 2) return FC_ERR_BAD_ARG;  // np.gradient needs two samples per axis
  cudaStream_t st = (cudaStream_t)stream;
  const int pitch = div_up(width, 32);
  const bool aligned = (width % 4 == 0) && ((reinterpret_cast<uintptr_t>(gray) & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(response) & 15) == 0);
  if (window_size == 5 && aligned) {
    // band height: amortise the 6 halo rows, but keep enough warps in flight on small inputs
    const int strips = div_up(width, kStripCols);
    int rows = 64;
    while (rows > 8 && (int64_t)strips * div_up(height, rows) * batch < 148 * 16) rows >>= 1;
    dim3 grid(div_up(strips, kStripWarps), div_up(height, rows), batch);
    if (mask)
      harris5_f32_kernel<true><<<grid, kStripWarps * 32, 0, st>>>(gray, height, width, rows, k, threshold, response, mask, pitch);
    else
      harris5_f32_kernel<false><<<grid, kStripWarps * 32, 0, st>>>(gray, height, width, rows, k, threshold, response, mask, pitch);
    note_launch();
    FC_RETURN_IF_LAUNCH_FAILED();
    return FC_OK;
  }
  int rc = launch_tiled<GRAD_CENTRAL>(gray, batch, height, width, window_size / 2, RESP_HARRIS, k, response, nullptr, st);
  if (rc != FC_OK) return rc;
  if (mask) return fc_threshold_mask_abs(response, batch, height, width, threshold, 0, mask, stream);
  return FC_OK;
}

extern "C" int fc_notebook_harris_response(const uint8_t* gray, int batch, int height, int width, int window_size,
                                           double k, double* response, void* stream) {
  if (bad_image_args(gray, batch, height, width) || !response) return FC_ERR_BAD_ARG;
  // the notebook takes any size: offset = int(window_size / 2), window = 2 * offset + 1 (harris.ipynb:170-176); only the
  // app's convolve2d_optimized path fails on even sizes
  if (window_size < 1 || window_size > 31) return FC_ERR_BAD_ARG;
  return launch_tiled<GRAD_KX>(gray, batch, height, width, window_size / 2, RESP_HARRIS, k, response, nullptr,
                               (cudaStream_t)stream);
}

extern "C" int fc_selftest_div25(unsigned long long* mismatches, void* stream) {
  if (!mismatches) return FC_ERR_BAD_ARG;
  FC_CUDA_TRY(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  selftest_div25_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(mismatches);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_selftest_divsqrt(unsigned long long* mismatches, void* stream) {
  if (!mismatches) return FC_ERR_BAD_ARG;
  FC_CUDA_TRY(cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  selftest_divsqrt_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(mismatches);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

static int launch_lambda_minus(const uint8_t* gray, int batch, int height, int width, double* eig, double* image_max,
                               bool fast, void* stream) {
  if (bad_image_args(gray, batch, height, width) || !eig) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* bits = reinterpret_cast<unsigned long long*>(image_max);
  if (bits) {
    fill_u64_kernel<<<div_up(batch, 256), 256, 0, st>>>(bits, batch, 0ull);
    note_launch();
    FC_RETURN_IF_LAUNCH_FAILED();
  }
  const bool aligned = (width % 4 == 0) && ((reinterpret_cast<uintptr_t>(gray) & 3) == 0) &&
                       ((reinterpret_cast<uintptr_t>(eig) & 15) == 0);
  if (fast && !aligned) return FC_ERR_UNSUPPORTED;
  if (aligned) {
    const int strips = div_up(width, kStripCols);
    int rows = 64;
    while (rows > 8 && (int64_t)strips * div_up(height, rows) * batch < 148 * 16) rows >>= 1;
    dim3 grid(div_up(strips, kStripWarps), div_up(height, rows), batch);
    const int hor_bytes = (int)(sizeof(uint4) * 10 * kStripThreads);
    if (fast) {
      FC_CUDA_TRY(cudaFuncSetAttribute(lambda5_packed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, hor_bytes));
      lambda5_packed_kernel<true><<<grid, kStripThreads, hor_bytes, st>>>(gray, height, width, rows, eig, bits);
    } else {
      FC_CUDA_TRY(cudaFuncSetAttribute(lambda5_packed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, hor_bytes));
      lambda5_packed_kernel<false><<<grid, kStripThreads, hor_bytes, st>>>(gray, height, width, rows, eig, bits);
    }
    note_launch();
    FC_RETURN_IF_LAUNCH_FAILED();
  } else {
    int rc = launch_tiled<GRAD_KX>(gray, batch, height, width, 2, RESP_LAMBDA_MINUS, 0.0, eig, bits, st);
    if (rc != FC_OK) return rc;
  }
  if (bits) {
    decode_max_kernel<<<div_up(batch, 256), 256, 0, st>>>(bits, batch);
    note_launch();
    FC_RETURN_IF_LAUNCH_FAILED();
  }
  return FC_OK;
}

extern "C" int fc_lambda_minus_response(const uint8_t* gray, int batch, int height, int width, double* eig,
                                        double* image_max, void* stream) {
  return launch_lambda_minus(gray, batch, height, width, eig, image_max, false, stream);
}

extern "C" int fc_lambda_minus_response_fast(const uint8_t* gray, int batch, int height, int width, double* eig,
                                             double* image_max, void* stream) {
  return launch_lambda_minus(gray, batch, height, width, eig, image_max, true, stream);
}

static int launch_threshold(const double* map, int B, int H, int W, double thr, double frac, const double* image_max,
                            int border, uint32_t* mask, cudaStream_t st) {
  if (bad_image_args(map, B, H, W) || !mask || border < 0) return FC_ERR_BAD_ARG;
  const int pitch = div_up(W, 32);
  const int64_t warps = (int64_t)B * H * div_up(pitch, 4);
  threshold_mask_kernel<<<(unsigned)div_up64(warps * 32, 256), 256, 0, st>>>(map, B, H, W, thr, frac, image_max, border,
                                                                           mask, pitch);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_threshold_mask_abs(const double* map, int batch, int height, int width, double threshold, int border,
                                     uint32_t* mask, void* stream) {
  return launch_threshold(map, batch, height, width, threshold, 0.0, nullptr, border, mask, (cudaStream_t)stream);
}

extern "C" int fc_threshold_mask_rel(const double* map, int batch, int height, int width, double fraction,
                                     const double* image_max, uint32_t* mask, void* stream) {
  if (!image_max) return FC_ERR_BAD_ARG;
  return launch_threshold(map, batch, height, width, 0.0, fraction, image_max, 0, mask, (cudaStream_t)stream);
}

extern "C" int64_t fc_scan_scratch_elems(int64_t n) { return n + scan_scratch_elems(n); }

extern "C" int fc_mask_scan(const uint32_t* mask, int batch, int height, int width, int32_t* row_offsets,
                            int32_t* scratch, void* stream) {
  if (bad_image_args(mask, batch, height, width) || !row_offsets || !scratch) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)batch * height;
  const int pitch = div_up(width, 32);
  mask_row_count_kernel<<<(unsigned)div_up64(rows * 32, 256), 256, 0, st>>>(mask, rows, pitch, scratch);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return exclusive_scan_i32(scratch, rows, row_offsets, scratch + rows, st);
}

// counts[i] = set bits of mask row i (the first half of fc_mask_scan), for callers that scan several masks at once
extern "C" int fc_mask_row_count(const uint32_t* mask, int batch, int height, int width, int32_t* counts, void* stream) {
  if (bad_image_args(mask, batch, height, width) || !counts) return FC_ERR_BAD_ARG;
  const int64_t rows = (int64_t)batch * height;
  mask_row_count_kernel<<<(unsigned)div_up64(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(mask, rows, div_up(width, 32), counts);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_mask_emit(const uint32_t* mask, const int32_t* row_offsets, int batch, int height, int width,
                            const double* map, int32_t* coords, double* values, void* stream) {
  if (bad_image_args(mask, batch, height, width) || !row_offsets || !coords) return FC_ERR_BAD_ARG;
  if (values && !map) return FC_ERR_BAD_ARG;
  const int64_t rows = (int64_t)batch * height;
  mask_emit_kernel<<<(unsigned)div_up64(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      mask, row_offsets, rows, height, width, div_up(width, 32), map, coords, values, 0);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_mask_emit_keypoints(const uint32_t* mask, const int32_t* row_offsets, int batch, int height, int width,
                                      int32_t* keypoints, void* stream) {
  if (bad_image_args(mask, batch, height, width) || !row_offsets || !keypoints) return FC_ERR_BAD_ARG;
  const int64_t rows = (int64_t)batch * height;
  mask_emit_kernel<<<(unsigned)div_up64(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      mask, row_offsets, rows, height, width, div_up(width, 32), nullptr, keypoints, nullptr, 1);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

// ------------------------------------------------------------------------------------------------
// convolve2d_optimized (APP/utils/harris_utils.py:33-80) for arbitrary float64 data
// ------------------------------------------------------------------------------------------------
namespace fc {

struct WindowProducts {
  const double* img;
  const double* ker;
  int H, W, ks, pad, y, x;
  __device__ __forceinline__ double operator()(int i) const {
    const int dy = i / ks, dx = i - dy * ks;
    const int yy = y + dy - pad, xx = x + dx - pad;
    const double v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[(size_t)yy * W + xx] : 0.0;
    return __dmul_rn(v, ker[i]);
  }
};

// numpy's pairwise summation (numpy/core/src/umath/loops_utils.h.src, pairwise_sum_DOUBLE) -- the
// reduction np.sum(..., axis=(2, 3)) performs over the k*k contiguous products of one output pixel.
__device__ double numpy_pairwise_sum(const WindowProducts& a, int lo, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, a(lo + i));
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a(lo + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __dadd_rn(numpy_pairwise_sum(a, lo, n2), numpy_pairwise_sum(a, lo + n2, n - n2));
}

__global__ void correlate_zero_pad_f64_kernel(const double* __restrict__ img, int H, int W,
                                              const double* __restrict__ ker, int ks, double* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const size_t off = (size_t)blockIdx.z * H * W;
  WindowProducts a{img + off, ker, H, W, ks, ks / 2, y, x};
  out[off + (size_t)y * W + x] = numpy_pairwise_sum(a, 0, ks * ks);
}

}  // namespace fc

extern "C" int fc_correlate_zero_pad_f64(const double* image, int batch, int height, int width, const double* kernel,
                                         int kernel_size, double* out, void* stream) {
  if (bad_image_args(image, batch, height, width) || !kernel || !out) return FC_ERR_BAD_ARG;
  if (kernel_size < 1 || kernel_size > 63 || (kernel_size & 1) == 0) return FC_ERR_BAD_ARG;
  dim3 block(32, 8), grid(div_up(width, 32), div_up(height, 8), batch);
  correlate_zero_pad_f64_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(image, height, width, kernel, kernel_size, out);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
