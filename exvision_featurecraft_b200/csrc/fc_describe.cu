// SIFT orientation assignment and descriptors (sm_100a).
//
//   gradient_field_kernel<T>    sift_gradient (APP/utils/keypoint_descriptor.py:72-79) of one pyramid level:
//                               magnitude, direction (degrees, Python % 360) and the 36-bin index
//                               np.round(direction * 36 / 360) (:140-142) as one byte per pixel (float64 mode).
//   gradient_codes(4)_kernel    certified mode: float32 magnitude | exact 5-degree direction code, one word per pixel.
//   orientation_kernel<T,LPK,NR,MODE>
//                               8 / 16 / 32 lanes per keypoint: zero-padded (2r+1)^2 window, weight = magnitude x
//                               unnormalised Gaussian table, 36-bin histogram (:144-163).  Every lane owns a private
//                               copy of the histogram in shared memory (no atomics, fixed summation order); bins are
//                               rounded to float32 and compared with float32(0.8) * max (:165-167).  kOriCoded: the
//                               certified float32 form (separable weights, bank-per-lane table, error-bound flag).
//   orientation_recheck_kernel  the flagged keypoints of certified mode in float64, one CTA each.
//   descriptor_kernel<T,OUT>    one warp per oriented keypoint: 16x16 nearest-neighbour rotated window with
//                               OpenCV's 10-bit fixed-point coordinates (:175-212), 4x4 cells x 8 bins with
//                               float32 cell histograms, normalise / clip [2^-10, 0.2] / normalise (:259-287).
//                               Two lanes per cell, 8 samples per lane (float64 mode and caller-supplied planes).
//   descriptor_coded_kernel<OUT>  the pipeline's kernel: half a warp per keypoint, a lane per cell, integer bins.
//   sift_layout_kernel          final image-major placement of descriptors and keypoint records.
#include <cuda_fp16.h>

#include "fc_common.cuh"

namespace fc {

template <typename T> struct Math;
template <> struct Math<float> {
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ float fmod_(float a, float b) { return fmodf(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct Math<double> {
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double fmod_(double a, double b) { return fmod(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

// Python's float % 360 (numpy remainder): fmod, then shift negatives up; can return exactly 360.0
template <typename T>
__device__ __forceinline__ T py_mod360(T a) {
  T r = Math<T>::fmod_(a, T(360));
  if (r < T(0)) r = Math<T>::add(r, T(360));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
gradient_field_kernel(const T* __restrict__ gauss, int L, int level, int H, int W, T* __restrict__ mag,
                      T* __restrict__ dir, uint8_t* __restrict__ obin) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= W) return;
  const T* g = gauss + ((size_t)b * L + level) * H * W;
  // true convolution with [-1, 0, 1], 'symm' border: gx = I[., x-1] - I[., x+1] with the edge sample repeated
  const int xl = x > 0 ? x - 1 : 0, xr = x < W - 1 ? x + 1 : W - 1;
  const int yu = y > 0 ? y - 1 : 0, yd = y < H - 1 ? y + 1 : H - 1;
  const T gx = g[(size_t)y * W + xl] - g[(size_t)y * W + xr];
  const T gy = g[(size_t)yu * W + x] - g[(size_t)yd * W + x];
  const T m = Math<T>::sqrt_(Math<T>::add(Math<T>::mul(gx, gx), Math<T>::mul(gy, gy)));
  // np.rad2deg(arctan2(gy, gx)) % 360
  const T deg = Math<T>::mul(Math<T>::atan2_(gy, gx), T(57.29577951308232087680));
  // |deg| <= 180, so Python's % 360 is the identity for deg >= 0 and one rounded + 360 otherwise
  const T d = deg < T(0) ? Math<T>::add(deg, T(360)) : deg;
  const size_t o = ((size_t)b * H + y) * W + x;
  mag[o] = m;
  dir[o] = d;
  // np.round(direction * 36 / 360): half to even, evaluated in the field's type on the stored direction
  if constexpr (sizeof(T) == 4) obin[o] = (uint8_t)(int)rintf(__fdiv_rn(__fmul_rn(d, 36.0f), 360.0f));
  else obin[o] = (uint8_t)(int)rint(__ddiv_rn(__dmul_rn((double)d, 36.0), 360.0));
}

// ------------------------------------------------------------------------------------------------
// Certified mode: one 32-bit word per pixel = magnitude | exact direction code
// ------------------------------------------------------------------------------------------------
// Both consumers of the direction only ever look at which side of a multiple of 5 degrees it lies on: the
// orientation histogram bins by np.round(direction * 36 / 360) (edges at 5, 15, ..; :140-142) and the descriptor by
// int(((direction - theta) % 360) * 8 / 360) with theta = 10 b + 5 (edges at theta + 45 j; :263-268).  So the gradient
// field is reduced to
//   bits  5..0   o   = the reference's 36-bin orientation index (0..36, half-to-even on exact ties)
//   bits  7..6   sub = 0: direction in [10 o - 5, 10 o), 1: in [10 o, 10 o + 5), 2: exactly 10 o + 5 (tie rounded down)
//   bits 31..8   the float32 magnitude without its sign, rounded to 16 mantissa bits (relative error 2^-17)
// i.e. the 5-degree sector is c = 2 o - 1 + sub.  Gradients are differences of the FLOAT64 level; the sector comes from
// a float32 arctangent of the octant-reduced angle in [0, 45] degrees (error < 2e-5 degrees) unless that angle lies
// within 1e-4 degrees of a multiple of 5: then -- and for vanishing gradients -- the reference's own float64
// expression decides (np.rad2deg(np.arctan2(gy, gx)) % 360), so that the codes are the float64 pipeline's codes.
__device__ __forceinline__ uint32_t direction_code_exact(double gy, double gx) {
  const double deg = __dmul_rn(atan2(gy, gx), 57.29577951308232087680);
  const double d = deg < 0.0 ? __dadd_rn(deg, 360.0) : deg;
  const int o = (int)rint(__ddiv_rn(__dmul_rn(d, 36.0), 360.0));
  const double centre = (double)(10 * o);
  uint32_t sub = d >= centre ? 1u : 0u;
  if (d == centre + 5.0) sub = 2u;
  return (sub << 6) | (uint32_t)o;
}

__device__ __forceinline__ uint32_t gradient_code(double gxd, double gyd) {
  const float x = (float)gxd, y = (float)gyd;
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float m;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(m) : "f"(fmaf(x, x, y * y)));
  const uint32_t mbits = ((__float_as_uint(m) + 64u) >> 7) << 8;
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(mx));
  const float a = mn * rcp;
  const float t = a * a;
  float p = -2.323167295e-01f;
  p = fmaf(p, t, 1.252680846e+00f);
  p = fmaf(p, t, -3.203576766e+00f);
  p = fmaf(p, t, 5.524598167e+00f);
  p = fmaf(p, t, -7.969067623e+00f);
  p = fmaf(p, t, 1.142854225e+01f);
  p = fmaf(p, t, -1.909660375e+01f);
  p = fmaf(p, t, 5.729574145e+01f);
  const float r = p * a;  // octant-reduced angle in [0, 45] degrees
  const float s = rintf(r * 0.2f);
  const bool unsure = !(mx >= 1e-30f) || fabsf(fmaf(s, -5.0f, r)) < 1e-4f;
  if (unsure) return mbits | direction_code_exact(gyd, gxd);
  int c = __float2int_rd(r * 0.2f);  // sector of the reduced angle, 0..8 (r is at least 1e-4 away from every edge)
  c = ay > ax ? 17 - c : c;          // reflections map 5-degree sectors onto 5-degree sectors
  c = x < 0.f ? 35 - c : c;
  c = y < 0.f ? 71 - c : c;
  const uint32_t o = (uint32_t)(c + 1) >> 1;
  return mbits | (((c & 1) ? 0u : 1u) << 6) | o;
}

// level = one float64 plane [B][n_planes][H][W] (plane index `plane_idx`); true convolution with [-1, 0, 1], 'symm' border
__global__ void __launch_bounds__(256)
gradient_codes_kernel(const double* __restrict__ levels, int n_planes, int plane_idx, int H, int W,
                      uint32_t* __restrict__ codes) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= W) return;
  const double* g = levels + ((size_t)b * n_planes + plane_idx) * H * W;
  const int xl = x > 0 ? x - 1 : 0, xr = x < W - 1 ? x + 1 : W - 1;
  const int yu = y > 0 ? y - 1 : 0, yd = y < H - 1 ? y + 1 : H - 1;
  const double gx = g[(size_t)y * W + xl] - g[(size_t)y * W + xr];
  const double gy = g[(size_t)yu * W + x] - g[(size_t)yd * W + x];
  codes[((size_t)b * H + y) * W + x] = gradient_code(gx, gy);
}

// four pixels per thread (W % 4 == 0, 16-byte aligned planes) and kCodeRows consecutive rows: the three rows a
// gradient needs roll through registers, so a row is loaded once per thread instead of three times (the kernel sat at
// 66 % of the L1 data-pipe limit with one row per thread)
static constexpr int kCodeRows = 8;

__global__ void __launch_bounds__(256)
gradient_codes4_kernel(const double* __restrict__ levels, int n_planes, int plane_idx, int H, int W,
                       uint32_t* __restrict__ codes) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y0 = blockIdx.y * kCodeRows;
  const int b = blockIdx.z;
  if (x >= W) return;
  const double* g = levels + ((size_t)b * n_planes + plane_idx) * H * W;
  uint32_t* out_p = codes + ((size_t)b * H + y0) * W + x;
  const int xl = x > 0 ? x - 1 : 0, xr = x + 4 < W ? x + 4 : W - 1;
  auto load4 = [&](int row, double2& a, double2& c) {
    const double* p = g + (size_t)row * W + x;
    a = *reinterpret_cast<const double2*>(p);
    c = *reinterpret_cast<const double2*>(p + 2);
  };
  double2 u0, u1, c0, c1, d0, d1;
  load4(y0 > 0 ? y0 - 1 : 0, u0, u1);
  load4(y0, c0, c1);
  const int y1 = min(y0 + kCodeRows, H);
#pragma unroll 2
  for (int y = y0; y < y1; ++y) {
    load4(y < H - 1 ? y + 1 : H - 1, d0, d1);
    const double* row = g + (size_t)y * W;
    const double left = row[xl], right = row[xr];
    uint4 out;
    out.x = gradient_code(left - c0.y, u0.x - d0.x);
    out.y = gradient_code(c0.x - c1.x, u0.y - d0.y);
    out.z = gradient_code(c0.y - c1.y, u1.x - d1.x);
    out.w = gradient_code(c1.x - right, u1.y - d1.y);
    *reinterpret_cast<uint4*>(out_p) = out;
    out_p += W;
    u0 = c0; u1 = c1; c0 = d0; c1 = d1;
  }
}

// 16-byte shared-memory vectors of the histogram type
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  using type = float4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float add_in_order(float s, float4 v) { return (((s + v.x) + v.y) + v.z) + v.w; }
};
template <> struct Vec16<double> {
  using type = double2;
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ double add_in_order(double s, double2 v) { return (s + v.x) + v.y; }
};

// ------------------------------------------------------------------------------------------------
static constexpr int kOriWarps = 4;
static constexpr int kOriBins = 36;
static constexpr int kOriPitch = 33;  // [bin][lane] with one pad column: the final column sums are conflict free

// LPK lanes share one keypoint (32 / LPK keypoints per warp): the per-lane fixed work (clearing and reducing the private
// histograms) is the same for every LPK, so small windows -- most keypoints sit in octave 0 with 15x15 or 21x21
// windows -- use 8 lanes and amortise it over four keypoints per warp.  Sums are in T: double in exact (f64) mode,
// float in f32 mode (whose parity bar is a tolerance), always in a fixed order (deterministic).
// MODE: kOriPlanes = magnitude plane + bin plane (float64 exact mode and the standalone entry point), kOriCoded =
// certified mode's magnitude | code words (gradient_codes_kernel).
// The weight table holds |dy| <= wry, |dx| <= wrx only (the window is clipped to the image, so a table larger than
// twice the image is never read): weight of offset (dy, dx) = weights[(dy + wry) * (2 wrx + 1) + dx + wrx].
// kOriCoded also certifies its float32 histogram: `tol` bounds the relative error of a bin against the float64 sum
// (magnitudes carry 2^-17, the float32 products and the running sums the rest); a keypoint with a bin within tol * max
// of the 0.8 * max cut is marked (bit 63) and appended to `recheck_list` for fc_orientation_recheck instead of being
// decided here (the list holds up to n_keypoints entries; recheck_count is zeroed by the launcher).
enum { kOriPlanes = 0, kOriCoded = 2 };
static constexpr unsigned long long kOriRecheckBit = 1ull << 63;

// `hist >= 0.8 * hist.max()` (APP/utils/keypoint_descriptor.py:165-167) with a float32 histogram.  numpy >= 2 (NEP 50, what
// this image has and the golden vectors were produced with): the Python float is weak, the product is float32(0.8) * max in
// float32.  numpy 1.26, which the reference's requirements.txt pins: the scalar product is float64 and is rounded to
// float32 for the comparison.  The two cuts differ by at most one float32 ulp (in ~20 % of the cases), i.e. only for a bin
// that sits on the cut; `legacy` selects the second form.  The certified float32 kernel needs no switch: whatever lies
// within its tolerance of the cut is decided by the float64 re-check, which has one.
__device__ __forceinline__ float orientation_cut(float mx, int legacy) {
  return legacy ? __double2float_rn(__dmul_rn(0.8, (double)mx)) : __fmul_rn(0.8f, mx);
}
struct FalseTag { static constexpr bool value = false; };
struct TrueTag { static constexpr bool value = true; };

template <typename T, int LPK, int NR, int MODE>
__global__ void __launch_bounds__(kOriWarps * 32)
orientation_kernel(const T* __restrict__ mag, const uint8_t* __restrict__ obin, int H, int W,
                   const int32_t* __restrict__ keypoints, int n_keypoints, const T* __restrict__ weights,
                   int radius, int wry, int wrx, float tol, unsigned long long* __restrict__ bin_mask,
                   int32_t* __restrict__ recheck_list, int32_t* __restrict__ recheck_count, int legacy_cut) {
  constexpr bool PACKED = MODE != kOriPlanes;
  constexpr int KPW = 32 / LPK;        // keypoints per warp
  // kOriPlanes: [keypoint][bin][lane of the keypoint], rows of LPK values (clearing and the final sums move 16 bytes at a time).
  // kOriCoded: [bin][lane of the warp] -- a lane's private bins always sit in the lane's own bank.  ncu had the kernel
  // at 96 % of the L1 data-pipe wavefront limit with 71 % of the wavefronts from shared memory, half of those bank
  // conflicts between the keypoints of a warp (their rows start 8 banks apart and the bin picks the rest).
  constexpr bool WIDE = PACKED && sizeof(T) == 4;
  constexpr int PITCH = WIDE ? 32 : LPK;
  constexpr int NB = (kOriBins + LPK - 1) / LPK;  // bins reduced by one lane
  constexpr int VEC = 16 / sizeof(T);  // values per 16-byte shared access
  constexpr int NROWS = (NR <= 2) ? 4 : 2;  // image rows whose samples are loaded before the first histogram update
  __shared__ __align__(16) T hist[kOriWarps][(kOriBins + 1) * 32];  // KPW * LPK = 32 values per bin either way
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPK, slot = lane / LPK;
  const int n = (blockIdx.x * kOriWarps + warp) * KPW + slot;
  const bool valid = n < n_keypoints;
  if ((blockIdx.x * kOriWarps + warp) * KPW >= n_keypoints) return;  // warp-uniform; no block-wide barriers below
  T* h = hist[warp] + (WIDE ? slot * LPK : slot * (kOriBins + 1) * LPK);  // bin b of lane `sub`: h[b * PITCH + sub]
  {
    using V = typename Vec16<T>::type;
    V* hv = reinterpret_cast<V*>(hist[warp]);  // the whole warp clears the warp's table
    for (int t = lane; t < (kOriBins + 1) * 32 / VEC; t += 32) hv[t] = Vec16<T>::zero();
  }
  __syncwarp();
  const int nn = valid ? n : 0;
  const int b = keypoints[3 * nn], ki = keypoints[3 * nn + 1], kj = keypoints[3 * nn + 2];
  const T* m = mag + (size_t)b * H * W;
  const uint8_t* ob = PACKED ? nullptr : obin + (size_t)b * H * W;
  // PACKED (float only): `mag` holds fc_gradient_codes' words = magnitude | direction code
  const uint32_t* pk = reinterpret_cast<const uint32_t*>(m);
  const int side = 2 * wrx + 1;
  // clip the window to the image (outside is zero padding = no contribution)
  const int r0 = max(ki - radius, 0), r1 = valid ? min(ki + radius, H - 1) : -1;
  const int c0 = max(kj - radius, 0), c1 = min(kj + radius, W - 1);
  // Row-wise walk: lane `sub` owns window columns c0+sub, c0+sub+LPK, ... of every row, so a sample costs two
  // gathers, one table load and one private shared-memory update -- no index division.  Two rows (2*NR samples
  // per lane) are loaded before the first dependent histogram update.
  // NR = column rounds that cover the window (template parameter, chosen by the launcher from the radius)
  const int woff = (wry - ki) * side + (wrx - kj);  // weight of image pixel (r, c): weights[woff + r*side + c]
  const int rounds = (c1 - c0 + LPK) / LPK;
  T* hme = h + sub;
  if (rounds <= NR) {
    // Per-lane base pointers advance by whole rows, so that every sample of the unrolled (NROWS x NR) block is a load
    // with an immediate offset under a predicate: ncu counted 33 instructions per sample when the addresses were formed
    // as r * W + c per sample (8 IMAD + 4 LEA of them), most of them 64-bit address arithmetic.
    const int cb = c0 + sub;
    bool cok[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) cok[u] = cb + u * LPK <= c1;
    const ptrdiff_t first = (ptrdiff_t)r0 * W + cb;
    const uint32_t* prow = pk + first;
    const T* mrow = m + first;
    const uint8_t* orow = PACKED ? nullptr : ob + first;
    const T* wrow = weights + (woff + r0 * side + cb);
    // kOriCoded: separable float32 weights, table = [wy: 2 wry + 1 | wx: 2 wrx + 1] with w(dy, dx) = wy[dy] * wx[dx]
    // (exp(-dy^2 / 2 sigma^2) / (2 pi sigma^2) and exp(-dx^2 / 2 sigma^2): two more float32 roundings per sample, covered
    // by `tol`).  The lane's column factors never change and a row needs one factor: one gather per sample instead of two.
    const float* wy_tab = reinterpret_cast<const float*>(weights) + (wry - ki);             // wy_tab[image row]
    const float* wx_tab = reinterpret_cast<const float*>(weights) + (2 * wry + 1) + (wrx - kj);  // wx_tab[image column]
    float cw[NR];
#pragma unroll
    for (int u = 0; u < NR; ++u) cw[u] = (PACKED && cok[u]) ? wx_tab[cb + u * LPK] : 0.f;
    if constexpr (WIDE) {
      // Certified float32 form.  A sample outside the window needs no dropped bin: its column factor is 0 and a masked
      // load returns the word 0 (magnitude 0), so it adds +0 to bin 0.  Whole groups of NROWS rows run without any row
      // test (ncu: predicates, default values and loop bookkeeping were 8 of 20 instructions per sample); one tail
      // group checks its rows.
      const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(hme);
      auto group = [&](auto check_rows, int rr) {
        constexpr bool CHECK = decltype(check_rows)::value;
        uint32_t p[NROWS][NR];
        float rw[NROWS];
#pragma unroll
        for (int v = 0; v < NROWS; ++v) {
          const bool row_ok = !CHECK || rr + v <= r1;
          rw[v] = row_ok ? wy_tab[rr + v] : 0.f;
#pragma unroll
          for (int u = 0; u < NR; ++u) p[v][u] = (row_ok && cok[u]) ? (prow + (ptrdiff_t)v * W)[u * LPK] : 0u;
        }
        prow += (ptrdiff_t)NROWS * W;
#pragma unroll
        for (int v = 0; v < NROWS; ++v)
#pragma unroll
          for (int u = 0; u < NR; ++u) {
            const uint32_t a = hbase + ((p[v][u] & 63u) << 7);  // 32 floats per bin row
            const float wv = __fmul_rn(__uint_as_float((p[v][u] & ~255u) >> 1), __fmul_rn(rw[v], cw[u]));
            float cur;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(a));
            cur = __fadd_rn(cur, wv);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(cur) : "memory");
          }
      };
      int rr = r0;
      for (; rr + NROWS - 1 <= r1; rr += NROWS) group(FalseTag(), rr);
      if (rr <= r1) group(TrueTag(), rr);
    } else
    for (int rr = r0; rr <= r1; rr += NROWS) {
      int bins[NROWS][NR];
      T ws[NROWS][NR];
#pragma unroll
      for (int v = 0; v < NROWS; ++v) {
        const bool row_ok = rr + v <= r1;
        const T* wv = wrow + v * side;
        float rw = 0.f;
        if (PACKED && row_ok) rw = wy_tab[rr + v];
#pragma unroll
        for (int u = 0; u < NR; ++u) {
          bins[v][u] = kOriBins;  // the dropped bin
          ws[v][u] = T(0);
          if (row_ok && cok[u]) {
            if constexpr (PACKED) {
              const uint32_t p = (prow + (ptrdiff_t)v * W)[u * LPK];
              bins[v][u] = (int)(p & 63u);
              ws[v][u] = (T)__fmul_rn(__uint_as_float((p & ~255u) >> 1), __fmul_rn(rw, cw[u]));
            } else {
              bins[v][u] = (orow + (ptrdiff_t)v * W)[u * LPK];
              ws[v][u] = Math<T>::mul((mrow + (ptrdiff_t)v * W)[u * LPK], wv[u * LPK]);
            }
          }
        }
      }
      prow += (ptrdiff_t)NROWS * W;
      mrow += (ptrdiff_t)NROWS * W;
      if (!PACKED) orow += (ptrdiff_t)NROWS * W;
      wrow += NROWS * side;
      if constexpr (PACKED && sizeof(T) == 4) {
        // private bin of this lane: one shift-add on the 32-bit shared address (the compiler otherwise rebuilds
        // (sub + bin * PITCH) * 4 + base for every sample); bin 36 is accumulated but never read
        const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(hme);
#pragma unroll
        for (int v = 0; v < NROWS; ++v)
#pragma unroll
          for (int u = 0; u < NR; ++u) {
            const uint32_t a = hbase + ((uint32_t)bins[v][u] << 7);  // WIDE layout: 32 floats per bin row
            float cur;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(a));
            cur = __fadd_rn(cur, (float)ws[v][u]);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(cur) : "memory");
          }
      } else {
#pragma unroll
        for (int v = 0; v < NROWS; ++v)
#pragma unroll
          for (int u = 0; u < NR; ++u) hme[bins[v][u] * PITCH] += ws[v][u];  // bin 36 is accumulated but never read
      }
    }
  } else {
    // window wider than the class was sized for (only reachable through the C ABI with a hand-picked radius)
    for (int rr = r0; rr <= r1; ++rr)
      for (int c = c0 + sub; c <= c1; c += LPK) {
        if constexpr (PACKED) {
          const uint32_t p = pk[rr * W + c];
          const float* wt = reinterpret_cast<const float*>(weights);
          hme[(p & 63u) * PITCH] += (T)__fmul_rn(__uint_as_float((p & ~255u) >> 1),
                                                 __fmul_rn(wt[wry - ki + rr], wt[2 * wry + 1 + wrx - kj + c]));
        } else {
          hme[ob[rr * W + c] * PITCH] += Math<T>::mul(m[rr * W + c], weights[woff + rr * side + c]);
        }
      }
  }
  __syncwarp();
  if constexpr (WIDE) {
    // Sums without a bank conflict: the eight lanes of a quarter warp read the eight 16-byte chunks of ONE bin row
    // (chunk c belongs to keypoint c / (LPK / 4)), lane group g = lane / 8 takes bins g, g + 4, ...; the chunks of a
    // keypoint are combined by shuffles, the maximum and the bin masks over the four lane groups.
    constexpr int CPS = LPK / 4;  // chunks per keypoint
    const int chunk = lane & 7, grp = lane >> 3;
    const int rslot = chunk / CPS;
    float f[9];
    float mx = 0.0f;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int bin = grp + 4 * q;  // < 36 for every q
      const float4 t = *reinterpret_cast<const float4*>(hist[warp] + bin * 32 + chunk * 4);
      float sacc = ((t.x + t.y) + t.z) + t.w;
#pragma unroll
      for (int o = 1; o < CPS; o <<= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      f[q] = sacc;
      mx = fmaxf(mx, sacc);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    const float cut = __fmul_rn(0.8f, mx);  // see the note on numpy's promotion rules below
    uint32_t lo = 0u, hi = 0u;
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      const int bin = grp + 4 * q;
      if (f[q] >= cut) {
        if (bin < 32) lo |= 1u << bin; else hi |= 1u << (bin - 32);
      }
      if (fabsf(f[q] - cut) <= tol * mx) hi |= 0x80000000u;
    }
    lo |= __shfl_xor_sync(0xffffffffu, lo, 8);
    hi |= __shfl_xor_sync(0xffffffffu, hi, 8);
    lo |= __shfl_xor_sync(0xffffffffu, lo, 16);
    hi |= __shfl_xor_sync(0xffffffffu, hi, 16);
    const int nr = (blockIdx.x * kOriWarps + warp) * KPW + rslot;
    if (grp == 0 && (chunk % CPS) == 0 && nr < n_keypoints) {
      bin_mask[nr] = (unsigned long long)lo | ((unsigned long long)hi << 32);
      if (hi & 0x80000000u) recheck_list[atomicAdd(recheck_count, 1)] = nr;
    }
    return;
  }
  // column sums: lane `sub` reduces bins sub, sub + LPK, ...; histogram entries are float32 in the reference (:158)
  float f[NB];
  float mx = 0.0f;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int bin = sub + q * LPK;
    f[q] = -1.0f;
    if (bin < kOriBins) {
      using V = typename Vec16<T>::type;
      const V* row = reinterpret_cast<const V*>(h + bin * PITCH);
      T sacc = T(0);
#pragma unroll
      for (int k = 0; k < LPK / VEC; ++k) sacc = Vec16<T>::add_in_order(sacc, row[k]);  // lane 0, 1, 2, ... of the keypoint
      f[q] = (float)sacc;
      mx = fmaxf(mx, f[q]);
    }
  }
#pragma unroll
  for (int o = LPK / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float cut = orientation_cut(mx, legacy_cut);
  uint32_t lo = 0u, hi = 0u;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int bin = sub + q * LPK;
    if (bin < kOriBins && f[q] >= cut) {
      if (bin < 32) lo |= 1u << bin; else hi |= 1u << (bin - 32);
    }
    if (MODE == kOriCoded && bin < kOriBins && fabsf(f[q] - cut) <= tol * mx) hi |= 0x80000000u;
  }
#pragma unroll
  for (int o = LPK / 2; o > 0; o >>= 1) {
    lo |= __shfl_xor_sync(0xffffffffu, lo, o);
    hi |= __shfl_xor_sync(0xffffffffu, hi, o);
  }
  if (sub == 0 && valid) {
    bin_mask[n] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    if (MODE == kOriCoded && (hi & 0x80000000u)) recheck_list[atomicAdd(recheck_count, 1)] = n;
  }
}

__global__ void popcount64_kernel(const unsigned long long* __restrict__ m, int n, int32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) counts[i] = __popcll(m[i]);
}

__global__ void orientation_emit_kernel(const int32_t* __restrict__ keypoints, const unsigned long long* __restrict__ m,
                                        const int32_t* __restrict__ offsets, int n, int32_t* __restrict__ oriented) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long bits = m[i];
  int pos = offsets[i];
  const int b = keypoints[3 * i], r = keypoints[3 * i + 1], c = keypoints[3 * i + 2];
  while (bits) {
    const int bin = __ffsll((long long)bits) - 1;
    bits &= bits - 1;
    reinterpret_cast<int4*>(oriented)[pos++] = make_int4(b, r, c, bin);
  }
}

// Certified mode: the keypoints orientation_kernel<.., kOriCoded> marked (bit 63) are decided here the way the float64
// pipeline decides them -- float64 magnitudes from the float64 level, float64 window weights, the exact bin codes,
// float64 sums rounded to float32 (:158-163), cut at float32(0.8) * max.  A small persistent grid walks the list the
// histogram kernel appended to, one CTA of 16 warps per keypoint (windows reach 155 x 155 samples: a single warp would
// need a millisecond for one of those); every lane owns a private histogram, sums are combined in a fixed order.
static constexpr int kRecheckWarps = 16;
static constexpr size_t kRecheckSmem = sizeof(double) * kRecheckWarps * (kOriBins + 1) * 33;

__global__ void __launch_bounds__(kRecheckWarps * 32)
orientation_recheck_kernel(const double* __restrict__ levels, int n_planes, int plane_idx, const uint32_t* __restrict__ codes,
                           int H, int W, const int32_t* __restrict__ keypoints, int n_keypoints,
                           const double* __restrict__ weights, int radius, int wry, int wrx,
                           unsigned long long* __restrict__ bin_mask, const int32_t* __restrict__ recheck_list,
                           const int32_t* __restrict__ recheck_count, int legacy_cut) {
  extern __shared__ __align__(16) double rc_hist[];  // [warp][bin 0..36][33]
  __shared__ float s_f[kOriBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* myh = rc_hist + (size_t)warp * (kOriBins + 1) * 33 + lane;
  const int n_list = min(*recheck_count, n_keypoints);
  for (int e = blockIdx.x; e < n_list; e += gridDim.x) {  // block-uniform
    const int n = recheck_list[e];
    for (int bn = 0; bn <= kOriBins; ++bn) myh[bn * 33] = 0.0;
    const int b = keypoints[3 * n], ki = keypoints[3 * n + 1], kj = keypoints[3 * n + 2];
    const double* g = levels + ((size_t)b * n_planes + plane_idx) * H * W;
    const uint32_t* pk = codes + (size_t)b * H * W;
    const int side = 2 * wrx + 1;
    const int r0 = max(ki - radius, 0), r1 = min(ki + radius, H - 1);
    const int c0 = max(kj - radius, 0), c1 = min(kj + radius, W - 1);
    const int ncol = c1 - c0 + 1, chunks = (ncol + 31) >> 5, items = chunks * (r1 - r0 + 1);
    const int woff = (wry - ki) * side + (wrx - kj);
    // a warp takes (row, 32-column chunk) items; the loads of four items are in flight before the first square root
    // (with one item at a time a 155 x 155 window is a chain of ~100 dependent global round trips per thread)
    constexpr int kUnroll = 4;
    for (int it0 = warp; it0 < items; it0 += kRecheckWarps * kUnroll) {
      double gx[kUnroll], gy[kUnroll], wv[kUnroll];
      uint32_t bn[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int it = it0 + u * kRecheckWarps;
        const int rr = it / chunks;
        const int r = r0 + rr, c = c0 + ((it - rr * chunks) << 5) + lane;
        gx[u] = gy[u] = wv[u] = 0.0;
        bn[u] = kOriBins;
        if (it < items && c <= c1) {
          const int ru = r > 0 ? r - 1 : 0, rd = r < H - 1 ? r + 1 : H - 1;
          const int cl = c > 0 ? c - 1 : 0, cr = c < W - 1 ? c + 1 : W - 1;
          gx[u] = g[(size_t)r * W + cl] - g[(size_t)r * W + cr];
          gy[u] = g[(size_t)ru * W + c] - g[(size_t)rd * W + c];
          wv[u] = weights[woff + r * side + c];
          bn[u] = pk[(size_t)r * W + c] & 63u;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const double m = sqrt(__dadd_rn(__dmul_rn(gx[u], gx[u]), __dmul_rn(gy[u], gy[u])));
        myh[bn[u] * 33] += __dmul_rn(m, wv[u]);  // bin 36 collects the masked lanes and is never read
      }
    }
    __syncthreads();
    // warp w sums bins w, w + 16, ..: lane l adds column l of the 16 private copies in warp order, then a fixed
    // shuffle tree over the lanes (deterministic)
    for (int bn = warp; bn < kOriBins; bn += kRecheckWarps) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < kRecheckWarps; ++w) a += rc_hist[((size_t)w * (kOriBins + 1) + bn) * 33 + lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) s_f[bn] = (float)a;
    }
    __syncthreads();
    if (warp == 0) {
      const float f0 = s_f[lane], f1 = lane + 32 < kOriBins ? s_f[lane + 32] : -1.0f;
      float mx = fmaxf(f0, f1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float cut = orientation_cut(mx, legacy_cut);
      const uint32_t lo = __ballot_sync(0xffffffffu, f0 >= cut);
      const uint32_t hi = __ballot_sync(0xffffffffu, lane + 32 < kOriBins && f1 >= cut);
      if (lane == 0) bin_mask[n] = (unsigned long long)lo | ((unsigned long long)hi << 32);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Final placement of the batched pipeline's outputs
// ------------------------------------------------------------------------------------------------
// The pipeline keeps its lists block-major: (octave, scale) blocks one after the other, each sorted by (image, row,
// col, bin).  The caller wants image-major.  From the two scans the pipeline already has -- row_offs (cumulative
// keypoints over the concatenated mask rows of all blocks) and ori_offs (cumulative oriented keypoints over all
// keypoints) -- this single-CTA kernel derives, per block i and image b, the list index OB(i, b) of the first oriented
// keypoint and its final row, plus the per-image totals.
struct LayoutBlocks {
  int n_blocks;
  int row_start[64];  // first row of block i in the concatenated row array
  int height[64];     // image rows per image in block i
};

__global__ void __launch_bounds__(256)
sift_layout_kernel(const int32_t* __restrict__ row_offs, const int32_t* __restrict__ ori_offs, const LayoutBlocks lb, int batch,
                   int2* __restrict__ remap, int32_t* __restrict__ image_counts) {
  extern __shared__ int s_start[];  // [batch + 1] first final row of every image
  auto first = [&](int i, int b) { return ori_offs[row_offs[lb.row_start[i] + b * lb.height[i]]]; };
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    int total = 0;
    for (int i = 0; i < lb.n_blocks; ++i) total += first(i, b + 1) - first(i, b);
    image_counts[b] = total;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < batch; ++b) { s_start[b] = run; run += image_counts[b]; }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    int run = s_start[b];
    for (int i = 0; i < lb.n_blocks; ++i) {
      const int ob = first(i, b);
      remap[i * batch + b] = make_int2(ob, run);
      run += first(i, b + 1) - ob;
    }
  }
}

// ------------------------------------------------------------------------------------------------
static constexpr int kDescWarps = 4;
static constexpr int kTrigStride = 2 + 32;

// Persistent: every CTA converts the host tables once (Gaussian mask, bin edges, the fixed-point column deltas as
// int32, cos / sin) into shared memory, then its warps walk the oriented keypoints with stride = warps in the grid.
// Input: separate magnitude / direction planes (float64 mode, and float32 planes of a caller's own stack).  The certified
// mode's code-word planes go through descriptor_coded_kernel below.
enum { kDescPlanes = 0, kDescCoded = 2 };

// Output placement: without `remap` descriptor n of this call goes to row n.  With it (the batched pipeline) the rows
// land directly in the caller's final order -- images one after the other, each image in the reference's emission order
// (octave, scale, row-major, bin): remap[image] = (index of the image's first oriented keypoint of this (octave, scale)
// block in the block-major list, its final row), see sift_layout_kernel; `points` then receives the matching keypoint
// records (int16 row, int16 col, uint8 octave, uint8 scale, uint8 orientation bin, 0) so that no sort or gather pass
// over the descriptors is needed afterwards.
struct DescPlacement {
  const int2* remap;  // [batch] or null
  uint2* points;      // final-order keypoint records or null
  int list_base;      // index of this call's first oriented keypoint in the block-major list
  int octave, scale;
};

template <typename OUT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void put(float* d, float a, float b, float c, float e) { *reinterpret_cast<float4*>(d) = make_float4(a, b, c, e); }
};
template <> struct Store4<double> {
  static __device__ __forceinline__ void put(double* d, double a, double b, double c, double e) {
    reinterpret_cast<double2*>(d)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(d)[1] = make_double2(c, e);
  }
};
template <> struct Store4<__half> {
  static __device__ __forceinline__ void put(__half* d, float a, float b, float c, float e) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, e);
    *reinterpret_cast<uint2*>(d) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
  }
};

template <typename T, typename OUT>
__global__ void __launch_bounds__(kDescWarps * 32)
descriptor_kernel(const T* __restrict__ mag, const T* __restrict__ dir, int H, int W, const int32_t* __restrict__ oriented,
                  int n_oriented, const double* __restrict__ gmask, const double* __restrict__ trig, OUT* __restrict__ desc,
                  const DescPlacement place) {
  using ACC = T;  // per-sample arithmetic: double in exact mode, float in f32 mode (tolerance-based parity)
  __shared__ int sXY[kDescWarps][32];       // x0[0..15], y0[0..15] per warp
  __shared__ ACC sHist[kDescWarps][9][32];  // per-lane private bins 0..8 (8 = dropped), [bin][lane]: conflict free
  __shared__ ACC sThr[8];                   // smallest rel with int(rel * 8 / 360.0) >= j + 1 (host computed, exact)
  __shared__ ACC sMask[256];
  __shared__ __align__(16) int sAd[36][16], sBd[36][16];  // rint(cos * x * 1024), rint(sin * x * 1024) per bin
  __shared__ double sCS[36][2];
  if (threadIdx.x < 8) sThr[threadIdx.x] = (ACC)trig[36 * kTrigStride + threadIdx.x];
  for (int t = threadIdx.x; t < 256; t += kDescWarps * 32) sMask[t] = (ACC)gmask[t];
  for (int t = threadIdx.x; t < 36 * 16; t += kDescWarps * 32) {
    const int bn = t >> 4, x = t & 15;
    sAd[bn][x] = (int)trig[bn * kTrigStride + 2 + x];
    sBd[bn][x] = (int)trig[bn * kTrigStride + 18 + x];
  }
  if (threadIdx.x < 72) sCS[threadIdx.x >> 1][threadIdx.x & 1] = trig[(threadIdx.x >> 1) * kTrigStride + (threadIdx.x & 1)];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = lane >> 1, half = lane & 1;
  const int cy = (cell >> 2) * 4 + half * 2, cx = (cell & 3) * 4;
  ACC* myh = &sHist[warp][0][lane];
  ACC wmask[2][4];  // this lane's 8 Gaussian weights never change
#pragma unroll
  for (int ry = 0; ry < 2; ++ry)
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) wmask[ry][rx] = sMask[(cy + ry) * 16 + cx + rx];
  const int warps_total = gridDim.x * kDescWarps;
  for (int n = blockIdx.x * kDescWarps + warp; n < n_oriented; n += warps_total) {  // no block-wide barrier inside
    const int4 kp = reinterpret_cast<const int4*>(oriented)[n];
    const int b = kp.x, ki = kp.y, kj = kp.z, bin = kp.w;
    const double theta = (double)(10 * bin + 5);  // (bin + 0.5) * 10 % 360, exact
    const double cs = sCS[bin][0], sn = sCS[bin][1];
    {
      // rotated_subimage: v_x = (c, s), v_y = (-s, c); s_x = j - c*7.5 - (-s)*7.5; s_y = i - s*7.5 - c*7.5
      // X0[y] = rint((M01*y + M02) * 1024) + 512 ; Y0[y] = rint((M11*y + M12) * 1024) + 512
      const double yy = (double)(lane & 15);
      double a;
      if (lane < 16) a = __dadd_rn(__dmul_rn(-sn, yy), __dsub_rn(__dsub_rn((double)kj, __dmul_rn(cs, 7.5)), __dmul_rn(-sn, 7.5)));
      else a = __dadd_rn(__dmul_rn(cs, yy), __dsub_rn(__dsub_rn((double)ki, __dmul_rn(sn, 7.5)), __dmul_rn(cs, 7.5)));
      sXY[warp][lane] = __double2int_rn(__dmul_rn(a, 1024.0)) + 512;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) myh[k * 32] = ACC(0);
    __syncwarp();
    const T* m = mag + (size_t)b * H * W;
    const T* d = dir + (size_t)b * H * W;
    const int4 ad4 = *reinterpret_cast<const int4*>(&sAd[bin][cx]);
    const int4 bd4 = *reinterpret_cast<const int4*>(&sBd[bin][cx]);
    const int ad[4] = {ad4.x, ad4.y, ad4.z, ad4.w}, bd[4] = {bd4.x, bd4.y, bd4.z, bd4.w};
    const ACC theta_a = (ACC)theta;
    // the rotated 16x16 window reaches at most 7.5 * sqrt(2) + 1 < 12 pixels from the keypoint
    const bool interior = (ki >= 12) && (ki < H - 12) && (kj >= 12) && (kj < W - 12);
    // all 8 gathers of the lane are issued before the first dependent histogram update
    T mv[2][4], dv[2][4];
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
      const int X0 = sXY[warp][cy + ry], Y0 = sXY[warp][16 + cy + ry];
#pragma unroll
      for (int rx = 0; rx < 4; ++rx) {
        const int X = (X0 + ad[rx]) >> 10;
        const int Y = (Y0 + bd[rx]) >> 10;
        const bool in = interior || (((unsigned)X < (unsigned)W) && ((unsigned)Y < (unsigned)H));
        const int o = in ? Y * W + X : 0;  // one image plane has fewer than 2^30 pixels (checked on the host)
        mv[ry][rx] = in ? m[o] : T(0);
        dv[ry][rx] = in ? d[o] : T(0);
      }
    }
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
#pragma unroll
      for (int rx = 0; rx < 4; ++rx) {
        const ACC wv = Math<ACC>::mul((ACC)mv[ry][rx], wmask[ry][rx]);
        // (dir - theta) % 360 with dir in [0, 360], theta in [5, 355]: |dir - theta| < 360, so Python's % is the
        // identity for non-negative values and one rounded +360 otherwise (can give exactly 360.0 -> bin 8, dropped)
        ACC rel = Math<ACC>::add((ACC)dv[ry][rx], -theta_a);
        if (rel < ACC(0)) rel = Math<ACC>::add(rel, ACC(360));
        int hb;
        if constexpr (sizeof(ACC) == 4) {
          // f32 mode: truncate rel / 45 and repair the one-off at the bin edges with a single compare
          hb = __float2int_rz(rel * 0.022222222f);
          hb += (rel >= 45.0f * (float)(hb + 1)) ? 1 : 0;
        } else {
          // bin = int(rel * 8 / 360.0): count of thresholds <= rel (binary search over the exact host table)
          hb = (rel >= sThr[3]) ? 4 : 0;
          hb += (rel >= sThr[hb + 1]) ? 2 : 0;
          hb += (rel >= sThr[hb]) ? 1 : 0;
          hb = (rel >= sThr[7]) ? 8 : hb;
        }
        myh[hb * 32] += wv;
      }
    }
    __syncwarp();
    // merge the two halves of the cell (rows 0-1 then rows 2-3), round to float32 like the reference's cell hist;
    // normalise / clip [2^-10, 0.2] / normalise (:281-287) in NT = double (exact mode) or float (f32 mode)
    using NT = T;
    NT f[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const ACC mine = myh[k * 32];
      const ACC other = __shfl_xor_sync(0xffffffffu, mine, 1);
      const ACC sum = half ? Math<ACC>::add(other, mine) : Math<ACC>::add(mine, other);
      // even lane keeps bins 0..3, odd lane bins 4..7 of the cell -> feature index cell*8 + half*4 + k
      if ((k >> 2) == half) f[k & 3] = (NT)(float)sum;
    }
    NT ss = f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const NT lo = NT(0.0009765625), hi = NT(0.2);  // np.finfo(np.float16).eps, 0.2
    NT s2 = NT(0);
    if constexpr (sizeof(NT) == 4) {
      // f32 mode: one reciprocal square root instead of a square root and four IEEE divisions; 0 * inf = NaN for an
      // all-zero window, like the reference's 0 / 0
      const float inv = rsqrtf(ss);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        NT v = f[k] * inv;
        if (v == v) v = v < lo ? lo : (v > hi ? hi : v);  // np.clip keeps NaN
        f[k] = v;
        s2 += v * v;
      }
    } else {
      // exact mode: IEEE square root and divisions as the straight-line sequences of fc_common.cuh (bit-identical to
      // sqrt / `/` for these operands, one refined reciprocal per divisor); an all-zero window gives NaN like the
      // reference's 0 / 0
      const NT nrm = sqrt_seq(ss);
      const NT rn = rcp_seq(nrm);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        NT v = div_seq(f[k], nrm, rn);
        if (v == v) v = v < lo ? lo : (v > hi ? hi : v);
        f[k] = v;
        s2 += v * v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    size_t row = (size_t)n;
    if (place.remap) {
      const int2 rm = place.remap[b];
      row = (size_t)(rm.y + (place.list_base + n - rm.x));
    }
    if (place.points && lane == 0)
      place.points[row] = make_uint2(((uint32_t)ki & 0xffffu) | ((uint32_t)kj << 16),
                                     (uint32_t)place.octave | ((uint32_t)place.scale << 8) | ((uint32_t)bin << 16));
    OUT* dst = desc + row * 128 + lane * 4;
    if constexpr (sizeof(NT) == 4) {
      const float inv2 = rsqrtf(s2);
      Store4<OUT>::put(dst, f[0] * inv2, f[1] * inv2, f[2] * inv2, f[3] * inv2);
    } else {
      const NT nrm2 = sqrt_seq(s2);
      const NT rn2 = rcp_seq(nrm2);
      Store4<OUT>::put(dst, div_seq(f[0], nrm2, rn2), div_seq(f[1], nrm2, rn2), div_seq(f[2], nrm2, rn2), div_seq(f[3], nrm2, rn2));
    }
    __syncwarp();  // the next keypoint rewrites sXY and the private bins
  }
}

// Certified-mode descriptors, the kernel the pipeline launches.  ncu showed the generic kernel above to be bound by
// instruction issue (68 % issue-active, ALU pipe 58 %, ~500 warp instructions per keypoint), not by its gathers: with
// two lanes per 4x4 cell every keypoint pays the per-warp fixed work (window origin, clearing and merging the private
// bins of two half cells, two 32-lane reductions) for only 8 samples per lane.  Here HALF a warp owns a keypoint and
// a lane owns a whole cell: 16 samples per lane, no merge step, 16-lane reductions, and the fixed work is shared by
// two keypoints per warp -- ~300 warp instructions per keypoint.  Arithmetic and summation order per cell (row-major
// over its 4x4 samples, float32) are otherwise those of descriptor_kernel<float, .>.
template <typename OUT>
__global__ void __launch_bounds__(kDescWarps * 32)
descriptor_coded_kernel(const uint32_t* __restrict__ codes, int H, int W, const int32_t* __restrict__ oriented, int n_oriented,
                        const double* __restrict__ gmask, const double* __restrict__ trig, OUT* __restrict__ desc,
                        const DescPlacement place) {
  __shared__ int sXY[kDescWarps][2][32];      // per half warp: x0[0..15], y0[0..15]
  __shared__ float sHist[kDescWarps][9][32];  // per-lane private bins 0..8 (8 = dropped), [bin][lane]: conflict free
  __shared__ float sMask[256];
  __shared__ __align__(16) int sAd[36][16], sBd[36][16];  // rint(cos * x * 1024), rint(sin * x * 1024) per bin
  __shared__ double sCS[36][2];
  for (int t = threadIdx.x; t < 256; t += kDescWarps * 32) sMask[t] = (float)gmask[t];
  for (int t = threadIdx.x; t < 36 * 16; t += kDescWarps * 32) {
    const int bn = t >> 4, x = t & 15;
    sAd[bn][x] = (int)trig[bn * kTrigStride + 2 + x];
    sBd[bn][x] = (int)trig[bn * kTrigStride + 18 + x];
  }
  if (threadIdx.x < 72) sCS[threadIdx.x >> 1][threadIdx.x & 1] = trig[(threadIdx.x >> 1) * kTrigStride + (threadIdx.x & 1)];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = lane >> 4, cell = lane & 15;  // half warp = keypoint slot, lane of it = cell (row-major 4x4)
  const int cy = (cell >> 2) * 4, cx = (cell & 3) * 4;
  float* myh = &sHist[warp][0][lane];
  float wmask[4][4];  // this lane's 16 Gaussian weights never change
#pragma unroll
  for (int ry = 0; ry < 4; ++ry)
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) wmask[ry][rx] = sMask[(cy + ry) * 16 + cx + rx];
  const int* xy = sXY[warp][hw];
  const int pairs_total = gridDim.x * kDescWarps;
  for (int pair = blockIdx.x * kDescWarps + warp; 2 * pair < n_oriented; pair += pairs_total) {  // no block-wide barrier inside
    const int n = 2 * pair + hw;
    const bool valid = n < n_oriented;
    const int4 kp = reinterpret_cast<const int4*>(oriented)[valid ? n : 2 * pair];
    const int b = kp.x, ki = kp.y, kj = kp.z, bin = kp.w;
    const double cs = sCS[bin][0], sn = sCS[bin][1];
    {
      // rotated_subimage: v_x = (c, s), v_y = (-s, c); s_x = j - c*7.5 - (-s)*7.5; s_y = i - s*7.5 - c*7.5
      // X0[y] = rint((M01*y + M02) * 1024) + 512 ; Y0[y] = rint((M11*y + M12) * 1024) + 512   (lane `cell` does row y = cell)
      const double yy = (double)cell;
      const double ax = __dadd_rn(__dmul_rn(-sn, yy), __dsub_rn(__dsub_rn((double)kj, __dmul_rn(cs, 7.5)), __dmul_rn(-sn, 7.5)));
      const double ay = __dadd_rn(__dmul_rn(cs, yy), __dsub_rn(__dsub_rn((double)ki, __dmul_rn(sn, 7.5)), __dmul_rn(cs, 7.5)));
      sXY[warp][hw][cell] = __double2int_rn(__dmul_rn(ax, 1024.0)) + 512;
      sXY[warp][hw][16 + cell] = __double2int_rn(__dmul_rn(ay, 1024.0)) + 512;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) myh[k * 32] = 0.f;
    __syncwarp();
    const uint32_t* m = codes + (size_t)b * H * W;
    const int4 ad4 = *reinterpret_cast<const int4*>(&sAd[bin][cx]);
    const int4 bd4 = *reinterpret_cast<const int4*>(&sBd[bin][cx]);
    const int ad[4] = {ad4.x, ad4.y, ad4.z, ad4.w}, bd[4] = {bd4.x, bd4.y, bd4.z, bd4.w};
    // the rotated 16x16 window reaches at most 7.5 * sqrt(2) + 1 < 12 pixels from the keypoint: when both keypoints of
    // the warp are that far inside the image (nearly always) no sample needs a bounds test
    const bool interior = (ki >= 12) && (ki < H - 12) && (kj >= 12) && (kj < W - 12);
    uint32_t pv[4][4];  // all 16 gathers of the lane are issued before the first dependent histogram update
    if (__all_sync(0xffffffffu, interior)) {
#pragma unroll
      for (int ry = 0; ry < 4; ++ry) {
        const int X0 = xy[cy + ry], Y0 = xy[16 + cy + ry];
#pragma unroll
        for (int rx = 0; rx < 4; ++rx) pv[ry][rx] = m[((Y0 + bd[rx]) >> 10) * W + ((X0 + ad[rx]) >> 10)];
      }
    } else {
#pragma unroll
      for (int ry = 0; ry < 4; ++ry) {
        const int X0 = xy[cy + ry], Y0 = xy[16 + cy + ry];
#pragma unroll
        for (int rx = 0; rx < 4; ++rx) {
          const int X = (X0 + ad[rx]) >> 10;
          const int Y = (Y0 + bd[rx]) >> 10;
          const bool in = ((unsigned)X < (unsigned)W) && ((unsigned)Y < (unsigned)H);
          pv[ry][rx] = in ? m[Y * W + X] : 0u;  // one image plane has fewer than 2^30 pixels (checked on the host)
        }
      }
    }
    // sector of (direction - theta) = (2 o + sub) - (2 bin + 2) mod 72 and bin = sector / 9.  72 = 8 * 9, so with the
    // offset +72 folded into the constant (sc in [0, 144]) the wrap-around is the `& 7` of the quotient:
    // ((sc * 57) >> 9) & 7 == (sc mod 72) / 9 for every sc < 200; shifted by 7 it is the byte offset of the lane's bin.
    const int cb = 70 - 2 * bin;
    const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(myh);
#pragma unroll
    for (int ry = 0; ry < 4; ++ry) {
#pragma unroll
      for (int rx = 0; rx < 4; ++rx) {
        const uint32_t p = pv[ry][rx];
        const float wv = __fmul_rn(__uint_as_float((p & ~255u) >> 1), wmask[ry][rx]);
        const uint32_t sc = 2u * (p & 63u) + ((p >> 6) & 3u) + (uint32_t)cb;
        const uint32_t a = hbase + (((sc * 57u) >> 2) & 0x380u);
        float cur;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cur) : "r"(a));
        cur = __fadd_rn(cur, wv);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(cur) : "memory");
      }
    }
    // normalise / clip [2^-10, 0.2] / normalise (:281-287); sums over the 16 lanes of the half warp
    float f[8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = myh[k * 32];
      ss = fmaf(f[k], f[k], ss);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float lo = 0.0009765625f, hi = 0.2f;  // np.finfo(np.float16).eps, 0.2
    const float inv = rsqrtf(ss);               // 0 * inf = NaN for an all-zero window, like the reference's 0 / 0
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = f[k] * inv;
      if (v == v) v = v < lo ? lo : (v > hi ? hi : v);  // np.clip keeps NaN
      f[k] = v;
      s2 = fmaf(v, v, s2);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (valid) {
      size_t row = (size_t)n;
      if (place.remap) {
        const int2 rm = place.remap[b];
        row = (size_t)(rm.y + (place.list_base + n - rm.x));
      }
      if (place.points && cell == 0)
        place.points[row] = make_uint2(((uint32_t)ki & 0xffffu) | ((uint32_t)kj << 16),
                                       (uint32_t)place.octave | ((uint32_t)place.scale << 8) | ((uint32_t)bin << 16));
      const float inv2 = rsqrtf(s2);
      OUT* dst = desc + row * 128 + cell * 8;
      Store4<OUT>::put(dst, f[0] * inv2, f[1] * inv2, f[2] * inv2, f[3] * inv2);
      Store4<OUT>::put(dst + 4, f[4] * inv2, f[5] * inv2, f[6] * inv2, f[7] * inv2);
    }
    __syncwarp();  // the next pair rewrites sXY and the private bins
  }
}

// rotated_subimage (APP/utils/keypoint_descriptor.py:175-212) as a standalone call: any theta, any size
// up to 64x64; cos/sin are computed by the caller with the reference's own libm calls.
__global__ void rotated_window_kernel(const double* __restrict__ image, int H, int W, double cx, double cy, double cs,
                                      double sn, int width, int height, double* __restrict__ out) {
  const int x = threadIdx.x, y = threadIdx.y;
  if (x >= width || y >= height) return;
  const double sx = __dsub_rn(__dsub_rn(cx, __dmul_rn(cs, (width - 1) / 2.0)), __dmul_rn(-sn, (height - 1) / 2.0));
  const double sy = __dsub_rn(__dsub_rn(cy, __dmul_rn(sn, (width - 1) / 2.0)), __dmul_rn(cs, (height - 1) / 2.0));
  const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(-sn, (double)y), sx), 1024.0)) + 512;
  const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(cs, (double)y), sy), 1024.0)) + 512;
  const int X = (X0 + __double2int_rn(__dmul_rn(__dmul_rn(cs, (double)x), 1024.0))) >> 10;
  const int Y = (Y0 + __double2int_rn(__dmul_rn(__dmul_rn(sn, (double)x), 1024.0))) >> 10;
  out[y * width + x] = (X >= 0 && X < W && Y >= 0 && Y < H) ? image[(size_t)Y * W + X] : 0.0;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_rotated_window(const double* image, int height, int width, double center_x, double center_y,
                                 double cos_theta, double sin_theta, int out_width, int out_height, double* out, void* stream) {
  if (!image || !out || height <= 0 || width <= 0) return FC_ERR_BAD_ARG;
  if (out_width < 1 || out_height < 1 || out_width * out_height > 1024) return FC_ERR_UNSUPPORTED;
  dim3 block(out_width, out_height);
  rotated_window_kernel<<<1, block, 0, (cudaStream_t)stream>>>(image, height, width, center_x, center_y, cos_theta,
                                                              sin_theta, out_width, out_height, out);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_gradient_field(const void* gauss, int dtype, int batch, int n_levels, int level, int height, int width,
                                 void* magnitude, void* direction, uint8_t* orientation_bin, void* stream) {
  if (!gauss || !magnitude || !direction || !orientation_bin || batch <= 0 || height <= 0 || width <= 0) return FC_ERR_BAD_ARG;
  if (level < 0 || level >= n_levels) return FC_ERR_BAD_ARG;
  dim3 grid(div_up(width, 256), height, batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FC_F32)
    gradient_field_kernel<float><<<grid, 256, 0, st>>>((const float*)gauss, n_levels, level, height, width, (float*)magnitude,
                                                       (float*)direction, orientation_bin);
  else if (dtype == FC_F64)
    gradient_field_kernel<double><<<grid, 256, 0, st>>>((const double*)gauss, n_levels, level, height, width,
                                                        (double*)magnitude, (double*)direction, orientation_bin);
  else
    return FC_ERR_BAD_ARG;
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

static int launch_orientation(const void* magnitude, const uint8_t* orientation_bin, int dtype, int mode, int batch,
                              int height, int width, const int32_t* keypoints, int n_keypoints, const void* weights,
                              int radius, int table_ry, int table_rx, float tol, uint64_t* bin_mask, int32_t* recheck_list,
                              int32_t* recheck_count, int legacy_cut, void* stream) {
  if (mode == kOriCoded && (!recheck_list || !recheck_count)) return FC_ERR_BAD_ARG;
  if (mode == kOriCoded) FC_CUDA_TRY(cudaMemsetAsync(recheck_count, 0, sizeof(int32_t), (cudaStream_t)stream));
  if (!magnitude || (mode == kOriPlanes && !orientation_bin) || !weights || batch <= 0 || height <= 0 || width <= 0 || radius < 0)
    return FC_ERR_BAD_ARG;
  // the table must cover every offset of the image-clipped window
  if (table_ry < (radius < height - 1 ? radius : height - 1) || table_rx < (radius < width - 1 ? radius : width - 1))
    return FC_ERR_BAD_ARG;
  if (n_keypoints == 0) return FC_OK;
  if (!keypoints || !bin_mask || n_keypoints < 0) return FC_ERR_BAD_ARG;
  if ((int64_t)height * width > 0x7fffffffLL) return FC_ERR_UNSUPPORTED;  // kernels index one image plane with int32
  if ((int64_t)(2 * table_ry + 1) * (2 * table_rx + 1) > 0x7fffffffLL) return FC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype != FC_F32 && dtype != FC_F64) return FC_ERR_BAD_ARG;
  // lanes per keypoint by window size: (2r+1)^2 <= 441 samples -> 8, <= 961 -> 16, else a whole warp; the number
  // of column rounds per row is a template parameter so that the row loop is fully unrolled and predicated
  const int lpk = radius <= 10 ? 8 : (radius <= 15 ? 16 : 32);
  const int nr = div_up(2 * radius + 1, lpk);
#define FC_ORI(T, LPK, NR, MD)                                                                                      \
  orientation_kernel<T, LPK, NR, MD><<<div_up(n_keypoints, kOriWarps * (32 / LPK)), kOriWarps * 32, 0, st>>>(        \
      (const T*)magnitude, orientation_bin, height, width, keypoints, n_keypoints, (const T*)weights, radius,      \
      table_ry, table_rx, tol, (unsigned long long*)bin_mask, recheck_list, recheck_count, legacy_cut)
#define FC_ORI_T(T, MD)                                              \
  do {                                                               \
    if (lpk == 8) {                                                  \
      if (nr <= 1) FC_ORI(T, 8, 1, MD); else if (nr == 2) FC_ORI(T, 8, 2, MD); else FC_ORI(T, 8, 3, MD); \
    } else if (lpk == 16) {                                          \
      FC_ORI(T, 16, 2, MD);                                          \
    } else {                                                         \
      if (nr <= 2) FC_ORI(T, 32, 2, MD); else if (nr == 3) FC_ORI(T, 32, 3, MD); else if (nr == 4) FC_ORI(T, 32, 4, MD); else FC_ORI(T, 32, 5, MD); \
    }                                                                \
  } while (0)
  if (mode == kOriCoded) FC_ORI_T(float, kOriCoded);
  else if (dtype == FC_F32) FC_ORI_T(float, kOriPlanes);
  else FC_ORI_T(double, kOriPlanes);
#undef FC_ORI_T
#undef FC_ORI
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_orientation_bins(const void* magnitude, const uint8_t* orientation_bin, int dtype, int batch, int height,
                                   int width, const int32_t* keypoints, int n_keypoints, const void* weights, int radius,
                                   int table_ry, int table_rx, uint64_t* bin_mask, int numpy_legacy_promotion, void* stream) {
  return launch_orientation(magnitude, orientation_bin, dtype, kOriPlanes, batch, height, width, keypoints, n_keypoints,
                            weights, radius, table_ry, table_rx, 0.f, bin_mask, nullptr, nullptr, numpy_legacy_promotion != 0,
                            stream);
}

extern "C" int fc_orientation_bins_coded(const uint32_t* codes, int batch, int height, int width, const int32_t* keypoints,
                                         int n_keypoints, const float* weights, int radius, int table_ry, int table_rx,
                                         double tolerance, uint64_t* bin_mask, int32_t* recheck_list, int32_t* recheck_count,
                                         void* stream) {
  if (!(tolerance >= 0.0)) return FC_ERR_BAD_ARG;
  return launch_orientation(codes, nullptr, FC_F32, kOriCoded, batch, height, width, keypoints, n_keypoints, weights, radius,
                            table_ry, table_rx, (float)tolerance, bin_mask, recheck_list, recheck_count, 0, stream);
}

extern "C" int fc_orientation_recheck(const double* levels, int n_planes, int plane, const uint32_t* codes, int batch,
                                      int height, int width, const int32_t* keypoints, int n_keypoints, const double* weights,
                                      int radius, int table_ry, int table_rx, uint64_t* bin_mask, const int32_t* recheck_list,
                                      const int32_t* recheck_count, int numpy_legacy_promotion, void* stream) {
  if (!recheck_list || !recheck_count) return FC_ERR_BAD_ARG;
  if (!levels || !codes || !weights || batch <= 0 || height <= 0 || width <= 0 || radius < 0 || plane < 0 || plane >= n_planes)
    return FC_ERR_BAD_ARG;
  if (table_ry < (radius < height - 1 ? radius : height - 1) || table_rx < (radius < width - 1 ? radius : width - 1))
    return FC_ERR_BAD_ARG;
  if (n_keypoints == 0) return FC_OK;
  if (!keypoints || !bin_mask || n_keypoints < 0) return FC_ERR_BAD_ARG;
  if ((int64_t)height * width > 0x7fffffffLL) return FC_ERR_UNSUPPORTED;
  FC_CUDA_TRY(cudaFuncSetAttribute(orientation_recheck_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRecheckSmem));
  orientation_recheck_kernel<<<n_keypoints < 148 * 2 ? n_keypoints : 148 * 2, kRecheckWarps * 32, kRecheckSmem, (cudaStream_t)stream>>>(
      levels, n_planes, plane, codes, height, width, keypoints, n_keypoints, weights, radius, table_ry, table_rx,
      (unsigned long long*)bin_mask, recheck_list, recheck_count, numpy_legacy_promotion != 0);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_gradient_codes(const double* levels, int batch, int n_planes, int plane, int height, int width,
                                 uint32_t* codes, void* stream) {
  if (!levels || !codes || batch <= 0 || height <= 0 || width <= 0 || plane < 0 || plane >= n_planes) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_ok = (width % 4 == 0) && (((reinterpret_cast<uintptr_t>(levels) | reinterpret_cast<uintptr_t>(codes)) & 15) == 0);
  if (vec_ok)
    gradient_codes4_kernel<<<dim3(div_up(width, 1024), div_up(height, kCodeRows), batch), 256, 0, st>>>(levels, n_planes, plane, height, width, codes);
  else
    gradient_codes_kernel<<<dim3(div_up(width, 256), height, batch), 256, 0, st>>>(levels, n_planes, plane, height, width, codes);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_orientation_scan(const uint64_t* bin_mask, int n_keypoints, int32_t* offsets, int32_t* scratch, void* stream) {
  if (!bin_mask || !offsets || !scratch || n_keypoints <= 0) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  popcount64_kernel<<<div_up(n_keypoints, 256), 256, 0, st>>>((const unsigned long long*)bin_mask, n_keypoints, scratch);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return exclusive_scan_i32(scratch, n_keypoints, offsets, scratch + n_keypoints, st);
}

extern "C" int fc_orientation_emit(const int32_t* keypoints, const uint64_t* bin_mask, const int32_t* offsets, int n_keypoints,
                                   int32_t* oriented, void* stream) {
  if (!keypoints || !bin_mask || !offsets || !oriented || n_keypoints <= 0) return FC_ERR_BAD_ARG;
  if (reinterpret_cast<uintptr_t>(oriented) & 15) return FC_ERR_BAD_ARG;
  orientation_emit_kernel<<<div_up(n_keypoints, 256), 256, 0, (cudaStream_t)stream>>>(
      keypoints, (const unsigned long long*)bin_mask, offsets, n_keypoints, oriented);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

static int launch_descriptors(const void* magnitude, const void* direction, int layout, int dtype, int batch, int height,
                              int width, const int32_t* oriented, int n_oriented, const double* gmask, const double* trig,
                              void* descriptors, int out_dtype, const fc_desc_placement* placement, void* stream) {
  DescPlacement place;
  place.remap = placement ? reinterpret_cast<const int2*>(placement->remap) : nullptr;
  place.points = placement ? reinterpret_cast<uint2*>(placement->points) : nullptr;
  place.list_base = placement ? placement->list_base : 0;
  place.octave = placement ? placement->octave : 0;
  place.scale = placement ? placement->scale : 0;
  if (placement && ((reinterpret_cast<uintptr_t>(placement->remap) | reinterpret_cast<uintptr_t>(placement->points)) & 7))
    return FC_ERR_BAD_ARG;
  const bool coded = layout == kDescCoded;
  if (!magnitude || (layout == kDescPlanes && !direction) || !gmask || !trig || batch <= 0 || height <= 0 || width <= 0) return FC_ERR_BAD_ARG;
  if (n_oriented == 0) return FC_OK;
  if (!oriented || !descriptors || n_oriented < 0) return FC_ERR_BAD_ARG;
  if ((int64_t)height * width > 0x3fffffffLL) return FC_ERR_UNSUPPORTED;  // kernels index one image plane with int32
  if (reinterpret_cast<uintptr_t>(oriented) & 15) return FC_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // persistent grid: as many CTAs as are resident at once (occupancy query per instantiation), fewer when there is less work
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    FC_CUDA_TRY(cudaGetDevice(&dev));
    FC_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
#define FC_LAUNCH(T, OUT)                                                                                       \
  do {                                                                                                            \
    static int per_sm = 0;                                                                                        \
    if (per_sm == 0)                                                                                              \
      FC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, descriptor_kernel<T, OUT>, kDescWarps * 32, 0)); \
    const int want = div_up(n_oriented, kDescWarps);                                                              \
    const int blocks = want < n_sm * per_sm ? want : n_sm * per_sm;                                               \
    descriptor_kernel<T, OUT><<<blocks, kDescWarps * 32, 0, st>>>((const T*)magnitude, (const T*)direction, height, \
                                                                     width, oriented, n_oriented, gmask, trig, (OUT*)descriptors, place); \
  } while (0)
#define FC_LAUNCH_CODED(OUT)                                                                                        \
  do {                                                                                                            \
    static int per_sm = 0;                                                                                        \
    if (per_sm == 0)                                                                                              \
      FC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, descriptor_coded_kernel<OUT>, kDescWarps * 32, 0)); \
    const int want = div_up(div_up(n_oriented, 2), kDescWarps);                                                   \
    const int blocks = want < n_sm * per_sm ? want : n_sm * per_sm;                                               \
    descriptor_coded_kernel<OUT><<<blocks, kDescWarps * 32, 0, st>>>((const uint32_t*)magnitude, height, width, oriented, \
                                                                    n_oriented, gmask, trig, (OUT*)descriptors, place); \
  } while (0)
  if (coded && out_dtype == FC_F32) FC_LAUNCH_CODED(float);
  else if (coded && out_dtype == FC_F64) FC_LAUNCH_CODED(double);
  else if (coded && out_dtype == FC_F16) FC_LAUNCH_CODED(__half);
  else if (coded) return FC_ERR_BAD_ARG;
  else if (dtype == FC_F32 && out_dtype == FC_F32) FC_LAUNCH(float, float);
  else if (dtype == FC_F32 && out_dtype == FC_F64) FC_LAUNCH(float, double);
  else if (dtype == FC_F64 && out_dtype == FC_F32) FC_LAUNCH(double, float);
  else if (dtype == FC_F64 && out_dtype == FC_F64) FC_LAUNCH(double, double);
  else return FC_ERR_BAD_ARG;
#undef FC_LAUNCH
#undef FC_LAUNCH_CODED
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_descriptors(const void* magnitude, const void* direction, int dtype, int batch, int height, int width,
                              const int32_t* oriented, int n_oriented, const double* gmask, const double* trig,
                              void* descriptors, int out_dtype, const fc_desc_placement* placement, void* stream) {
  return launch_descriptors(magnitude, direction, kDescPlanes, dtype, batch, height, width, oriented, n_oriented, gmask, trig,
                            descriptors, out_dtype, placement, stream);
}

extern "C" int fc_descriptors_coded(const uint32_t* codes, int batch, int height, int width, const int32_t* oriented,
                                    int n_oriented, const double* gmask, const double* trig, void* descriptors, int out_dtype,
                                    const fc_desc_placement* placement, void* stream) {
  return launch_descriptors(codes, nullptr, kDescCoded, FC_F32, batch, height, width, oriented, n_oriented, gmask, trig,
                            descriptors, out_dtype, placement, stream);
}

extern "C" int fc_sift_layout(const int32_t* row_offsets, const int32_t* oriented_offsets, const int* block_row_start_host,
                              const int* block_height_host, int n_blocks, int batch, int32_t* remap, int32_t* image_counts,
                              void* stream) {
  if (!row_offsets || !oriented_offsets || !block_row_start_host || !block_height_host || !remap || !image_counts)
    return FC_ERR_BAD_ARG;
  if (n_blocks < 1 || n_blocks > 64 || batch < 1 || batch > 8192) return FC_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(remap) & 7) return FC_ERR_BAD_ARG;
  LayoutBlocks lb;
  lb.n_blocks = n_blocks;
  for (int i = 0; i < 64; ++i) {
    lb.row_start[i] = i < n_blocks ? block_row_start_host[i] : 0;
    lb.height[i] = i < n_blocks ? block_height_host[i] : 0;
  }
  sift_layout_kernel<<<1, 256, (batch + 1) * sizeof(int), (cudaStream_t)stream>>>(
      row_offsets, oriented_offsets, lb, batch, reinterpret_cast<int2*>(remap), image_counts);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
