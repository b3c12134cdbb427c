// Opt-in variants of the corner response (BASELINE.json north_star: Sobel gradients, Gaussian window, 3x3 NMS) --
// everything the REFERENCE does not do and that therefore defaults OFF (SURVEY.md trap 1: the app uses np.gradient or
// K_X, a 5x5 box window, and no non-maximum suppression).  One shared-memory-tiled kernel: gradient tile -> weighted
// structure-tensor sums -> response tile -> (optional) 3x3 NMS + threshold mask.
//
// Each component keeps the border rule of the library call it stands for:
//   FC_GRAD_CENTRAL   np.gradient (one-sided at the border), APP/FeatureCraft_Backend.py:336
//   FC_GRAD_KX        zero-padded correlation with K_X / K_X^T, APP/FeatureCraft_Backend.py:405-421
//   FC_GRAD_SOBEL     cv2.Sobel(gray, CV_64F, ., ., ksize=3), BORDER_REFLECT_101  (moravec notebook cell, :107-135 uses ksize 5)
//   FC_WINDOW_BOX     ones window, products outside the image are zero (convolve2d_optimized's zero padding)
//   FC_WINDOW_GAUSS   cv2.GaussianBlur(product, (n, n), sigma), BORDER_REFLECT_101: products at mirrored positions
// With FC_GRAD_CENTRAL / FC_GRAD_KX + FC_WINDOW_BOX the sums are exact integers (or quarters) in float64, so the result
// equals the fast reference kernels of fc_corner.cu bit for bit (tests/test_corner_extras_gpu.py).
#include "fc_common.cuh"

namespace fc {

static constexpr int kExTileX = 32, kExTileY = 16, kExMaxR = 15;

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n - 2;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

template <int GRAD>
__device__ __forceinline__ double2 gradient_at(const uint8_t* __restrict__ g, int H, int W, int y, int x) {
  auto px = [&](int yy, int xx) { return (double)g[(size_t)yy * W + xx]; };
  if constexpr (GRAD == FC_GRAD_CENTRAL) {
    // np.gradient returns (d/drow, d/dcol); the app names them Ix, Iy (APP/FeatureCraft_Backend.py:336)
    const double d0 = y == 0 ? px(1, x) - px(0, x) : (y == H - 1 ? px(H - 1, x) - px(H - 2, x) : (px(y + 1, x) - px(y - 1, x)) * 0.5);
    const double d1 = x == 0 ? px(y, 1) - px(y, 0) : (x == W - 1 ? px(y, W - 1) - px(y, W - 2) : (px(y, x + 1) - px(y, x - 1)) * 0.5);
    return make_double2(d0, d1);
  } else if constexpr (GRAD == FC_GRAD_KX) {
    const int kx[5][5] = {{-1, -2, 0, 2, 1}, {-2, -3, 0, 3, 2}, {-3, -5, 0, 5, 3}, {-2, -3, 0, 3, 2}, {-1, -2, 0, 2, 1}};
    int sx = 0, sy = 0;
#pragma unroll
    for (int a = -2; a <= 2; ++a)
#pragma unroll
      for (int b = -2; b <= 2; ++b) {
        const int yy = y + a, xx = x + b;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const int v = g[(size_t)yy * W + xx];
          sx += kx[a + 2][b + 2] * v;
          sy += kx[b + 2][a + 2] * v;
        }
      }
    return make_double2((double)sx, (double)sy);
  } else {
    const int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H), xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    const double gx = (px(ym, xp) - px(ym, xm)) + 2.0 * (px(y, xp) - px(y, xm)) + (px(yp, xp) - px(yp, xm));
    const double gy = (px(yp, xm) - px(ym, xm)) + 2.0 * (px(yp, x) - px(ym, x)) + (px(yp, xp) - px(ym, xp));
    return make_double2(gx, gy);
  }
}

struct WindowTaps {
  double w[2 * kExMaxR + 1];
};

template <int GRAD>
__global__ void __launch_bounds__(256)
corner_ex_kernel(const uint8_t* __restrict__ gray, int H, int W, int R, int window, const __grid_constant__ WindowTaps taps,
                 int response, double k, int nms, double threshold, double* __restrict__ out, uint32_t* __restrict__ mask,
                 int mask_pitch) {
  extern __shared__ __align__(16) unsigned char ex_smem[];
  const int E = nms ? 1 : 0;                              // response halo for the 3x3 comparison
  const int gw = kExTileX + 2 * (E + R), gh = kExTileY + 2 * (E + R);
  const int rw = kExTileX + 2 * E, rh = kExTileY + 2 * E;
  double2* sG = reinterpret_cast<double2*>(ex_smem);      // [gh][gw] gradients of the in-image positions
  double* sR = reinterpret_cast<double*>(sG + gw * gh);   // [rh][rw] responses
  const int b = blockIdx.z;
  const uint8_t* img = gray + (size_t)b * H * W;
  const int x0 = blockIdx.x * kExTileX, y0 = blockIdx.y * kExTileY;
  const int gx0 = x0 - E - R, gy0 = y0 - E - R;
  for (int i = threadIdx.x; i < gw * gh; i += 256) {
    const int yy = gy0 + i / gw, xx = gx0 + i % gw;
    sG[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? gradient_at<GRAD>(img, H, W, yy, xx) : make_double2(0.0, 0.0);
  }
  __syncthreads();
  const double inv_area = 1.0 / (double)((2 * R + 1) * (2 * R + 1));
  for (int i = threadIdx.x; i < rw * rh; i += 256) {
    const int ry = i / rw, rx = i % rw;
    const int y = y0 - E + ry, x = x0 - E + rx;
    double res = -INFINITY;
    if (y >= 0 && y < H && x >= 0 && x < W) {
      double sxx = 0.0, sxy = 0.0, syy = 0.0;
      for (int dy = -R; dy <= R; ++dy) {
        int yy = y + dy;
        if (yy < 0 || yy >= H) {
          if (window == FC_WINDOW_BOX) continue;
          yy = reflect101(yy, H);
        }
        for (int dx = -R; dx <= R; ++dx) {
          int xx = x + dx;
          if (xx < 0 || xx >= W) {
            if (window == FC_WINDOW_BOX) continue;
            xx = reflect101(xx, W);
          }
          const double2 g = sG[(yy - gy0) * gw + (xx - gx0)];
          if (window == FC_WINDOW_BOX) {
            sxx += g.x * g.x;
            sxy += g.x * g.y;
            syy += g.y * g.y;
          } else {
            const double wgt = __dmul_rn(taps.w[dy + R], taps.w[dx + R]);
            sxx = __dadd_rn(sxx, __dmul_rn(wgt, __dmul_rn(g.x, g.x)));
            sxy = __dadd_rn(sxy, __dmul_rn(wgt, __dmul_rn(g.x, g.y)));
            syy = __dadd_rn(syy, __dmul_rn(wgt, __dmul_rn(g.y, g.y)));
          }
        }
      }
      if (response == FC_RESP_HARRIS) {
        res = harris_from_sums(sxx, sxy, syy, k);
      } else {
        // the app divides its box sums by 25 = window area before eigvalsh (APP/FeatureCraft_Backend.py:428-431);
        // Gaussian weights already sum to one
        if (window == FC_WINDOW_BOX) {
          const double area = (double)((2 * R + 1) * (2 * R + 1));
          sxx = __ddiv_rn(sxx, area); sxy = __ddiv_rn(sxy, area); syy = __ddiv_rn(syy, area);
        }
        res = dsterf_2x2_min(sxx, sxy, syy);
      }
    }
    sR[i] = res;
  }
  (void)inv_area;
  __syncthreads();
  const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;  // 8 warps, two rows each: one ballot = one mask word
  for (int ty = ty0; ty < kExTileY; ty += 8) {
    const int y = y0 + ty, x = x0 + tx;
    const bool inside = y < H && x < W;
    const double r = sR[(ty + E) * rw + tx + E];
    if (inside) out[((size_t)b * H + y) * W + x] = r;
    if (mask) {
      bool hit = inside && r > threshold;
      if (hit && nms) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) hit &= r >= sR[(ty + E + dy) * rw + tx + E + dx];  // outside the image: -inf
      }
      const uint32_t bits = __ballot_sync(0xffffffffu, hit);
      if (tx == 0 && y < H) mask[((size_t)b * H + y) * mask_pitch + (x0 >> 5)] = bits;
    }
  }
}

// 3x3 non-maximum suppression + threshold of an existing response map (any of the response kernels):
// bit set iff R > threshold and R >= every neighbour inside the image (R == 3x3 dilation of R)
__global__ void __launch_bounds__(256)
nms3x3_mask_kernel(const double* __restrict__ resp, int H, int W, double threshold, uint32_t* __restrict__ mask, int pitch) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  const double* r = resp + (size_t)b * H * W;
  bool hit = false;
  if (y < H && x < W) {
    const double c = r[(size_t)y * W + x];
    hit = c > threshold;
    if (hit) {
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) hit &= c >= r[(size_t)yy * W + xx];
        }
    }
  }
  const uint32_t bits = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && y < H) mask[((size_t)b * H + y) * pitch + blockIdx.x] = bits;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_corner_response_ex(const uint8_t* gray, int batch, int height, int width, int gradient, int window,
                                     int window_size, const double* window_taps_host, int response, double k, int nms,
                                     double threshold, double* out, uint32_t* mask, void* stream) {
  if (!gray || !out || batch <= 0 || height <= 0 || width <= 0) return FC_ERR_BAD_ARG;
  if (gradient != FC_GRAD_CENTRAL && gradient != FC_GRAD_KX && gradient != FC_GRAD_SOBEL) return FC_ERR_BAD_ARG;
  if (window != FC_WINDOW_BOX && window != FC_WINDOW_GAUSS) return FC_ERR_BAD_ARG;
  if (response != FC_RESP_HARRIS && response != FC_RESP_LAMBDA_MINUS) return FC_ERR_BAD_ARG;
  if (window_size < 1 || (window_size & 1) == 0 || window_size > 2 * kExMaxR + 1) return FC_ERR_BAD_ARG;
  if (window == FC_WINDOW_GAUSS && !window_taps_host) return FC_ERR_BAD_ARG;
  if (nms && !mask) return FC_ERR_BAD_ARG;
  if (gradient == FC_GRAD_CENTRAL && (height < 2 || width < 2)) return FC_ERR_BAD_ARG;
  const int R = window_size / 2;
  if (window == FC_WINDOW_GAUSS && (height <= R + 1 || width <= R + 1)) return FC_ERR_UNSUPPORTED;  // single reflection only
  WindowTaps taps;
  for (int i = 0; i < 2 * kExMaxR + 1; ++i) taps.w[i] = (window == FC_WINDOW_GAUSS && i < window_size) ? window_taps_host[i] : 1.0;
  const int E = nms ? 1 : 0;
  const size_t smem = sizeof(double2) * (size_t)(kExTileX + 2 * (E + R)) * (kExTileY + 2 * (E + R)) +
                      sizeof(double) * (size_t)(kExTileX + 2 * E) * (kExTileY + 2 * E);
  dim3 grid(div_up(width, kExTileX), div_up(height, kExTileY), batch);
  cudaStream_t st = (cudaStream_t)stream;
  const int pitch = div_up(width, 32);
#define FC_EX(G)                                                                                                      \
  do {                                                                                                                \
    FC_CUDA_TRY(cudaFuncSetAttribute(corner_ex_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    corner_ex_kernel<G><<<grid, 256, smem, st>>>(gray, height, width, R, window, taps, response, k, nms, threshold, out, \
                                                 mask, pitch);                                                        \
  } while (0)
  if (gradient == FC_GRAD_CENTRAL) FC_EX(FC_GRAD_CENTRAL);
  else if (gradient == FC_GRAD_KX) FC_EX(FC_GRAD_KX);
  else FC_EX(FC_GRAD_SOBEL);
#undef FC_EX
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_nms3x3_mask(const double* response, int batch, int height, int width, double threshold, uint32_t* mask,
                              void* stream) {
  if (!response || !mask || batch <= 0 || height <= 0 || width <= 0) return FC_ERR_BAD_ARG;
  dim3 grid(div_up(width, 32), div_up(height, 8), batch);
  nms3x3_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(response, height, width, threshold, mask, div_up(width, 32));
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
