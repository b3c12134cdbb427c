// Opt-in sub-pixel refinement of SIFT keypoints: localize_keypoint of the app (APP/utils/SIFT_scale_space.py:213-261)
// with the acceptance tests its caller keeps commented out (:89-103) -- a 3-D quadratic fit to the DoG stack around each
// extremum.  The reference never runs it (SURVEY.md trap 1), so nothing on the default path calls this kernel.
#include "fc_common.cuh"

namespace fc {

// solve H o = -J for a symmetric 3x3 H by Gaussian elimination with partial pivoting (np.linalg.inv(H).dot(J) goes
// through LAPACK's LU as well); false when a pivot vanishes (numpy raises LinAlgError there)
__device__ __forceinline__ bool solve3(double (&a)[3][4]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int r = c + 1; r < 3; ++r)
      if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (a[p][c] == 0.0) return false;
    if (p != c)
      for (int j = 0; j < 4; ++j) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
    for (int r = c + 1; r < 3; ++r) {
      const double f = a[r][c] / a[c][c];
      for (int j = c; j < 4; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int r = 2; r >= 0; --r) {
    double s = a[r][3];
    for (int j = r + 1; j < 3; ++j) s -= a[r][j] * a[j][3];
    a[r][3] = s / a[r][r];
  }
  return true;
}

__global__ void __launch_bounds__(128)
refine_keypoints_kernel(const double* __restrict__ gauss, int L, int H, int W, const int32_t* __restrict__ keypoints, int n,
                        int k, double contrast_th, double ratio_th, double* __restrict__ offsets,
                        double* __restrict__ contrast, double* __restrict__ ratio, uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = keypoints[3 * i], y = keypoints[3 * i + 1], x = keypoints[3 * i + 2];
  const size_t plane = (size_t)H * W;
  const double* g = gauss + (size_t)b * L * plane;
  // D[yy][xx][s] = DoG_s = G_{s+1} - G_s
  auto D = [&](int yy, int xx, int s) { return g[(size_t)(s + 1) * plane + (size_t)yy * W + xx] - g[(size_t)s * plane + (size_t)yy * W + xx]; };
  const double c0 = D(y, x, k);
  const double dx = (D(y, x + 1, k) - D(y, x - 1, k)) / 2.0;
  const double dy = (D(y + 1, x, k) - D(y - 1, x, k)) / 2.0;
  const double ds = (D(y, x, k + 1) - D(y, x, k - 1)) / 2.0;
  const double dxx = D(y, x + 1, k) - 2 * c0 + D(y, x - 1, k);
  const double dxy = ((D(y + 1, x + 1, k) - D(y + 1, x - 1, k)) - (D(y - 1, x + 1, k) - D(y - 1, x - 1, k))) / 4;
  const double dxs = ((D(y, x + 1, k + 1) - D(y, x - 1, k + 1)) - (D(y, x + 1, k - 1) - D(y, x - 1, k - 1))) / 4;
  const double dyy = D(y + 1, x, k) - 2 * c0 + D(y - 1, x, k);
  const double dys = ((D(y + 1, x, k + 1) - D(y - 1, x, k + 1)) - (D(y + 1, x, k - 1) - D(y - 1, x, k - 1))) / 4;
  const double dss = D(y, x, k + 1) - 2 * c0 + D(y, x, k - 1);
  double a[3][4] = {{dxx, dxy, dxs, -dx}, {dxy, dyy, dys, -dy}, {dxs, dys, dss, -ds}};
  const bool ok = solve3(a);
  const double ox = ok ? a[0][3] : NAN, oy = ok ? a[1][3] : NAN, os = ok ? a[2][3] : NAN;
  offsets[3 * i] = ox;
  offsets[3 * i + 1] = oy;
  offsets[3 * i + 2] = os;
  // the caller's commented-out tests (:89-101): offset, interpolated contrast, principal-curvature ratio
  const double con = c0 + 0.5 * (dx * ox + dy * oy + ds * os);
  const double tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
  const double r = (tr * tr) / det;
  contrast[i] = con;
  ratio[i] = r;
  bool kp = ok;
  if (fmax(ox, fmax(oy, os)) > 0.5) kp = false;
  if (fabs(con) < contrast_th) kp = false;
  if (r > ratio_th) kp = false;
  keep[i] = kp ? 1 : 0;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_refine_keypoints(const double* gauss, int batch, int n_levels, int height, int width, const int32_t* keypoints,
                                   int n_keypoints, int k, double contrast_th, double ratio_th, double* offsets,
                                   double* contrast, double* ratio, uint8_t* keep, void* stream) {
  if (!gauss || batch <= 0 || height < 3 || width < 3) return FC_ERR_BAD_ARG;
  if (k < 1 || k + 2 >= n_levels) return FC_ERR_BAD_ARG;  // needs DoG k-1 .. k+1 = levels k-1 .. k+2
  if (n_keypoints == 0) return FC_OK;
  if (!keypoints || !offsets || !contrast || !ratio || !keep || n_keypoints < 0) return FC_ERR_BAD_ARG;
  refine_keypoints_kernel<<<div_up(n_keypoints, 128), 128, 0, (cudaStream_t)stream>>>(
      gauss, n_levels, height, width, keypoints, n_keypoints, k, contrast_th, ratio_th, offsets, contrast, ratio, keep);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
