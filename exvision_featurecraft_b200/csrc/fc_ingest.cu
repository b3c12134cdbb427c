// The steps in front of the hot path: channel handling at load time and sift_resize.
//
//   bgr_to_rgb_kernel      convert_BGR_to_RGB (APP/utils/helper_functions.py:12-14): img[..., ::-1] of the cv2.imread result.
//   rgb_to_gray_f64_kernel convert_to_grayscale (APP/utils/keypoint_descriptor.py:36-39): float weighted sum, no cast.
//   blur_axis_kernel       scipy.ndimage.gaussian_filter's 1-D correlation along one axis, 'mirror' border   } sift_resize
//   zoom_linear_kernel     scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True)                         } (:11-33)
//
// sift_resize is skimage.transform.resize(img, newshape, anti_aliasing=True) in the reference (scikit-image 0.24):
// img_as_float, a Gaussian low-pass with sigma = (in/out - 1) / 2 per axis (ndimage 'mirror' = skimage 'reflect'),
// then ndi.zoom(order=1, grid_mode=True).  scikit-image is not in this image, so the pair of kernels is pinned
// against the scipy calls skimage makes (tests/test_ingest_gpu.py), not against skimage itself ("parity unpinned"
// for that one library call).  All arithmetic float64, like the reference.
#include "fc_common.cuh"

namespace fc {

__global__ void bgr_to_rgb_kernel(const uint8_t* __restrict__ bgr, int64_t n_pixels, uint8_t* __restrict__ rgb) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_pixels; i += stride) {
    const uint8_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    rgb[3 * i] = r;
    rgb[3 * i + 1] = g;
    rgb[3 * i + 2] = b;
  }
}

// np.dot(image[..., :3], [0.2989, 0.5870, 0.1140]) on float64 pixels with `channels` >= 3 interleaved channels
__global__ void rgb_to_gray_f64_kernel(const double* __restrict__ rgb, int channels, int64_t n_pixels, double* __restrict__ gray) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_pixels; i += stride) {
    const double* p = rgb + (size_t)i * channels;
    gray[i] = __dadd_rn(__dadd_rn(__dmul_rn(p[0], 0.2989), __dmul_rn(p[1], 0.5870)), __dmul_rn(p[2], 0.1140));
  }
}

// ndimage 'mirror': d c b | a b c d | c b a (the edge sample is not repeated)
__device__ __forceinline__ int mirror_index(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n - 2;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - i;
}

// out[.., i, ..] = sum_k w[k] * in[.., mirror(i + k - r), ..] along `axis` of a [planes][H][W][C] array (C interleaved
// channels), evaluated like ndimage's correlate1d evaluates a symmetric filter (ni_filters.c): centre tap first, then
// the pairs (in[i - d] + in[i + d]) * w[d] from the outermost distance inwards; explicit roundings, no contraction
__global__ void __launch_bounds__(256)
blur_axis_kernel(const double* __restrict__ in, int H, int W, int C, int axis, const double* __restrict__ taps, int radius,
                 double* __restrict__ out) {
  const int64_t per_plane = (int64_t)H * W * C;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_plane) return;
  const double* src = in + (size_t)blockIdx.y * per_plane;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W);
  const int y = (int)(i / ((int64_t)C * W));
  auto at = [&](int d) {
    return axis == 0 ? src[((size_t)mirror_index(y + d, H) * W + x) * C + c] : src[((size_t)y * W + mirror_index(x + d, W)) * C + c];
  };
  double acc = __dmul_rn(at(0), taps[radius]);
  for (int d = radius; d >= 1; --d) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(at(-d), at(d)), taps[radius - d]));
  out[(size_t)blockIdx.y * per_plane + i] = acc;
}

// ndimage.zoom(order=1, mode='mirror', grid_mode=True): output pixel o samples input coordinate (o + 0.5) * (n_in /
// n_out) - 0.5 by linear interpolation between floor and floor + 1, both mirrored into range
__global__ void __launch_bounds__(256)
zoom_linear_kernel(const double* __restrict__ in, int H, int W, int C, int H2, int W2, double* __restrict__ out) {
  const int64_t per_out = (int64_t)H2 * W2 * C;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_out) return;
  const double* src = in + (size_t)blockIdx.y * H * W * C;
  const int c = (int)(i % C);
  const int x = (int)((i / C) % W2);
  const int y = (int)(i / ((int64_t)C * W2));
  // (explicit roundings: a fused multiply-add here moves the sample position by an ulp of ~1000, i.e. 1e-13)
  const double sy = __dsub_rn(__dmul_rn(__dadd_rn((double)y, 0.5), __ddiv_rn((double)H, (double)H2)), 0.5);
  const double sx = __dsub_rn(__dmul_rn(__dadd_rn((double)x, 0.5), __ddiv_rn((double)W, (double)W2)), 0.5);
  const double fy = floor(sy), fx = floor(sx);
  const double ty = __dsub_rn(sy, fy), tx = __dsub_rn(sx, fx);
  const int y0 = mirror_index((int)fy, H), y1 = mirror_index((int)fy + 1, H);
  const int x0 = mirror_index((int)fx, W), x1 = mirror_index((int)fx + 1, W);
  const double v00 = src[((size_t)y0 * W + x0) * C + c], v01 = src[((size_t)y0 * W + x1) * C + c];
  const double v10 = src[((size_t)y1 * W + x0) * C + c], v11 = src[((size_t)y1 * W + x1) * C + c];
  // separable evaluation, rows first then columns (the spline filter's axis order)
  const double a = __dadd_rn(__dmul_rn(__dsub_rn(1.0, ty), v00), __dmul_rn(ty, v10));
  const double b = __dadd_rn(__dmul_rn(__dsub_rn(1.0, ty), v01), __dmul_rn(ty, v11));
  out[(size_t)blockIdx.y * per_out + i] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, tx), a), __dmul_rn(tx, b));
}

__global__ void u8_to_unit_f64_kernel(const uint8_t* __restrict__ src, int64_t n, double* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = (double)src[i] / 255.0;  // skimage.util.img_as_float
}

}  // namespace fc

using namespace fc;

static unsigned grid_for(int64_t n) { return (unsigned)(div_up64(n, 256) < 148 * 32 ? div_up64(n, 256) : 148 * 32); }

extern "C" int fc_bgr_to_rgb_u8(const uint8_t* bgr, int batch, int height, int width, uint8_t* rgb, void* stream) {
  if (!bgr || !rgb || batch <= 0 || height <= 0 || width <= 0 || bgr == rgb) return FC_ERR_BAD_ARG;
  const int64_t n = (int64_t)batch * height * width;
  bgr_to_rgb_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(bgr, n, rgb);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_rgb_to_gray_f64(const double* rgb, int batch, int height, int width, int channels, double* gray, void* stream) {
  if (!rgb || !gray || batch <= 0 || height <= 0 || width <= 0 || channels < 3) return FC_ERR_BAD_ARG;
  const int64_t n = (int64_t)batch * height * width;
  rgb_to_gray_f64_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(rgb, channels, n, gray);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_u8_to_unit_f64(const uint8_t* src, int64_t n, double* dst, void* stream) {
  if (!src || !dst || n <= 0) return FC_ERR_BAD_ARG;
  u8_to_unit_f64_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, n, dst);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_blur_axis_f64(const double* in, int batch, int height, int width, int channels, int axis, const double* taps,
                                int radius, double* out, void* stream) {
  if (!in || !out || !taps || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || radius < 0 || in == out)
    return FC_ERR_BAD_ARG;
  if (axis != 0 && axis != 1) return FC_ERR_BAD_ARG;
  const int64_t per_plane = (int64_t)height * width * channels;
  blur_axis_kernel<<<dim3((unsigned)div_up64(per_plane, 256), batch), 256, 0, (cudaStream_t)stream>>>(in, height, width, channels,
                                                                                                   axis, taps, radius, out);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}

extern "C" int fc_zoom_linear_f64(const double* in, int batch, int height, int width, int channels, int out_height, int out_width,
                                  double* out, void* stream) {
  if (!in || !out || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || out_height <= 0 || out_width <= 0)
    return FC_ERR_BAD_ARG;
  const int64_t per_out = (int64_t)out_height * out_width * channels;
  zoom_linear_kernel<<<dim3((unsigned)div_up64(per_out, 256), batch), 256, 0, (cudaStream_t)stream>>>(in, height, width, channels,
                                                                                                  out_height, out_width, out);
  note_launch();
  FC_RETURN_IF_LAUNCH_FAILED();
  return FC_OK;
}
