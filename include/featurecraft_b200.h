/*
 * featurecraft_b200.h -- C ABI of libfeaturecraft_b200.so (sm_100a only).
 *
 * Drop-in boundary for the feature-extraction hot path of exVision-FeatureCraft.  The reference is
 * pure Python (no FFI of its own), so each entry point below names the reference function whose
 * arithmetic it replaces (file:line relative to the reference root; APP = FeatureCraft-PyQt-App).
 * The host-side mirror of the reference's Python signatures lives in exvision_featurecraft_b200/
 * and binds these symbols with ctypes (INTEGRATION.md shows the stub a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates result buffers and never synchronises: the caller owns all memory
 *     and the stream (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - images are dense row-major, batched as [batch][height][width] with identical shapes;
 *   - return value: FC_OK or a negative fc_status; fc_last_cuda_error() gives the CUDA code behind
 *     FC_ERR_CUDA;
 *   - there is no CPU fallback: without a CUDA device every compute call returns FC_ERR_CUDA.
 */
#ifndef FEATURECRAFT_B200_H
#define FEATURECRAFT_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define FC_API __attribute__((visibility("default")))
#else
#define FC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum fc_status {
  FC_OK = 0,
  FC_ERR_BAD_ARG = -1,   /* null pointer, non-positive size, unsupported parameter            */
  FC_ERR_CAPACITY = -2,  /* an output buffer is too small.  Output sizes of this path are data dependent and known
                            only after the device counted them, so it is the host-side drivers (BatchExtractor,
                            BatchCornerDetector) that report it -- with the required count -- after their one
                            synchronisation and BEFORE anything is written to the caller's buffers */
  FC_ERR_CUDA = -3,      /* a CUDA runtime call or launch failed                              */
  FC_ERR_UNSUPPORTED = -4
} fc_status;

/* element type of a scale-space stack (FC_F32: the bulk stack of certified mode, FC_F64: the float64 mode) */
typedef enum fc_dtype { FC_F32 = 0, FC_F64 = 1, FC_F16 = 2 /* descriptor export only */ } fc_dtype;

/* keypoint filter: FC_FILTER_APP keeps every extremum (APP/utils/SIFT_scale_space.py:89-105),
 * FC_FILTER_SCRIPT applies the contrast + edge-ratio tests of
 * implementation_without_ui/SIFT/SIFT.py:202-215. */
typedef enum fc_filter { FC_FILTER_APP = 0, FC_FILTER_SCRIPT = 1 } fc_filter;

/* Opt-in variants of the corner response.  The reference uses none of them (SURVEY.md trap 1), so every host-side
 * default is the first enumerator: np.gradient / K_X gradients, box window, no non-maximum suppression. */
typedef enum fc_gradient { FC_GRAD_CENTRAL = 0, FC_GRAD_KX = 1, FC_GRAD_SOBEL = 2 } fc_gradient;
typedef enum fc_window { FC_WINDOW_BOX = 0, FC_WINDOW_GAUSS = 1 } fc_window;
typedef enum fc_response { FC_RESP_HARRIS = 0, FC_RESP_LAMBDA_MINUS = 1 } fc_response;

FC_API int fc_version(void);
FC_API const char* fc_status_string(int status);
FC_API int fc_last_cuda_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
FC_API unsigned long long fc_launch_count(void);
/* which kernel family the scale-space calls of this process took (introspection for tests and bench.py): launches of
 * the packed float32 blur kernel, of its fused blur + DoG + search form, of the two-level float64 kernel, of the
 * generic level-by-level kernel, and fc_extrema_repair calls that worked from shared-memory patches */
typedef enum fc_path {
  FC_PATH_PACKED_BLUR = 0, FC_PATH_FUSED_SEARCH = 1, FC_PATH_MID_F64 = 2, FC_PATH_GENERIC_OCTAVE = 3, FC_PATH_REPAIR_PATCH = 4,
  FC_PATH_COUNT = 5
} fc_path;
FC_API unsigned long long fc_path_count(int path);

/* ---------------------------------------------------------------------------------------------
 * Corner detectors
 * ------------------------------------------------------------------------------------------- */

/* convert_to_gray, APP/utils/harris_utils.py:4-9: (0.2989 r + 0.5870 g + 0.1140 b) truncated to u8.
 * rgb is [batch][h][w][3] uint8. */
FC_API int fc_rgb_to_gray_u8(const uint8_t* rgb, int batch, int height, int width, uint8_t* gray, void* stream);

/* convolve2d_optimized, APP/utils/harris_utils.py:33-80: zero-padded CORRELATION of float64 images
 * with a square odd kernel (device pointer, kernel_size^2 float64); the k*k products of each pixel
 * are reduced in numpy's pairwise order so float data reproduce np.sum(..., axis=(2, 3)). */
FC_API int fc_correlate_zero_pad_f64(const double* image, int batch, int height, int width, const double* kernel,
                              int kernel_size, double* out, void* stream);

/* apply_harris_detector_vectorized, APP/FeatureCraft_Backend.py:336-358: np.gradient central
 * differences, window_size x window_size ones window (zero padded), R = det - k*trace^2, bit-exact.
 * response: [batch][h][w] float64.  mask (optional, may be NULL): bit (x & 31) of word
 * [b][y][x >> 5] is set iff R > threshold; row pitch = (width + 31) / 32 words. */
FC_API int fc_harris_response(const uint8_t* gray, int batch, int height, int width, int window_size, double k,
                       double threshold, double* response, uint32_t* mask, void* stream);

/* harris_detection, implementation_without_ui/Corner Detection/harris.ipynb:150-201: same response
 * with the 5x5 K_X gradient (zero padded).  The window is 2 * int(window_size / 2) + 1 wide, so even sizes are
 * accepted like in the notebook (:170-176).  The notebook only visits window_size/2 <= y < H - ...;
 * pass that as `border` to fc_threshold_mask_abs. */
FC_API int fc_notebook_harris_response(const uint8_t* gray, int batch, int height, int width, int window_size, double k,
                                double* response, void* stream);

/* The corner response with its opt-in components (north_star: Sobel gradients, Gaussian window, 3x3 NMS), one
 * shared-memory-tiled kernel: gradient (fc_gradient) -> structure-tensor products -> window (fc_window: ones window of
 * window_size^2 with zero padding like the reference, or the separable Gaussian window_taps_host[window_size] with
 * BORDER_REFLECT_101 like cv2.GaussianBlur) -> response (Harris det - k tr^2, or the smallest eigenvalue of the window
 * MEAN) -> optional mask: R > threshold, with nms != 0 also R >= its 8 neighbours.  FC_GRAD_CENTRAL + FC_WINDOW_BOX +
 * FC_RESP_HARRIS equals fc_harris_response and FC_GRAD_KX + FC_WINDOW_BOX(5) + FC_RESP_LAMBDA_MINUS equals
 * fc_lambda_minus_response bit for bit; the other combinations are checked against cv2.Sobel / cv2.GaussianBlur. */
FC_API int fc_corner_response_ex(const uint8_t* gray, int batch, int height, int width, int gradient, int window,
                          int window_size, const double* window_taps_host, int response, double k, int nms, double threshold,
                          double* out, uint32_t* mask, void* stream);

/* 3x3 non-maximum suppression + threshold of a resident response map: bit set iff R > threshold and R >= every
 * neighbour inside the image.  Opt-in: the reference thresholds only (APP/FeatureCraft_Backend.py:358, :441). */
FC_API int fc_nms3x3_mask(const double* response, int batch, int height, int width, double threshold, uint32_t* mask,
                   void* stream);

/* apply_lambda_minus_vectorized, APP/FeatureCraft_Backend.py:402-437: K_X gradients, 5x5 ones
 * window / 25, smallest eigenvalue with LAPACK's dsterf/dlae2 rounding sequence, bit-exact.
 * eig: [batch][h][w] float64; image_max: [batch] float64 (max over each image, for
 * threshold = pct * max, :440). */
FC_API int fc_lambda_minus_response(const uint8_t* gray, int batch, int height, int width, double* eig,
                             double* image_max, void* stream);
/* Opt-in fast form of the same map (default OFF everywhere; north_star's bar for responses is 1e-4 relative): the smaller
 * eigenvalue as det / larger eigenvalue, with the determinant Sxx Syy - Sxy^2 and the radicand (Sxx - Syy)^2 + 4 Sxy^2 of
 * the integer window sums formed exactly in 64-bit integers, so no cancellation occurs: agrees with
 * fc_lambda_minus_response to a few float64 ulps (tests: <= 1e-13 relative) but not bit for bit, at about half the
 * float64 work.  Needs width % 4 == 0 and the alignments of the fast path (FC_ERR_UNSUPPORTED otherwise). */
FC_API int fc_lambda_minus_response_fast(const uint8_t* gray, int batch, int height, int width, double* eig,
                                         double* image_max, void* stream);

/* Device self-test: counts the integers |n| <= 2^33 for which the division-free n/25 used by the lambda-minus
 * kernel differs from IEEE division (must be 0).  mismatches: one device uint64. */
FC_API int fc_selftest_div25(unsigned long long* mismatches, void* stream);
/* Device self-test of the straight-line float64 division / square-root sequences the lambda-minus kernel uses instead of
 * the library calls (so that the pixels of a thread overlap): compares them with __ddiv_rn / __dsqrt_rn on 2^32
 * pseudo-random operands of the kernel's range and writes the number of mismatches (must be 0). */
FC_API int fc_selftest_divsqrt(unsigned long long* mismatches, void* stream);

/* np.argwhere(map > thr) as a bit mask (APP/FeatureCraft_Backend.py:358, :268).  Pixels closer than
 * `border` to any image edge are never set. */
FC_API int fc_threshold_mask_abs(const double* map, int batch, int height, int width, double threshold, int border,
                          uint32_t* mask, void* stream);
/* map > fraction * image_max[b] (APP/FeatureCraft_Backend.py:440-441); the product is rounded once. */
FC_API int fc_threshold_mask_rel(const double* map, int batch, int height, int width, double fraction,
                          const double* image_max, uint32_t* mask, void* stream);

/* Row-major compaction of a bit mask (the order np.argwhere / np.where produce).
 * fc_mask_scan: row_offsets[i] = number of set bits in rows [0, i) of the whole batch, i in
 * [0, batch*height]; so row_offsets[batch*height] is the total and row_offsets[b*height] the start
 * of image b.  scratch must hold fc_scan_scratch_elems(batch*height) int32. */
FC_API int64_t fc_scan_scratch_elems(int64_t n);
FC_API int fc_mask_scan(const uint32_t* mask, int batch, int height, int width, int32_t* row_offsets, int32_t* scratch,
                 void* stream);
/* Copy n int32 from device memory to PINNED host memory by a kernel store (UVA pointer), not by the D2H copy
 * engine: the way the host mirrors read output sizes without queueing behind bulk result transfers. */
FC_API int fc_publish_i32(const int32_t* src, int32_t* dst_host_pinned, int n, void* stream);
/* The pieces of fc_mask_scan / fc_orientation_scan on their own, for the batched driver: the per-row (per-keypoint)
 * counts of several masks are written into ONE concatenated array, scanned with one call, and the totals of all
 * blocks (the scan values at the block boundaries, indices given by the host) reach pinned host memory with one
 * launch.  Row offsets taken from such a scan are positions in the concatenated output list, so fc_mask_emit* /
 * fc_orientation_emit are called with the base pointer of that list.
 * fc_exclusive_scan_i32: out[i] = sum(in[0..i)), i in [0, n]; scratch: fc_scan_scratch_elems(n) int32. */
FC_API int fc_mask_row_count(const uint32_t* mask, int batch, int height, int width, int32_t* counts, void* stream);
FC_API int fc_popcount_u64(const uint64_t* masks, int64_t n, int32_t* counts, void* stream);
FC_API int fc_exclusive_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* scratch, void* stream);
FC_API int fc_publish_gather_i32(const int32_t* src, const int64_t* indices_host, int n, int32_t* dst_host_pinned, void* stream);

/* coords: [total][2] int32 (row, col); values (optional): map gathered at those pixels. */
FC_API int fc_mask_emit(const uint32_t* mask, const int32_t* row_offsets, int batch, int height, int width,
                 const double* map, int32_t* coords, double* values, void* stream);
/* same order, but (image, row, col) int32 triples: the keypoint records the SIFT stages consume
 * (np.argwhere(mask) of represent_keypoints' bool images, APP/utils/keypoint_descriptor.py:42-69, :144). */
FC_API int fc_mask_emit_keypoints(const uint32_t* mask, const int32_t* row_offsets, int batch, int height, int width,
                           int32_t* keypoints, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SIFT scale space (APP/utils/SIFT_scale_space.py)
 * ------------------------------------------------------------------------------------------- */

/* One octave of generate_gaussian_images_in_octave (:109-160) for a batch of equally sized bases:
 * G_i = base (*) K_i for the n_levels unnormalised, truncated Gaussians whose separable taps are
 * given in taps_host (concatenated, lengths in tap_len_host; 'same' / 'symm' border), except that
 * for first_is_identity != 0 level 0 is the base itself (octaves > 0, :136-137).
 * gauss: [batch][n_levels][h][w] of `dtype`.  The 3x3x3 extremum rule of get_keypoints (:78-88) is
 * evaluated on DoG_k = G_{k+1} - G_k for k = 1 .. n_levels-3 and written as bit masks
 * extrema: [n_levels-3][batch][h][(w+31)/32] (one batch of masks per k, ready for fc_mask_scan); with filter == FC_FILTER_SCRIPT the contrast and
 * edge-ratio tests are applied too.  base has element type `dtype`. */
FC_API int fc_gauss_octave(const void* base, int dtype, int batch, int height, int width, int n_levels,
                    const double* taps_host, const int* tap_len_host, int first_is_identity, int filter,
                    double contrast_th, double ratio_th, void* gauss, uint32_t* extrema, void* stream);

/* the extrema half of fc_gauss_octave on an existing [batch][n_levels][h][w] stack (extrema == NULL there
 * skips it): lets a caller time / re-run the two kernels separately. */
FC_API int fc_gauss_octave_extrema(const void* gauss, int dtype, int batch, int height, int width, int n_levels,
                            int filter, double contrast_th, double ratio_th, uint32_t* extrema, void* stream);

/* get_keypoints (APP/utils/SIFT_scale_space.py:60-106 / SIFT.py:172-222) called directly on a stack of
 * n_dog >= 3 DoG images [batch][n_dog][h][w]: extrema [n_dog-2][batch][h][(w+31)/32], plane k-1 for DoG k. */
FC_API int fc_dog_extrema(const void* dog, int dtype, int batch, int height, int width, int n_dog, int filter,
                   double contrast_th, double ratio_th, uint32_t* extrema, void* stream);

/* Opt-in sub-pixel refinement (north_star; the reference leaves it commented out, APP/utils/SIFT_scale_space.py:
 * 89-103): localize_keypoint (:213-261) for n keypoints (image, row, col) of DoG level k of a float64 Gaussian stack
 * [batch][n_levels][h][w] (interior keypoints, as get_keypoints emits them).  offsets[i] = -inv(H) J in (x, y, s)
 * order, contrast[i] = D + J.offset / 2, ratio[i] = tr^2 / det of the spatial Hessian, keep[i] = the commented-out
 * acceptance: max(offset) <= 0.5 and |contrast| >= contrast_th and not ratio > ratio_th. */
FC_API int fc_refine_keypoints(const double* gauss, int batch, int n_levels, int height, int width, const int32_t* keypoints,
                        int n_keypoints, int k, double contrast_th, double ratio_th, double* offsets, double* contrast,
                        double* ratio, uint8_t* keep, void* stream);

/* rescale(gray, 2, anti_aliasing=False) in front of the pyramid (APP/utils/keypoint_descriptor.py:296):
 * bilinear, pixel-centre aligned, mirror border (== scipy.ndimage.zoom(order=1, mode='mirror',
 * grid_mode=True); scikit-image itself is not in this image).  image: [batch][h][w]
 * float64; out: [batch][2h][2w] of `dtype`. */
FC_API int fc_upsample2(const double* image, int batch, int height, int width, int dtype, void* out, void* stream);

/* img[::2, ::2] (APP/utils/SIFT_scale_space.py:209) of level `level` of a [batch][n_levels][h][w]
 * stack into a dense [batch][(h+1)/2][(w+1)/2] array. */
FC_API int fc_downsample2(const void* gauss, int dtype, int batch, int n_levels, int level, int height, int width,
                   void* out, void* stream);

/* float64 image -> scale-space dtype (identity copy for FC_F64) */
FC_API int fc_convert_f64(const double* src, int64_t n, int dtype, void* dst, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Certified mode of the scale space: float32 bulk arithmetic, float64 decisions.
 * The float32 stack of fc_gauss_octave is kept for the 3x3x3 comparisons; every comparison whose outcome the
 * float32 error bound cannot guarantee is re-evaluated on float64 values, so the keypoint lists are the ones the
 * float64 path (and the reference, APP/utils/SIFT_scale_space.py:78-88) produces.
 * ------------------------------------------------------------------------------------------- */

/* float32 copy of a float64 array; value_bound[0] (float) receives an upper bound of max |src| -- the scale the
 * error bounds below are relative to. */
FC_API int fc_convert_f64_bound(const double* src, int64_t n, float* dst, float* value_bound, void* stream);

/* The two base-image producers of certified mode, each writing the float64 and the float32 copy in one pass:
 * fc_upsample2 (even width; also value_bound as fc_convert_f64_bound) and fc_downsample2 of a float64 stack. */
FC_API int fc_upsample2_dual(const double* image, int batch, int height, int width, double* out64, float* out32,
                      float* value_bound, void* stream);
FC_API int fc_downsample2_dual(const double* gauss, int batch, int n_levels, int level, int height, int width,
                        double* out64, float* out32, void* stream);

/* Levels first_level .. first_level + level_count - 1 of the octave in float64 (same operator as fc_gauss_octave,
 * SIFT_scale_space.py:132,144): out [batch][level_count][h][w].  Certified mode keeps levels 1 .. s: the gradient
 * fields are taken from them (APP/utils/keypoint_descriptor.py:126, :241) and level s is the next octave's base
 * (SIFT_scale_space.py:209). */
FC_API int fc_gauss_levels_f64(const double* base, int batch, int height, int width, int n_levels, const double* taps_host,
                        const int* tap_len_host, int first_level, int level_count, double* out, void* stream);

/* get_keypoints on the float32 stack with an error margin: a centre that beats its 26 neighbours by more than
 * eps = eps_rel * value_bound[0] is an extremum of the float64 DoG too and is written to `extrema`; one that loses by
 * more than eps is not; anything closer (every tie included; with FC_FILTER_SCRIPT every hit, whose contrast / edge
 * tests are float64 decisions as well) is appended to amb as int32 (image, k, row, col) with its mask bit clear.
 * amb_count[0] is zeroed first and counts ALL such entries: if it exceeds amb_cap the list is incomplete and the
 * caller must fall back to the float64 kernels for this octave. */
FC_API int fc_gauss_octave_extrema_certified(const float* gauss, int batch, int height, int width, int n_levels, int filter,
                                      double contrast_th, double ratio_th, const float* value_bound, double eps_rel,
                                      uint32_t* extrema, int32_t* amb, int amb_cap, int32_t* amb_count, void* stream);

/* fc_gauss_octave (float32, default radii 3/5/6/9/13 only) and fc_gauss_octave_extrema_certified fused: the float32
 * levels never leave shared memory -- nothing but the search reads them in certified mode -- and the search runs on the
 * DoG tiles in place.  Same outputs as the two calls (extrema [2][batch][h][(w+31)/32] is zeroed here, amb / amb_count
 * as above); FC_ERR_UNSUPPORTED for other tap sets or images below 4 x 4: use the two calls then. */
FC_API int fc_gauss_octave_search(const float* base, int batch, int height, int width, int n_levels, const double* taps_host,
                           const int* tap_len_host, int first_is_identity, int filter, const float* value_bound,
                           double eps_rel, uint32_t* extrema, int32_t* amb, int amb_cap, int32_t* amb_count, void* stream);

/* Decides the amb entries on float64 DoG values and sets their bits in `extrema`: levels mid_first ..
 * mid_first + mid_count - 1 are read from `mid` ([batch][mid_count][h][w], fc_gauss_levels_f64), level 0 of an
 * octave > 0 (first_is_identity) from `base`, every other level is blurred on demand from `base` (float64) around the
 * entry.  taps_host / tap_len_host as for fc_gauss_octave. */
FC_API int fc_extrema_repair(const double* base, int first_is_identity, const double* mid, int mid_first, int mid_count,
                      int batch, int height, int width, int n_levels, const double* taps_host, const int* tap_len_host,
                      int filter, double contrast_th, double ratio_th, const int32_t* amb, const int32_t* amb_count,
                      int amb_cap, uint32_t* extrema, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SIFT orientation + descriptor (APP/utils/keypoint_descriptor.py)
 * ------------------------------------------------------------------------------------------- */

/* sift_gradient (:72-79) of level `level` of a [batch][n_levels][h][w] stack: magnitude and
 * direction (degrees, wrapped with Python's % 360), stored in `dtype`, each [batch][h][w], plus
 * orientation_bin = np.round(direction * 36 / 360) (:140-142; 0..36, half to even) as uint8. */
FC_API int fc_gradient_field(const void* gauss, int dtype, int batch, int n_levels, int level, int height, int width,
                      void* magnitude, void* direction, uint8_t* orientation_bin, void* stream);

/* Certified mode's gradient field of plane `plane` of a float64 [batch][n_planes][h][w] stack: one word per pixel,
 *   bits 5..0  the reference's orientation bin np.round(direction * 36 / 360) (:140-142; 0..36),
 *   bits 7..6  which side of the bin centre the direction lies on (0 below, 1 at / above, 2 exactly on the upper
 *              edge of a tie rounded down), so that 2 * bin - 1 + side is the exact 5-degree sector of the direction,
 *   bits 31..8 the float32 magnitude (sign dropped) rounded to 16 mantissa bits.
 * Sectors are decided in float32 unless the direction lies within 1e-4 degrees of a multiple of 5: those pixels
 * evaluate the reference's float64 expression, so the codes are the float64 pipeline's. */
FC_API int fc_gradient_codes(const double* levels, int batch, int n_planes, int plane, int height, int width,
                      uint32_t* codes, void* stream);

/* dog_keypoints_orientations (:116-172) for one (octave, scale): keypoints are (image, row, col)
 * int32 triples in reference order; `weights` is gaussian_filter_kernel(sigma) restricted to the offsets a window
 * clipped to the image can reach: a dense (2 table_ry + 1) x (2 table_rx + 1) device table of element type `dtype`,
 * entry [dy + table_ry][dx + table_rx], table_ry >= min(radius, h - 1), table_rx >= min(radius, w - 1).
 * For keypoint n the 36-bin float32 histogram is built (sums in `dtype`) and bin_mask[n] receives bit b for every
 * bin with hist[b] >= cut (:165-167); cut = float32(0.8) * max in float32 (numpy >= 2, NEP 50: the default and what the
 * golden vectors were produced with) or, with numpy_legacy_promotion != 0, float32(0.8 * float64(max)) (numpy 1.26, the
 * version the reference pins): the two differ by at most one float32 ulp. */
FC_API int fc_orientation_bins(const void* magnitude, const uint8_t* orientation_bin, int dtype, int batch, int height,
                        int width, const int32_t* keypoints, int n_keypoints, const void* weights, int radius,
                        int table_ry, int table_rx, uint64_t* bin_mask, int numpy_legacy_promotion, void* stream);
/* certified mode: float32 histograms from fc_gradient_codes' words; a keypoint with a bin within
 * tolerance * max of the cut gets bit 63 of its mask set instead of a decision and its index appended to
 * recheck_list (int32 [n_keypoints]; recheck_count[0] is zeroed first) ...
 * `weights` here is the SEPARABLE float32 form of the same window: (2 table_ry + 1) row factors
 * exp(-dy^2 / 2 sigma^2) / (2 pi sigma^2) followed by (2 table_rx + 1) column factors exp(-dx^2 / 2 sigma^2); the
 * tolerance has to cover the extra float32 roundings (see sift.orientation_tolerance). */
FC_API int fc_orientation_bins_coded(const uint32_t* codes, int batch, int height, int width, const int32_t* keypoints,
                              int n_keypoints, const float* weights, int radius, int table_ry, int table_rx,
                              double tolerance, uint64_t* bin_mask, int32_t* recheck_list, int32_t* recheck_count,
                              void* stream);
/* ... and this call decides exactly those keypoints with float64 magnitudes (from the float64 level), float64
 * weights and float64 sums, clearing bit 63: bin_mask then equals the float64 path's. */
FC_API int fc_orientation_recheck(const double* levels, int n_planes, int plane, const uint32_t* codes, int batch,
                           int height, int width, const int32_t* keypoints, int n_keypoints, const double* weights,
                           int radius, int table_ry, int table_rx, uint64_t* bin_mask, const int32_t* recheck_list,
                           const int32_t* recheck_count, int numpy_legacy_promotion, void* stream);

/* Expand bin masks into oriented keypoints in reference order: offsets[n] = popcount prefix
 * (exclusive, offsets[n_keypoints] = total; scratch as for fc_mask_scan), then
 * oriented[m] = (image, row, col, bin). */
FC_API int fc_orientation_scan(const uint64_t* bin_mask, int n_keypoints, int32_t* offsets, int32_t* scratch, void* stream);
FC_API int fc_orientation_emit(const int32_t* keypoints, const uint64_t* bin_mask, const int32_t* offsets, int n_keypoints,
                        int32_t* oriented, void* stream);

/* extract_sift_descriptors (:230-289) for one (octave, scale): 16x16 nearest-neighbour rotated
 * window (cv2.warpAffine fixed-point rule, :175-212), Gaussian mask `gmask` (16x16 float64, device,
 * from get_gaussian_mask :215-227), 4x4x8 float32 cell histograms, normalise / clip / normalise.
 * trig: device table [37][2 + 32] float64; rows 0..35 per orientation bin: cos, sin of theta*3.14159/180 and
 * the rint(cos*x*1024), rint(sin*x*1024) columns for x = 0..15 (computed on the host with the
 * reference's own libm calls); row 36, entries 0..7: smallest float64 rel with int(rel*8/360.0) >= j+1.
 * descriptors: [n_oriented][128] float32 (out_dtype FC_F32) or float64; fc_descriptors_coded also writes FC_F16
 * (host export at half the bytes: element error <= 2^-11 relative, i.e. <= 5e-4 L2 on a unit vector). */
/* Optional output placement of the descriptor calls (NULL = descriptor n of the call goes to row n).  The batched
 * pipeline keeps its keypoint lists block-major -- (octave, scale) blocks, each sorted by (image, row, col, bin) --
 * while the reference's output is image-major (per image: octave, scale, row-major, bin).  With a placement the rows
 * land in that final order directly: remap[image] = (list index of the image's first oriented keypoint of this block,
 * its final row) from fc_sift_layout, list_base = list index of the call's first keypoint; `points` (optional)
 * receives the keypoint record of every row: int16 row, int16 col, uint8 octave, uint8 scale, uint8 orientation bin,
 * uint8 0 (the reference's 5-tuple (i, j, octave, scale, (bin + 0.5) * 10), APP/utils/keypoint_descriptor.py:167-170). */
typedef struct fc_desc_placement {
  const int32_t* remap; /* device, [batch][2] */
  void* points;         /* device, [total oriented][8 bytes], or NULL */
  int list_base;
  int octave, scale;
} fc_desc_placement;

FC_API int fc_descriptors(const void* magnitude, const void* direction, int dtype, int batch, int height, int width,
                   const int32_t* oriented, int n_oriented, const double* gmask, const double* trig,
                   void* descriptors, int out_dtype, const fc_desc_placement* placement, void* stream);
/* remap tables ([n_blocks][batch][2] int32) and per-image totals ([batch] int32) for fc_desc_placement, from the
 * pipeline's two scans: row_offsets = cumulative keypoints over the concatenated mask rows of all blocks (block i
 * starts at row block_row_start_host[i] and has block_height_host[i] rows per image), oriented_offsets = cumulative
 * oriented keypoints over all keypoints (fc_orientation_scan / fc_exclusive_scan_i32). */
FC_API int fc_sift_layout(const int32_t* row_offsets, const int32_t* oriented_offsets, const int* block_row_start_host,
                   const int* block_height_host, int n_blocks, int batch, int32_t* remap, int32_t* image_counts, void* stream);

/* certified mode: the same descriptors from fc_gradient_codes' words; the bin of a sample is integer arithmetic on
 * the exact direction sector (no float rounding near a bin edge). */
FC_API int fc_descriptors_coded(const uint32_t* codes, int batch, int height, int width, const int32_t* oriented,
                         int n_oriented, const double* gmask, const double* trig, void* descriptors, int out_dtype,
                         const fc_desc_placement* placement, void* stream);

/* rotated_subimage (:175-212) on its own: nearest-neighbour window of out_width x out_height (<= 1024
 * samples) around (center_x, center_y) with cv2.warpAffine's 10-bit fixed-point coordinates; the caller
 * passes cos/sin of theta*3.14159/180.  image and out are float64. */
FC_API int fc_rotated_window(const double* image, int height, int width, double center_x, double center_y,
                      double cos_theta, double sin_theta, int out_width, int out_height, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * In front of the hot path: channel handling at load time and sift_resize
 * ------------------------------------------------------------------------------------------- */

/* convert_BGR_to_RGB (APP/utils/helper_functions.py:12-14), applied to the cv2.imread result in load_image
 * (APP/FeatureCraft_Backend.py:105-106): [batch][h][w][3] uint8, channel order reversed (out of place). */
FC_API int fc_bgr_to_rgb_u8(const uint8_t* bgr, int batch, int height, int width, uint8_t* rgb, void* stream);

/* convert_to_grayscale (APP/utils/keypoint_descriptor.py:36-39): np.dot(image[..., :3], [0.2989, 0.5870, 0.1140]) on
 * float64 pixels with `channels` >= 3 interleaved channels; no cast. */
FC_API int fc_rgb_to_gray_f64(const double* rgb, int batch, int height, int width, int channels, double* gray, void* stream);

/* skimage.util.img_as_float of uint8 data (value / 255), the first thing skimage.transform.resize does */
FC_API int fc_u8_to_unit_f64(const uint8_t* src, int64_t n, double* dst, void* stream);

/* The two halves of sift_resize (APP/utils/keypoint_descriptor.py:11-33 = skimage.transform.resize(img, newshape,
 * anti_aliasing=True)) as the scipy.ndimage calls scikit-image 0.24 makes, on [batch][h][w][channels] float64:
 * fc_blur_axis_f64 = gaussian_filter's 1-D correlation with `taps` (2 radius + 1 device doubles) along axis 0 (rows) or
 * 1 (columns), 'mirror' border; fc_zoom_linear_f64 = zoom(order=1, mode='mirror', grid_mode=True) to out_height x
 * out_width. */
FC_API int fc_blur_axis_f64(const double* in, int batch, int height, int width, int channels, int axis, const double* taps,
                     int radius, double* out, void* stream);
FC_API int fc_zoom_linear_f64(const double* in, int batch, int height, int width, int channels, int out_height, int out_width,
                       double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Matching (APP/utils/keypoint_descriptor.py:330-349)
 * ------------------------------------------------------------------------------------------- */

/* cv2.BFMatcher().knnMatch(query, train, k=2) + the ratio test: for each of n_query float32
 * descriptors the two nearest of n_train in L2, ties to the lower train index.
 * Candidate search: fp16 tcgen05 GEMM (|a-b|^2 = |a|^2 + |b|^2 - 2ab) with a fused running top-2;
 * the two winners are then re-evaluated in exact float32 so distances and the ratio test match an
 * fp32 brute-force matcher.  knn_idx: [n_query][2] int32, knn_dist: [n_query][2] float32,
 * good: [n_query] uint8 = dist0 < ratio * dist1 (evaluated in double, strict).  workspace must hold
 * fc_match_workspace_bytes(n_query, n_train) bytes. */
FC_API int64_t fc_match_workspace_bytes(int n_query, int n_train);
/* byte offset inside the workspace of four uint32 diagnostics written by fc_match_knn2: bits of max |t|^2,
 * fp16-overflow flag (1 = whole call re-done exactly), number of rows re-checked exactly, reserved. */
FC_API int64_t fc_match_workspace_stats_offset(int n_query, int n_train);
FC_API int fc_match_knn2(const float* query, int n_query, const float* train, int n_train, double ratio,
                  int32_t* knn_idx, float* knn_dist, uint8_t* good, void* workspace, void* stream);
/* exact float32 brute force on CUDA cores (same outputs); the verification twin of fc_match_knn2 */
FC_API int fc_match_knn2_exact(const float* query, int n_query, const float* train, int n_train, double ratio,
                        int32_t* knn_idx, float* knn_dist, uint8_t* good, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FEATURECRAFT_B200_H */
