"""GPU-backed twins of FeatureCraft-PyQt-App/utils/harris_utils.py (same names and arguments).

numpy in, numpy out -- like the reference -- with the arithmetic done by libfeaturecraft_b200.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..corners import rgb_to_gray
from ..device import ptr, stream_ptr, to_device


def convert_to_gray(img_RGB):
    """harris_utils.py:4-9 -- weighted sum truncated to uint8; 2-D input is just cast."""
    img_RGB = np.asarray(img_RGB)
    if img_RGB.ndim == 3:
        if img_RGB.dtype != np.uint8:
            raise TypeError("the device path converts uint8 RGB images (what the app loads)")
        return rgb_to_gray(img_RGB).cpu().numpy()
    return img_RGB.astype(np.uint8)


def padding_matrix(matrix, width, height, pad_size):
    """harris_utils.py:12-30 -- float64 zero frame (device memset + strided copy)."""
    m = to_device(np.asarray(matrix), torch.float64)
    out = torch.zeros((height + 2 * pad_size, width + 2 * pad_size), dtype=torch.float64, device=m.device)
    out[pad_size:pad_size + height, pad_size:pad_size + width] = m
    return out.cpu().numpy()


def convolve2d_optimized(input_matrix, convolution_kernel, mode="same"):
    """harris_utils.py:33-80 -- zero-padded correlation; `mode` is ignored, as in the reference."""
    img = np.asarray(input_matrix)
    ker = np.asarray(convolution_kernel, dtype=np.float64)
    if img.ndim != 2:
        raise ValueError("not enough values to unpack (expected 2, got %d)" % img.ndim)
    ks = ker.shape[0]
    if ker.ndim != 2 or ker.shape != (2 * (ks // 2) + 1,) * 2:
        raise ValueError("cannot reshape array: kernel must be square with odd size")
    d_img = to_device(img, torch.float64)
    d_ker = to_device(ker)
    out = torch.empty_like(d_img)
    h, w = img.shape
    _lib.call("fc_correlate_zero_pad_f64", ptr(d_img), 1, h, w, ptr(d_ker), ks, ptr(out), stream_ptr())
    return out.cpu().numpy()
