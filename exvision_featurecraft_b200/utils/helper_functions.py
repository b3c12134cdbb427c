"""GPU-backed twins of FeatureCraft-PyQt-App/utils/helper_functions.py (same names and arguments)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..device import ptr, stream_ptr, to_device
from .harris_utils import convert_to_gray  # noqa: F401  (helper_functions.py:4-9 duplicates harris_utils.convert_to_gray)


def convert_BGR_to_RGB(img_BGR_nd_arr):
    """helper_functions.py:12-14 -- img[..., ::-1] of a cv2.imread result (APP/FeatureCraft_Backend.py:105-106).
    uint8 [H,W,3] numpy array or device tensor; returns the same kind."""
    on_device = isinstance(img_BGR_nd_arr, torch.Tensor)
    t = to_device(img_BGR_nd_arr if on_device else np.asarray(img_BGR_nd_arr), torch.uint8)
    if t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("expected an [H, W, 3] uint8 image")
    out = torch.empty_like(t)
    _lib.call("fc_bgr_to_rgb_u8", ptr(t), 1, t.shape[0], t.shape[1], ptr(out), stream_ptr())
    return out if on_device else out.cpu().numpy()


def load_image(file_path):
    """BackendClass.load_image's arithmetic (APP/FeatureCraft_Backend.py:103-106): cv2.imread(path, 1) on the host (file
    decoding is outside the device path), BGR -> RGB on the device.  Returns the uint8 RGB device tensor."""
    import cv2

    img = cv2.imread(file_path, 1)
    if img is None:
        raise FileNotFoundError(file_path)
    return convert_BGR_to_RGB(to_device(img, torch.uint8))
