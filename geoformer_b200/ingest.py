"""Image ingest for the matcher wrappers — GPU version of
``eval_tool/immatch/utils/data_io.py::load_gray_scale_tensor_cv`` (lines 48-62) and ``resize_im`` (16-26).

The decode (cv2.imread) stays on the host; the uint8 image is uploaded from pinned memory (4x fewer bytes than the
fp32 tensor) and resized + normalised on the GPU by a kernel that reproduces cv2.resize(INTER_LINEAR, uint8)
bit-exactly, so the model sees the same tensor as with the reference loader."""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch

from . import ops


def resize_dims(wo: int, ho: int, imsize=None, dfactor: int = 1, value_to_scale: Callable = max,
                aspan: bool = False) -> Tuple[int, int, Tuple[float, float]]:
    """Target size and the (x, y) up-scale factors (data_io.py:16-26): scale so that value_to_scale(w, h) == imsize
    when it is larger, then floor both sides to a multiple of dfactor."""
    wt, ht = wo, ho
    if imsize and (aspan or value_to_scale(wo, ho) > imsize) and imsize > 0:
        scale = imsize / value_to_scale(wo, ho)
        ht, wt = int(round(ho * scale)), int(round(wo * scale))
    wt, ht = int(wt // dfactor * dfactor), int(ht // dfactor * dfactor)
    return wt, ht, (wo / wt, ho / ht)


def gray_to_tensor(im: np.ndarray, device: torch.device, imsize=None, dfactor: int = 8,
                   value_to_scale: Callable = min, aspan: bool = False, out: Optional[torch.Tensor] = None):
    """uint8 [ho, wo] grayscale image -> (fp32 [1,1,ht,wt] CUDA tensor in [0,1], (sx, sy)).
    out: optional fp32 [ht, wt] destination on the device (e.g. one image of a batch tensor) - then `out` is returned."""
    assert im.dtype == np.uint8 and im.ndim == 2
    ho, wo = im.shape
    wt, ht, scale = resize_dims(wo, ho, imsize=imsize, dfactor=dfactor, value_to_scale=value_to_scale, aspan=aspan)
    host = torch.from_numpy(np.ascontiguousarray(im))
    src = (host.pin_memory() if device.type == "cuda" else host).to(device, non_blocking=True)
    if out is not None:
        assert tuple(out.shape) == (ht, wt), (tuple(out.shape), (ht, wt))
        dst = out
    else:
        dst = torch.empty((1, 1, ht, wt), device=device, dtype=torch.float32)
    ops.ensure_init(device)
    ops.resize_gray_u8(src, dst if out is not None else dst[0, 0])
    return dst, scale


def load_gray_scale_tensor_gpu(im_path: str, device: Union[str, torch.device], imsize=None, dfactor: int = 8,
                               enhanced: bool = False, value_to_scale: Callable = min, aspan: bool = False):
    """Drop-in for load_gray_scale_tensor_cv(im_path, device, imsize, dfactor, ...) -> (tensor, scale)."""
    import cv2
    im = cv2.imread(im_path, cv2.IMREAD_GRAYSCALE)
    if im is None:
        raise FileNotFoundError(im_path)
    return gray_to_tensor(im, torch.device(device), imsize, dfactor, value_to_scale, aspan)
