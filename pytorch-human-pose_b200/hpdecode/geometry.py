"""Image <-> network-input geometry of the inference path, computed by libhpdecode.so (csrc/input.cu).

Python twins of /root/reference/src/base/transforms/utils.py:25-97 as far as the decode path uses them:
``get_multi_scale_size`` / ``get_affine_transform`` are float64 host code inside the library (the matrix comes
out bit-identical to cv2.getAffineTransform); ``resize_align_multi_scale`` + ToTensor + Normalize
(model.py:45-50,70-76) are ONE device kernel, ``prepare_input``, that replays cv2.warpAffine's fixed-point bilinear
arithmetic, so the tensor handed to the network is bit-identical to the reference's.
"""
import ctypes
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

COCO_FLIP_INDEX = list(ops.COCO_FLIP_INDEX)   # /root/reference/src/keypoints/transforms.py:11
MEAN = (0.485, 0.456, 0.406)                  # model.py:48
STD = (0.229, 0.224, 0.225)


def get_multi_scale_size(image, input_size: int, current_scale: float, min_scale: float):
    """utils.py:60-87.  ``image``: array [h,w,3] or an (h, w) pair.  -> ((w_resized, h_resized), center, scale)."""
    h, w = image.shape[:2] if hasattr(image, "shape") else image
    size, center, scale = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_double * 2)()
    _lib.check(_lib.lib().hpd_multi_scale_size(int(h), int(w), int(input_size), float(current_scale), float(min_scale),
                                               size, center, scale), "hpd_multi_scale_size")
    return (size[0], size[1]), (center[0], center[1]), (scale[0], scale[1])


def get_affine_transform(center, scale, rot: float, output_size, shift=(0, 0), inverse: bool = False) -> np.ndarray:
    """utils.py:25-57 for the call shapes of this path (rot = 0, shift = 0): float64 [2,3]."""
    if rot != 0 or tuple(shift) != (0, 0):
        raise _lib.HpdError("get_affine_transform: the decode path only uses rot = 0, shift = (0, 0)")
    c = (ctypes.c_double * 2)(float(center[0]), float(center[1]))
    s = (ctypes.c_double * 2)(float(scale[0]), float(scale[1]))
    o = (ctypes.c_int32 * 2)(int(output_size[0]), int(output_size[1]))
    m = (ctypes.c_double * 6)()
    _lib.check(_lib.lib().hpd_get_affine_transform(c, s, o, int(bool(inverse)), m), "hpd_get_affine_transform")
    return np.array(m[:], np.float64).reshape(2, 3)


def prepare_input(images: Sequence, input_size: int, device, current_scale: float = 1, min_scale: float = 1,
                  mean=MEAN, std=STD):
    """Batched ``InferenceKeypointsModel.prepare_input`` (model.py:70-76) on the device.

    images: uint8 [h,w,3] arrays (host) or CUDA tensors, all mapping to the SAME resized size (group them with
    ``get_multi_scale_size`` first).  Returns (x [B,3,H,W] float32 on ``device``, centers, scales)."""
    device = torch.device(device)
    geo = [get_multi_scale_size(im, input_size, current_scale, min_scale) for im in images]
    size = geo[0][0]
    if any(g[0] != size for g in geo):
        raise _lib.HpdError("prepare_input: images of one call must share the resized size; got %s" % sorted({g[0] for g in geo}))
    dev_imgs, mats = [], []
    for im, (_, center, scale) in zip(images, geo):
        t = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im))
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise _lib.HpdError("prepare_input: images must be uint8 [h,w,3]")
        dev_imgs.append(t.to(device, non_blocking=True).contiguous())
        mats.append(get_affine_transform(center, scale, 0, size).ravel())
    x = torch.ops.hpd.prepare_input(dev_imgs, torch.from_numpy(np.stack(mats)), size[1], size[0], list(mean), list(std))
    return x, [g[1] for g in geo], [g[2] for g in geo]


def group_by_resized_size(shapes_hw: Sequence[Tuple[int, int]], input_size: int, current_scale: float = 1,
                          min_scale: float = 1) -> dict:
    """{(w_resized, h_resized): [indices]} -- images that can share one batched network call."""
    groups: dict = {}
    for i, hw in enumerate(shapes_hw):
        groups.setdefault(get_multi_scale_size(hw, input_size, current_scale, min_scale)[0], []).append(i)
    return groups
