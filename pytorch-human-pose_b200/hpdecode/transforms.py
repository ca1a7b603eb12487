"""Host-side geometry used by the result objects (a "next" row):
back-projection of decoded coordinates to the raw image.

Mirrors /root/reference/src/base/transforms/utils.py:5-57 (``affine_transform``,
``get_affine_transform``) and transforms.py:11 (``COCO_FLIP_INDEX``).  The 2x3 matrix is obtained
from cv2.getAffineTransform exactly as the reference does when OpenCV is importable; otherwise the
same three-point system is solved with numpy (float64).
"""
import numpy as np

COCO_FLIP_INDEX = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]

try:  # the reference depends on OpenCV; it is optional here
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None


def affine_transform(point, transform_matrix: np.ndarray) -> np.ndarray:
    """utils.py:5-8: [x, y] -> M @ [x, y, 1]."""
    return (transform_matrix @ np.array([point[0], point[1], 1.0]))[:2]


def _perp_third(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def _three_point_affine(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    if _cv2 is not None:
        return _cv2.getAffineTransform(src, dst)
    A = np.concatenate([src.astype(np.float64), np.ones((3, 1))], axis=1)
    return np.linalg.solve(A, dst.astype(np.float64)).T


def get_affine_transform(center, scale, rot: float, output_size, shift=(0, 0), inverse: bool = False) -> np.ndarray:
    """utils.py:25-57: similarity transform mapping the (center, scale) box onto ``output_size``."""
    shift = np.array(shift)
    scale = np.array(scale)
    center = np.array(center)
    dst_w, dst_h = output_size[0], output_size[1]
    ang = np.pi * rot / 180
    sn, cs = np.sin(ang), np.cos(ang)
    half = -scale[0] / 2
    src_dir = (0 * cs - half * sn, 0 * sn + half * cs)
    dst_dir = np.array([0, -dst_w / 2], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center + scale * shift
    src[1] = center + src_dir + scale * shift
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2] = _perp_third(src[0], src[1])
    dst[2] = _perp_third(dst[0], dst[1])
    if inverse:
        src, dst = dst, src
    return _three_point_affine(src, dst)
