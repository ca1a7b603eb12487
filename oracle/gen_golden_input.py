"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/input_cases.npz by running the reference's own input-side and
back-projection functions in the build container: resize_align_multi_scale / get_affine_transform / affine_transform
from the UNMODIFIED /root/reference/src/base/transforms/utils.py (cv2 4.13), the torchvision transform of
/root/reference/src/keypoints/model.py:45-50, and the loop of transform_coords (results.py:158-171; results.py itself
cannot be imported here, see oracle/ref_runner.py).

    python oracle/gen_golden_input.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = os.environ.get("HP_REFERENCE", "/root/reference")

# (h, w, input_size, seed): landscape, portrait, square, upscaling, odd sizes, tiny
CASES = [(480, 640, 512, 1), (640, 427, 512, 2), (512, 512, 512, 3), (200, 333, 512, 4), (719, 1280, 512, 5),
         (1333, 800, 640, 6), (97, 61, 256, 7), (375, 500, 384, 8)]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def image_of(h, w, seed):
    """Seeded uint8 test image with smooth structure plus noise (so sub-pixel weights matter)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(xx / 17.0 + seed) * np.cos(yy / 23.0)
    img = base[..., None] + rng.integers(-27, 28, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def coords_of(size_wh, seed, persons=5, K=17):
    rng = np.random.default_rng(1000 + seed)
    xy = np.stack([rng.integers(0, size_wh[0], (persons, K)), rng.integers(0, size_wh[1], (persons, K))], -1)
    return (xy + rng.choice([0.25, 0.75], xy.shape)).astype(np.float32)


def main():
    sys.path.insert(0, REFERENCE_ROOT)
    import cv2
    import torchvision
    import torchvision.transforms as T
    from src.base.transforms.utils import affine_transform, get_affine_transform, resize_align_multi_scale
    transform = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out = {"versions": np.array([cv2.__version__, torchvision.__version__, np.__version__])}
    for i, (h, w, input_size, seed) in enumerate(CASES):
        img = image_of(h, w, seed)
        resized, center, scale = resize_align_multi_scale(img, input_size, 1, 1)
        size = (resized.shape[1], resized.shape[0])
        x = transform(resized).numpy()
        M = get_affine_transform(center, scale, 0, size)
        Minv = get_affine_transform(center, scale, 0, size, inverse=True)
        kpts = coords_of(size, seed)
        back32 = kpts.copy()                     # results.py:165-170: float32 array, float64 values written into it
        back64 = kpts.astype(np.float64)         # the empty-scene fallback's float64 pseudo-person takes this path
        for p in range(kpts.shape[0]):
            for k in range(kpts.shape[1]):
                back32[p, k, :2] = affine_transform(kpts[p, k, :2].tolist(), Minv)
                back64[p, k, :2] = affine_transform(back64[p, k, :2].tolist(), Minv)
        out.update({f"c{i}_hw_in_seed": np.array([h, w, input_size, seed]), f"c{i}_size": np.array(size),
                    f"c{i}_center": np.array(center), f"c{i}_scale": np.array(scale, np.float64), f"c{i}_M": M,
                    f"c{i}_Minv": Minv, f"c{i}_warped_sha": np.array(sha(resized)), f"c{i}_x_sha": np.array(sha(x)),
                    f"c{i}_x_head": x[:, :2, :8].copy(), f"c{i}_kpts": kpts, f"c{i}_back32": back32, f"c{i}_back64": back64})
        print(i, (h, w), "->", size, center, scale, flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "input_cases.npz"), **out)


if __name__ == "__main__":
    main()
