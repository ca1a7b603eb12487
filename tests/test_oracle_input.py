"""CPU suite: the input-side / back-projection oracle (oracle/input_oracle.py) against the goldens recorded from
the reference's own functions + cv2 + torchvision (oracle/gen_golden_input.py), against the live reference when
/root/reference is present, and the library's HOST geometry entry points (hpd_multi_scale_size,
hpd_get_affine_transform: float64 host code, no GPU needed) against both."""
import ctypes
import os
import sys

import numpy as np
import pytest

from oracle import input_oracle
from oracle.gen_golden_input import CASES, coords_of, image_of, sha

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_cases.npz"))
REF = os.environ.get("HP_REFERENCE", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "src", "base", "transforms", "utils.py")),
                                     reason="build container only (/root/reference)")


@pytest.mark.parametrize("i", range(len(CASES)))
def test_input_oracle_matches_reference_goldens(i):
    h, w, input_size, seed = CASES[i]
    img = image_of(h, w, seed)
    x, center, scale, size, M = input_oracle.prepare_input(img, input_size)
    assert tuple(GOLD[f"c{i}_size"]) == size and tuple(GOLD[f"c{i}_center"]) == center
    assert np.array_equal(GOLD[f"c{i}_scale"], np.array(scale, np.float64))
    assert np.array_equal(GOLD[f"c{i}_M"], M)                                   # float64, bit-exact (cv2's LU order)
    Minv = input_oracle.get_affine_transform(center, scale, size, inverse=True)
    assert np.array_equal(GOLD[f"c{i}_Minv"], Minv)
    assert sha(input_oracle.warp_affine(img, M, size)) == str(GOLD[f"c{i}_warped_sha"])   # uint8, bit-exact
    assert sha(x) == str(GOLD[f"c{i}_x_sha"]) and np.array_equal(x[:, :2, :8], GOLD[f"c{i}_x_head"])
    kpts = GOLD[f"c{i}_kpts"]
    back = input_oracle.affine_points(kpts.reshape(-1, 2), Minv).reshape(kpts.shape)
    assert np.array_equal(back.astype(np.float32), GOLD[f"c{i}_back32"])
    assert np.array_equal(back, GOLD[f"c{i}_back64"])


def _lib():
    from hpdecode import _lib
    return _lib.lib()


@pytest.mark.parametrize("i", range(len(CASES)))
def test_library_host_geometry_matches_goldens(i):
    L = _lib()
    h, w, input_size, _ = CASES[i]
    size, center, scale = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_double * 2)()
    assert L.hpd_multi_scale_size(h, w, input_size, 1.0, 1.0, size, center, scale) == 0
    assert tuple(size) == tuple(GOLD[f"c{i}_size"]) and tuple(center) == tuple(GOLD[f"c{i}_center"])
    assert np.array_equal(np.array(scale[:]), GOLD[f"c{i}_scale"])
    c = (ctypes.c_double * 2)(*[float(v) for v in center])
    for inverse, key in ((0, "M"), (1, "Minv")):
        m = (ctypes.c_double * 6)()
        assert L.hpd_get_affine_transform(c, scale, size, inverse, m) == 0
        assert np.array_equal(np.array(m[:]).reshape(2, 3), GOLD[f"c{i}_{key}"])


def test_library_host_geometry_many_sizes_against_oracle():
    """Every (h, w) class the size logic distinguishes (w < h, w >= h, multiples of 64 or not, scales != 1)."""
    L = _lib()
    rng = np.random.default_rng(5)
    for _ in range(400):
        h, w = (int(v) for v in rng.integers(40, 2200, 2))
        input_size = int(rng.choice([256, 384, 512, 640]))
        cur, mn = (float(v) for v in rng.choice([(1, 1), (0.5, 0.5), (1.0, 0.5), (1.5, 0.5), (2.0, 1.0)]))
        size, center, scale = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_double * 2)()
        assert L.hpd_multi_scale_size(h, w, input_size, cur, mn, size, center, scale) == 0
        want = input_oracle.get_multi_scale_size(h, w, input_size, cur, mn)
        assert (tuple(size), tuple(center), tuple(scale)) == want, (h, w, input_size, cur, mn)
        c = (ctypes.c_double * 2)(*[float(v) for v in center])
        for inverse in (0, 1):
            m = (ctypes.c_double * 6)()
            L.hpd_get_affine_transform(c, scale, size, inverse, m)
            assert np.array_equal(np.array(m[:]).reshape(2, 3),
                                  input_oracle.get_affine_transform(want[1], want[2], want[0], bool(inverse)))


def test_batched_geometry_entry_point_equals_the_single_image_ones():
    """hpd_prepare_geometry (one call for n images) == hpd_multi_scale_size + hpd_get_affine_transform per image."""
    from hpdecode import geometry
    rng = np.random.default_rng(11)
    shapes = [(int(h), int(w)) for h, w in rng.integers(40, 1500, (50, 2))]
    for cur, mn in ((1, 1), (1.5, 0.5)):
        sizes, centers, scales, fwd, inv = geometry.prepare_geometry(shapes, 512, cur, mn)
        for i, hw in enumerate(shapes):
            size, center, scale = geometry.get_multi_scale_size(hw, 512, cur, mn)
            assert tuple(sizes[i]) == size and tuple(centers[i]) == center and tuple(scales[i]) == scale
            assert np.array_equal(fwd[i].reshape(2, 3), geometry.get_affine_transform(center, scale, 0, size))
            assert np.array_equal(inv[i].reshape(2, 3), geometry.get_affine_transform(center, scale, 0, size, inverse=True))
    groups = geometry.group_by_resized_size([(480, 640), (640, 480), (479, 640)], 512)
    assert groups == {(704, 512): [0, 2], (512, 704): [1]}


@needs_reference
def test_input_oracle_against_live_reference_cv2_and_torchvision():
    sys.path.insert(0, REF)
    import cv2
    import torchvision.transforms as T
    from src.base.transforms.utils import affine_transform, get_affine_transform, get_multi_scale_size, resize_align_multi_scale
    tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    rng = np.random.default_rng(9)
    for t in range(25):
        h, w = (int(v) for v in rng.integers(50, 900, 2))
        input_size = int(rng.choice([256, 320, 512]))
        cur, mn = (1, 1) if t % 3 else (1.5, 0.5)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        size, center, scale = get_multi_scale_size(img, input_size, cur, mn)
        assert input_oracle.get_multi_scale_size(h, w, input_size, cur, mn) == (size, center, tuple(float(s) for s in scale))
        for inv in (False, True):
            assert np.array_equal(get_affine_transform(center, scale, 0, size, inverse=inv),
                                  input_oracle.get_affine_transform(center, scale, size, inv))
        M = get_affine_transform(center, scale, 0, size)
        want = cv2.warpAffine(img, M, size)
        got = input_oracle.warp_affine(img, M, size)
        assert np.array_equal(want, got)
        assert np.array_equal(tf(want).numpy(), input_oracle.to_tensor_normalize(got))
        Minv = get_affine_transform(center, scale, 0, size, inverse=True)
        pts = coords_of(size, t, persons=2).reshape(-1, 2)
        ref = np.array([affine_transform(p.tolist(), Minv) for p in pts])
        assert np.array_equal(ref, input_oracle.affine_points(pts, Minv))
    # resize_align_multi_scale as a whole, as model.py:73 calls it
    img = image_of(333, 517, 3)
    resized, center, scale = resize_align_multi_scale(img, 512, 1, 1)
    x, c2, s2, size, _ = input_oracle.prepare_input(img, 512)
    assert np.array_equal(tf(resized).numpy(), x) and center == c2 and tuple(float(s) for s in scale) == s2
