"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the input side of the inference path.

Follows, line by line in spirit (not in code):
  get_multi_scale_size      /root/reference/src/base/transforms/utils.py:60-87
  get_affine_transform      /root/reference/src/base/transforms/utils.py:25-57 (rot = 0, shift = 0) with
                            cv2.getAffineTransform = cv::solve(DECOMP_LU) on the 6x6 system (OpenCV 4.13,
                            modules/imgproc/src/imgwarp.cpp, modules/core/src/matrix_decomp.cpp: LUImpl)
  warp_affine               cv2.warpAffine(img, M, dsize) as resize_align_multi_scale calls it (utils.py:96):
                            INTER_LINEAR, BORDER_CONSTANT 0; OpenCV's fixed-point path (AB_BITS 10, INTER_BITS 5,
                            INTER_REMAP_COEF_BITS 15)
  to_tensor_normalize       T.ToTensor + T.Normalize, /root/reference/src/keypoints/model.py:45-50
  affine_points             transform_coords / affine_transform, results.py:158-171 + utils.py:5-8: np.dot of the
                            float64 2x3 matrix with [x, y, 1.0]

PINNED: tests/test_oracle_input.py checks every function against the reference's own functions, cv2 4.13 and
torchvision 0.26 in the build container (bit-exact), and against tests/golden/input_cases.npz everywhere.
Allowed importers: tests/, __graft_entry__.smoke(), bench.py's checker legs.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)   # model.py:48
STD = (0.229, 0.224, 0.225)


def get_multi_scale_size(h: int, w: int, input_size: int, current_scale: float, min_scale: float):
    center = (int(w / 2.0 + 0.5), int(h / 2.0 + 0.5))
    min_input_size = int((min_scale * input_size + 63) // 64 * 64)
    if w < h:
        w_r = int(min_input_size * current_scale / min_scale)
        h_r = int(int((min_input_size / w * h + 63) // 64 * 64) * current_scale / min_scale)
        scale = (float(w), h_r / w_r * w)
    else:
        h_r = int(min_input_size * current_scale / min_scale)
        w_r = int(int((min_input_size / h * w + 63) // 64 * 64) * current_scale / min_scale)
        scale = (w_r / h_r * h, float(h))
    return (w_r, h_r), center, scale


def _lu_solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """OpenCV's LUImpl<double>: partial pivoting, elimination with alpha = A[j][i] * (-1/A[i][i])."""
    A = A.astype(np.float64).copy()
    b = b.astype(np.float64).copy()
    m = A.shape[0]
    for i in range(m):
        k = i
        for j in range(i + 1, m):
            if abs(A[j, i]) > abs(A[k, i]):
                k = j
        if k != i:
            A[[i, k], i:] = A[[k, i], i:]
            b[[i, k]] = b[[k, i]]
        d = -1 / A[i, i]
        for j in range(i + 1, m):
            alpha = A[j, i] * d
            for c in range(i + 1, m):
                A[j, c] += alpha * A[i, c]
            b[j] += alpha * b[i]
    for i in range(m - 1, -1, -1):
        s = b[i]
        for c in range(i + 1, m):
            s -= A[i, c] * b[c]
        b[i] = s / A[i, i]
    return b


def get_affine_transform(center, scale, output_size, inverse: bool = False) -> np.ndarray:
    cx, cy = float(center[0]), float(center[1])
    dst_w, dst_h = output_size
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = (cx, cy)
    src[1] = (cx + 0.0, cy + (-float(scale[0]) / 2))
    dst[0] = (dst_w * 0.5, dst_h * 0.5)
    dst[1] = (dst_w * 0.5 + 0.0, dst_h * 0.5 + float(np.float32(-dst_w / 2)))
    for pts in (src, dst):
        d = pts[0] - pts[1]
        pts[2] = pts[1] + np.array([-d[1], d[0]], np.float32)
    if inverse:
        src, dst = dst, src
    A = np.zeros((6, 6))
    b = np.zeros(6)
    for i in range(3):
        A[2 * i, 0:3] = (src[i, 0], src[i, 1], 1)
        A[2 * i + 1, 3:6] = (src[i, 0], src[i, 1], 1)
        b[2 * i], b[2 * i + 1] = dst[i]
    return _lu_solve(A, b).reshape(2, 3)


def _cv_round_sat(v):
    return np.clip(np.rint(v), -2.0 ** 31, 2.0 ** 31 - 1).astype(np.int64)


def warp_affine(img: np.ndarray, M: np.ndarray, dsize) -> np.ndarray:
    """img uint8 [H,W,C]; M float64 2x3 forward matrix; dsize (w, h)."""
    dw, dh = dsize
    H, W, C = img.shape
    m = np.asarray(M, np.float64).ravel().copy()
    D = m[0] * m[4] - m[1] * m[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = m[4] * D, m[0] * D
    m[0] = A11
    m[1] *= -D
    m[3] *= -D
    m[4] = A22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    xs = np.arange(dw, dtype=np.float64)
    ys = np.arange(dh, dtype=np.float64)
    adelta = _cv_round_sat(m[0] * xs * 1024)
    bdelta = _cv_round_sat(m[3] * xs * 1024)
    X0 = _cv_round_sat((m[1] * ys + m[2]) * 1024) + 16
    Y0 = _cv_round_sat((m[4] * ys + m[5]) * 1024) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    ax, ay = X & 31, Y & 31
    w = [(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32]
    pad = np.zeros((H + 2, W + 2, C), np.int64)
    pad[1:-1, 1:-1] = img

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = pad[np.clip(yy, -1, H) + 1, np.clip(xx, -1, W) + 1]
        return np.where(ok[..., None], v, 0)

    acc = (tap(sy, sx) * w[0][..., None] + tap(sy, sx + 1) * w[1][..., None] + tap(sy + 1, sx) * w[2][..., None] +
           tap(sy + 1, sx + 1) * w[3][..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def to_tensor_normalize(img: np.ndarray, mean=MEAN, std=STD) -> np.ndarray:
    """uint8 [H,W,3] -> float32 [3,H,W]: (x / 255 - mean) / std, every step rounded to float32."""
    x = img.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    mean = np.asarray(mean, np.float32)[:, None, None]
    std = np.asarray(std, np.float32)[:, None, None]
    return ((x - mean) / std).astype(np.float32)


def prepare_input(image: np.ndarray, input_size: int, current_scale: float = 1, min_scale: float = 1):
    """model.py:70-76: (x [3,h,w] float32, center, scale, size_resized (w,h), M)."""
    size, center, scale = get_multi_scale_size(image.shape[0], image.shape[1], input_size, current_scale, min_scale)
    M = get_affine_transform(center, scale, size)
    return to_tensor_normalize(warp_affine(image, M, size)), center, scale, size, M


def _fma(a: float, b: float, c: float) -> float:
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def affine_points(xy: np.ndarray, M: np.ndarray) -> np.ndarray:
    """np.dot(M, [x, y, 1.0]) for every row of xy [n,2] (float64 in, float64 out).

    NumPy hands the 2x3 @ 3 product to OpenBLAS dgemv; on the AVX-512 / Haswell kernels of OpenBLAS 0.3.30 (the
    build container, numpy 2.3) each output is  fma(m2, 1.0, fma(m0, x, m1*y))  -- found by exhaustive search over
    accumulation orders, 5000/5000 random cases (tests/test_oracle_input.py re-checks it against np.dot)."""
    M = np.asarray(M, np.float64)
    out = np.empty((len(xy), 2))
    for i, (x, y) in enumerate(np.asarray(xy, np.float64)):
        for r in range(2):
            out[i, r] = _fma(M[r, 2], 1.0, _fma(M[r, 0], x, M[r, 1] * y))
    return out
