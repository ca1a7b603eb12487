"""ctypes binding of libhpdecode.so (include/hpdecode.h).

The library is the product: there is NO CPU or eager-PyTorch fallback.  ``lib()`` raises if the
shared object is missing; every op raises if it is handed non-CUDA tensors.
"""
import ctypes
import os

HPD_MAX_KPTS = 32
HPD_MAX_PEOPLE = 32
HPD_MAX_EMB = 2
HPD_MAX_SCALES = 4
HPD_ABI_VERSION = 2
HPD_F32, HPD_F16 = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhpdecode.so")

EXPORTS = (
    "hpd_abi_version", "hpd_last_error_string", "hpd_workspace_bytes", "hpd_aggregate_nms", "hpd_nms",
    "hpd_topk", "hpd_group", "hpd_adjust_refine", "hpd_decode", "hpd_last_launch_count", "hpd_resize_bilinear",
    "hpd_record_layout", "hpd_multi_scale_size", "hpd_get_affine_transform", "hpd_prepare_input", "hpd_prepare_geometry",
)


class HpdMap(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_b", ctypes.c_int64), ("stride_c", ctypes.c_int64),
                ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("dtype", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


class HpdScaleInputs(ctypes.Structure):
    _fields_ = [("hm_lo", HpdMap), ("hm_hi", HpdMap), ("tag", HpdMap),
                ("hm_lo_f", HpdMap), ("hm_hi_f", HpdMap), ("tag_f", HpdMap)]


class HpdParams(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("num_kpts", ctypes.c_int32), ("out_h", ctypes.c_int32),
                ("out_w", ctypes.c_int32), ("emb", ctypes.c_int32), ("max_people", ctypes.c_int32),
                ("num_scales", ctypes.c_int32), ("tag_scale", ctypes.c_int32), ("do_adjust", ctypes.c_int32),
                ("do_refine", ctypes.c_int32), ("tags_preflipped", ctypes.c_int32), ("force_generic", ctypes.c_int32),
                ("batches_in_flight", ctypes.c_int32), ("reserved_", ctypes.c_int32),
                ("det_thr", ctypes.c_double), ("tag_thr", ctypes.c_double),
                ("flip_index", ctypes.c_int32 * HPD_MAX_KPTS), ("joints_order", ctypes.c_int32 * HPD_MAX_KPTS)]


class HpdBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "agg_hm", "agg_tags", "nms_mask", "nms_wmax", "hm_wmax", "scores_k", "idx_k", "coords_k", "tags_k",
        "poses", "person_scores", "n_person", "flags", "tag_bmin", "tag_bmax", "records", "inv_affine")]


class HpdRecordLayout(ctypes.Structure):
    _fields_ = [("row_bytes", ctypes.c_int64), ("off_coco", ctypes.c_int64), ("off_poses", ctypes.c_int64),
                ("off_person_scores", ctypes.c_int64), ("off_n_person", ctypes.c_int64), ("off_flags", ctypes.c_int64),
                ("coco_stride", ctypes.c_int32), ("reserved_", ctypes.c_int32)]


class HpdImage(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride_row", ctypes.c_int64), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("m", ctypes.c_double * 6)]


class HpdError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libhpdecode.so or fail loudly (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HpdError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  hpdecode has no CPU / eager fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.hpd_abi_version.restype = ctypes.c_int
    L.hpd_last_error_string.restype = ctypes.c_char_p
    L.hpd_last_launch_count.restype = ctypes.c_int
    P, S, B, V = ctypes.POINTER(HpdParams), ctypes.POINTER(HpdScaleInputs), ctypes.POINTER(HpdBuffers), ctypes.c_void_p
    L.hpd_workspace_bytes.argtypes = [P, ctypes.POINTER(ctypes.c_size_t)]
    L.hpd_aggregate_nms.argtypes = [P, S, B, V]
    L.hpd_nms.argtypes = [P, B, V, V]
    L.hpd_topk.argtypes = [P, B, V]
    L.hpd_group.argtypes = [P, B, V]
    L.hpd_adjust_refine.argtypes = [P, B, V, ctypes.c_size_t, V]
    L.hpd_decode.argtypes = [P, S, B, V, ctypes.c_size_t, V]
    L.hpd_resize_bilinear.argtypes = [ctypes.POINTER(HpdMap), ctypes.c_int, ctypes.c_int, V, ctypes.c_int,
                                      ctypes.c_int, V]
    I32P, F64P, F32P = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_float)
    L.hpd_record_layout.argtypes = [P, ctypes.POINTER(HpdRecordLayout)]
    L.hpd_multi_scale_size.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, I32P,
                                       I32P, F64P]
    L.hpd_get_affine_transform.argtypes = [F64P, F64P, I32P, ctypes.c_int, F64P]
    L.hpd_prepare_geometry.argtypes = [ctypes.c_int, I32P, I32P, ctypes.c_int, ctypes.c_double, ctypes.c_double, I32P, I32P, F64P,
                                       F64P, F64P]
    L.hpd_prepare_input.argtypes = [ctypes.POINTER(HpdImage), ctypes.c_int, V, ctypes.c_int, ctypes.c_int, F32P, F32P, V]
    for n in EXPORTS:
        if n not in ("hpd_last_error_string",):
            getattr(L, n).restype = ctypes.c_int
    if L.hpd_abi_version() != HPD_ABI_VERSION:
        raise HpdError(f"libhpdecode ABI {L.hpd_abi_version()} != binding {HPD_ABI_VERSION}")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        raise HpdError(f"{what} failed (code {rc}): {lib().hpd_last_error_string().decode()}")
