"""CPU suite: libhpdecode.so loads, exports every symbol include/hpdecode.h declares, and the ctypes
structs have the C layout (checked with a gcc-compiled probe).  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hpdecode.h")


def _declared():
    src = open(HEADER).read()
    return re.findall(r"HPD_EXPORT\s+[\w\s\*]+?\b(hpd_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    from hpdecode import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    names = _declared()
    assert len(names) >= 11 and set(names) == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert _lib.lib().hpd_abi_version() == _lib.HPD_ABI_VERSION


def test_ctypes_structs_match_c_layout(tmp_path):
    from hpdecode import _lib
    probe = tmp_path / "probe.c"
    probe.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hpdecode.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                     "sizeof(HpdMap),sizeof(HpdScaleInputs),sizeof(HpdParams),sizeof(HpdBuffers),"
                     "offsetof(HpdParams,det_thr),offsetof(HpdParams,joints_order),sizeof(HpdRecordLayout),sizeof(HpdImage),"
                     "offsetof(HpdImage,m),offsetof(HpdBuffers,records));return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.HpdMap), ctypes.sizeof(_lib.HpdScaleInputs), ctypes.sizeof(_lib.HpdParams),
            ctypes.sizeof(_lib.HpdBuffers), _lib.HpdParams.det_thr.offset, _lib.HpdParams.joints_order.offset,
            ctypes.sizeof(_lib.HpdRecordLayout), ctypes.sizeof(_lib.HpdImage), _lib.HpdImage.m.offset,
            _lib.HpdBuffers.records.offset]
    assert got == want


def test_argument_validation_without_gpu():
    """Validation happens before any launch, so the error path is testable on a CPU box."""
    from hpdecode import _lib, ops
    L = _lib.lib()
    p = ops.make_params(1, 17, 16, 16, 1, 30, 0.05, 0.5)     # 256 px < 64*30: torch topk regime differs
    b = _lib.HpdBuffers()
    rc = L.hpd_topk(ctypes.byref(p), ctypes.byref(b), None)
    assert rc == 1 and b"64*max_people" in L.hpd_last_error_string()
    p = ops.make_params(1, 17, 256, 256, 3, 30, 0.05, 0.5)
    assert L.hpd_group(ctypes.byref(p), ctypes.byref(b), None) == 1 and b"emb" in L.hpd_last_error_string()
    p = ops.make_params(1, 17, 256, 256, 1, 40, 0.05, 0.5)
    assert L.hpd_group(ctypes.byref(p), ctypes.byref(b), None) == 1 and b"max_people" in L.hpd_last_error_string()
    p = ops.make_params(1, 17, 256, 256, 1, 30, 0.05, 0.5)
    assert L.hpd_group(ctypes.byref(p), ctypes.byref(b), None) == 1 and b"required" in L.hpd_last_error_string()
    n = ctypes.c_size_t(0)
    assert L.hpd_workspace_bytes(ctypes.byref(p), ctypes.byref(n)) == 0 and n.value > 0


def test_new_entry_points_validate_before_any_launch():
    """ABI v2 additions: record layout, host geometry and the input kernel reject bad arguments with HPD_EINVAL and
    a message -- all before a launch, so this runs on a CPU box."""
    from hpdecode import _lib, ops
    L = _lib.lib()
    lay = _lib.HpdRecordLayout()
    for K, M, E in ((17, 30, 2), (8, 20, 1), (32, 32, 2), (1, 1, 1)):
        p = ops.make_params(1, K, 256, 256, E, M, 0.05, 0.5)
        assert L.hpd_record_layout(ctypes.byref(p), ctypes.byref(lay)) == 0
        assert lay.coco_stride == 3 * K + 1 and lay.off_coco == 0 and lay.off_poses == 8 * M * (3 * K + 1)
        assert lay.off_person_scores == lay.off_poses + 4 * M * K * (3 + E) and lay.off_n_person == lay.off_person_scores + 4 * M
        assert lay.off_flags == lay.off_n_person + 4 and lay.row_bytes % 8 == 0 and lay.row_bytes >= lay.off_flags + 4
    p = ops.make_params(1, 40, 256, 256, 1, 30, 0.05, 0.5)
    assert L.hpd_record_layout(ctypes.byref(p), ctypes.byref(lay)) == 1 and b"out of range" in L.hpd_last_error_string()
    size, center, scale = (ctypes.c_int32 * 2)(), (ctypes.c_int32 * 2)(), (ctypes.c_double * 2)()
    assert L.hpd_multi_scale_size(0, 640, 512, 1.0, 1.0, size, center, scale) == 1
    assert L.hpd_multi_scale_size(480, 640, 512, 1.0, 0.0, size, center, scale) == 1
    assert L.hpd_get_affine_transform(None, scale, size, 0, (ctypes.c_double * 6)()) == 1
    img = (_lib.HpdImage * 1)()
    mean, std = (ctypes.c_float * 3)(0.5, 0.5, 0.5), (ctypes.c_float * 3)(0.2, 0.2, 0.2)
    assert L.hpd_prepare_input(img, 1, None, 64, 64, mean, std, None) == 1                     # no output buffer
    assert L.hpd_prepare_input(img, 1, ctypes.c_void_p(8), 64, 64, mean, std, None) == 1       # image 0 has no pointer
    assert b"image 0" in L.hpd_last_error_string()
    img[0].ptr, img[0].h, img[0].w, img[0].stride_row = 8, 10, 10, 20                          # stride < 3 * w
    assert L.hpd_prepare_input(img, 1, ctypes.c_void_p(8), 64, 64, mean, std, None) == 1
    m = _lib.HpdMap()
    m.ptr, m.h, m.w, m.dtype = 8, 4, 4, 7                                                      # unknown dtype code
    assert L.hpd_resize_bilinear(ctypes.byref(m), 1, 1, ctypes.c_void_p(8), 8, 8, None) == 1


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from hpdecode import BottomUpDecoder, MPPEHeatmapParser, _lib
    with pytest.raises(_lib.HpdError):
        BottomUpDecoder()
    with pytest.raises(_lib.HpdError):
        MPPEHeatmapParser(17)
    with pytest.raises(_lib.HpdError):
        torch.ops.hpd.nms(torch.zeros(1, 17, 64, 64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pytorch-human-pose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in src and "py_port" not in src and "hpd_oracle" not in src, f
