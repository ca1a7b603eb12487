"""Build-container only: the C++ oracle and the Python port against the UNMODIFIED reference
(/root/reference/src/keypoints/grouping.py imported as it lies) on fresh seeds.  Skipped on the
GPU box, where the reference does not exist; the committed goldens carry the pin there."""
import numpy as np
import pytest

from hpdecode import synth
from oracle import ref_runner

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="/root/reference not present")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("gen,kw,size,M,det,tthr", [
    ("netlike", dict(batch=1, size=192, flip=True, seed=101), 192, 30, 0.05, 0.5),
    ("crowd", dict(batch=1, size=192, persons=10, flip=True, seed=102, quantised=True), 192, 30, 0.05, 0.5),
    ("crowd", dict(batch=1, size=256, persons=25, flip=False, seed=103), 256, 20, 0.1, 1.0),
    ("crowd", dict(batch=1, size=256, persons=30, flip=True, seed=104, tag_spread=0.3), 256, 30, 0.05, 0.5),
])
def test_oracles_match_unmodified_reference(gen, kw, size, M, det, tthr, oracle):
    from oracle import py_port
    scales = getattr(synth, gen)(**kw)
    img = synth.image_slice(scales, 0)
    hm_t, tg_t = ref_runner.aggregate_torch(img, (size, size))
    R = ref_runner.parse_reference(hm_t, tg_t, M, det, tthr)
    hm_o, tg_o = oracle.aggregate(img, (size, size))
    assert np.array_equal(_bits(hm_o), _bits(hm_t.numpy())) and np.array_equal(_bits(tg_o), _bits(tg_t.numpy()))
    O = oracle.parse(hm_o, tg_o, M, det, tthr)
    assert np.array_equal(O["idx_k"], R["idx_k"])
    assert np.array_equal(_bits(O["grouped_joints"]), _bits(R["grouped_joints"]))
    assert np.array_equal(_bits(O["person_scores"]), _bits(R["person_scores"]))
    gj, ps = py_port.parse(hm_t, tg_t, M, det, tthr)
    assert np.array_equal(gj, R["grouped_joints"]) and np.array_equal(ps, R["person_scores"])


def test_library_affine_matrix_matches_reference_function():
    """hpdecode.geometry.get_affine_transform (host code of libhpdecode.so) vs the reference's
    base/transforms/utils.py:25-57 loaded from /root/reference: float64, bit-exact, random centers / scales."""
    import importlib.util
    from hpdecode.geometry import get_affine_transform
    spec = importlib.util.spec_from_file_location("ref_tu", ref_runner.REFERENCE_ROOT + "/src/base/transforms/utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(0)
    for _ in range(200):
        c = (int(rng.integers(100, 900)), int(rng.integers(100, 700)))
        s = (float(rng.uniform(200, 1200)), float(rng.uniform(200, 1200)))
        o = (int(rng.choice([512, 640, 704])), int(rng.choice([512, 384, 640])))
        for inv in (False, True):
            assert np.array_equal(ref.get_affine_transform(c, s, 0, o, inverse=inv), get_affine_transform(c, s, 0, o, inverse=inv))
