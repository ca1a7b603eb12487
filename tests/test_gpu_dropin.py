"""-m gpu: the Python shell around the decode as the reference's callers use it -- input preparation on the device
(model.py:70-76), ``InferenceKeypointsModel.__call__(raw_image, annot)`` (model.py:78-111, called like
bin/eval.py:28), ``from_preds`` / back-projection (results.py:158-263), the batched evaluation loop
(bin/eval.py:18-49) and the validation-time ``KeypointsResult.set_preds`` (results.py:94-124).

Oracles: oracle/input_oracle.py (pinned against cv2 / torchvision / the reference's own functions, goldens in
tests/golden/input_cases.npz) for the geometry, oracle/hpd_oracle.cpp for the decode.  Network outputs are captured
while the model runs and handed to the oracle as they are (conv outputs are not reproducible across calls).
"""
import os

import numpy as np
import pytest
import torch
from torch import nn

from oracle import input_oracle
from oracle.gen_golden_input import CASES, image_of, sha

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_cases.npz"))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


class CaptureNet(nn.Module):
    """Wraps a HigherHRNet and keeps every call's (input, outputs) on the host."""

    def __init__(self, net):
        super().__init__()
        self.net = net
        self.calls = []

    def forward(self, x):
        (lo, hi), tag = self.net(x)
        self.calls.append((x.detach().cpu().numpy().copy(), lo.detach().float().cpu().numpy().copy(),
                           hi.detach().float().cpu().numpy().copy(), tag.detach().float().cpu().numpy().copy()))
        return [lo, hi], tag


def _net(C=32, seed=0):
    from hpdecode import synth_net
    torch.manual_seed(seed)
    return CaptureNet(synth_net.HigherHRNet(17, C))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_prepare_input_matches_reference_goldens(i):
    """uint8 image -> normalised network input: bit-identical to resize_align_multi_scale + ToTensor + Normalize."""
    from hpdecode import geometry
    h, w, input_size, seed = CASES[i]
    img = image_of(h, w, seed)
    x, centers, scales = geometry.prepare_input([img], input_size, "cuda:0")
    assert tuple(centers[0]) == tuple(GOLD[f"c{i}_center"]) and np.array_equal(np.array(scales[0]), GOLD[f"c{i}_scale"])
    got = x[0].cpu().numpy()
    assert got.shape[1:] == tuple(GOLD[f"c{i}_size"])[::-1]
    assert np.array_equal(got[:, :2, :8], GOLD[f"c{i}_x_head"])
    assert sha(got) == str(GOLD[f"c{i}_x_sha"])
    want, _, _, _, _ = input_oracle.prepare_input(img, input_size)
    assert np.array_equal(_bits(got), _bits(want))


def test_prepare_input_batched_and_against_cv2():
    """Several images of one resized size in one launch (raw sizes differ), 40 images to cross the 32-per-launch
    chunk; and, where OpenCV is importable, the live cv2.warpAffine."""
    from hpdecode import geometry
    rng = np.random.default_rng(3)
    shapes = [(480, 640), (479, 640), (481, 641), (240, 320)] * 10            # all -> (704, 512) at input_size 512
    assert len(geometry.group_by_resized_size(shapes, 512)) == 1
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    x, centers, scales = geometry.prepare_input(imgs, 512, "cuda:0")
    assert x.shape == (40, 3, 512, 704)
    xs = x.cpu().numpy()
    for j in (0, 1, 2, 3, 31, 32, 39):
        want, c, s, size, M = input_oracle.prepare_input(imgs[j], 512)
        assert np.array_equal(_bits(xs[j]), _bits(want)) and c == centers[j] and s == scales[j]
        try:
            import cv2
        except ImportError:
            continue
        warped = cv2.warpAffine(imgs[j], M, size)
        assert np.array_equal(_bits(xs[j]), _bits(input_oracle.to_tensor_normalize(warped)))
    with pytest.raises(Exception):
        geometry.prepare_input([imgs[0], np.zeros((100, 300, 3), np.uint8)], 512, "cuda:0")     # different resized sizes


@pytest.mark.parametrize("i", range(len(CASES)))
def test_back_projection_epilogue_matches_reference_goldens(i):
    """get_final_kpts_coords / transform_coords (results.py:158-201) through the decode's record epilogue:
    float32 in -> the reference's float32 array, float64 in (the fallback pseudo-person) -> float64."""
    from hpdecode import InferenceKeypointsResult
    from hpdecode.results import transform_coords
    size, center, scale = tuple(GOLD[f"c{i}_size"]), tuple(GOLD[f"c{i}_center"]), tuple(GOLD[f"c{i}_scale"])
    kpts = GOLD[f"c{i}_kpts"]
    got32 = InferenceKeypointsResult.get_final_kpts_coords(kpts, center, scale, size)
    assert got32.dtype == np.float32 and np.array_equal(_bits(got32), _bits(GOLD[f"c{i}_back32"]))
    got64 = InferenceKeypointsResult.get_final_kpts_coords(kpts.astype(np.float64), center, scale, size)
    assert got64.dtype == np.float64 and np.array_equal(got64, GOLD[f"c{i}_back64"])
    one = transform_coords(kpts[1], center, scale, size)
    assert np.array_equal(_bits(one), _bits(GOLD[f"c{i}_back32"][1]))


def _oracle_result(oracle, calls, flip, size_wh, center, scale, M=30, det=0.05, tthr=0.5, b=0, tag_scale=0, n_scales=1):
    """What the reference returns for image b, from the captured network outputs (one (normal[, flipped]) pair per scale)."""
    W, H = size_wh
    per = 2 if flip else 1
    scales = []
    for s in range(n_scales):
        _, lo, hi, tag = calls[s * per]
        d = {"hm_lo": lo[b], "hm_hi": hi[b], "tag": tag[b]}
        if flip:
            _, lo_f, hi_f, tag_f = calls[s * per + 1]
            d.update(hm_lo_f=lo_f[b], hm_hi_f=hi_f[b], tag_f=tag_f[b])
        scales.append(d)
    hm_o, tg_o = oracle.aggregate(scales, (H, W), tag_scale=tag_scale)
    ref = oracle.parse(hm_o, tg_o, M, det, tthr)
    minv = input_oracle.get_affine_transform(center, scale, (W, H), inverse=True)
    gj = ref["grouped_joints"]
    coords = input_oracle.affine_points(gj[..., :2].reshape(-1, 2), minv).reshape(gj.shape[0], -1, 2).astype(np.float32)
    return hm_o, tg_o, ref, coords


@pytest.mark.parametrize("flip", [False, True])
def test_model_call_is_a_drop_in_for_the_reference(flip, oracle):
    """model(raw_image, annot) exactly like bin/eval.py:28 / bin/inference.py:57 on a uint8 image."""
    from hpdecode import InferenceKeypointsModel, InferenceKeypointsResult
    net = _net()
    model = InferenceKeypointsModel(net, det_thr=0.05, tag_thr=0.5, use_flip=flip, input_size=256, max_num_people=30,
                                    device="cuda:0")
    raw = image_of(300, 421, 11)
    x, center, scale = model.prepare_input(raw)                       # the reference's helper, same signature
    want_x, c2, s2, size, _ = input_oracle.prepare_input(raw, 256)
    assert x.shape[0] == 1 and np.array_equal(_bits(x[0].cpu().numpy()), _bits(want_x)) and center == c2 and scale == s2
    result = model(raw, annot=None)
    assert isinstance(result, InferenceKeypointsResult) and model.model_input_shape == (size[1], size[0])
    assert np.array_equal(_bits(net.calls[0][0][0]), _bits(want_x))                  # the network saw the reference's tensor
    if flip:
        assert np.array_equal(net.calls[1][0], net.calls[0][0][..., ::-1])
    hm_o, tg_o, ref, coords = _oracle_result(oracle, net.calls, flip, size, center, scale)
    gj = ref["grouped_joints"]
    assert result.raw_image is raw and result.annot is None and result.limbs == model.limbs
    assert result.kpts_coords.dtype == np.float32 and np.array_equal(_bits(result.kpts_coords), _bits(coords))
    assert np.array_equal(_bits(result.kpts_scores), _bits(gj[..., 2])) and np.array_equal(_bits(result.kpts_tags), _bits(gj[..., 3:]))
    assert np.array_equal(_bits(result.obj_scores), _bits(ref["person_scores"]))
    assert np.array_equal(_bits(result.kpts_heatmaps), _bits(hm_o)) and np.array_equal(_bits(result.tags_heatmaps), _bits(tg_o[..., 0]))
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    want_img = ((want_x.transpose(1, 2, 0) * std + mean) * 255).astype(np.uint8)            # base/transforms/base.py:33-42
    assert result.model_input_image.dtype == np.uint8 and np.array_equal(result.model_input_image, want_img)
    assert (result.det_thr, result.tag_thr) == (0.05, 0.5)


def test_from_preds_nonempty_coordinates(oracle):
    """from_preds with the reference's calling convention (flip-averaged heatmaps, [tag, un-flipped flip tag]) and
    a non-trivial back-projection (480x640 raw image -> 704x512 network input)."""
    from hpdecode import InferenceKeypointsResult, synth
    FL = synth.COCO_FLIP_INDEX
    size, center, scale = input_oracle.get_multi_scale_size(480, 640, 512, 1, 1)           # (704, 512)
    W, H = size
    rng = np.random.default_rng(5)
    K = 17
    s = {"hm_lo": rng.standard_normal((1, K, H // 4, W // 4)).astype(np.float32) * 0.1,
         "hm_hi": rng.standard_normal((1, K, H // 2, W // 2)).astype(np.float32) * 0.1,
         "tag": rng.standard_normal((1, K, H // 4, W // 4)).astype(np.float32)}
    for p in range(6):
        for k in range(K):
            y, x = int(rng.integers(4, H // 4 - 4)), int(rng.integers(4, W // 4 - 4))
            s["hm_lo"][0, k, y, x] += 0.7 + 0.2 * rng.random()
            s["tag"][0, k, y - 2:y + 3, x - 2:x + 3] = 2.0 * p
    s.update({k + "_f": np.ascontiguousarray(v[..., ::-1][:, FL]) for k, v in list(s.items())})   # a consistent flipped run
    t = {k: torch.from_numpy(v) for k, v in s.items()}
    hms = [(t["hm_lo"] + torch.flip(t["hm_lo_f"], [3])[:, FL]) / 2, (t["hm_hi"] + torch.flip(t["hm_hi_f"], [3])[:, FL]) / 2]
    tags = [t["tag"], torch.flip(t["tag_f"], [3])[:, FL].contiguous()]
    x_in = torch.zeros(3, H, W)
    res = InferenceKeypointsResult.from_preds(None, None, x_in, [h.cuda() for h in hms], [g.cuda() for g in tags], [],
                                              scale, center, 0.05, 0.5, 30)
    hm_o, tg_o = oracle.aggregate([{k: v[0] for k, v in s.items()}], (H, W))
    ref = oracle.parse(hm_o, tg_o, 30, 0.05, 0.5)
    gj = ref["grouped_joints"]
    assert gj.shape[0] >= 6 and not ref["fallback"]
    minv = input_oracle.get_affine_transform(center, scale, size, inverse=True)
    coords = input_oracle.affine_points(gj[..., :2].reshape(-1, 2), minv).reshape(gj.shape[0], K, 2).astype(np.float32)
    assert np.array_equal(_bits(res.kpts_coords), _bits(coords))
    assert np.abs(res.kpts_coords - gj[..., :2]).max() > 1.0                                 # the projection really moved them
    assert np.array_equal(_bits(res.kpts_scores), _bits(gj[..., 2])) and np.array_equal(_bits(res.obj_scores), _bits(ref["person_scores"]))
    assert np.array_equal(_bits(res.kpts_heatmaps), _bits(hm_o))


class _FakeCoco:
    """What evaluate_dataset touches of CocoKeypointsDataset: images_filepaths, load_image, __len__."""

    def __init__(self, shapes):
        self.images = [image_of(h, w, 20 + i) for i, (h, w) in enumerate(shapes)]
        self.images_filepaths = ["/data/coco/images/val2017/%012d.jpg" % (139 + 7 * i) for i in range(len(shapes))]

    def __len__(self):
        return len(self.images)

    def load_image(self, idx):
        return self.images[idx]


def test_evaluate_dataset_batched_equals_the_per_image_loop(oracle):
    """bin/eval.py:18-49: the batched loop (groups by resized size, one decode per group, COCO dicts straight from
    the records) against the reference's loop shape -- model(raw_image, None) image by image -- decoded by the
    oracle from the captured network outputs."""
    from hpdecode import InferenceKeypointsModel, evaluate_dataset_batched
    ds = _FakeCoco([(240, 320), (320, 240), (239, 320), (256, 256), (320, 241)])
    net = _net(seed=2)
    model = InferenceKeypointsModel(net, use_flip=True, input_size=256, device="cuda:0")
    got = evaluate_dataset_batched(model, ds, batch_size=4)
    calls = list(net.calls)
    # replay what the batched loop did, group by group, against the oracle
    from hpdecode import geometry
    want, ci = {}, 0
    for i0 in (0, 4):
        idxs = list(range(i0, min(i0 + 4, len(ds))))
        groups = geometry.group_by_resized_size([ds.images[i].shape[:2] for i in idxs], 256)
        for size, members in groups.items():
            pair = calls[ci:ci + 2]
            ci += 2
            for b, j in enumerate(members):
                i = idxs[j]
                _, center, scale = input_oracle.get_multi_scale_size(*ds.images[i].shape[:2], 256, 1, 1)
                _, _, ref, coords = _oracle_result(oracle, pair, True, size, center, scale, b=b)
                want[i] = [{"image_id": 139 + 7 * i, "category_id": 1,
                            "keypoints": np.concatenate([coords[p].astype(np.float64), np.ones((17, 1))], 1).ravel().tolist(),
                            "score": float(ref["person_scores"][p])} for p in range(coords.shape[0])]
    flat = [r for i in range(len(ds)) for r in want[i]]
    assert len(got) == len(flat) and got == flat
    assert all(len(r["keypoints"]) == 51 and r["keypoints"][2::3] == [1.0] * 17 for r in got)


def test_predict_batch_equals_single_calls_in_layout():
    """predict_batch returns one InferenceKeypointsResult per image in input order, whatever the grouping."""
    from hpdecode import InferenceKeypointsModel, result_to_coco
    ds = _FakeCoco([(240, 320), (320, 240), (240, 319)])
    model = InferenceKeypointsModel(_net(seed=3), use_flip=False, input_size=256, device="cuda:0")
    res = model.predict_batch(ds.images)
    assert len(res) == 3 and all(r.raw_image is im for r, im in zip(res, ds.images))
    assert res[0].kpts_heatmaps is None and res[0].kpts_coords.shape[1:] == (17, 2)
    recs = result_to_coco(5, res[1])
    assert len(recs) == len(res[1].obj_scores) and recs[0]["keypoints"][0] == float(res[1].kpts_coords[0, 0, 0])


def test_multi_scale_inference_matches_oracle(oracle):
    """Real multi-scale inference (test_scales 0.5 / 1.0 / 1.5 + flip): inputs prepared per scale with the
    reference's size logic (utils.py:60-97), heatmaps averaged over scales, tags from scale 1.0."""
    from hpdecode import InferenceKeypointsModel
    net = _net(seed=4)
    model = InferenceKeypointsModel(net, use_flip=True, input_size=512, device="cuda:0", test_scales=(0.5, 1.0, 1.5))
    raw = image_of(360, 480, 31)
    result = model(raw, None)
    sizes = [input_oracle.get_multi_scale_size(360, 480, 512, s, 0.5)[0] for s in (0.5, 1.0, 1.5)]
    assert [c[0].shape[-2:] for c in net.calls[::2]] == [(h, w) for (w, h) in sizes]
    for s, call in zip((0.5, 1.0, 1.5), net.calls[::2]):
        want_x = input_oracle.prepare_input(raw, 512, s, 0.5)[0]
        assert np.array_equal(_bits(call[0][0]), _bits(want_x))
    size, center, scale = input_oracle.get_multi_scale_size(360, 480, 512, 1.0, 0.5)
    hm_o, tg_o, ref, coords = _oracle_result(oracle, net.calls, True, size, center, scale, tag_scale=1, n_scales=3)
    assert np.array_equal(_bits(result.kpts_heatmaps), _bits(hm_o))
    assert np.array_equal(_bits(result.kpts_coords), _bits(coords))
    assert np.array_equal(_bits(result.obj_scores), _bits(ref["person_scores"]))


@pytest.mark.parametrize("half", [False, True])
def test_set_preds_validation_path(half, oracle):
    """KeypointsResult.set_preds (results.py:94-124) as module.py:100-110 builds it: E = 1, max_num_people 20,
    det_thr 0.1, tag_thr 1.0, per-stage resized heatmaps.  half=True: fp16 network outputs (autocast, module.py:78);
    the oracle is the reference's float32 path on the up-cast tensors."""
    from hpdecode import KeypointsResult, synth
    scales = synth.crowd(1, 256, persons=9, flip=False, seed=77)
    s = {k: torch.from_numpy(v).cuda() for k, v in scales[0].items()}
    if half:
        s = {k: v.half() for k, v in s.items()}
    up = {k: v.float().cpu().numpy() for k, v in s.items()}
    image = torch.zeros(3, 256, 256)
    r = KeypointsResult(image, [s["hm_lo"], s["hm_hi"]], s["tag"], [], max_num_people=20, det_thr=0.1, tag_thr=1.0)
    r.set_preds()
    hm_o, tg_o = oracle.aggregate([{k: v[0] for k, v in up.items()}], (256, 256))
    ref = oracle.parse(hm_o, tg_o, 20, 0.1, 1.0)
    gj = ref["grouped_joints"]
    assert gj.shape[0] >= 5
    assert np.array_equal(_bits(r.kpts_coords), _bits(gj[..., :2])) and np.array_equal(_bits(r.kpts_scores), _bits(gj[..., 2]))
    assert np.array_equal(_bits(r.kpts_tags), _bits(gj[..., 3:])) and r.kpts_tags.shape[-1] == 1
    assert np.array_equal(_bits(r.obj_scores), _bits(ref["person_scores"]))
    assert np.array_equal(_bits(r.tags_heatmaps), _bits(tg_o)) and r.tags_heatmaps.shape == (17, 256, 256, 1)
    # results.py:121-124: [K, H, W, stages] -- stage 1 resized to stage 2's size, then both to the image
    lo_up = oracle.resize_bilinear(up["hm_lo"][0], 128, 128)
    want = np.stack([oracle.resize_bilinear(lo_up, 256, 256), oracle.resize_bilinear(up["hm_hi"][0], 256, 256)], -1)
    assert r.kpts_heatmaps.shape == (17, 256, 256, 2) and np.array_equal(_bits(r.kpts_heatmaps), _bits(want))
    assert r.model_input_image.shape == (256, 256, 3) and r.model_input_image.dtype == np.uint8
