"""-m gpu: the CUDA path (through the C ABI / torch.ops.hpd) against the CPU oracle, stage by stage.

Bars (BASELINE.json north_star): NMS survivors, top-k indices and person/joint assignment bit-exact;
float outputs within 1e-5 relative -- in practice every float below is compared BIT-exactly too,
because the kernels replay the reference's float32/float64 operation order.
"""
import numpy as np
import pytest
import torch

from hpdecode import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # the tolerance north_star states for float outputs


def _dev(scales):
    return [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in scales]


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


CASES = [
    # name, generator, kwargs, size, det_thr, tag_thr, M
    ("netlike192_flip", synth.netlike, dict(batch=2, size=192, flip=True, seed=1), 192, 0.05, 0.5, 30),
    ("netlike192_noflip", synth.netlike, dict(batch=1, size=192, flip=False, seed=2), 192, 0.05, 0.5, 30),
    ("crowd192_flip", synth.crowd, dict(batch=2, size=192, persons=8, flip=True, seed=3), 192, 0.05, 0.5, 30),
    ("crowd256_q", synth.crowd, dict(batch=1, size=256, persons=20, flip=True, seed=4, quantised=True), 256, 0.05, 0.5, 30),
    ("crowd256_q_noflip", synth.crowd, dict(batch=1, size=256, persons=30, flip=False, seed=5, quantised=True), 256, 0.05, 0.5, 30),
    ("crowd256_val", synth.crowd, dict(batch=1, size=256, persons=12, flip=False, seed=6), 256, 0.1, 1.0, 20),
    ("netlike512_flip", synth.netlike, dict(batch=1, size=512, flip=True, seed=7), 512, 0.05, 0.5, 30),
    ("crowd512_30", synth.crowd, dict(batch=2, size=512, persons=30, flip=True, seed=8), 512, 0.05, 0.5, 30),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_decode_matches_oracle(case, oracle):
    from hpdecode import BottomUpDecoder
    name, gen, kw, size, det, tthr, M = case
    scales = gen(**kw)
    B = kw["batch"]
    dec = BottomUpDecoder(17, M, det, tthr, "cuda:0")
    res = dec.decode(_dev(scales), (size, size))
    torch.cuda.synchronize()
    agg, tags = res.agg_hm.cpu().numpy(), res.agg_tags.cpu().numpy()
    bufs = res.bufs
    out = res.to_numpy()
    for b in range(B):
        hm_o, tg_o = oracle.aggregate(synth.image_slice(scales, b), (size, size))
        assert np.array_equal(_bits(agg[b]), _bits(hm_o)), f"{name}[{b}] aggregated heatmaps"
        assert np.array_equal(_bits(tags[b]), _bits(tg_o)), f"{name}[{b}] aggregated tags"
        # NMS survivors (bit mask) and word maxima
        nm, keep = oracle.nms(hm_o)
        mask = bufs.nms_mask[b].cpu().numpy().view(np.uint32)
        bits = ((mask[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(17, size, -1)[:, :, :size]
        assert np.array_equal(bits.astype(np.uint8), keep), f"{name}[{b}] NMS survivors"
        wm = bufs.nms_wmax[b].cpu().numpy()
        assert np.array_equal(wm, nm.reshape(17, size, -1, 32).max(-1)), f"{name}[{b}] nms word max"
        hx = bufs.hm_wmax[b].cpu().numpy()
        assert np.array_equal(hx, hm_o.reshape(17, size, -1, 32).max(-1)), f"{name}[{b}] raw word max"
        ref = oracle.parse(hm_o, tg_o, M, det, tthr)
        assert np.array_equal(bufs.idx_k[b].cpu().numpy(), ref["idx_k"]), f"{name}[{b}] top-k indices"
        assert np.array_equal(_bits(bufs.scores_k[b].cpu().numpy()), _bits(ref["scores_k"])), f"{name}[{b}] top-k scores"
        assert np.array_equal(bufs.coords_k[b].cpu().numpy(), ref["coords_k"]), f"{name}[{b}] top-k coords"
        assert np.array_equal(_bits(bufs.tags_k[b].cpu().numpy()), _bits(ref["tags_k"])), f"{name}[{b}] top-k tags"
        gj, ps = out[b]
        assert gj.shape == ref["grouped_joints"].shape, f"{name}[{b}] person count {gj.shape} vs {ref['grouped_joints'].shape}"
        # assignment: which (person, joint) slots are filled and with which integer peak
        assert np.array_equal(gj[..., 2] != 0, ref["grouped_joints"][..., 2] != 0), f"{name}[{b}] joint assignment"
        np.testing.assert_allclose(gj, ref["grouped_joints"], rtol=RTOL, atol=0, err_msg=f"{name}[{b}] grouped joints")
        np.testing.assert_allclose(ps, ref["person_scores"], rtol=RTOL, atol=0, err_msg=f"{name}[{b}] person scores")
        assert np.array_equal(_bits(gj), _bits(ref["grouped_joints"])), f"{name}[{b}] grouped joints (bit-exact)"
        assert np.array_equal(_bits(ps), _bits(ref["person_scores"])), f"{name}[{b}] person scores (bit-exact)"


@pytest.mark.parametrize("C,size,B", [(32, 512, 2), (48, 384, 1)])
def test_decode_of_random_init_higherhrnet_outputs(C, size, B, oracle):
    """BASELINE configs 0-2: heatmaps / tags produced by a default-init HigherHRNet on synthetic images (flip
    test), decoded on the device and by the oracle from the SAME in-memory tensors (conv outputs are not
    bit-reproducible across devices).  hm_lo / tag are strided channel-slice views."""
    from hpdecode import BottomUpDecoder, synth_net
    scale = synth_net.network_outputs(B, size, flip=True, seed=3, C=C, device="cuda:0", chunk=2)
    assert B == 1 or not scale["hm_lo"].is_contiguous()
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    res = dec.decode([scale], (size, size))
    out = res.to_numpy()
    host = {k: v.contiguous().cpu().numpy() for k, v in scale.items()}
    for b in range(B):
        hm_o, tg_o = oracle.aggregate([{k: v[b] for k, v in host.items()}], (size, size))
        assert np.array_equal(_bits(res.agg_hm[b].cpu().numpy()), _bits(hm_o))
        assert np.array_equal(_bits(res.agg_tags[b].cpu().numpy()), _bits(tg_o))
        ref = oracle.parse(hm_o, tg_o, 30, 0.05, 0.5)
        assert np.array_equal(res.bufs.idx_k[b].cpu().numpy(), ref["idx_k"])
        assert np.array_equal(_bits(out[b][0]), _bits(ref["grouped_joints"]))
        assert np.array_equal(_bits(out[b][1]), _bits(ref["person_scores"]))


def test_multiscale_decode_matches_oracle(oracle):
    """BASELINE config 3 end to end at a test size: scales 0.5/1.0/1.5 + flip, heatmaps averaged over scales,
    tags from scale 1.0, then the normal decode."""
    from hpdecode import BottomUpDecoder
    size = 512
    scales = synth.netlike(2, size, True, seed=71, scales=(0.5, 1.0, 1.5))
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    res = dec.decode(_dev(scales), (size, size), tag_scale=1)
    out = res.to_numpy()
    for b in range(2):
        hm_o, tg_o = oracle.aggregate(synth.image_slice(scales, b), (size, size), tag_scale=1)
        assert np.array_equal(_bits(res.agg_hm[b].cpu().numpy()), _bits(hm_o))
        ref = oracle.parse(hm_o, tg_o, 30, 0.05, 0.5)
        assert np.array_equal(res.bufs.idx_k[b].cpu().numpy(), ref["idx_k"])
        assert np.array_equal(_bits(out[b][0]), _bits(ref["grouped_joints"]))
        assert np.array_equal(_bits(out[b][1]), _bits(ref["person_scores"]))


def test_pipeline_gives_the_same_results_as_sequential_decode():
    """DecodePipeline (batches in flight on several streams, buffer sets re-used) vs plain decode()."""
    from hpdecode import BottomUpDecoder
    from hpdecode.decoder import DecodePipeline
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    batches = [_dev(synth.crowd(2, 256, persons=5 + 3 * i, flip=True, seed=80 + i)) for i in range(5)]
    want = []
    for s in batches:
        r = dec.decode(s, (256, 256), slot=99)
        want.append([(g.copy(), p.copy()) for g, p in r.to_numpy()])
    pipe = DecodePipeline(dec, depth=2)
    got = []
    keep = []
    for s in batches:
        holder = {}
        pipe.submit(s, (256, 256), after_tail=lambda ln, res, h=holder: h.update(packed=res.packed().clone()))
        keep.append(holder)
    pipe.drain()
    torch.cuda.synchronize()
    for h in keep:
        got.append(type(r).unpack(h["packed"].cpu().numpy(), 30, 17, 2))
    for w, g in zip(want, got):
        for (wg, wp), (gg, gp) in zip(w, g):
            assert np.array_equal(_bits(wg), _bits(gg)) and np.array_equal(_bits(wp), _bits(gp))


def test_nonsquare_other_joint_count_and_m20(oracle):
    """Edge shapes through the whole path: H != W (both kernels' strides), K = 8 joints (identity joint order),
    M = 20 with the validation thresholds, E = 1, batch with very different person counts."""
    from hpdecode import BottomUpDecoder
    H, W, K, M = 256, 384, 8, 20
    rng = np.random.default_rng(7)
    B = 3
    lo = (rng.standard_normal((B, K, H // 4, W // 4)) * 0.05).astype(np.float32)
    hi = (rng.standard_normal((B, K, H // 2, W // 2)) * 0.05).astype(np.float32)
    tag = (rng.standard_normal((B, K, H // 4, W // 4)) * 0.02).astype(np.float32)
    for b, persons in enumerate((0, 3, 25)):          # image 0 stays below det_thr -> empty-scene fallback
        for p in range(persons):
            for k in range(K):
                y, x = rng.integers(3, H // 4 - 3), rng.integers(3, W // 4 - 3)
                lo[b, k, y, x] += 0.6 + 0.3 * rng.random()
                hi[b, k, 2 * y:2 * y + 2, 2 * x:2 * x + 2] += 0.5
                tag[b, k, y - 1:y + 2, x - 1:x + 2] = 1.5 * p
    lo[0] = np.minimum(lo[0], 0.01)
    hi[0] = np.minimum(hi[0], 0.01)
    scale = {"hm_lo": lo, "hm_hi": hi, "tag": tag}
    dec = BottomUpDecoder(K, M, 0.1, 1.0, "cuda:0")
    res = dec.decode([{k: torch.from_numpy(v).cuda() for k, v in scale.items()}], (H, W))
    out = res.to_numpy()
    assert int(res.flags[0].item()) == 1 and out[0][0].dtype == np.float64 and out[0][0].shape == (1, K, 4)
    for b in range(B):
        hm_o, tg_o = oracle.aggregate([{k: v[b] for k, v in scale.items()}], (H, W), flip_index=list(range(K)))
        assert np.array_equal(_bits(res.agg_hm[b].cpu().numpy()), _bits(hm_o))
        assert np.array_equal(_bits(res.agg_tags[b].cpu().numpy()), _bits(tg_o))
        ref = oracle.parse(hm_o, tg_o, M, 0.1, 1.0)
        assert np.array_equal(res.bufs.idx_k[b].cpu().numpy(), ref["idx_k"])
        if ref["fallback"]:
            from hpdecode.decoder import _finish
            gj, ps = _finish(ref["grouped_joints"], ref["person_scores"], 1)
            assert np.array_equal(out[b][0], gj) and np.array_equal(out[b][1], ps)
        else:
            assert np.array_equal(_bits(out[b][0]), _bits(ref["grouped_joints"]))
            assert np.array_equal(_bits(out[b][1]), _bits(ref["person_scores"]))


@pytest.mark.parametrize("M", [5, 32])
def test_extreme_max_people(M, oracle):
    """max_num_people at its smallest sensible value and at the library's limit (32 = warp width)."""
    from hpdecode import BottomUpDecoder
    scales = synth.crowd(2, 256, persons=28, flip=True, seed=90 + M, tag_spread=1.1)
    dec = BottomUpDecoder(17, M, 0.05, 0.5, "cuda:0")
    res = dec.decode(_dev(scales), (256, 256))
    out = res.to_numpy()
    for b in range(2):
        hm_o, tg_o = oracle.aggregate(synth.image_slice(scales, b), (256, 256))
        ref = oracle.parse(hm_o, tg_o, M, 0.05, 0.5)
        assert np.array_equal(res.bufs.idx_k[b].cpu().numpy(), ref["idx_k"])
        assert out[b][0].shape == ref["grouped_joints"].shape
        assert np.array_equal(_bits(out[b][0]), _bits(ref["grouped_joints"]))
        assert np.array_equal(_bits(out[b][1]), _bits(ref["person_scores"]))


def test_stale_result_is_refused():
    """a DecodeResult views the decoder's cached buffer set; reading it after a later decode of the same
    shape and slot raises instead of silently returning the newer batch."""
    from hpdecode import BottomUpDecoder
    from hpdecode._lib import HpdError
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    a = _dev(synth.crowd(1, 192, persons=4, flip=True, seed=1))
    b = _dev(synth.crowd(1, 192, persons=7, flip=True, seed=2))
    first = dec.decode(a, (192, 192))
    kept = first.to_numpy()
    second = dec.decode(b, (192, 192))
    with pytest.raises(HpdError):
        first.to_numpy()
    assert len(second.to_numpy()[0][0]) != 0 and len(kept[0][0]) != 0
    other = dec.decode(a, (192, 192), slot=1)          # distinct slots can be held together
    assert len(second.to_numpy()) == 1 and len(other.to_numpy()) == 1
