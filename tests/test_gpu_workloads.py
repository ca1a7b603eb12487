"""-m gpu: the BASELINE workloads at their real shapes, decoded on the device and by the oracle from the SAME
in-memory tensors (conv outputs are not bit-reproducible across devices).

* configs[2] -- exactly what bench.py times: ``synth_net.network_outputs(64, 512, flip=True, seed=1)``; batch 64
  takes ``topk_kernel`` (rows > 512) and ``agg_nms_x2_kernel<2,4>``, which the small-batch tests never reach.
* configs[3] -- HigherHRNet-W48 640x640, test scales 0.5 / 1.0 / 1.5 + flip, through the whole decode.
* the ``torch.ops.hpd.decode`` custom op (the op north_star names) against ``BottomUpDecoder.decode``.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _check_image(oracle, host_scales, b, res, out, out_hw, tag_scale=0, M=30, det=0.05, tthr=0.5, maps=True):
    hm_o, tg_o = oracle.aggregate([{k: v[b] for k, v in s.items()} for s in host_scales], out_hw, tag_scale=tag_scale)
    if maps:
        assert np.array_equal(_bits(res.agg_hm[b].cpu().numpy()), _bits(hm_o)), f"image {b}: aggregated heatmaps"
        assert np.array_equal(_bits(res.agg_tags[b].cpu().numpy()), _bits(tg_o)), f"image {b}: aggregated tags"
    ref = oracle.parse(hm_o, tg_o, M, det, tthr)
    assert np.array_equal(res.bufs.idx_k[b].cpu().numpy(), ref["idx_k"]), f"image {b}: top-k indices"
    assert np.array_equal(_bits(res.bufs.scores_k[b].cpu().numpy()), _bits(ref["scores_k"])), f"image {b}: top-k scores"
    gj, ps = out[b]
    assert gj.shape == ref["grouped_joints"].shape, f"image {b}: person count"
    assert np.array_equal(_bits(gj), _bits(ref["grouped_joints"])), f"image {b}: grouped joints"
    assert np.array_equal(_bits(ps), _bits(ref["person_scores"])), f"image {b}: person scores"


def test_bench_workload_b64_all_images_match_oracle(oracle):
    """BASELINE configs[2] as bench.py builds it (seed 1, rank 0): all 64 images, every output bit-exact."""
    from hpdecode import BottomUpDecoder, synth_net
    B, S = 64, 512
    scale = synth_net.network_outputs(B, S, flip=True, seed=1, C=32, device="cuda:0")
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    res = dec.decode([scale], (S, S))
    out = res.to_numpy()
    host = {k: v.contiguous().cpu().numpy() for k, v in scale.items()}
    persons = 0
    for b in range(B):
        _check_image(oracle, [host], b, res, out, (S, S), maps=(b % 8 == 0))
        persons += len(out[b][0])
    assert persons > 0


def test_bench_workload_through_the_pipeline(oracle):
    """The same workload through DecodePipeline (what bench.py's timed region runs): 3 submits on 2 lanes,
    the packed rows of the last submit checked against the oracle for 8 images."""
    from hpdecode import BottomUpDecoder, synth_net
    from hpdecode.decoder import DecodePipeline
    B, S = 64, 512
    scale = synth_net.network_outputs(B, S, flip=True, seed=1, C=32, device="cuda:0")
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    pipe = DecodePipeline(dec, depth=2)
    last = None
    for _ in range(3):
        last = pipe.submit([scale], (S, S))
    pipe.drain()
    torch.cuda.synchronize()
    out = last.to_numpy()
    host = {k: v.contiguous().cpu().numpy() for k, v in scale.items()}
    for b in range(0, B, 8):
        _check_image(oracle, [host], b, last, out, (S, S), maps=False)


def test_config3_w48_640_multiscale_matches_oracle(oracle):
    """BASELINE configs[3] at its real shape: W48, 640x640, scales 0.5/1.0/1.5 (inputs 320/640/960) + flip."""
    from hpdecode import BottomUpDecoder, synth_net
    B, S = 2, 640
    ins = [synth_net.network_outputs(B, int(round(S * sc / 64.0)) * 64, flip=True, seed=5, C=48, device="cuda:0", chunk=2)
           for sc in (0.5, 1.0, 1.5)]
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    res = dec.decode(ins, (S, S), tag_scale=1)
    out = res.to_numpy()
    host = [{k: v.contiguous().cpu().numpy() for k, v in s.items()} for s in ins]
    for b in range(B):
        _check_image(oracle, host, b, res, out, (S, S), tag_scale=1)


def test_decode_custom_op_matches_decoder():
    """torch.ops.hpd.decode (device pointers in, device tensors out) == BottomUpDecoder.decode, output by output."""
    from hpdecode import BottomUpDecoder, synth
    scales = synth.crowd(3, 256, persons=9, flip=True, seed=17)
    dev = [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in scales]
    s = dev[0]
    outs = torch.ops.hpd.decode([s["hm_lo"]], [s["hm_hi"]], [s["tag"]], [s["hm_lo_f"]], [s["hm_hi_f"]], [s["tag_f"]],
                                256, 256, 30, 0.05, 0.5, True, True, 0)
    agg_hm, agg_tags, poses, person_scores, n_person, flags, scores_k, idx_k, coords_k, tags_k = outs
    res = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0").decode(dev, (256, 256))
    b = res.bufs
    for got, want in ((agg_hm, b.agg_hm), (agg_tags, b.agg_tags), (n_person, b.n_person), (flags, b.flags),
                      (scores_k, b.scores_k), (idx_k, b.idx_k), (coords_k, b.coords_k), (tags_k, b.tags_k),
                      (person_scores, b.person_scores)):
        assert torch.equal(got, want)
    for i in range(3):
        P = int(n_person[i])
        assert P > 0 and torch.equal(poses[i, :P], b.poses[i, :P])
    # no flip -> E = 1
    outs1 = torch.ops.hpd.decode([s["hm_lo"]], [s["hm_hi"]], [s["tag"]], [], [], [], 256, 256, 30, 0.05, 0.5, True, True, 0)
    assert outs1[1].shape == (3, 17, 256, 256, 1) and outs1[2].shape == (3, 30, 17, 4)
    # CPU tensors are refused loudly: there is no fallback
    from hpdecode._lib import HpdError
    with pytest.raises((HpdError, RuntimeError)):
        torch.ops.hpd.decode([s["hm_lo"].cpu()], [s["hm_hi"].cpu()], [s["tag"].cpu()], [], [], [], 256, 256, 30, 0.05, 0.5,
                             True, True, 0)


def test_pipeline_inputs_may_be_dropped_after_submit():
    """submit() records the inputs on the lane's stream, so the caller can free them at once and keep
    allocating on its own stream without corrupting the in-flight batch."""
    from hpdecode import BottomUpDecoder, synth
    from hpdecode.decoder import DecodePipeline
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    host = synth.crowd(4, 256, persons=7, flip=True, seed=23)
    want = dec.decode([{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in host], (256, 256), slot=7).to_numpy()
    want = [(g.copy(), p.copy()) for g, p in want]
    pipe = DecodePipeline(dec, depth=2)
    results = []
    for _ in range(4):
        fresh = [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in host]
        torch.cuda.current_stream().synchronize()
        results.append(pipe.submit(fresh, (256, 256)))
        del fresh
        # same-size allocations on the caller's stream right away: without record_stream they could reuse the blocks
        junk = [torch.full((4, 17, 64, 64), float("nan"), device="cuda") for _ in range(12)]
        del junk
    pipe.drain()
    torch.cuda.synchronize()
    got = results[-1].to_numpy()
    for (wg, wp), (gg, gp) in zip(want, got):
        assert np.array_equal(_bits(wg), _bits(gg)) and np.array_equal(_bits(wp), _bits(gp))


def test_pipeline_with_cuda_graphs_and_records_ring(oracle):
    """What bench.py's timed region runs: DecodePipeline with one CUDA-graph replay per submit and the lanes' result
    records aimed at a ring.  Static inputs (graphs bake the pointers in), 7 submits over 3 lanes, small batch with
    ``batches_in_flight`` >= 8 semantics exercised separately by the top-k tests."""
    from hpdecode import BottomUpDecoder, ops, synth
    from hpdecode.decoder import DecodePipeline, Records
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    host = synth.crowd(4, 256, persons=11, flip=True, seed=31)
    dev = [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in host]
    want = dec.decode(dev, (256, 256), slot=9).records.clone()
    row = ops.record_layout(17, 30, 2).row_bytes
    ring = torch.zeros((3, 4, row), device="cuda", dtype=torch.uint8)
    pipe = DecodePipeline(dec, depth=3, records_ring=ring, use_graphs=True)
    for _ in range(7):
        pipe.submit(dev, (256, 256))
    pipe.drain()
    torch.cuda.synchronize()
    assert pipe.graph_replays == 7
    for lane in range(3):
        assert torch.equal(ring[lane], want)
    rec = Records(ring[1].cpu().numpy(), 30, 17, 2)
    for b in range(4):
        hm_o, tg_o = oracle.aggregate(synth.image_slice(host, b), (256, 256))
        ref = oracle.parse(hm_o, tg_o, 30, 0.05, 0.5)
        gj, ps = rec.image(b)
        assert np.array_equal(_bits(gj), _bits(ref["grouped_joints"])) and np.array_equal(_bits(ps), _bits(ref["person_scores"]))


def test_pipeline_passes_the_back_projection_matrices():
    """DecodePipeline.submit(inv_affine=...) writes the same records (COCO section back-projected per image) as
    BottomUpDecoder.decode(inv_affine=...), eagerly and under CUDA graphs (the matrices are then baked-in pointers,
    so the same device tensor must be passed again)."""
    from hpdecode import BottomUpDecoder, geometry, synth
    from hpdecode.decoder import DecodePipeline
    dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
    host = synth.crowd(2, 256, persons=6, flip=True, seed=41)
    dev = [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in host]
    minv = np.stack([geometry.get_affine_transform((320, 240), (640.0, 480.0), 0, (256, 256), inverse=True).ravel(),
                     geometry.get_affine_transform((100, 90), (250.0, 250.0), 0, (256, 256), inverse=True).ravel()])
    want = dec.decode(dev, (256, 256), slot=5, inv_affine=minv).records.clone()
    plain = dec.decode(dev, (256, 256), slot=6).records.clone()
    assert not torch.equal(want, plain)                                  # the projection really changed the records
    pipe = DecodePipeline(dec, depth=2)
    got = [pipe.submit(dev, (256, 256), inv_affine=minv) for _ in range(3)]
    pipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(got[-1].records, want) and torch.equal(got[-2].records, want)
    minv_dev = torch.from_numpy(minv).cuda()
    gpipe = DecodePipeline(dec, depth=2, use_graphs=True)
    out = [gpipe.submit(dev, (256, 256), inv_affine=minv_dev) for _ in range(4)]
    gpipe.drain()
    torch.cuda.synchronize()
    assert torch.equal(out[-1].records, want) and torch.equal(out[-2].records, want)
