"""Batched bottom-up decoder: the public entry a pipeline calls once per batch.

``BottomUpDecoder.decode`` takes the raw HigherHRNet outputs of a batch (optionally of the flipped
run and of several test scales) as CUDA tensors and enqueues the whole path -- fused aggregation +
NMS, top-k, grouping, adjust, refine -- on the current stream through ``hpd_decode``.  Nothing is
copied to the host until ``DecodeResult.to_numpy()`` is called, and then only the pose lists.

The reference decodes one image per call (results.py:233-234 strips the batch dim); here the
batch is a superset: image b of a batch gives exactly what the reference returns for that image.
"""
from collections import OrderedDict
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


class DecodeResult:
    """Device-side result of one batch; ``to_numpy`` gives the reference's per-image return values."""

    def __init__(self, bufs: ops.DecodeBuffers):
        self.bufs = bufs
        self.agg_hm = bufs.agg_hm          # [B,K,H,W]
        self.agg_tags = bufs.agg_tags      # [B,K,H,W,E]
        self.poses = bufs.poses            # [B,M,K,3+E]
        self.person_scores = bufs.person_scores
        self.n_person = bufs.n_person
        self.flags = bufs.flags

    def packed(self) -> torch.Tensor:
        """One device tensor [B, M*K*(3+E) + M + 2] holding poses, person scores, count and flag,
        so the device->host read of a batch is a single copy."""
        B = self.poses.shape[0]
        return torch.cat([self.poses.reshape(B, -1), self.person_scores,
                          self.n_person.to(torch.float32)[:, None], self.flags.to(torch.float32)[:, None]], dim=1)

    @staticmethod
    def unpack(packed: np.ndarray, M: int, K: int, E: int) -> List[Tuple[np.ndarray, np.ndarray]]:
        out = []
        D = 3 + E
        for row in packed:
            P = int(row[-2])
            fb = int(row[-1]) & 1
            poses = row[: M * K * D].reshape(M, K, D)[:P].copy()
            scores = row[M * K * D: M * K * D + M][:P].copy()
            out.append(_finish(poses, scores, fb))
        return out

    def to_numpy(self) -> List[Tuple[np.ndarray, np.ndarray]]:
        """[(grouped_joints [P,K,3+E], person_scores [P]) per image] -- MPPEHeatmapParser.parse's return."""
        B, M, K, D = self.poses.shape
        return self.unpack(self.packed().cpu().numpy(), M, K, D - 3)


def _finish(poses: np.ndarray, scores: np.ndarray, fallback: int):
    if fallback:
        # grouping.py:262-269 builds this pseudo-person in float64 with score 0.01; the kernel wrote
        # the float32 image of it.  Coordinates and tags are exact in both; restore the dtype/score.
        poses = poses.astype(np.float64)
        poses[..., 2] = 0.01
        scores = poses[..., 2].mean(1)
    return poses, scores


class BottomUpDecoder:
    def __init__(self, num_kpts: int = 17, max_num_people: int = 30, det_thr: float = 0.05, tag_thr: float = 0.5,
                 device="cuda:0", adjust: bool = True, refine: bool = True, max_cached_shapes: int = 4):
        if not torch.cuda.is_available():
            raise ops._lib.HpdError("hpdecode needs a CUDA device (sm_100a); there is no CPU fallback")
        ops._lib.lib()
        self.num_kpts, self.max_num_people = num_kpts, max_num_people
        self.det_thr, self.tag_thr = det_thr, tag_thr
        self.adjust, self.refine = adjust, refine
        self.device = torch.device(device)
        self._bufs = OrderedDict()
        self.max_cached_shapes = max_cached_shapes

    def buffers(self, B, H, W, E, slot=0) -> ops.DecodeBuffers:
        """Buffer set for one (shape, slot), kept in a small LRU: bottom-up inference sees many padded
        (H, W) (resize_align_multi_scale keeps the aspect ratio) and a set is ~80 MB per image at 512-class
        sizes, so at most ``max_cached_shapes`` distinct shapes stay resident (slots of one shape are evicted
        together; an evicted set is freed once the stream work that uses it has run, by the caching allocator)."""
        shape = (B, H, W, E)
        slots = self._bufs.get(shape)
        if slots is None:
            while len(self._bufs) >= self.max_cached_shapes:
                self._bufs.popitem(last=False)
            slots = self._bufs[shape] = {}
        else:
            self._bufs.move_to_end(shape)
        if slot not in slots:
            slots[slot] = ops.DecodeBuffers(B, self.num_kpts, H, W, E, self.max_num_people, self.device)
        return slots[slot]

    def decode(self, scales: Sequence[dict], out_hw: Tuple[int, int], tag_scale: int = 0,
               tags_preflipped: bool = False, slot: int = 0) -> DecodeResult:
        """scales: one dict per test scale with CUDA float32 tensors hm_lo, hm_hi, tag [B,K,h,w] and,
        for the flip test, hm_lo_f, hm_hi_f, tag_f (raw outputs of the flipped forward).

        The returned DecodeResult VIEWS the decoder's cached buffer set for this (shape, slot): it is
        overwritten by the next decode of the same shape and slot.  Read it (``to_numpy()``), ``clone()``
        what must outlive that, or pass distinct ``slot`` values for results that are held together."""
        H, W = out_hw
        B = scales[0]["hm_lo"].shape[0]
        E = 2 if scales[tag_scale].get("tag_f") is not None else 1
        bufs = self.buffers(B, H, W, E, slot)
        params = ops.make_params(B, self.num_kpts, H, W, E, self.max_num_people, self.det_thr, self.tag_thr,
                                 self.adjust, self.refine, len(scales), tag_scale, tags_preflipped=tags_preflipped)
        ops.run_decode(scales, bufs, params)
        return DecodeResult(bufs)

    def decode_maps(self, agg_hm: torch.Tensor, agg_tags: torch.Tensor) -> DecodeResult:
        """Aggregated maps [B,K,H,W], [B,K,H,W,E] -> poses (the MPPEHeatmapParser.parse entry)."""
        B, K, H, W = agg_hm.shape
        E = agg_tags.shape[4]
        bufs = ops.DecodeBuffers(B, K, H, W, E, self.max_num_people, agg_hm.device, agg_hm.contiguous(),
                                 agg_tags.contiguous())
        params = ops.make_params(B, K, H, W, E, self.max_num_people, self.det_thr, self.tag_thr, self.adjust,
                                 self.refine)
        ops.run_decode(None, bufs, params)
        return DecodeResult(bufs)


class DecodePipeline:
    """Keeps ``depth`` batches in flight on one GPU.

    The path has one bandwidth-bound kernel (fused aggregation + NMS) followed by three latency-bound
    ones (top-k, grouping, refine) that occupy a handful of SMs each.  Every lane owns a buffer set,
    a normal-priority stream for the aggregation kernel and a high-priority stream for the rest, so
    batch i+1's aggregation overlaps batch i's tail and the small kernels get SM slots as soon as
    they are runnable.  ``submit`` returns the lane's DecodeResult; it stays valid until the lane is
    reused ``depth`` submits later (call ``result.ready.synchronize()`` or ``drain()`` before reading).
    Lifetime contract for the inputs: ``submit`` records their use on the lane's stream
    (``Tensor.record_stream``), so the caller may drop them as soon as it returns.
    """

    def __init__(self, decoder: "BottomUpDecoder", depth: int = 3, split_priority: bool = False):
        self.dec = decoder
        self.depth = max(1, depth)
        dev = decoder.device
        self.lanes = []
        for i in range(self.depth):
            s_agg = torch.cuda.Stream(device=dev, priority=0)
            # split_priority: the tail kernels run on their own high-priority stream (measured slightly
            # slower on B200 than one stream per lane, so it is off by default)
            s_tail = torch.cuda.Stream(device=dev, priority=-1) if split_priority else s_agg
            self.lanes.append({"slot": i, "s_agg": s_agg, "s_tail": s_tail, "agg_done": torch.cuda.Event(),
                               "ready": torch.cuda.Event(), "used": False})
        self._next = 0

    def submit(self, scales: Sequence[dict], out_hw: Tuple[int, int], tag_scale: int = 0, before_agg=None,
               after_tail=None) -> DecodeResult:
        """Enqueue one batch.  ``before_agg(lane)`` / ``after_tail(lane, result)`` run on the lane's streams
        (e.g. the H2D copies of the inputs and the D2H copy of the packed poses)."""
        ln = self.lanes[self._next]
        self._next = (self._next + 1) % self.depth
        H, W = out_hw
        B = scales[0]["hm_lo"].shape[0]
        E = 2 if scales[tag_scale].get("tag_f") is not None else 1
        d = self.dec
        bufs = d.buffers(B, H, W, E, slot=ln["slot"])
        params = ops.make_params(B, d.num_kpts, H, W, E, d.max_num_people, d.det_thr, d.tag_thr, d.adjust, d.refine,
                                 len(scales), tag_scale)
        cur = torch.cuda.current_stream(d.device)
        ln["s_agg"].wait_stream(cur)
        # The inputs were allocated on the caller's stream but are read on the lane's: tell the caching
        # allocator, or a caller that drops them right after submit() (net outputs -> submit -> next forward)
        # could get their memory back while the aggregation kernel is still reading it.
        for s in scales:
            for t in s.values():
                if t is not None and t.is_cuda:
                    t.record_stream(ln["s_agg"])
        with torch.cuda.stream(ln["s_agg"]):
            if ln["used"]:
                ln["s_agg"].wait_event(ln["ready"])     # the lane's previous batch has been consumed
            if before_agg is not None:
                scales = before_agg(ln) or scales      # tensors made here belong to the lane's stream already
            ops.run_stage("aggregate_nms", bufs, params, scales=scales)
            if ln["s_tail"] is not ln["s_agg"]:
                ln["agg_done"].record()
        res = DecodeResult(bufs)
        with torch.cuda.stream(ln["s_tail"]):
            if ln["s_tail"] is not ln["s_agg"]:
                ln["s_tail"].wait_event(ln["agg_done"])
            for st in ("topk", "group", "adjust_refine"):
                ops.run_stage(st, bufs, params)
            if after_tail is not None:
                after_tail(ln, res)
            ln["ready"].record()
        ln["used"] = True
        res.ready = ln["ready"]
        return res

    def drain(self):
        """Make the current stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream(self.dec.device)
        for ln in self.lanes:
            if ln["used"]:
                cur.wait_event(ln["ready"])
