"""PyTorch custom ops over the C ABI: device pointers in, device tensors out, no host round trip.

``torch.ops.hpd.*`` are registered with ``torch.library.custom_op``; each one fills ctypes structs
with ``tensor.data_ptr()`` / strides and the current CUDA stream and calls libhpdecode.so.  PyTorch is
plumbing here (memory, streams); all arithmetic happens in the library's sm_100a kernels.

Stage ops mirror the reference methods one to one (paths relative to /root/reference):
  hpd::aggregate_nms  model.py:85-96 + results.py:225-230 (+ grouping.py:80-83 fused)
  hpd::nms            MPPEHeatmapParser.nms          grouping.py:80-83
  hpd::topk           MPPEHeatmapParser.top_k        grouping.py:147-170
  hpd::group          MPPEHeatmapParser.match_by_tag grouping.py:85-145 (+ fallback :262-269)
  hpd::adjust_refine  adjust / person score / refine grouping.py:172-191,276,193-250
  hpd::decode         all of the above in one call   (from_preds, results.py:225-238)
"""
import ctypes
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from ._lib import HpdBuffers, HpdImage, HpdMap, HpdParams, HpdRecordLayout, HpdScaleInputs

COCO_FLIP_INDEX = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]  # transforms.py:11
JOINTS_ORDER_17 = [0, 1, 2, 3, 4, 5, 6, 11, 12, 7, 8, 9, 10, 13, 14, 15, 16]   # grouping.py:63-65

_launches = 0


def launches_total() -> int:
    """Kernel launches enqueued through this module since import (bench accounting)."""
    return _launches


def _require_cuda(t: Tensor, name: str):
    if not t.is_cuda:
        raise _lib.HpdError(f"{name} must be a CUDA tensor: hpdecode has no CPU path")
    if t.dtype != torch.float32:
        raise _lib.HpdError(f"{name} must be float32 (got {t.dtype})")


def _map(t: Optional[Tensor], name: str, allow_half: bool = False) -> HpdMap:
    m = HpdMap()
    if t is None:
        return m
    if allow_half and t.is_cuda and t.dtype == torch.float16:
        m.dtype = _lib.HPD_F16      # network outputs under fp16 autocast (module.py:78): widened inside the kernel
    else:
        _require_cuda(t, name)
    if t.dim() != 4:
        raise _lib.HpdError(f"{name} must be [B,K,h,w]")
    if t.stride(3) != 1 or t.stride(2) != t.shape[3]:
        raise _lib.HpdError(f"{name}: rows must be contiguous (strides {t.stride()}); call .contiguous()")
    m.ptr = t.data_ptr()
    m.stride_b, m.stride_c = t.stride(0), t.stride(1)
    m.h, m.w = t.shape[2], t.shape[3]
    return m


def make_params(B: int, K: int, H: int, W: int, E: int, M: int, det_thr: float, tag_thr: float, adjust: bool = True,
                refine: bool = True, num_scales: int = 1, tag_scale: int = 0, flip_index=None,
                joints_order=None, tags_preflipped: bool = False) -> HpdParams:
    p = HpdParams()
    p.batch, p.num_kpts, p.out_h, p.out_w, p.emb, p.max_people = B, K, H, W, E, M
    p.num_scales, p.tag_scale = num_scales, tag_scale
    p.do_adjust, p.do_refine = int(adjust), int(refine)
    p.tags_preflipped = int(tags_preflipped)
    p.det_thr, p.tag_thr = float(det_thr), float(tag_thr)
    if flip_index is None:
        flip_index = COCO_FLIP_INDEX if K == 17 else list(range(K))
    if joints_order is None:
        joints_order = JOINTS_ORDER_17 if K == 17 else list(range(K))
    for k in range(min(K, _lib.HPD_MAX_KPTS)):
        p.flip_index[k] = flip_index[k]
        p.joints_order[k] = joints_order[k]
    return p


_layouts = {}


def record_layout(K: int, M: int, E: int) -> HpdRecordLayout:
    """hpd_record_layout for (num_kpts, max_people, emb): byte offsets of one image's result record."""
    key = (K, M, E)
    if key not in _layouts:
        p = HpdParams()
        p.num_kpts, p.max_people, p.emb = K, M, E
        L = HpdRecordLayout()
        _lib.check(_lib.lib().hpd_record_layout(ctypes.byref(p), ctypes.byref(L)), "hpd_record_layout")
        _layouts[key] = L
    return _layouts[key]


class DecodeBuffers:
    """Device buffers of one decode call (HpdBuffers).  Re-usable across calls of the same shape.

    ``records`` [B,row_bytes] uint8 receives one result record per image (HpdRecordLayout) from the epilogue of
    the last kernel; ``inv_affine`` (optional float64 [B,6] on the device) is the per-image matrix the COCO
    section's coordinates are back-projected with (None = network-input coordinates)."""

    def __init__(self, B: int, K: int, H: int, W: int, E: int, M: int, device, agg_hm: Optional[Tensor] = None,
                 agg_tags: Optional[Tensor] = None):
        self.shape = (B, K, H, W, E, M)
        wpr = (W + 31) // 32
        f32 = dict(device=device, dtype=torch.float32)
        i32 = dict(device=device, dtype=torch.int32)
        self.agg_hm = agg_hm if agg_hm is not None else torch.empty((B, K, H, W), **f32)
        self.agg_tags = agg_tags if agg_tags is not None else torch.empty((B, K, H, W, E), **f32)
        self.nms_mask = torch.empty((B, K, H, wpr), **i32)
        self.nms_wmax = torch.empty((B, K, H, wpr), **f32)
        self.hm_wmax = torch.empty((B, K, H, wpr), **f32)
        self.tag_bmin = torch.empty((B, K, (H + 3) // 4, wpr), **f32)
        self.tag_bmax = torch.empty((B, K, (H + 3) // 4, wpr), **f32)
        self.scores_k = torch.empty((B, K, M), **f32)
        self.idx_k = torch.empty((B, K, M), **i32)
        self.coords_k = torch.empty((B, K, M, 2), **i32)
        self.tags_k = torch.empty((B, K, M, E), **f32)
        self.poses = torch.empty((B, M, K, 3 + E), **f32)
        self.person_scores = torch.empty((B, M), **f32)
        self.n_person = torch.empty((B,), **i32)
        self.flags = torch.empty((B,), **i32)
        self.records = torch.empty((B, record_layout(K, M, E).row_bytes), device=device, dtype=torch.uint8)
        self.inv_affine = None
        self.workspace = None

    @classmethod
    def for_grouping(cls, scores_k: Tensor, coords_k: Tensor, tags_k: Tensor) -> "DecodeBuffers":
        """Only what hpd_group touches (top-k inputs, pose outputs); the map-sized buffers stay empty."""
        B, K, M, E = tags_k.shape
        dev = tags_k.device
        self = cls.__new__(cls)
        self.shape = (B, K, 0, 0, E, M)
        zf = torch.empty((0,), device=dev, dtype=torch.float32)
        zi = torch.empty((0,), device=dev, dtype=torch.int32)
        self.agg_hm = self.agg_tags = self.nms_wmax = self.hm_wmax = self.tag_bmin = self.tag_bmax = zf
        self.person_scores = zf
        self.nms_mask = self.idx_k = zi
        self.scores_k, self.coords_k, self.tags_k = scores_k, coords_k, tags_k
        self.poses = torch.empty((B, M, K, 3 + E), device=dev, dtype=torch.float32)
        self.n_person = torch.empty((B,), device=dev, dtype=torch.int32)
        self.flags = torch.empty((B,), device=dev, dtype=torch.int32)
        self.records = self.inv_affine = None
        self.workspace = None
        return self

    def struct(self) -> HpdBuffers:
        s = HpdBuffers()
        for name, _ in HpdBuffers._fields_:
            t = getattr(self, name)
            if t is None:
                continue
            if name == "inv_affine" and (t.dtype != torch.float64 or t.numel() != 6 * self.shape[0] or not t.is_cuda):
                raise _lib.HpdError("inv_affine must be a CUDA float64 tensor [B,6]")
            if not t.is_contiguous():
                raise _lib.HpdError(f"buffer {name} must be contiguous")
            setattr(s, name, t.data_ptr())
        return s

    def ensure_workspace(self, params: HpdParams):
        n = ctypes.c_size_t(0)
        _lib.check(_lib.lib().hpd_workspace_bytes(ctypes.byref(params), ctypes.byref(n)), "hpd_workspace_bytes")
        if self.workspace is None or self.workspace.numel() < n.value:
            self.workspace = torch.empty((max(n.value, 8),), device=self.agg_hm.device, dtype=torch.uint8)
        return self.workspace


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _scales_struct(scales: Sequence[dict]):
    arr = (HpdScaleInputs * len(scales))()
    keep = []
    for i, s in enumerate(scales):
        for name in ("hm_lo", "hm_hi", "tag", "hm_lo_f", "hm_hi_f", "tag_f"):
            t = s.get(name)
            setattr(arr[i], name, _map(t, name, allow_half=True))
            keep.append(t)
    return arr, keep


_capture_launches = 0


def _count(L):
    global _launches, _capture_launches
    n = L.hpd_last_launch_count()
    _launches += n
    _capture_launches += n


def launches_in_last_capture() -> int:
    """Launches enqueued since the previous call of this function (DecodePipeline brackets a graph capture with
    two calls and books that many launches per replay)."""
    global _capture_launches
    n, _capture_launches = _capture_launches, 0
    return n


def add_launches(n: int):
    global _launches
    _launches += n


# ------------------------------------------------------------------------------------------------
# functional core (used by the custom ops below and by hpdecode.decoder)
# ------------------------------------------------------------------------------------------------
def run_decode(scales: Optional[Sequence[dict]], bufs: DecodeBuffers, params: HpdParams):
    """hpd_decode on the current stream.  scales=None -> bufs.agg_hm / agg_tags are the inputs."""
    L = _lib.lib()
    dev = bufs.agg_hm.device
    ws = bufs.ensure_workspace(params)
    with torch.cuda.device(dev):
        if scales is not None:
            arr, _keep = _scales_struct(scales)
            rc = L.hpd_decode(ctypes.byref(params), arr, ctypes.byref(bufs.struct()), ws.data_ptr(), ws.numel(),
                              _stream_ptr(dev))
        else:
            rc = L.hpd_decode(ctypes.byref(params), None, ctypes.byref(bufs.struct()), ws.data_ptr(), ws.numel(),
                              _stream_ptr(dev))
    _lib.check(rc, "hpd_decode")
    _count(L)
    return bufs


def run_stage(stage: str, bufs: DecodeBuffers, params: HpdParams, scales=None, nms_out: Optional[Tensor] = None):
    L = _lib.lib()
    dev = bufs.agg_hm.device
    st = _stream_ptr(dev)
    b = bufs.struct()
    with torch.cuda.device(dev):
        if stage == "aggregate_nms":
            arr, _keep = _scales_struct(scales)
            rc = L.hpd_aggregate_nms(ctypes.byref(params), arr, ctypes.byref(b), st)
        elif stage == "nms":
            rc = L.hpd_nms(ctypes.byref(params), ctypes.byref(b), nms_out.data_ptr() if nms_out is not None else None, st)
        elif stage == "topk":
            rc = L.hpd_topk(ctypes.byref(params), ctypes.byref(b), st)
        elif stage == "group":
            rc = L.hpd_group(ctypes.byref(params), ctypes.byref(b), st)
        elif stage == "adjust_refine":
            ws = bufs.ensure_workspace(params)
            rc = L.hpd_adjust_refine(ctypes.byref(params), ctypes.byref(b), ws.data_ptr(), ws.numel(), st)
        else:
            raise ValueError(stage)
    _lib.check(rc, f"hpd_{stage}")
    _count(L)
    return bufs


def _scales_from_lists(hm_lo, hm_hi, tag, hm_lo_f, hm_hi_f, tag_f):
    n = len(hm_lo)
    flip = len(hm_lo_f) > 0
    out = []
    for i in range(n):
        d = {"hm_lo": hm_lo[i], "hm_hi": hm_hi[i], "tag": tag[i] if i < len(tag) else None}
        if flip:
            d.update(hm_lo_f=hm_lo_f[i], hm_hi_f=hm_hi_f[i], tag_f=tag_f[i] if i < len(tag_f) else None)
        out.append(d)
    return out


# ------------------------------------------------------------------------------------------------
# torch.library custom ops
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("hpd::decode", mutates_args=())
def decode_op(hm_lo: Sequence[Tensor], hm_hi: Sequence[Tensor], tag: Sequence[Tensor], hm_lo_f: Sequence[Tensor],
              hm_hi_f: Sequence[Tensor], tag_f: Sequence[Tensor], out_h: int, out_w: int, max_people: int,
              det_thr: float, tag_thr: float, adjust: bool, refine: bool, tag_scale: int) -> List[Tensor]:
    """Network outputs (one list entry per test scale) -> [agg_hm, agg_tags, poses, person_scores, n_person,
    flags, scores_k, idx_k, coords_k, tags_k]."""
    for name, ts in (("hm_lo", hm_lo), ("hm_hi", hm_hi), ("tag", tag), ("hm_lo_f", hm_lo_f), ("hm_hi_f", hm_hi_f),
                     ("tag_f", tag_f)):
        for t in ts:
            if not (t.is_cuda and t.dtype == torch.float16):
                _require_cuda(t, name)
    scales = _scales_from_lists(hm_lo, hm_hi, tag, hm_lo_f, hm_hi_f, tag_f)
    B, K = hm_lo[0].shape[:2]
    E = 2 if len(tag_f) > 0 else 1
    bufs = DecodeBuffers(B, K, out_h, out_w, E, max_people, hm_lo[0].device)
    params = make_params(B, K, out_h, out_w, E, max_people, det_thr, tag_thr, adjust, refine, len(scales), tag_scale)
    run_decode(scales, bufs, params)
    return [bufs.agg_hm, bufs.agg_tags, bufs.poses, bufs.person_scores, bufs.n_person, bufs.flags, bufs.scores_k,
            bufs.idx_k, bufs.coords_k, bufs.tags_k]


@torch.library.custom_op("hpd::parse", mutates_args=())
def parse_op(agg_hm: Tensor, agg_tags: Tensor, max_people: int, det_thr: float, tag_thr: float, adjust: bool,
             refine: bool) -> List[Tensor]:
    """Aggregated maps [B,K,H,W] / [B,K,H,W,E] -> [poses, person_scores, n_person, flags, scores_k, idx_k,
    coords_k, tags_k]   (MPPEHeatmapParser.parse, grouping.py:252-283)."""
    _require_cuda(agg_hm, "agg_hm")
    _require_cuda(agg_tags, "agg_tags")
    B, K, H, W = agg_hm.shape
    E = agg_tags.shape[4]
    bufs = DecodeBuffers(B, K, H, W, E, max_people, agg_hm.device, agg_hm.contiguous(), agg_tags.contiguous())
    params = make_params(B, K, H, W, E, max_people, det_thr, tag_thr, adjust, refine)
    run_decode(None, bufs, params)
    return [bufs.poses, bufs.person_scores, bufs.n_person, bufs.flags, bufs.scores_k, bufs.idx_k, bufs.coords_k,
            bufs.tags_k]


@torch.library.custom_op("hpd::nms", mutates_args=())
def nms_op(hm: Tensor) -> Tensor:
    """MPPEHeatmapParser.nms on [B,K,H,W] (grouping.py:80-83): x where it equals its 5x5 max, else x*0."""
    _require_cuda(hm, "hm")
    B, K, H, W = hm.shape
    bufs = DecodeBuffers(B, K, H, W, 1, 1, hm.device, hm.contiguous(), torch.empty((0,), device=hm.device))
    params = make_params(B, K, H, W, 1, 1, 0.0, 0.0)
    out = torch.empty_like(bufs.agg_hm)
    run_stage("nms", bufs, params, nms_out=out)
    return out


@torch.library.custom_op("hpd::topk", mutates_args=())
def topk_op(agg_hm: Tensor, agg_tags: Tensor, max_people: int) -> List[Tensor]:
    """MPPEHeatmapParser.top_k on [B,K,H,W] / [B,K,H,W,E] (grouping.py:147-170):
    [tags_k, coords_k, scores_k, idx_k]."""
    _require_cuda(agg_hm, "agg_hm")
    _require_cuda(agg_tags, "agg_tags")
    B, K, H, W = agg_hm.shape
    E = agg_tags.shape[4]
    bufs = DecodeBuffers(B, K, H, W, E, max_people, agg_hm.device, agg_hm.contiguous(), agg_tags.contiguous())
    params = make_params(B, K, H, W, E, max_people, 0.0, 0.0)
    run_stage("nms", bufs, params)
    run_stage("topk", bufs, params)
    return [bufs.tags_k, bufs.coords_k, bufs.scores_k, bufs.idx_k]


@torch.library.custom_op("hpd::group", mutates_args=())
def group_op(tags_k: Tensor, coords_k: Tensor, scores_k: Tensor, det_thr: float, tag_thr: float,
             out_h: int, out_w: int) -> List[Tensor]:
    """MPPEHeatmapParser.match_by_tag on [B,K,M,E] / [B,K,M,2] / [B,K,M] (grouping.py:85-145):
    [poses [B,M,K,3+E], n_person [B], flags [B]]."""
    _require_cuda(tags_k, "tags_k")
    _require_cuda(scores_k, "scores_k")
    B, K, M, E = tags_k.shape
    dev = tags_k.device
    params = make_params(B, K, out_h, out_w, E, M, det_thr, tag_thr)
    bufs = DecodeBuffers.for_grouping(scores_k.contiguous(), coords_k.contiguous().int(), tags_k.contiguous())
    run_stage("group", bufs, params)
    return [bufs.poses, bufs.n_person, bufs.flags]


@torch.library.custom_op("hpd::resize_bilinear", mutates_args=())
def resize_bilinear_op(x: Tensor, out_h: int, out_w: int) -> Tensor:
    """F.interpolate(x, size=[out_h, out_w], mode='bilinear', align_corners=False) with torch's CPU
    arithmetic (results.py:46-67) on a [B,C,h,w] CUDA tensor."""
    m = _map(x, "x")
    B, C = x.shape[:2]
    out = torch.empty((B, C, out_h, out_w), device=x.device, dtype=torch.float32)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        rc = L.hpd_resize_bilinear(ctypes.byref(m), B, C, out.data_ptr(), out_h, out_w, _stream_ptr(x.device))
    _lib.check(rc, "hpd_resize_bilinear")
    _count(L)
    return out


@torch.library.custom_op("hpd::prepare_input", mutates_args=())
def prepare_input_op(images: Sequence[Tensor], matrices: Tensor, out_h: int, out_w: int, mean: Sequence[float],
                     std: Sequence[float]) -> Tensor:
    """cv2.warpAffine(image, M, (out_w, out_h)) + ToTensor + Normalize (utils.py:96, model.py:45-50) for a batch of
    uint8 [h,w,3] CUDA images; ``matrices``: float64 [B,6] on the HOST (the forward 2x3 matrices).  -> [B,3,out_h,out_w]."""
    if len(images) == 0 or matrices.is_cuda or matrices.dtype != torch.float64 or matrices.numel() != 6 * len(images):
        raise _lib.HpdError("prepare_input: need >= 1 image and a host float64 [B,6] matrix tensor")
    dev = images[0].device
    arr = (HpdImage * len(images))()
    mats = matrices.reshape(-1, 6).tolist()
    for i, t in enumerate(images):
        if not t.is_cuda or t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3 or t.stride(2) != 1 or t.stride(1) != 3:
            raise _lib.HpdError("prepare_input: image %d must be a CUDA uint8 [h,w,3] tensor with packed pixels" % i)
        arr[i].ptr, arr[i].stride_row, arr[i].h, arr[i].w = t.data_ptr(), t.stride(0), t.shape[0], t.shape[1]
        for j in range(6):
            arr[i].m[j] = mats[i][j]
    out = torch.empty((len(images), 3, out_h, out_w), device=dev, dtype=torch.float32)
    mean_c, std_c = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.hpd_prepare_input(arr, len(images), out.data_ptr(), out_h, out_w, mean_c, std_c, _stream_ptr(dev))
    _lib.check(rc, "hpd_prepare_input")
    _count(L)
    return out


@prepare_input_op.register_fake
def _(images, matrices, out_h, out_w, mean, std):
    return images[0].new_empty((len(images), 3, out_h, out_w), dtype=torch.float32)


@resize_bilinear_op.register_fake
def _(x, out_h, out_w):
    return x.new_empty((x.shape[0], x.shape[1], out_h, out_w))


@decode_op.register_fake
def _(hm_lo, hm_hi, tag, hm_lo_f, hm_hi_f, tag_f, out_h, out_w, max_people, det_thr, tag_thr, adjust, refine,
      tag_scale):
    B, K = hm_lo[0].shape[:2]
    E = 2 if len(tag_f) > 0 else 1
    M = max_people
    f = lambda *s: hm_lo[0].new_empty(s)
    i = lambda *s: hm_lo[0].new_empty(s, dtype=torch.int32)
    return [f(B, K, out_h, out_w), f(B, K, out_h, out_w, E), f(B, M, K, 3 + E), f(B, M), i(B), i(B), f(B, K, M),
            i(B, K, M), i(B, K, M, 2), f(B, K, M, E)]


@parse_op.register_fake
def _(agg_hm, agg_tags, max_people, det_thr, tag_thr, adjust, refine):
    B, K = agg_hm.shape[:2]
    E, M = agg_tags.shape[4], max_people
    f = lambda *s: agg_hm.new_empty(s)
    i = lambda *s: agg_hm.new_empty(s, dtype=torch.int32)
    return [f(B, M, K, 3 + E), f(B, M), i(B), i(B), f(B, K, M), i(B, K, M), i(B, K, M, 2), f(B, K, M, E)]


@nms_op.register_fake
def _(hm):
    return torch.empty_like(hm)


@topk_op.register_fake
def _(agg_hm, agg_tags, max_people):
    B, K = agg_hm.shape[:2]
    E, M = agg_tags.shape[4], max_people
    return [agg_hm.new_empty((B, K, M, E)), agg_hm.new_empty((B, K, M, 2), dtype=torch.int32), agg_hm.new_empty((B, K, M)),
            agg_hm.new_empty((B, K, M), dtype=torch.int32)]


@group_op.register_fake
def _(tags_k, coords_k, scores_k, det_thr, tag_thr, out_h, out_w):
    B, K, M, E = tags_k.shape
    return [tags_k.new_empty((B, M, K, 3 + E)), tags_k.new_empty((B,), dtype=torch.int32),
            tags_k.new_empty((B,), dtype=torch.int32)]
