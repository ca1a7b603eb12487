"""Drop-in for the decode part of /root/reference/src/keypoints/results.py.

Kept: ``BaseKeypointsResult.match_heatmaps_size / resize_heatmaps_list / resize_heatmaps``
(:46-67), ``InferenceKeypointsResult.from_preds`` (:203-263) with its fields, and
``KeypointsResult.set_preds`` (:94-124).  Plotting and OKS (visualisation / evaluation consumers)
are out of scope.  ``from_preds`` hands the already flip-averaged heatmaps and the list of tag
maps to ONE fused device call (aggregation + NMS + top-k + grouping + adjust + refine).
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import ops
from .decoder import _finish
from .transforms import get_affine_transform, affine_transform


class BaseKeypointsResult:
    @classmethod
    def match_heatmaps_size(cls, heatmaps: List[Tensor]) -> List[Tensor]:
        h, w = heatmaps[-1].shape[-2:]
        return [torch.ops.hpd.resize_bilinear(hm, h, w) for hm in heatmaps[:-1]] + [heatmaps[-1]]

    @classmethod
    def resize_heatmaps_list(cls, heatmaps: List[Tensor], h: int, w: int) -> List[Tensor]:
        return [torch.ops.hpd.resize_bilinear(hm, h, w) for hm in heatmaps]

    @classmethod
    def resize_heatmaps(cls, heatmaps: Tensor, h: int, w: int) -> Tensor:
        return torch.ops.hpd.resize_bilinear(heatmaps, h, w)


def _decode_preaveraged(kpts_heatmaps: List[Tensor], tags_heatmaps: List[Tensor], img_h: int, img_w: int,
                        max_num_people: int, det_thr: float, tag_thr: float):
    """results.py:225-238 in one device call.  kpts_heatmaps = [stage1, stage2] (flip averaging
    already applied by the model, model.py:87-90); tags_heatmaps = [tag] or [tag, unflipped flip tag]."""
    if len(kpts_heatmaps) != 2:
        raise ops._lib.HpdError("hpdecode handles the two-stage HigherHRNet head (got %d stages)" % len(kpts_heatmaps))
    if len(tags_heatmaps) not in (1, 2):
        raise ops._lib.HpdError("1 or 2 tag maps expected")
    dev = kpts_heatmaps[0].device
    if not kpts_heatmaps[0].is_cuda:
        dev = torch.device("cuda:0")
    f = lambda t: t.to(dev, torch.float32)
    scale = {"hm_lo": f(kpts_heatmaps[0]), "hm_hi": f(kpts_heatmaps[1]), "tag": f(tags_heatmaps[0])}
    if len(tags_heatmaps) == 2:
        scale["tag_f"] = f(tags_heatmaps[1])
    B, K = scale["hm_lo"].shape[:2]
    E = len(tags_heatmaps)
    bufs = ops.DecodeBuffers(B, K, img_h, img_w, E, max_num_people, dev)
    params = ops.make_params(B, K, img_h, img_w, E, max_num_people, det_thr, tag_thr, True, True,
                             tags_preflipped=True)
    ops.run_decode([scale], bufs, params)
    return bufs


def transform_coords(kpts_coords: np.ndarray, center, scale, output_size) -> np.ndarray:
    """results.py:158-171."""
    out = kpts_coords.copy()
    mat = get_affine_transform(center, scale, 0, output_size, inverse=True)
    for i in range(kpts_coords.shape[0]):
        out[i, :2] = affine_transform(kpts_coords[i, :2].tolist(), mat)
    return out


@dataclass
class InferenceKeypointsResult(BaseKeypointsResult):
    raw_image: np.ndarray
    annot: Optional[list]
    model_input_image: np.ndarray
    kpts_heatmaps: np.ndarray
    tags_heatmaps: np.ndarray
    kpts_coords: np.ndarray
    kpts_scores: np.ndarray
    kpts_tags: np.ndarray
    obj_scores: np.ndarray
    limbs: list
    det_thr: float
    tag_thr: float

    @classmethod
    def get_final_kpts_coords(cls, kpts_coords: np.ndarray, center, scale, hm_size) -> np.ndarray:
        """results.py:189-201."""
        if len(kpts_coords) == 0:
            return kpts_coords
        return np.stack([transform_coords(p, center, scale, hm_size) for p in kpts_coords])

    @classmethod
    def from_preds(cls, raw_image: np.ndarray, annot, model_input_image, kpts_heatmaps: List[Tensor],
                   tags_heatmaps: List[Tensor], limbs, scale, center, det_thr: float = 0.05, tag_thr: float = 0.5,
                   max_num_people: int = 30) -> "InferenceKeypointsResult":
        """results.py:203-263.  ``model_input_image`` may be the normalised tensor [3,H,W] (only its
        spatial size is used here; un-normalising for plots is a visualisation concern) or an array."""
        if isinstance(model_input_image, Tensor):
            img_h, img_w = model_input_image.shape[-2:]
            model_input_image_npy = model_input_image
        else:
            img_h, img_w = model_input_image.shape[:2]
            model_input_image_npy = model_input_image
        bufs = _decode_preaveraged(kpts_heatmaps, tags_heatmaps, img_h, img_w, max_num_people, det_thr, tag_thr)
        P = int(bufs.n_person[0].item())
        grouped_joints, obj_scores = _finish(bufs.poses[0, :P].cpu().numpy(), bufs.person_scores[0, :P].cpu().numpy(),
                                             int(bufs.flags[0].item()) & 1)
        kpts_coords = grouped_joints[..., :2]
        kpts_scores = grouped_joints[..., 2]
        kpts_tags = grouped_joints[..., 3:]
        kpts_coords = cls.get_final_kpts_coords(kpts_coords, center, scale, (img_w, img_h))
        return cls(raw_image=raw_image, annot=annot, model_input_image=model_input_image_npy,
                   kpts_heatmaps=bufs.agg_hm[0].cpu().numpy(), tags_heatmaps=bufs.agg_tags[0, ..., 0].cpu().numpy(),
                   kpts_coords=kpts_coords, kpts_scores=kpts_scores, kpts_tags=kpts_tags, obj_scores=obj_scores,
                   limbs=limbs, det_thr=det_thr, tag_thr=tag_thr)


class KeypointsResult(BaseKeypointsResult):
    """results.py:70-124 (validation-time caller): one tag map, E = 1."""

    def __init__(self, model_input_image, kpts_heatmaps: List[Tensor], tags_heatmaps: Tensor, limbs,
                 max_num_people: int = 30, det_thr: float = 0.05, tag_thr: float = 0.5):
        self.model_input_image = model_input_image
        self._kpts_heatmaps = kpts_heatmaps
        self._tags_heatmaps = tags_heatmaps
        self.num_kpts = kpts_heatmaps[0].shape[1]
        self.limbs = limbs
        self.max_num_people = max_num_people
        self.det_thr = det_thr
        self.tag_thr = tag_thr

    def set_preds(self):
        img_h, img_w = self.model_input_image.shape[-2:] if isinstance(self.model_input_image, Tensor) \
            else self.model_input_image.shape[:2]
        hms = [h[:1].float() for h in self._kpts_heatmaps]
        bufs = _decode_preaveraged(hms, [self._tags_heatmaps[:1].float()], img_h, img_w, self.max_num_people,
                                   self.det_thr, self.tag_thr)
        P = int(bufs.n_person[0].item())
        grouped_joints, obj_scores = _finish(bufs.poses[0, :P].cpu().numpy(), bufs.person_scores[0, :P].cpu().numpy(),
                                             int(bufs.flags[0].item()) & 1)
        self.kpts_coords = grouped_joints[..., :2]
        self.kpts_scores = grouped_joints[..., 2]
        self.kpts_tags = grouped_joints[..., 3:]
        self.obj_scores = obj_scores
        # results.py:121-124: per-stage heatmaps resized to the image, stages last
        stages = self.match_heatmaps_size([h.to(bufs.agg_hm.device) for h in hms])
        stacked = torch.stack(stages, dim=1)[0]                       # [stages, K, h, w]
        resized = self.resize_heatmaps(stacked, img_h, img_w)
        self.kpts_heatmaps = resized.permute(1, 2, 3, 0).cpu().numpy()
        self.tags_heatmaps = bufs.agg_tags[0].cpu().numpy()
