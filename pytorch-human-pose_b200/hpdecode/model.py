"""Drop-in for the flip-test logic of /root/reference/src/keypoints/model.py:78-111.

``InferenceKeypointsModel`` wraps any stock-PyTorch HigherHRNet-style ``net`` returning
``([hm_lo, hm_hi], tag)``; the convolutions are NOT part of this library.  Unlike the reference,
the flip averaging (model.py:85-96) is not done with torch ops: the raw outputs of the normal and
the flipped forward go straight into the fused aggregation kernel, which applies the W-flip and
the COCO joint permutation while it loads its tiles.
"""
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor, nn

from .decoder import BottomUpDecoder, DecodeResult


class InferenceKeypointsModel:
    def __init__(self, net: nn.Module, det_thr: float = 0.05, tag_thr: float = 0.5, use_flip: bool = False,
                 input_size: int = 512, max_num_people: int = 30, device: str = "cuda:0", num_kpts: int = 17):
        self.net = net.to(device).eval()
        self.det_thr, self.tag_thr = det_thr, tag_thr
        self.use_flip = use_flip
        self.input_size = input_size
        self.max_num_people = max_num_people
        self.device = torch.device(device)
        self.decoder = BottomUpDecoder(num_kpts, max_num_people, det_thr, tag_thr, device)

    @torch.no_grad()
    def forward_decode(self, x: Tensor) -> DecodeResult:
        """x: normalised input batch [B,3,H,W] on the device.  Returns the device-side result."""
        h, w = x.shape[-2:]
        (hm_lo, hm_hi), tag = self.net(x)
        scale = {"hm_lo": hm_lo.float(), "hm_hi": hm_hi.float(), "tag": tag.float()}
        if self.use_flip:   # model.py:85-94, fused into the aggregation kernel
            (fl_lo, fl_hi), fl_tag = self.net(torch.flip(x, [3]))
            scale.update(hm_lo_f=fl_lo.float(), hm_hi_f=fl_hi.float(), tag_f=fl_tag.float())
        return self.decoder.decode([scale], (h, w))

    def __call__(self, x: Tensor):
        """Batched superset of model.py:78: per image (grouped_joints, person_scores)."""
        return self.forward_decode(x.to(self.device)).to_numpy()
