"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference decode in the build container.

Imports /root/reference/src/keypoints/grouping.py as it lies (with oracle/refshim on the path
so that ``import munkres`` resolves to the restated stand-in) and replays the torch calls of
results.py:225-234 and model.py:85-96 around it (those two modules cannot be imported here:
they pull pycocotools / albumentations / matplotlib / torchinfo / mlflow).

/root/reference exists only in the build container; nothing in the ``-m gpu`` tests, smoke()
or bench.py imports this module.  It is used by oracle/gen_golden.py and by the
container-only tests that pin the C++ oracle against the reference.
"""
import os
import sys

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("HP_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))

COCO_FLIP_INDEX = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "keypoints", "grouping.py"))


def _import_grouping():
    shim = os.path.join(_HERE, "refshim")
    for p in (shim, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.keypoints import grouping  # noqa: the reference module, unmodified
    return grouping


def _interp(x, h, w):
    return torch.nn.functional.interpolate(x, size=[h, w], mode="bilinear", align_corners=False)


def aggregate_torch(scales, out_hw, tag_scale=0):
    """model.py:85-96 + results.py:225-230 replayed with the same torch CPU calls.

    scales: list of dicts of [K,h,w] float32 arrays (one image).  Returns torch tensors
    (hm [K,H,W], tags [K,H,W,E]).  Multi-scale (len > 1) is the extension of SURVEY 8(a):
    per-scale full-res maps, torch.stack(...).mean(0); tags from scales[tag_scale].
    """
    H, W = out_hw
    per_scale = []
    tags_list = None
    for si, s in enumerate(scales):
        hms = [torch.from_numpy(np.ascontiguousarray(s["hm_lo"]))[None], torch.from_numpy(np.ascontiguousarray(s["hm_hi"]))[None]]
        tag = torch.from_numpy(np.ascontiguousarray(s["tag"]))[None]
        if s.get("hm_lo_f") is not None:
            fl = [torch.from_numpy(np.ascontiguousarray(s["hm_lo_f"]))[None], torch.from_numpy(np.ascontiguousarray(s["hm_hi_f"]))[None]]
            for i in range(2):
                hms[i] = (hms[i] + torch.flip(fl[i], [3])[:, COCO_FLIP_INDEX]) / 2
            tags = [tag, torch.flip(torch.from_numpy(np.ascontiguousarray(s["tag_f"]))[None], [3])[:, COCO_FLIP_INDEX]]
        else:
            tags = [tag]
        h, w = hms[-1].shape[-2:]
        hms = [_interp(hms[0], h, w), hms[1]]
        avg = torch.stack(hms).mean(dim=0)
        per_scale.append(_interp(avg, H, W))
        if si == tag_scale:
            tags_list = tags
    hm = per_scale[0] if len(per_scale) == 1 else torch.stack(per_scale).mean(dim=0)
    tg = torch.stack([_interp(t, H, W) for t in tags_list], dim=4)
    return hm[0], tg[0]


def parse_reference(hm: torch.Tensor, tags: torch.Tensor, M=30, det_thr=0.05, tag_thr=0.5, adjust=True, refine=True):
    """MPPEHeatmapParser.parse plus the intermediates of top_k / match_by_tag."""
    g = _import_grouping()
    parser = g.MPPEHeatmapParser(hm.shape[0], max_num_people=M, det_thr=det_thr, tag_thr=tag_thr)
    nmsd = parser.nms(hm.unsqueeze(0))[0]
    scores_k, idx_k = nmsd.view(hm.shape[0], -1).topk(M, dim=1)
    tags_k, coords_k, scores_k2 = parser.top_k(hm, tags)
    assert np.array_equal(scores_k.numpy(), scores_k2)
    matched = parser.match_by_tag(tags_k, coords_k, scores_k2)
    grouped, person_scores = parser.parse(hm, tags, adjust=adjust, refine=refine)
    return dict(nms=nmsd.numpy(), idx_k=idx_k.numpy().astype(np.int32), tags_k=tags_k, coords_k=coords_k,
                scores_k=scores_k2, matched=matched, grouped_joints=grouped, person_scores=person_scores)
