"""COCO keypoint result records from decoded poses (a "next" row: the feed of
/root/reference/src/keypoints/bin/eval.py:28-47).  Host-side, a few hundred bytes per person; the
batched decoder produces the inputs for a whole batch at once, so the dataset loop can finally be
fed B images per call instead of one."""
from typing import Iterable, List, Sequence, Tuple

import numpy as np

from .results import InferenceKeypointsResult


def coco_records(image_id: int, kpts_coords: np.ndarray, obj_scores: np.ndarray) -> List[dict]:
    """eval.py:31-47: one record per person, keypoints = [x, y, 1] * K, score = the person score."""
    out = []
    for kpts, score in zip(kpts_coords, obj_scores):
        flat = np.zeros((len(kpts) * 3,))
        flat[::3] = kpts[:, 0]
        flat[1::3] = kpts[:, 1]
        flat[2::3] = 1
        out.append({"image_id": int(image_id), "category_id": 1, "keypoints": flat.tolist(),
                    "score": np.asarray(score).mean().item()})
    return out


def batch_to_coco(image_ids: Sequence[int], decoded: Iterable[Tuple[np.ndarray, np.ndarray]],
                  centers: Sequence, scales: Sequence, hm_size: Tuple[int, int]) -> List[dict]:
    """decoded: DecodeResult.to_numpy() of a batch; centers / scales: per image, from the reference's
    resize_align_multi_scale (base/transforms/utils.py:89-97); hm_size = (W, H) of the network input.
    Coordinates are back-projected to the raw image exactly like results.py:189-201,244."""
    records = []
    for image_id, (grouped, scores), c, s in zip(image_ids, decoded, centers, scales):
        coords = InferenceKeypointsResult.get_final_kpts_coords(grouped[..., :2], c, s, hm_size)
        records.extend(coco_records(image_id, coords, scores))
    return records
