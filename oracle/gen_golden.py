"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, build container only) on the seeded synthetic inputs of oracle/golden_cases.py.

    python oracle/gen_golden.py [case ...]

Each file stores what the reference produced: top-k indices/scores/coords/tags, match_by_tag output,
final (grouped_joints, person_scores) and SHA-256 digests of the aggregated heatmaps, tag maps and
NMS'd map (the maps themselves are too large to commit), plus the digest of the inputs so a test
can tell "generator not reproducible on this platform" from "decoder wrong".
Versions used are recorded in tests/golden/MANIFEST.json.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-human-pose_b200"))

from hpdecode import synth  # noqa: E402
from oracle import golden_cases, ref_runner  # noqa: E402


def main():
    import torch
    assert ref_runner.available(), "needs /root/reference"
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    manifest = {"torch": torch.__version__, "numpy": np.__version__, "reference": "thawro/pytorch-human-pose @ /root/reference",
                "munkres": "restated 1.1.4 (oracle/refshim/munkres.py) -- parity unpinned", "cases": {}}
    only = set(sys.argv[1:])                       # optional: regenerate just these cases
    man_path = os.path.join(out_dir, "MANIFEST.json")
    if only and os.path.isfile(man_path):
        manifest["cases"] = json.load(open(man_path))["cases"]
    for name in golden_cases.CASES:
        if only and name not in only:
            continue
        scales, size, M, det, tthr = golden_cases.make_inputs(name)
        img = synth.image_slice(scales, 0)
        hm, tg = ref_runner.aggregate_torch(img, (size, size))
        R = ref_runner.parse_reference(hm, tg, M, det, tthr)
        gj = np.asarray(R["grouped_joints"])
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            inputs_digest=golden_cases.inputs_digest(scales), agg_hm_sha=golden_cases.sha(hm.numpy()),
            agg_tags_sha=golden_cases.sha(tg.numpy()), nms_sha=golden_cases.sha(R["nms"]),
            idx_k=R["idx_k"], scores_k=R["scores_k"], coords_k=R["coords_k"], tags_k=R["tags_k"],
            matched=np.asarray(R["matched"]), grouped_joints=gj, person_scores=np.asarray(R["person_scores"]))
        manifest["cases"][name] = dict(persons=int(gj.shape[0]), dtype=str(gj.dtype), filled=int((gj[..., 2] != 0).sum()))
        print(name, manifest["cases"][name], flush=True)
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
