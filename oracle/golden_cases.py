"""TEST INFRASTRUCTURE ONLY -- the list of golden cases shared by oracle/gen_golden.py and tests/."""
import hashlib

import numpy as np


def _synth():
    from hpdecode import synth
    return synth


# name -> (generator name, kwargs, output size, M, det_thr, tag_thr)
CASES = {
    "netlike192_flip": ("netlike", dict(batch=1, size=192, flip=True, seed=21), 192, 30, 0.05, 0.5),
    "netlike192_noflip": ("netlike", dict(batch=1, size=192, flip=False, seed=22), 192, 30, 0.05, 0.5),
    "crowd192_flip": ("crowd", dict(batch=1, size=192, persons=8, flip=True, seed=23), 192, 30, 0.05, 0.5),
    "crowd256_q_flip": ("crowd", dict(batch=1, size=256, persons=20, flip=True, seed=24, quantised=True), 256, 30, 0.05, 0.5),
    "crowd256_q_noflip": ("crowd", dict(batch=1, size=256, persons=30, flip=False, seed=25, quantised=True), 256, 30, 0.05, 0.5),
    "crowd256_val_m20": ("crowd", dict(batch=1, size=256, persons=12, flip=False, seed=26), 256, 20, 0.1, 1.0),
    "empty256_fallback": ("netlike", dict(batch=1, size=256, flip=True, seed=27, negative_channels=tuple(range(17))), 256, 30, 0.05, 0.5),
    "crowd512_30_flip": ("crowd", dict(batch=1, size=512, persons=30, flip=True, seed=28), 512, 30, 0.05, 0.5),
    "netlike512_flip": ("netlike", dict(batch=1, size=512, flip=True, seed=29), 512, 30, 0.05, 0.5),
    # edge cases of the parser's parameters (CPU suite only: they widen what the oracle is pinned on)
    "crowd256_m5": ("crowd", dict(batch=1, size=256, persons=12, flip=True, seed=30), 256, 5, 0.05, 0.5),        # persons beyond max_num_people
    "crowd256_m32": ("crowd", dict(batch=1, size=256, persons=30, flip=True, seed=31), 256, 32, 0.05, 0.5),
    "crowd256_tight_thr": ("crowd", dict(batch=1, size=256, persons=10, flip=True, seed=32, tag_spread=0.6), 256, 30, 0.3, 0.2),   # many unmatched -> new persons
    "netlike256_some_negative": ("netlike", dict(batch=1, size=256, flip=True, seed=33, negative_channels=(0, 5, 9)), 256, 30, 0.05, 0.5),   # +-0 tails next to full channels
    "crowd192_q_dense": ("crowd", dict(batch=1, size=192, persons=30, flip=False, seed=34, quantised=True, missing_frac=0.4), 192, 30, 0.05, 0.5),   # ties + many missing joints (refine)
}


def make_inputs(name):
    gen, kw, size, M, det, tthr = CASES[name]
    scales = getattr(_synth(), gen)(**kw)
    return scales, size, M, det, tthr


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs_digest(scales) -> str:
    h = hashlib.sha256()
    for s in scales:
        for k in sorted(s):
            h.update(k.encode())
            h.update(np.ascontiguousarray(s[k]).tobytes())
    return h.hexdigest()
