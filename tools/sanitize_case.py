"""Small decode cases for compute-sanitizer (memcheck / racecheck): every kernel, both aggregation kernels,
partial warps (W % 128 != 0), tie-heavy top-k (exact heap replay), multi-scale, E = 1 and E = 2."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-human-pose_b200"))
import torch
from hpdecode import BottomUpDecoder, synth

dev = lambda sc: [{k: torch.from_numpy(v).cuda() for k, v in s.items()} for s in sc]
dec = BottomUpDecoder(17, 30, 0.05, 0.5, "cuda:0")
for name, sc, hw, ts in [
    ("crowd192_flip", synth.crowd(2, 192, persons=8, flip=True, seed=3), (192, 192), 0),
    ("crowd256_q_noflip", synth.crowd(1, 256, persons=30, flip=False, seed=5, quantised=True), (256, 256), 0),
    ("netlike256_ms", synth.netlike(1, 256, True, seed=7, scales=(1.0, 1.5)), (256, 256), 0),
]:
    res = dec.decode(dev(sc), hw, tag_scale=ts)
    torch.cuda.synchronize()
    print(name, [len(g) for g, _ in res.to_numpy()])
print("done")
