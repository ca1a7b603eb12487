"""One batched hpd_prepare_input call (64 raw 480x640 images -> 512x704 network inputs) for ncu / timing.
usage: python tools/profile_prepare_input.py   (prints the CUDA-event time of the kernel and its HBM rate)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-human-pose_b200"))
import numpy as np
import torch
from hpdecode import geometry

rng = np.random.default_rng(0)
imgs = [torch.from_numpy(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)).cuda() for _ in range(64)]
for _ in range(3):
    x, _, _ = geometry.prepare_input(imgs, 512, "cuda:0")
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(10):
    x, _, _ = geometry.prepare_input(imgs, 512, "cuda:0")
t1.record()
torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 10
nbytes = x.numel() * 4 + sum(i.numel() for i in imgs)
print("prepare_input: %d images -> %s in %.3f ms per call (host marshalling included), %.0f GB/s of output + input bytes"
      % (len(imgs), tuple(x.shape), ms, nbytes / ms / 1e6))
