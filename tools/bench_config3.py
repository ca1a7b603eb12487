"""BASELINE config 3: HigherHRNet-W48 640x640, test scales 0.5/1.0/1.5 + flip, batch 32, scale-aggregated decode.
Not the headline bench (bench.py); a timing of the multi-scale path (generic aggregation kernel) for profiles/."""
import json, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-human-pose_b200"))
import torch
from hpdecode import BottomUpDecoder, ops, synth_net

B, S, scales = int(os.environ.get("B", 32)), 640, (0.5, 1.0, 1.5)
dev = torch.device("cuda:0")
ins = []
for i, sc in enumerate(scales):
    s_in = int(round(S * sc / 64.0)) * 64
    # one network, inputs of different size (the reference would resize one image; random images of each size here)
    ins.append(synth_net.network_outputs(B, s_in, flip=True, seed=5, C=48, device=dev, chunk=4))
dec = BottomUpDecoder(17, 30, 0.05, 0.5, dev)
bufs = dec.buffers(B, S, S, 2)
params = ops.make_params(B, 17, S, S, 2, 30, 0.05, 0.5, num_scales=3, tag_scale=1)
stages = ("aggregate_nms", "topk", "group", "adjust_refine")
for _ in range(3):
    for st in stages:
        ops.run_stage(st, bufs, params, scales=ins)
torch.cuda.synchronize()
K = 10
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
for i in range(K):
    for j, st in enumerate(stages):
        ev[i][j].record()
        ops.run_stage(st, bufs, params, scales=ins)
    ev[i][4].record()
torch.cuda.synchronize()
ms = [statistics.mean(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(K)) for j in range(4)]
f = 2
read = sum(4 * 17 * f * ((int(round(S * sc / 64)) * 16) ** 2 + (int(round(S * sc / 64)) * 32) ** 2) for sc in scales) + 4 * 17 * f * 160 * 160
write = 4 * 17 * S * S * 3
tot = sum(ms)
print(json.dumps({"workload": "config 3: W48 640x640, scales 0.5/1.0/1.5 + flip, batch %d" % B, "images_per_s": B / (tot * 1e-3),
                  "ms_per_step": tot, "stage_ms": dict(zip(stages, ms)),
                  "agg_algorithmic_GBps": B * (read + write) / (ms[0] * 1e-3) / 1e9, "algorithmic_MB_per_image": (read + write) / 1e6,
                  "persons_per_image": float(bufs.n_person.float().mean())}))
