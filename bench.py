#!/usr/bin/env python
"""Benchmark of the bottom-up decode path (BASELINE.json: decoded images/s, 512x512, HigherHRNet-W32, flip).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg2|cfg1|cfg3|cfg4] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the whole decode path (fused aggregation+NMS -> top-k -> grouping -> adjust/refine with the
result-record epilogue) over one batch of synthetic network outputs.  At N > 1 images shard by rank with no
collective on the hot path; the result records are gathered on rank 0 with NCCL on a separate stream, one
collective per group of steps.  Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.

Workloads (BASELINE.json configs): cfg2 (default) W32 512x512 flip batch 64; cfg1 the same at batch 1;
cfg3 W48 640x640 test scales 0.5/1.0/1.5 + flip batch 32; cfg4 planted 30-person crowds 512x512 batch 64.
--scaling weak: every rank decodes its own batch per step; strong: ONE batch split [r*B/N, (r+1)*B/N) (cfg2 as written).
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-human-pose_b200"))

import numpy as np  # noqa: E402

K_JOINTS, MAX_PEOPLE, DET_THR, TAG_THR = 17, 30, 0.05, 0.5
METRIC = "decoded images/s (512x512, HigherHRNet-W32, flip)"

WORKLOADS = {
    # name: arch (HRNet width), size, batch, test scales, inputs, batches in flight
    "cfg2": dict(arch=32, size=512, batch=64, scales=(1.0,), inputs="hrnet", streams=12, metric=METRIC,
                 text="HigherHRNet-W32 {size}x{size}, batch {batch} per GPU, flip test, single scale (BASELINE configs[2], sharded by image)"),
    "cfg1": dict(arch=32, size=512, batch=1, scales=(1.0,), inputs="hrnet", streams=1,
                 metric="decoded images/s (512x512, HigherHRNet-W32, flip, batch 1)",
                 text="HigherHRNet-W32 {size}x{size}, batch {batch}, flip test, single scale, one batch in flight (BASELINE configs[1])"),
    "cfg3": dict(arch=48, size=640, batch=32, scales=(0.5, 1.0, 1.5), inputs="hrnet", streams=4,
                 metric="decoded images/s (640x640, HigherHRNet-W48, test scales 0.5/1.0/1.5 + flip)",
                 text="HigherHRNet-W48 {size}x{size}, test scales 0.5/1.0/1.5 + flip, batch {batch} per GPU, scale-aggregated decode (BASELINE configs[3])"),
    "cfg4": dict(arch=32, size=512, batch=64, scales=(1.0,), inputs="crowd", streams=12,
                 metric="decoded images/s (512x512, planted 30-person crowds, flip)",
                 text="crowded-scene stress: planted 30-person maps {size}x{size}, batch {batch} per GPU, top-k 30, max 30 people (BASELINE configs[4])"),
}


def scale_input_size(size: int, s: float) -> int:
    return int(round(size * s / 64.0)) * 64


def algorithmic_bytes_per_image(size: int, scales=(1.0,), flip: bool = True) -> int:
    """reads 4*K*f*sum_s(q_s^2+h_s^2) + 4*K*f*q_1^2 (tags, scale 1 only), writes 4*K*S^2*(1+E)."""
    f = 2 if flip else 1
    read = 0
    for s in scales:
        n = scale_input_size(size, s)
        read += 4 * K_JOINTS * f * ((n // 4) ** 2 + (n // 2) ** 2)
    read += 4 * K_JOINTS * f * (size // 4) ** 2
    return read + 4 * K_JOINTS * size * size * (1 + f)


def make_inputs(batch: int, size: int, seed: int, unique: int = 8, kind: str = "netlike"):
    """`unique` distinct synthetic images tiled to the batch (generation is CPU-bound; addresses differ,
    so tiling does not make the decode cheaper)."""
    from hpdecode import synth
    if kind == "crowd":   # BASELINE config 4: planted 30-person scenes, grouping-bound
        base = synth.crowd(min(unique, batch), size, persons=30, flip=True, seed=seed)[0]
    else:
        base = synth.netlike(min(unique, batch), size, flip=True, seed=seed)[0]
    reps = -(-batch // min(unique, batch))
    return {k: np.ascontiguousarray(np.concatenate([v] * reps)[:batch]) for k, v in base.items()}


def resolve(args):
    wl = dict(WORKLOADS[args.workload])
    for key in ("batch", "size", "inputs", "streams"):
        if getattr(args, key) is not None:
            wl[key] = getattr(args, key)
    wl["name"] = args.workload
    wl["tag_scale"] = wl["scales"].index(1.0)
    return wl


def workload_config(wl, args, per_gpu_batch=None):
    B = wl["batch"] if per_gpu_batch is None else per_gpu_batch
    per_img = algorithmic_bytes_per_image(wl["size"], wl["scales"])
    out_mb = B * 4 * K_JOINTS * wl["size"] ** 2 * 3 / 1e6
    in_mb = B * per_img / 1e6 - out_mb
    l2 = ("inputs (%.0f MB/batch) and outputs (%.0f MB/batch) exceed the 126 MB L2; no flush needed" % (in_mb, out_mb)
          if in_mb + out_mb > 4 * 126 else
          "working set %.0f MB/batch is comparable to the 126 MB L2 and is NOT flushed between steps: batches in flight "
          "cycle through %d buffer sets; use the default batch for roofline numbers" % (in_mb + out_mb, wl["streams"]))
    return {"workload": wl["text"].format(size=wl["size"], batch=wl["batch"]), "name": wl["name"],
            "batch_per_gpu": B, "size": wl["size"], "test_scales": list(wl["scales"]), "flip": True,
            "max_people": MAX_PEOPLE, "inputs": wl["inputs"], "scaling_mode": args.scaling, "l2": l2}


# ------------------------------------------------------------------------------------------------
# CPU side: the reference arm / cpu_baseline (oracle/py_port.py = the reference's torch-CPU + NumPy calls)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import torch
    torch.set_num_threads(1)
    from oracle import py_port
    img_scales, size, tag_scale = args
    gj, ps = py_port.decode_image(img_scales, (size, size), MAX_PEOPLE, DET_THR, TAG_THR, tag_scale)
    return gj.shape[0]


class CpuPool:
    """One process per host core, each decoding whole images with the Python port (how the reference runs:
    one image per call, single-threaded NumPy/Python; the pool is the fair multi-core figure)."""

    def __init__(self, size: int, cores: int, host_scales, tag_scale: int):
        import multiprocessing as mp
        self.size, self.cores, self.tag_scale = size, cores, tag_scale
        n = host_scales[0]["hm_lo"].shape[0]
        self.images = [[{k: v[i] for k, v in s.items()} for s in host_scales] for i in range(n)]
        self.pool = mp.get_context("spawn").Pool(cores)
        self.pool.map(_cpu_worker, [(self.images[0], size, tag_scale)] * cores)     # import torch + warm caches, untimed

    def step(self, n_images: int) -> float:
        t = time.perf_counter()
        self.pool.map(_cpu_worker, [(self.images[i % len(self.images)], self.size, self.tag_scale) for i in range(n_images)],
                      chunksize=1)
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def produce_inputs(wl, batch, seed, dev):
    """One dict of tensors per test scale, on ``dev`` (GPU for the hrnet inputs when there is one).  hrnet inputs:
    ONE default-init network (weights from seed 1) for every rank, the images from ``seed`` -- rank r decodes its own
    images of the same model, like a sharded deployment (rank 0 / N = 1: seed 1 -> images from seed 2, exactly
    synth_net.network_outputs(B, S, seed=1), the batch tests/test_gpu_workloads.py checks image by image)."""
    import torch
    if wl["inputs"] == "hrnet":
        # BASELINE: heatmaps / tags "produced by random-init HigherHRNet weights on synthetic images"
        from hpdecode import synth_net
        return [synth_net.network_outputs(batch, scale_input_size(wl["size"], s), flip=True, seed=1, C=wl["arch"],
                                          device=dev, chunk=8 if wl["size"] <= 512 else 4, image_seed=seed + 1)
                for s in wl["scales"]]
    if len(wl["scales"]) != 1:
        raise SystemExit("--inputs netlike|crowd are single-scale generators")
    host = make_inputs(batch, wl["size"], seed=seed, kind=wl["inputs"])
    return [{k: torch.from_numpy(v).to(dev) for k, v in host.items()}]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = resolve(args)
    cores = min(host_cores(), 32)
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"      # producing the inputs is not part of the timed path
    ins = produce_inputs(wl, min(8, wl["batch"]), 1, dev)
    host_scales = [{k: v.contiguous().cpu().numpy() for k, v in s.items()} for s in ins]
    pool = CpuPool(wl["size"], cores, host_scales, wl["tag_scale"])
    sample = cores                                     # one image per core per step (~6 s of wall clock at 512x512)
    for _ in range(args.warmup):
        pool.step(sample)
    times = [pool.step(sample) for _ in range(args.steps)]
    pool.close()
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, args),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step (one per core), oracle/py_port.py = the reference's "
                                   "torch-CPU/NumPy/munkres calls; the reference itself is Python and cannot travel"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polling SM clock and throttle reasons every 20 ms.  It needs a few hundred ms to attach, so it
    is started early; only the samples stamped inside [mark_start(), stop()] -- the timed regions -- are kept."""
    Q = "timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def mark_start(self):
        self.t0 = time.time()

    def __init__(self, index: int):
        self.t0 = time.time()
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        import atexit
        atexit.register(self._kill)       # never leave the poller behind if the run dies before stop()

    def _kill(self):
        if self.p is not None and self.p.poll() is None:
            self.p.kill()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        t1 = time.time()
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        import datetime
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not (self.t0 - 0.02 <= ts <= t1 + 0.02):
                    continue
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to the GPU while the pinned staging buffers are allocated
    (first-touch then places the pages on the GPU's NUMA node).  Returns (note for the JSON line, affinity to restore)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        local = cpus & allowed
        if local and local != allowed:
            os.sched_setaffinity(0, local)
            return "allocated while bound to the %d GPU-local CPUs of %d" % (len(local), len(allowed)), allowed
        return "GPU lists all %d allowed CPUs as local (one NUMA node): nothing to bind" % len(allowed), None
    except Exception as e:      # noqa: BLE001
        return "no NVML affinity (%s)" % type(e).__name__, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hpdecode import BottomUpDecoder, DecodePipeline, ops
    from hpdecode.decoder import DecodeResult
    from hpdecode.parallel import gather_rows, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note, old_affinity = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local) if rank == 0 else None     # attaches while the inputs are produced

    wl = resolve(args)
    S, scales_cfg, tag_scale = wl["size"], wl["scales"], wl["tag_scale"]
    if args.scaling == "strong":
        # cfg2 as written: ONE batch, images [r*B/N, (r+1)*B/N) on rank r.  Every rank builds the same batch
        # (same seed) and keeps its shard, so N ranks together decode exactly the N = 1 batch.
        total_batch = wl["batch"]
        b0, b1 = shard_range(total_batch, rank, world)
        if b1 - b0 < 1 or total_batch % world:
            raise SystemExit("--scaling strong needs the batch to be a multiple of the number of GPUs")
        full = produce_inputs(wl, total_batch, 1, dev)
        resident = [{k: v[b0:b1] for k, v in s.items()} for s in full]
        B = b1 - b0
    else:
        B = wl["batch"]
        total_batch = B * world
        resident = produce_inputs(wl, B, 1 + rank + args.seed_offset, dev)
    torch.cuda.synchronize()
    pinned = [{k: v.contiguous().cpu().pin_memory() for k, v in s.items()} for s in resident]
    host = [{k: v[: min(8, B)].numpy() for k, v in s.items()} for s in pinned]
    E = 2
    dec = BottomUpDecoder(K_JOINTS, MAX_PEOPLE, DET_THR, TAG_THR, dev)
    params = ops.make_params(B, K_JOINTS, S, S, E, MAX_PEOPLE, DET_THR, TAG_THR, num_scales=len(scales_cfg), tag_scale=tag_scale)
    ROW = ops.record_layout(K_JOINTS, MAX_PEOPLE, E).row_bytes

    # DecodePipeline keeps NS batches in flight: batch i+1's bandwidth-bound aggregation kernel overlaps batch
    # i's latency-bound top-k / grouping / refine kernels.  Lane i's records land in ring[i]; every G lanes the block
    # ring[g*G:(g+1)*G] is gathered on rank 0 by ONE collective on a separate stream (and, in the e2e region, copied to
    # pinned host memory), so no decode stream ever waits for a peer.
    NS = max(1, wl["streams"])
    if args.streams is None and wl["streams"] > 1 and B < 32:
        NS = min(32, max(NS, 192 // B))       # small per-GPU batches (strong scaling): more, shorter batches in flight
    G = max(1, NS // 2) if args.gather_every is None else max(1, min(args.gather_every, NS))
    while NS % G:
        G -= 1
    n_groups = NS // G
    ring = torch.empty((NS, B, ROW), device=dev, dtype=torch.uint8)
    # one graph replay per step instead of eight launches: the host thread stays far ahead of the GPU even when
    # several ranks (and their NCCL proxy threads) share the box's cores
    use_graphs = args.graphs != "off"
    pipe = DecodePipeline(dec, depth=NS, records_ring=ring, use_graphs=use_graphs)
    comm = torch.cuda.Stream(device=dev)
    gathered = [torch.empty((world, G * B * ROW), device=dev, dtype=torch.uint8) if (world > 1 and rank == 0) else None
                for _ in range(n_groups)]
    result_host = [torch.empty((world, G * B * ROW), dtype=torch.uint8).pin_memory() if rank == 0 else None
                   for _ in range(n_groups)]
    staging = [[{k: torch.empty(v.shape, device=dev, dtype=v.dtype) for k, v in s.items()} for s in resident] for _ in range(NS)]
    if old_affinity is not None:
        os.sched_setaffinity(0, old_affinity)
    group_done = [torch.cuda.Event() for _ in range(n_groups)]
    lane_done = [torch.cuda.Event() for _ in range(NS)]
    state = {"to_host": False, "pending": [], "expected_records": None}
    main = torch.cuda.current_stream(dev)
    bufs = dec.buffers(B, S, S, E, slot=NS)            # a buffer set of its own for the sequential stage timing

    def flush_group(lanes):
        """Gather (and optionally copy to the host) the records of these consecutive lanes on the comm stream."""
        if not lanes:
            return
        g = lanes[0] // G
        for i in lanes:
            comm.wait_event(lane_done[i])
        with torch.cuda.stream(comm):
            block = ring[lanes[0]: lanes[0] + len(lanes)].view(-1)
            src = block
            if world > 1:
                out = gathered[g][:, : block.numel()] if rank == 0 else None
                if len(lanes) != G and rank == 0:
                    out = torch.empty((world, block.numel()), device=dev, dtype=torch.uint8)
                gather_rows(block, out, dst=0)
                src = out
            if state["to_host"] and rank == 0:
                dst = result_host[g][: (world if world > 1 else 1), : block.numel()]
                dst.copy_(src.view(dst.shape), non_blocking=True)
            group_done[g].record(comm)
        for i in lanes:
            pipe.lanes[i]["gate"] = group_done[g]       # the lane's records may be overwritten only after this

    def after_tail(ln, res):
        if args.no_gather:          # diagnosis only: what a step costs without the result gather
            return
        i = ln["slot"]
        lane_done[i].record()
        state["pending"].append(i)
        if len(state["pending"]) == G or i == NS - 1:
            flush_group(state["pending"])
            state["pending"] = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, submit):
        """n pipelined steps; returns device time (ms) from fork to join (decode lanes + gathers) on the main stream."""
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        pipe._next = 0
        t0.record(main)
        for _ in range(n):
            submit()
        flush_group(state["pending"])
        state["pending"] = []
        pipe.drain()
        t1.record(main)
        barrier()
        return t0.elapsed_time(t1)

    def value_step():
        pipe.submit(resident, (S, S), tag_scale=tag_scale, after_tail=after_tail)

    def parity_check():
        """Before anything is timed: decode this rank's batch once and compare images of it with the CPU oracle
        (bit-exact grouped joints and person scores); at N > 1 rank 0 also compares every rank's gathered rows
        with a checksum the owning rank computed locally."""
        from oracle import cpu_oracle      # the checker, never the thing measured
        res = dec.decode(resident, (S, S), tag_scale=tag_scale, slot=NS + 1)
        rows = res.records.cpu().numpy()
        out = DecodeResult.unpack(rows, MAX_PEOPLE, K_JOINTS, E)
        n_check = min(max(1, 8 // world), len(host[0]["hm_lo"]), B)
        ok = True
        for b in range(n_check):
            hm_o, tg_o = cpu_oracle.aggregate([{k: v[b] for k, v in s.items()} for s in host], (S, S), tag_scale=tag_scale)
            ref = cpu_oracle.parse(hm_o, tg_o, MAX_PEOPLE, DET_THR, TAG_THR)
            gj, ps = out[b]
            ok = ok and gj.shape == ref["grouped_joints"].shape and \
                np.array_equal(np.asarray(gj, np.float32).view(np.uint32), ref["grouped_joints"].view(np.uint32)) and \
                np.array_equal(np.asarray(ps, np.float32).view(np.uint32), ref["person_scores"].view(np.uint32))
        info = {"checked": n_check * world, "ok": bool(ok), "against": "oracle/hpd_oracle.cpp (grouped joints + person scores, bit-exact)"}
        state["expected_records"] = res.records.clone()      # what every pipelined step must reproduce for this batch
        if world > 1:
            digest = torch.frombuffer(bytearray(hashlib.sha256(rows.tobytes()).digest()), dtype=torch.uint8).to(dev)
            digests = [torch.empty_like(digest) for _ in range(world)]
            dist.all_gather(digests, digest)
            flag = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            out_rows = torch.empty((world, B * ROW), device=dev, dtype=torch.uint8) if rank == 0 else None
            gather_rows(res.records.view(-1), out_rows, dst=0)
            info["ok"] = bool(flag.item())
            if rank == 0:
                g = out_rows.cpu().numpy()
                same = [hashlib.sha256(g[r].tobytes()).digest() == bytes(digests[r].cpu().tolist()) for r in range(world)]
                info["gathered_rows_match_rank_checksums"] = bool(all(same))
                info["ok"] = info["ok"] and all(same)
        return info

    parity = parity_check()
    if rank == 0 and not parity["ok"]:
        print(json.dumps({"metric": wl["metric"], "error": "parity check against the oracle FAILED; nothing was timed",
                          "parity": parity}), flush=True)
    if not parity["ok"]:
        raise SystemExit(3)

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    W = max(args.warmup, 3)
    run_steps(W * NS, value_step)
    if sampler:
        sampler.mark_start()
    # per-stage durations: K sequential steps, one batch in flight, CUDA events on the launching stream
    stages = DecodePipeline.STAGES
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(args.steps)]
    for st in stages:
        ops.run_stage(st, bufs, params, scales=resident)
    for i in range(args.steps):
        for j, st in enumerate(stages):
            evs[i][j].record()
            ops.run_stage(st, bufs, params, scales=resident)
        evs[i][len(stages)].record()
    barrier()
    stage_ms = [statistics.mean(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(args.steps)) for j in range(len(stages))]
    seq_ms = evs[0][0].elapsed_time(evs[-1][len(stages)]) / args.steps
    l0 = ops.launches_total()
    ms_total = run_steps(args.steps, value_step)
    launches = ops.launches_total() - l0
    # the timed steps really decoded: every lane's records equal the ones checked against the oracle above
    lanes_used = min(NS, args.steps)
    parity["pipelined_records_identical"] = bool(all(torch.equal(ring[i], state["expected_records"]) for i in range(lanes_used)))
    if world > 1:
        flag = torch.tensor([int(parity["pipelined_records_identical"])], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity["pipelined_records_identical"] = bool(flag.item())
    parity["ok"] = parity["ok"] and parity["pipelined_records_identical"]

    # ---- e2e: host buffers, H2D of the inputs and D2H of the result records inside the timed region ---------
    state["to_host"] = True

    def e2e_step():
        def h2d(ln):
            st = staging[ln["slot"]]
            for s_dst, s_src in zip(st, pinned):
                for k in s_dst:
                    s_dst[k].copy_(s_src[k], non_blocking=True)
            return st

        pipe.submit(resident, (S, S), tag_scale=tag_scale, before_agg=h2d, after_tail=after_tail)

    run_steps(2 * NS, e2e_step)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = run_steps(e2e_steps, e2e_step)
    clocks = sampler.stop() if sampler else None
    e2e_ok = None
    if rank == 0:   # the host copy of the last full group really holds every rank's records
        last = host_records_sane(result_host[0], world, G, B, ROW)
        e2e_ok = last

    # max over ranks
    per_rank_ms = [ms_total]
    if world > 1:
        mine = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        every = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        per_rank_ms = [float(x.item()) for x in every]
        t = torch.tensor([ms_total, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = t.tolist()

    if rank == 0:
        n_person = bufs.n_person.cpu().numpy()
        peak, peak_src = measured_peak()
        bytes_launch = B * algorithmic_bytes_per_image(S, scales_cfg)
        achieved = bytes_launch / (stage_ms[0] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.isfile(tp) and wl["name"] in ("cfg2", "cfg3") and B == WORKLOADS[wl["name"]]["batch"]:
            try:
                tj = json.load(open(tp))
                traffic = tj.get("agg_nms_dram_bytes_per_launch" if wl["name"] == "cfg2" else "agg_nms_ms_dram_bytes_per_launch")
            except Exception:
                traffic = None
        imgs_per_step = total_batch
        line = {
            "metric": wl["metric"], "value": imgs_per_step * args.steps / (ms_total * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": ("synthetic: outputs of a default-init HigherHRNet-W%d (stock PyTorch, hpdecode/synth_net.py) on seeded N(0,1) "
                     "images, flipped forward included; no dataset/checkpoint offline" % wl["arch"]) if wl["inputs"] == "hrnet" else
                    "synthetic (%s)" % ("planted 30-person Gaussian peaks with per-person tags, SURVEY 8(d)-5" if wl["inputs"] == "crowd"
                                        else "seeded smooth random fields with default-init HigherHRNet value ranges"),
            "config": workload_config(wl, args, B),
            "roofline": {"bound": "hbm", "kernel": "agg_nms (fused aggregation + NMS)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": stage_ms[0],
                         "timing": "CUDA events around each of the K launches, one batch in flight"},
            "stage_ms": dict(zip(stages, stage_ms)), "sequential_ms_per_step": seq_ms, "streams": NS,
            "cuda_graphs": bool(use_graphs), "gather": {"steps_per_collective": G, "stream": "dedicated", "row_bytes": ROW},
            "grouping": {"kernel_ms": stage_ms[2], "latency_us_per_image": 1e3 * stage_ms[2],
                         "amortised_us_per_image": 1e3 * stage_ms[2] / B, "resident_warps": 4 * B,
                         "sm_occupancy_pct": 100.0 * 4 * B / (torch.cuda.get_device_properties(dev).multi_processor_count * 64),
                         "note": "one CTA of four warps per image; all images of a batch run concurrently, so the kernel duration is "
                                 "each image's latency"},
            "persons_per_image": float(n_person.mean()),
            "e2e": {"value": imgs_per_step * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for s in pinned for v in s.values())) * world,
                    "d2h_bytes_per_step": int(world * B * ROW), "steps": e2e_steps,
                    "host_copy_holds_every_rank": e2e_ok, "pinned_staging": numa_note},
            "gpu_launches": launches, "clocks": clocks, "parity": parity,
            "per_rank_ms_per_step": [round(x / args.steps, 4) for x in per_rank_ms],
        }
        if args.no_gather:
            line["invalid"] = "--no-gather: diagnosis run, the result records were not gathered"
        if world == 1 and not args.no_cpu_baseline:
            cores = min(host_cores(), 32)
            pool = CpuPool(S, cores, host, tag_scale)
            sample = cores * 2 if S <= 512 and len(scales_cfg) == 1 else cores
            sec = pool.step(sample)
            pool.close()
            line["cpu_baseline"] = {"value": sample / sec, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} images over {cores} processes in {sec:.1f} s (oracle/py_port.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_records_sane(host_block, world, G, B, ROW) -> bool:
    """Every rank's slice of the host copy parses as records with a sane person count (the D2H really carried
    world * G * B rows, not just rank 0's)."""
    from hpdecode.decoder import Records
    n = world if world > 1 else 1
    raw = host_block[:n].numpy().reshape(n * G * B, ROW)
    rec = Records(raw, MAX_PEOPLE, K_JOINTS, 2)
    return bool(((rec.n_person >= 1) & (rec.n_person <= MAX_PEOPLE)).all())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inputs", default=None, choices=["hrnet", "netlike", "crowd"],
                    help="hrnet: outputs of a default-init HigherHRNet on seeded random images; netlike / crowd: CPU-generated fields")
    ap.add_argument("--streams", type=int, default=None, help="batches in flight per GPU (1 = strictly sequential)")
    ap.add_argument("--gather-every", type=int, default=None, help="steps per gather collective (default: streams / 2)")
    ap.add_argument("--seed-offset", type=int, default=0, help="diagnosis: rank r's inputs use seed 1 + r + offset")
    ap.add_argument("--no-gather", action="store_true", help="diagnosis: skip the gather of the result records (the line is then not a valid bench line)")
    ap.add_argument("--graphs", default="auto", choices=["auto", "on", "off"],
                    help="CUDA graphs for the per-step launches of the pipelined regions (auto = on)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
