#!/usr/bin/env python
"""Benchmark of the bottom-up decode path (BASELINE.json: decoded images/s, 512x512, HigherHRNet-W32, flip).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the whole decode path (fused aggregation+NMS -> top-k -> grouping -> adjust/refine)
over one batch of synthetic network outputs; at N > 1 every rank decodes its own batch (images shard
naturally, weak scaling) and the packed pose lists are gathered on rank 0 with NCCL inside the step.
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
import argparse
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


def algorithmic_bytes_per_image(size: int, flip: bool) -> int:
    """reads 4*K*f*(q^2+h^2) + 4*K*f*q^2, writes 4*K*S^2*(1+E)."""
    f = 2 if flip else 1
    q, h = size // 4, size // 2
    return 4 * K_JOINTS * f * (q * q + h * h) + 4 * K_JOINTS * f * q * q + 4 * K_JOINTS * size * size * (1 + f)


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


# ------------------------------------------------------------------------------------------------
# CPU side: the reference arm / cpu_baseline (oracle/py_port.py = the reference's torch-CPU + NumPy calls)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import torch
    torch.set_num_threads(1)
    from oracle import py_port
    img, size = args
    gj, ps = py_port.decode_image([img], (size, size), MAX_PEOPLE, DET_THR, TAG_THR)
    return gj.shape[0]


class CpuPool:
    """One process per host core, each decoding whole images with the Python port (how the reference runs:
    one image per call, single-threaded NumPy/Python; the pool is the fair multi-core figure)."""

    def __init__(self, size: int, cores: int, inputs):
        import multiprocessing as mp
        self.size, self.cores = size, cores
        self.images = [{k: v[i] for k, v in inputs.items()} for i in range(inputs["hm_lo"].shape[0])]
        self.pool = mp.get_context("spawn").Pool(cores)
        self.pool.map(_cpu_worker, [(self.images[0], size)] * cores)     # import torch + warm caches, untimed

    def step(self, n_images: int) -> float:
        t = time.perf_counter()
        self.pool.map(_cpu_worker, [(self.images[i % len(self.images)], self.size) for i in range(n_images)], chunksize=1)
        return time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), 32)
    if args.inputs == "hrnet":
        import torch
        from hpdecode import synth_net
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"      # producing the inputs is not part of the timed path
        r = synth_net.network_outputs(min(8, args.batch), args.size, flip=True, seed=1, C=32, device=dev, chunk=2)
        inputs = {k: v.contiguous().cpu().numpy() for k, v in r.items()}
    else:
        inputs = make_inputs(min(8, args.batch), args.size, seed=1, kind=args.inputs)
    pool = CpuPool(args.size, cores, inputs)
    sample = cores                                     # one image per core per step (~6 s of wall clock)
    for _ in range(args.warmup):
        pool.step(sample)
    times = [pool.step(sample) for _ in range(args.steps)]
    pool.close()
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step (one per core), oracle/py_port.py = the reference's "
                                   "torch-CPU/NumPy/munkres calls; the reference itself is Python and cannot travel"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    in_mb = args.batch * (algorithmic_bytes_per_image(args.size, True) - 4 * K_JOINTS * args.size * args.size * 3) / 1e6
    out_mb = args.batch * 4 * K_JOINTS * args.size * args.size * 3 / 1e6
    l2 = ("inputs (%.0f MB/batch) and outputs (%.0f MB/batch) exceed the 126 MB L2; no flush needed" % (in_mb, out_mb)
          if in_mb + out_mb > 4 * 126 else
          "working set %.0f MB/batch is comparable to the 126 MB L2 and is NOT flushed between steps: use the default "
          "batch for roofline numbers" % (in_mb + out_mb))
    return {"workload": f"HigherHRNet-W32 {args.size}x{args.size}, batch {args.batch} per GPU, flip test, single scale "
                        "(BASELINE configs[2], sharded by image)",
            "batch_per_gpu": args.batch, "size": args.size, "flip": True, "max_people": MAX_PEOPLE, "inputs": args.inputs,
            "l2": l2}


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hpdecode import BottomUpDecoder, ops
    from hpdecode.parallel import gather_packed_equal

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local) if rank == 0 else None     # attaches while the inputs are produced

    B, S = args.batch, args.size
    if args.inputs == "hrnet":
        # BASELINE: heatmaps / tags "produced by random-init HigherHRNet weights on synthetic images"
        from hpdecode import synth_net
        resident = synth_net.network_outputs(B, S, flip=True, seed=1 + rank, C=32, device=dev)
        torch.cuda.synchronize()
        pinned = {k: v.contiguous().cpu().pin_memory() for k, v in resident.items()}
        host = {k: v[: min(8, B)].numpy() for k, v in pinned.items()}
    else:
        host = make_inputs(B, S, seed=1 + rank, kind=args.inputs)
        pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
        resident = {k: v.to(dev) for k, v in pinned.items()}
    dec = BottomUpDecoder(K_JOINTS, MAX_PEOPLE, DET_THR, TAG_THR, dev)
    params = ops.make_params(B, K_JOINTS, S, S, 2, MAX_PEOPLE, DET_THR, TAG_THR)
    F = MAX_PEOPLE * K_JOINTS * 5 + MAX_PEOPLE + 2
    from hpdecode.decoder import DecodeResult

    # DecodePipeline keeps NS batches in flight: batch i+1's bandwidth-bound aggregation kernel overlaps batch
    # i's latency-bound top-k / grouping / refine kernels (high-priority streams, a few SMs each).
    from hpdecode.decoder import DecodePipeline
    NS = max(1, args.streams)
    pipe = DecodePipeline(dec, depth=NS, split_priority=args.split_priority)
    extra = [{"gathered": torch.empty((world * B, F), device=dev) if (world > 1 and rank == 0) else None,
              "staging": {k: torch.empty(v.shape, device=dev, dtype=v.dtype) for k, v in resident.items()},
              "result_host": torch.empty((B, F), dtype=torch.float32).pin_memory()} for _ in range(NS)]
    stages = ("aggregate_nms", "topk", "group", "adjust_refine")
    main = torch.cuda.current_stream(dev)
    bufs = dec.buffers(B, S, S, 2, slot=0)

    def finish(ln, res):
        packed = res.packed()
        if world > 1:
            gather_packed_equal(packed, extra[ln["slot"]]["gathered"], dst=0)
        return packed

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n, submit):
        """n pipelined steps; returns device time (ms) from fork to join on the main stream."""
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0.record(main)
        for _ in range(n):
            submit()
        pipe.drain()
        t1.record(main)
        barrier()
        return t0.elapsed_time(t1)

    def value_step():
        pipe.submit([resident], (S, S), after_tail=finish)

    def parity_check():
        """Before anything is timed: decode this rank's batch once and compare images of it with the CPU oracle
        (bit-exact grouped joints and person scores); at N > 1 rank 0 also compares every rank's gathered rows
        with a checksum the owning rank computed locally."""
        import hashlib
        from oracle import cpu_oracle      # the checker, never the thing measured
        res = dec.decode([resident], (S, S), slot=NS)
        packed = res.packed()
        rows = packed.cpu().numpy()
        out = DecodeResult.unpack(rows, MAX_PEOPLE, K_JOINTS, 2)
        n_check = min(max(1, 8 // world), len(host["hm_lo"]), B)
        ok = True
        for b in range(n_check):
            hm_o, tg_o = cpu_oracle.aggregate([{k: v[b] for k, v in host.items()}], (S, S))
            ref = cpu_oracle.parse(hm_o, tg_o, MAX_PEOPLE, DET_THR, TAG_THR)
            gj, ps = out[b]
            ok = ok and gj.shape == ref["grouped_joints"].shape and \
                np.array_equal(np.asarray(gj, np.float32).view(np.uint32), ref["grouped_joints"].view(np.uint32)) and \
                np.array_equal(np.asarray(ps, np.float32).view(np.uint32), ref["person_scores"].view(np.uint32))
        info = {"checked": n_check * world, "ok": bool(ok), "against": "oracle/hpd_oracle.cpp (grouped joints + person scores, bit-exact)"}
        if world > 1:
            digest = torch.frombuffer(bytearray(hashlib.sha256(rows.tobytes()).digest()), dtype=torch.uint8).to(dev)
            digests = [torch.empty_like(digest) for _ in range(world)]
            dist.all_gather(digests, digest)
            flag = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            gathered = torch.empty((world * B, F), device=dev) if rank == 0 else None
            gather_packed_equal(packed, gathered, dst=0)
            info["ok"] = bool(flag.item())
            if rank == 0:
                g = gathered.cpu().numpy()
                same = [hashlib.sha256(g[r * B:(r + 1) * B].tobytes()).digest() == bytes(digests[r].cpu().tolist())
                        for r in range(world)]
                info["gathered_rows_match_rank_checksums"] = bool(all(same))
                info["ok"] = info["ok"] and all(same)
        return info

    parity = parity_check()
    if rank == 0 and not parity["ok"]:
        print(json.dumps({"metric": METRIC, "error": "parity check against the oracle FAILED; nothing was timed", "parity": parity}),
              flush=True)
    if not parity["ok"]:
        raise SystemExit(3)

    # ---- value: inputs resident in HBM -------------------------------------------------------------
    run_steps(max(args.warmup, 3) * NS, value_step)
    if sampler:
        sampler.mark_start()
    # per-stage durations: K sequential steps, one batch in flight, CUDA events on the launching stream
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(args.steps)]
    for i in range(args.steps):
        for j, st in enumerate(stages):
            evs[i][j].record()
            ops.run_stage(st, bufs, params, scales=[resident])
        evs[i][len(stages)].record()
    barrier()
    stage_ms = [statistics.mean(evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(args.steps)) for j in range(len(stages))]
    seq_ms = evs[0][0].elapsed_time(evs[-1][len(stages)]) / args.steps
    l0 = ops.launches_total()
    ms_total = run_steps(args.steps, value_step)
    launches = ops.launches_total() - l0

    # ---- e2e: host buffers, H2D of the inputs and D2H of the pose lists inside the timed region ---------
    def e2e_step():
        def h2d(ln):
            st = extra[ln["slot"]]["staging"]
            for k in st:
                st[k].copy_(pinned[k], non_blocking=True)
            return [st]

        def d2h(ln, res):
            packed = finish(ln, res)
            src = extra[ln["slot"]]["gathered"] if (world > 1 and rank == 0) else packed
            extra[ln["slot"]]["result_host"].copy_(src[:B], non_blocking=True)

        pipe.submit([resident], (S, S), before_agg=h2d, after_tail=d2h)

    run_steps(2 * NS, e2e_step)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = run_steps(e2e_steps, e2e_step)
    clocks = sampler.stop() if sampler else None

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms = t.tolist()

    if rank == 0:
        n_person = bufs.n_person.cpu().numpy()
        peak, peak_src = measured_peak()
        bytes_launch = B * algorithmic_bytes_per_image(S, True)
        achieved = bytes_launch / (stage_ms[0] * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.isfile(tp):
            try:
                traffic = json.load(open(tp)).get("agg_nms_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms_total * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": ("synthetic: outputs of a default-init HigherHRNet-W32 (stock PyTorch, hpdecode/synth_net.py) on seeded N(0,1) "
                     "images, flipped forward included; no dataset/checkpoint offline") if args.inputs == "hrnet" else
                    "synthetic (seeded smooth random fields with default-init HigherHRNet value ranges)",
            "config": workload_config(args),
            "roofline": {"bound": "hbm", "kernel": "agg_nms (fused aggregation + NMS)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": stage_ms[0],
                         "timing": "CUDA events around each of the K launches, one batch in flight"},
            "stage_ms": dict(zip(stages, stage_ms)), "sequential_ms_per_step": seq_ms, "streams": NS,
            "grouping": {"kernel_ms": stage_ms[2], "latency_us_per_image": 1e3 * stage_ms[2],
                         "amortised_us_per_image": 1e3 * stage_ms[2] / B, "resident_warps": 4 * B,
                         "sm_occupancy_pct": 100.0 * 4 * B / (torch.cuda.get_device_properties(dev).multi_processor_count * 64),
                         "note": "one CTA of four warps per image; all images of a batch run concurrently, so the kernel duration is "
                                 "each image's latency"},
            "persons_per_image": float(n_person.mean()),
            "e2e": {"value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(sum(v.numel() * 4 for v in pinned.values())) * world,
                    "d2h_bytes_per_step": int(B * F * 4), "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = min(host_cores(), 32)
            pool = CpuPool(S, cores, {k: v[:8] for k, v in host.items()})
            sample = cores * 2
            sec = pool.step(sample)
            pool.close()
            line["cpu_baseline"] = {"value": sample / sec, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} images over {cores} processes in {sec:.1f} s (oracle/py_port.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-priority", action="store_true")
    ap.add_argument("--inputs", default="hrnet", choices=["hrnet", "netlike", "crowd"],
                    help="hrnet: outputs of a default-init HigherHRNet-W32 on seeded random images; netlike: CPU-generated fields")
    ap.add_argument("--streams", type=int, default=8, help="batches in flight per GPU (1 = strictly sequential)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
