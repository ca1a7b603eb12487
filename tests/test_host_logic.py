"""CPU suite: host-side logic -- synthetic generators, result packing, sharding and the
world_size-2 gather over gloo (the N>1 path of bench.py without GPUs)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hpdecode import synth
from hpdecode.decoder import DecodeResult, _finish
from hpdecode.parallel import gather_packed, shard_range


def test_synth_is_deterministic_and_shaped():
    a = synth.netlike(2, 192, True, seed=3)
    b = synth.netlike(2, 192, True, seed=3)
    assert all(np.array_equal(a[0][k], b[0][k]) for k in a[0])
    assert a[0]["hm_lo"].shape == (2, 17, 48, 48) and a[0]["hm_hi"].shape == (2, 17, 96, 96)
    assert set(a[0]) == {"hm_lo", "hm_hi", "tag", "hm_lo_f", "hm_hi_f", "tag_f"}
    c = synth.crowd(1, 256, persons=5, flip=False, seed=1)
    assert set(c[0]) == {"hm_lo", "hm_hi", "tag"} and c[0]["hm_lo"].dtype == np.float32
    ms = synth.netlike(1, 640, True, seed=1, scales=(0.5, 1.0, 1.5))
    assert [s["hm_hi"].shape[-1] for s in ms] == [160, 320, 480]


def test_unpack_and_fallback_dtype():
    M, K, E = 4, 17, 2
    D = 3 + E
    row = np.zeros(M * K * D + M + 2, np.float32)
    row[: 2 * K * D] = np.arange(2 * K * D)
    row[M * K * D: M * K * D + 2] = [0.5, 0.25]
    row[-2] = 2
    poses, scores = DecodeResult.unpack(row[None], M, K, E)[0]
    assert poses.shape == (2, K, D) and poses.dtype == np.float32 and list(scores) == [0.5, 0.25]
    fb = np.zeros((1, K, D), np.float32)
    fb[..., 2] = np.float32(0.01)
    p, s = _finish(fb, np.zeros(1, np.float32), 1)
    assert p.dtype == np.float64 and p[0, 0, 2] == 0.01 and s[0] == np.full((1, K), 0.01).mean(1)[0]


def test_shard_range_partitions():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(5, rank, world)
    packed = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 3)[b:e].clone()
    out = gather_packed(packed, dst=0)
    if rank == 0:
        q.put(out.numpy())
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_packed_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got, np.arange(15, dtype=np.float32).reshape(5, 3))


def _worker_equal(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hpdecode.parallel import gather_packed_equal
    rows = torch.arange(3 * 4, dtype=torch.float32).reshape(3, 4) + 100 * rank
    out = torch.full((world * 3, 4), -1.0) if rank == 0 else None
    got = gather_packed_equal(rows, out, dst=0)
    if rank == 0:
        assert got is out
        q.put(out.numpy())
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_packed_equal_world2_gloo():
    """The equal-shard gather bench.py uses at N > 1: rank r's rows land at [r*b, (r+1)*b) of the preallocated tensor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_equal, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    base = np.arange(12, dtype=np.float32).reshape(3, 4)
    assert np.array_equal(got, np.concatenate([base, base + 100]))


def test_back_projection_and_coco_records():
    """results.py:158-201 + eval.py:31-47 on the host: identity-like transform keeps coordinates, records
    have the COCO layout."""
    from hpdecode.coco import batch_to_coco, coco_records
    from hpdecode.transforms import affine_transform, get_affine_transform
    m = get_affine_transform((256, 256), (512.0, 512.0), 0, (512, 512), inverse=True)
    assert m.shape == (2, 3) and np.allclose(affine_transform([10.0, 20.0], m), [10.0, 20.0], atol=1e-4)
    m2 = get_affine_transform((320, 240), (640.0, 480.0), 0, (512, 384), inverse=True)      # 0.8x network input
    assert np.allclose(affine_transform([256.0, 192.0], m2), [320.0, 240.0], atol=1e-3)
    grouped = np.zeros((2, 17, 4), np.float32)
    grouped[..., 0] = 100.0
    grouped[..., 1] = 50.0
    recs = batch_to_coco([7], [(grouped, np.array([0.5, 0.25], np.float32))], [(256, 256)], [(512.0, 512.0)], (512, 512))
    assert len(recs) == 2 and recs[0]["image_id"] == 7 and recs[0]["category_id"] == 1
    assert len(recs[0]["keypoints"]) == 51 and recs[0]["keypoints"][2] == 1 and abs(recs[0]["keypoints"][0] - 100.0) < 1e-3
    assert recs[1]["score"] == 0.25
    assert coco_records(1, np.zeros((0, 17, 2)), np.zeros((0,))) == []


def test_division_by_three_as_two_fmas_is_the_ieee_quotient(tmp_path):
    """csrc/aggregate_nms_ms.cuh divides the 3-scale sum by 3 with two FMAs around RN(1/3); tools/check_div3.c
    compares that with x / 3.0f bit for bit (here: every mantissa and sign in four binades; --full = all 2^32)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    exe = str(tmp_path / "check_div3")
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_div3.c")
    subprocess.run([cc, "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert " 0 mismatches" in r.stdout
