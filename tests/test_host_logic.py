"""CPU suite: host-side logic -- synthetic generators, result packing, sharding and the
world_size-2 gather over gloo (the N>1 path of bench.py without GPUs)."""
import os

import numpy as np
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


def test_records_view_and_fallback_dtype():
    """Records: the host view of the device-written result records (HpdRecordLayout)."""
    from hpdecode import ops
    from hpdecode.decoder import Records
    M, K, E = 4, 17, 2
    D = 3 + E
    L = ops.record_layout(K, M, E)
    assert L.coco_stride == 3 * K + 1 and L.row_bytes % 8 == 0 and L.off_coco == 0
    assert L.off_poses == 8 * M * (3 * K + 1) and L.off_person_scores == L.off_poses + 4 * M * K * D
    raw = np.zeros((2, L.row_bytes), np.uint8)
    rec = Records(raw, M, K, E)
    rec.poses[1, :2] = np.arange(2 * K * D, dtype=np.float32).reshape(2, K, D)
    rec.person_scores[1, :2] = [0.5, 0.25]
    rec.n_person[1] = 2
    rec.coco[1, 0, :3] = [10.5, 20.25, 1.0]
    rec.coco[1, 0, 3 * K] = 0.5
    again = Records(raw.copy(), M, K, E)                       # the views write through to the raw bytes
    poses, scores = again.image(1)
    assert poses.shape == (2, K, D) and poses.dtype == np.float32 and list(scores) == [0.5, 0.25]
    assert poses[1, 3, 2] == float(K * D + 3 * D + 2)
    assert again.final_coords(1).dtype == np.float32 and again.final_coords(1)[0, 0].tolist() == [10.5, 20.25]
    recs = again.coco_records(1, 42)
    assert len(recs) == 2 and recs[0]["image_id"] == 42 and recs[0]["category_id"] == 1 and recs[0]["score"] == 0.5
    assert len(recs[0]["keypoints"]) == 3 * K and recs[0]["keypoints"][:3] == [10.5, 20.25, 1.0]
    assert again.image(0)[0].shape == (0, K, D) and DecodeResult.unpack(raw, M, K, E)[1][0].shape == (2, K, D)
    fb = np.zeros((1, K, D), np.float32)
    fb[..., 2] = np.float32(0.01)
    p, s_ = _finish(fb, np.zeros(1, np.float32), 1)
    assert p.dtype == np.float64 and p[0, 0, 2] == 0.01 and s_[0] == np.full((1, K), 0.01).mean(1)[0]


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


def _worker_rows(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hpdecode.parallel import gather_rows
    block = (torch.arange(40, dtype=torch.int32) % 251 + 7 * rank).to(torch.uint8)     # one rank's records, flattened
    out = torch.zeros((world, 40), dtype=torch.uint8) if rank == 0 else None
    got = gather_rows(block, out, dst=0)
    if rank == 0:
        assert got is out
        q.put(out.numpy())
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_world2_gloo():
    """bench.py's gather of the result records: rank r's flat uint8 block lands in row r on rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_rows, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    base = (np.arange(40) % 251).astype(np.uint8)
    assert np.array_equal(got, np.stack([base, base + 7]))


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


def test_result_to_coco_layout():
    """bin/eval.py:31-47 from a result object (the reference's per-image path)."""
    from types import SimpleNamespace
    from hpdecode.coco import image_id_of, result_to_coco
    coords = np.zeros((2, 17, 2), np.float32)
    coords[..., 0], coords[..., 1] = 100.5, 50.25
    recs = result_to_coco(7, SimpleNamespace(kpts_coords=coords, obj_scores=np.array([0.5, 0.25], np.float32)))
    assert len(recs) == 2 and recs[0]["image_id"] == 7 and recs[0]["category_id"] == 1
    assert len(recs[0]["keypoints"]) == 51 and recs[0]["keypoints"][:3] == [100.5, 50.25, 1.0] and recs[1]["score"] == 0.25
    assert image_id_of("/data/coco/images/val2017/000000000139.jpg") == 139


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
