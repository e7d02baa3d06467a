"""world_size-2 gloo tests (CPU) of the multi-process host logic: shard maps, the one exchange step
of the point-range-sharded MSM (all_gather of 65-byte partials + local G1 additions in the library)."""
import os
import random
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        import golden_data as g
        from oracle import bn254 as o

        pkg = load_package()
        sh = __import__("rust_kzg_bn254_b200.sharding", fromlist=["x"])
        pts = g.srs_points_string()[:101]
        rnd = random.Random(99)
        sc = [rnd.randrange(o.R) for _ in range(101)]
        first, count = sh.shard_range(101, rank, world)
        part = o.msm(pts[first : first + count], sc[first : first + count])  # stands in for the GPU partial
        xy, inf = pkg.g1_to_abi([part])
        total = sh.reduce_g1_partials(pkg, sh.all_gather_bytes(xy + inf))
        ok = total == o.msm(pts, sc)
        # identity partial on one rank must be handled
        part2 = None if rank == 0 else pts[3]
        xy, inf = pkg.g1_to_abi([part2])
        ok = ok and sh.reduce_g1_partials(pkg, sh.all_gather_bytes(xy + inf)) == pts[3]
        q.put((rank, ok, first, count))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package

    load_package()
    sh = __import__("rust_kzg_bn254_b200.sharding", fromlist=["x"])
    for total in (0, 1, 7, 1024, (1 << 26) + 5):
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                first, count = sh.shard_range(total, r, world)
                assert first == nxt and count >= 0
                nxt += count
            assert nxt == total


def test_point_range_sharded_msm_exchange_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randrange(2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert sorted((f, c) for _, _, f, c in res) == [(0, 51), (51, 50)]
