#!/usr/bin/env python3
"""Accumulate-kernel variants through the whole MSM, fast (ctypes only, a few seconds per variant):
    python scripts/variant_check.py [variant ...]          (default: 0 28 34 31 33 35 32 29 30 23 25)
Each variant runs in its own process (KZGB_ACC_VARIANT is read once per process): synthetic SRS tau^i G of 2^16 points,
commit_coeff of pseudo-random scalars compared with the closed form (sum s_i tau^i) G from the CPU oracle, an
all-equal-scalars vector (hot buckets spanning many chunks), then kzgb_bench_msm (total / accumulate milliseconds).
One JSON line per variant."""
import ctypes as C
import json
import os
import random
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAU = 2480609854371098259468018140899271569021640719453669963486734696239309822386
LOGN = 16


def child():
    from __graft_entry__ import load_package
    from oracle import bn254 as o

    t0 = time.time()
    pkg = load_package()
    n = 1 << LOGN
    srs = pkg.SRS.synthetic(n, TAU)
    eng = srs.engine
    kzg = pkg.KZG()
    rnd = random.Random(29)
    out = {"variant": int(os.environ.get("KZGB_ACC_VARIANT", "0")), "waves": os.environ.get("KZGB_ACC_WAVES", "4")}
    ok = True
    for name, sc in (("random", [rnd.randrange(o.R) for _ in range(n)]), ("all_equal", [0x1234567890ABCDEF1234567890ABCDEF] * n),
                     ("sparse", [(i % 97 == 0) * rnd.randrange(o.R) for i in range(n)])):
        got = kzg.commit_coeff_form(pkg.PolynomialCoeffForm(sc), srs)
        acc = 0
        for s in reversed(sc):
            acc = (acc * TAU + s) % o.R
        want = o.g1_mul(o.G1_GEN, acc)
        out[name] = got == want
        ok = ok and out[name]
    tot, accm = C.c_double(0), C.c_double(0)
    eng.check(pkg.lib.kzgb_bench_msm(eng.h, n, 10, C.byref(tot), C.byref(accm)))
    out.update({"exact": ok, "msm_ms": round(tot.value, 4), "accumulate_ms": round(accm.value, 4), "wall_s": round(time.time() - t0, 2)})
    print(json.dumps(out), flush=True)


def main():
    if os.environ.get("KZGB_VARIANT_CHILD"):
        return child()
    variants = sys.argv[1:] or ["0", "28", "34", "31", "33", "35", "32", "29", "30", "23", "25"]
    for v in variants:
        env = dict(os.environ, KZGB_ACC_VARIANT=v, KZGB_VARIANT_CHILD="1")
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or json.dumps({"variant": int(v), "error": (r.stderr or "")[-400:]}), flush=True)


if __name__ == "__main__":
    main()
