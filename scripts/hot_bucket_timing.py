"""Worst-case digit distributions through kzgb_msm_srs at n = 2^19 (wall time per call incl. the 16 MiB H2D)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
n = 1 << 19
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, 2480609854371098259468018140899271569021640719453669963486734696239309822386, engine=eng)
srs.precompute(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
import random
rnd = random.Random(1)
cases = {
    "random": [rnd.randrange(R) for _ in range(256)] * (n // 256),
    "all equal": [0x2F0E1D3C4B5A69788796A5B4C3D2E1F00112233445566778899AABBCCDDEEFF % R] * n,
    "all r-1": [R - 1] * n,
    "two values": [5, R - 7] * (n // 2),
}
out = C.create_string_buffer(64); inf = C.c_uint8(0)
for name, sc in cases.items():
    buf = pkg.fr_to_mont_bytes(sc[:256]) * (n // 256) if name == "random" else pkg.fr_to_mont_bytes(sc[:2]) * (n // 2)
    eng.check(pkg.lib.kzgb_msm_srs(eng.h, buf, n, out, C.byref(inf)))
    t0 = time.perf_counter()
    for _ in range(5):
        eng.check(pkg.lib.kzgb_msm_srs(eng.h, buf, n, out, C.byref(inf)))
    print(f"{name:12s} {(time.perf_counter() - t0) / 5 * 1e3:8.2f} ms per MSM")
