"""Fr NTT timings (kzgb_bench_ntt): batched 1024 x 2^16 and single 2^19 / 2^16 / 2^22 transforms."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package(); lib = pkg.lib
eng = pkg.Engine(0)
for logn, batch, reps in ((16, 1024, 6), (19, 1, 40), (16, 1, 40), (22, 1, 10), (12, 4096, 10), (20, 64, 4)):
    ms = C.c_double(0)
    eng.check(lib.kzgb_bench_ntt(eng.h, logn, batch, reps, C.byref(ms)))
    n = batch << logn
    muls = (logn - 1) * n / 2 + n / 2  # alternating forward / inverse: + n for every inverse
    print(f"2^{logn} x {batch}: {ms.value:.4f} ms  {64.0*n/ms.value/1e6:.1f} GB/s algorithmic  {muls/ms.value/1e6:.2f} G Fr-mul/s")
