"""Times the isolated SRS MSM (kzgb_bench_msm) -- total pipeline and the accumulate kernel alone."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 19
n = 1 << logn
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, 2480609854371098259468018140899271569021640719453669963486734696239309822386, engine=eng)
srs.precompute(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
tot, acc = C.c_double(0), C.c_double(0)
eng.check(pkg.lib.kzgb_bench_msm(eng.h, n, 10, C.byref(tot), C.byref(acc)))
print(f"n=2^{logn} msm_total={tot.value:.3f} ms accumulate={acc.value:.3f} ms "
      f"-> {10 * 16 * n / acc.value / 1e6:.1f} GFqmul/s")
