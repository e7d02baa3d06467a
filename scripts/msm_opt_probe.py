"""Isolated 2^logn SRS MSM (kzgb_bench_msm) under kzgb_set_option settings.  usage: msm_opt_probe.py logn [name=value ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 19
for kv in sys.argv[2:]:
    assert pkg.lib.kzgb_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])) == 0, kv
n = 1 << logn
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, 2480609854371098259468018140899271569021640719453669963486734696239309822386, engine=eng)
srs.precompute(n, 0)
tot, acc = C.c_double(0), C.c_double(0)
eng.check(pkg.lib.kzgb_bench_msm(eng.h, n, 10, C.byref(tot), C.byref(acc)))
print(f"n=2^{logn} {' '.join(sys.argv[2:])} msm_total={tot.value:.3f} ms accumulate={acc.value:.3f} ms")
