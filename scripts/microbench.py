"""Integer-pipe microbenchmarks: 0 IMAD, 1 IMAD.WIDE, 2 carry-chained IMAD.WIDE.X, 3 Fq mul (CIOS)."""
import ctypes as C, sys
sys.path.insert(0, '.')
from __graft_entry__ import load_package
pkg = load_package(); eng = pkg.Engine(0)
for kind in range(4):
    v = C.c_double(0); eng.check(pkg.lib.kzgb_microbench(eng.h, kind, C.byref(v))); print(kind, "%.4e" % v.value)
