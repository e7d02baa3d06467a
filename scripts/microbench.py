import ctypes as C, sys
sys.path.insert(0,'.')
from __graft_entry__ import load_package
pkg=load_package(); eng=pkg.Engine(0)
for kind in range(6):
    v=C.c_double(0); eng.check(pkg.lib.kzgb_microbench(eng.h, kind, C.byref(v))); print(kind, "%.4e"%v.value)
