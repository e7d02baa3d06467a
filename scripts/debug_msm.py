"""Step-by-step exercise of the MSM path with a progress log (gpurun_out/debug_msm.log) -- run under `timeout`."""
import faulthandler, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/debug_msm.log", "w")
def say(*a):
    print(*a, file=LOG, flush=True); print(*a, flush=True)
faulthandler.dump_traceback_later(40, repeat=True, file=LOG)
from __graft_entry__ import load_package
from oracle import bn254 as o
pkg = load_package()
say("loaded")
TAU = o.SYNTH_TAU
eng = pkg.Engine(0)
kzg = pkg.KZG()
def closed(v):
    acc = 0
    for s in reversed(v):
        acc = (acc * TAU + s) % o.R
    return o.g1_mul(o.G1_GEN, acc)
rnd = random.Random(3)
for logn in (2, 6, 10, 12, 16):
    n = 1 << logn
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    say("srs", n)
    sc = [rnd.randrange(o.R) for _ in range(n)]
    t0 = time.time()
    got = kzg.commit_coeff_form(pkg.PolynomialCoeffForm(sc), srs)
    say("commit_coeff", n, got == closed(sc), round(time.time() - t0, 3))
    pts = srs.points(0, min(n, 300))
    got = pkg.g1_lincomb(pts, sc[: len(pts)], eng)
    say("lincomb", len(pts), got == o.msm(pts, sc[: len(pts)]) if len(pts) <= 64 else "skipped-compare", round(time.time() - t0, 3))
say("done")
