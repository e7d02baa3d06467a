"""A/B of the large-blob pipeline inside ONE process: for every setting ("B:name=value,name=value") prints blobs/s of R repetitions
(HBM-resident blobs and host blobs), interleaved so that drift shows.  usage: pipeline_ab.py R setting [setting ...]"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
from bench import TAU, make_blob
pkg = load_package(); lib = pkg.lib
R = int(sys.argv[1]); settings = sys.argv[2:]
n = 1 << 19
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, TAU, engine=eng); srs.precompute(n, 0)
Bmax = max(int(s.split(":")[0]) for s in settings)
host = [torch.from_numpy(make_blob(n, i)).pin_memory() for i in range(Bmax)]
dev = [h.to("cuda") for h in host]
touched = set()
def apply(spec, reset=False):
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        touched.add(k)
        assert lib.kzgb_set_option(k.encode(), -1 if reset and k in ("lane_wait", "hash_mb", "device_hash") else (0 if reset else int(v))) == 0, kv
res = {s: [] for s in settings}
for rep in range(R):
    for s in settings:
        B, _, spec = s.partition(":"); B = int(B)
        lens = (C.c_size_t * B)(*[n * 32] * B)
        hp = (C.c_void_p * B)(*[t.data_ptr() for t in host[:B]]); dp = (C.c_void_p * B)(*[t.data_ptr() for t in dev[:B]])
        cm, pf = C.create_string_buffer(32 * B), C.create_string_buffer(32 * B)
        apply(spec)
        out = []
        for fn in (lambda: lib.kzgb_commit_and_prove_blobs_dev(eng.h, dp, hp, lens, B, cm, pf), lambda: lib.kzgb_commit_and_prove_blobs(eng.h, hp, lens, B, cm, pf)):
            steps = max(2, 96 // B)
            eng.check(fn()); eng.check(fn())
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(steps): eng.check(fn())
            out.append(B * steps / (time.perf_counter() - t0))
        apply(spec, reset=True)
        res[s].append(out)
for s in settings:
    print(s, " resident:", " ".join(f"{a:.1f}" for a, _ in res[s]), " e2e:", " ".join(f"{b:.1f}" for _, b in res[s]), flush=True)
