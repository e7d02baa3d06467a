"""Timeline of one kzgb_commit_and_prove_blobs_dev step (16 x 2^19-Fr blobs) from kzgb_trace_begin/end:
per MSM the device times of sort begin / accumulate begin / accumulate end / MSM end and the host hand-off times.
usage: python scripts/trace_pipeline.py [blobs] [NAME=VALUE options ...]  -> gpurun_out/trace_pipeline.txt"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
from bench import TAU, make_blob
pkg = load_package(); lib = pkg.lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for kv in sys.argv[2:]:
    k, _, v = kv.partition("="); assert lib.kzgb_set_option(k.encode(), int(v)) == 0
n = 1 << 19
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, TAU, engine=eng); srs.precompute(n, 0)
host = [torch.from_numpy(make_blob(n, i)).pin_memory() for i in range(B)]
dev = [h.to("cuda") for h in host]
lens = (C.c_size_t * B)(*[n * 32] * B)
hp = (C.c_void_p * B)(*[t.data_ptr() for t in host]); dp = (C.c_void_p * B)(*[t.data_ptr() for t in dev])
cm, pf = C.create_string_buffer(32 * B), C.create_string_buffer(32 * B)
for _ in range(3):
    eng.check(lib.kzgb_commit_and_prove_blobs_dev(eng.h, dp, hp, lens, B, cm, pf))
eng.check(lib.kzgb_trace_begin(eng.h))
eng.check(lib.kzgb_commit_and_prove_blobs_dev(eng.h, dp, hp, lens, B, cm, pf))
cap = 4096; buf = (C.c_double * (4 * cap))(); nrec = C.c_size_t(0)
eng.check(lib.kzgb_trace_end(eng.h, buf, cap, C.byref(nrec)))
recs = [(int(buf[4*i]), int(buf[4*i+1]), buf[4*i+2], buf[4*i+3]) for i in range(nrec.value)]
# group per lane in order: 0,1,2,3,10 then later 11,12
out = open("gpurun_out/trace_pipeline.txt", "w")
def say(*a):
    print(*a); print(*a, file=out)
lanes = sorted({r[0] for r in recs})
msms = []
for ln in lanes:
    cur = None
    for r in [x for x in recs if x[0] == ln]:
        if r[1] == 0:
            cur = {"lane": ln, "sort": r[2], "h_enq0": r[3]}; msms.append(cur)
        elif cur is not None:
            if r[1] == 1: cur["acc0"] = r[2]
            elif r[1] == 2: cur["acc1"] = r[2]
            elif r[1] == 3: cur["end"] = r[2]
            elif r[1] == 10: cur["h_enq"] = r[3]
            elif r[1] == 11: cur["h_wake"] = r[3]
            elif r[1] == 12: cur["h_done"] = r[3]
msms.sort(key=lambda m: m["sort"])
say("lane  sort_begin  acc_begin  acc_end   msm_end | sort_ms  acc_ms  tail_ms | host: enq  wake  done  (ms since begin)")
for m in msms:
    say(f"{m['lane']:4d} {m['sort']:10.3f} {m['acc0']:10.3f} {m['acc1']:9.3f} {m['end']:9.3f} | {m['acc0']-m['sort']:7.3f} {m['acc1']-m['acc0']:7.3f} {m['end']-m['acc1']:7.3f} |"
        f" {m.get('h_enq',-1):8.3f} {m.get('h_wake',-1):8.3f} {m.get('h_done',-1):8.3f}")
T = max(m["end"] for m in msms)
# coverage of [0, T] by accumulate intervals
ev = sorted([(m["acc0"], 1) for m in msms] + [(m["acc1"], -1) for m in msms])
cov = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0, 4: 0.0}; last = 0.0; depth = 0
for t, d in ev:
    cov[min(depth, 4)] += t - last; last = t; depth += d
cov[0] += T - last
say(f"step {T:.3f} ms for {B} blobs = {T/B:.3f} ms/blob; time with k accumulate kernels in flight: " + ", ".join(f"{k}: {v:.2f} ms" for k, v in cov.items()))
say("mean sort %.3f  acc %.3f  tail %.3f  host gap (msm_end -> next sort_begin on the lane) %.3f" % (
    np.mean([m['acc0']-m['sort'] for m in msms]), np.mean([m['acc1']-m['acc0'] for m in msms]), np.mean([m['end']-m['acc1'] for m in msms]),
    np.mean([b['sort']-a['end'] for ln in lanes for a, b in zip([m for m in msms if m['lane']==ln][:-1], [m for m in msms if m['lane']==ln][1:])])))
