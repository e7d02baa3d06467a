"""Latency of device-side transcript hashing (k_fs_midstate_long) for B blobs of 2^19 Fr, alone on the GPU and next to the
MSM pipeline: runs the headline batch with device_hash = 0 / k and prints blobs/s.
usage: device_hash_bench.py B [name=value ...] k1[:lanes] k2[:lanes] ...   (name=value: kzgb_set_option; k:0 = the warp-per-transcript kernel)"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
from bench import TAU, make_blob
pkg = load_package(); lib = pkg.lib
B = int(sys.argv[1])
for kv in [a for a in sys.argv[2:] if "=" in a]:
    assert lib.kzgb_set_option(kv.split("=")[0].encode(), int(kv.split("=")[1])) == 0, kv
ks = [a for a in sys.argv[2:] if "=" not in a] or ["0"]
n = 1 << 19
eng = pkg.Engine(0)
srs = pkg.SRS.synthetic(n, TAU, engine=eng); srs.precompute(n, 0)
host = [torch.from_numpy(make_blob(n, i)).pin_memory() for i in range(B)]
dev = [h.to("cuda") for h in host]
lens = (C.c_size_t * B)(*[n * 32] * B)
hp = (C.c_void_p * B)(*[t.data_ptr() for t in host]); dp = (C.c_void_p * B)(*[t.data_ptr() for t in dev])
cm, pf = C.create_string_buffer(32 * B), C.create_string_buffer(32 * B)
ref = None
for tok in ks:
    k, _, lanes = tok.partition(":")
    k = int(k)
    lib.kzgb_set_option(b"device_hash", k)
    lib.kzgb_set_option(b"device_hash_lanes", int(lanes or 1))
    for path, fn in (("resident", lambda: lib.kzgb_commit_and_prove_blobs_dev(eng.h, dp, hp, lens, B, cm, pf)),
                     ("e2e", lambda: lib.kzgb_commit_and_prove_blobs(eng.h, hp, lens, B, cm, pf))):
        eng.check(fn())
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): eng.check(fn())
        dt = (time.perf_counter() - t0) / 3
        if ref is None: ref = (cm.raw, pf.raw)
        print(f"B={B} device_hash={tok} {path}: {dt*1e3:.1f} ms/step = {B/dt:.1f} blobs/s  same={ (cm.raw, pf.raw) == ref }", flush=True)
