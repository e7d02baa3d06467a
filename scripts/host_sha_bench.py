"""Aggregate host SHA-256 throughput (hashlib / OpenSSL, releases the GIL) vs thread count -- the bound of the
8-GPU blob-batch configuration (every 16 MiB blob is hashed once on the host for its Fiat-Shamir transcript)."""
import hashlib, os, threading, time
buf = os.urandom(16 << 20)
for nt in (1, 4, 8, 16, 32, 64):
    reps = 6
    def work():
        for _ in range(reps):
            hashlib.sha256(buf).digest()
    th = [threading.Thread(target=work) for _ in range(nt)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print(f"threads={nt:3d}  {nt * reps * len(buf) / dt / 1e9:6.2f} GB/s aggregate  ({dt / reps * 1e3:.1f} ms per 16 MiB per thread)")
