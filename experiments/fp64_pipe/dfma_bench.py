#!/usr/bin/env python3
"""FP64-pipe Fq arithmetic on the device: self-test + rates (csrc/dfma.cu).  ctypes only (no torch import), so the
whole run is a few seconds:   python scripts/dfma_bench.py [out.json] [--probe-only]

  pipe_mix_ms           do DFMA / DADD / IMAD.WIDE share a pipe?  kernel times with the two warp halves of every
                        scheduler running different instruction kinds (kzgb_pipe_mix_probe)

  dfma_per_s            independent DFMA chains (B200 nominal: 148 SM x 64 /clk)
  fqmul_int_per_s       the integer (IMAD.WIDE) Montgomery multiplication on every warp
  fqmul_dfma_per_s      the DFMA multiplication on every warp
  hybrid[...]           half of the warps of every scheduler integer, half DFMA, for several work splits:
                        total multiplications/s of both kinds together
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    t0 = time.time()
    lib = C.CDLL(os.path.join(ROOT, "rust-kzg-bn254_b200", "libkzgbn254_b200.so"))
    lib.kzgb_dfma_microbench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.kzgb_dfma_selftest.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    out = {}

    def emit():
        out["wall_s"] = round(time.time() - t0, 2)
        line = json.dumps(out)
        print(line, flush=True)
        if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
            with open(sys.argv[1], "w") as f:
                f.write(line + "\n")

    lib.kzgb_pipe_mix_probe.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    names = ["wide|idle", "dfma|idle", "wide|dfma", "wide|wide", "dfma|dfma", "dadd|idle", "dadd|dfma", "dadd|wide"]
    out["pipe_mix_ms"] = {}
    for mix, nm in enumerate(names):
        v = C.c_double(0)
        rc = lib.kzgb_pipe_mix_probe(0, mix, 2048, C.byref(v))
        out["pipe_mix_ms"][nm] = round(v.value, 4) if rc == 0 else f"rc={rc}"
    out["pipe_mix_note"] = "2048 x 128 operations per active thread, 148 x 8 blocks of 256; the two halves share every scheduler"
    emit()
    if "--probe-only" in sys.argv:
        return

    bad = C.c_uint32(0xFFFFFFFF)
    rc = lib.kzgb_dfma_selftest(0, 148 * 128 * 4, 12, C.byref(bad))
    out["selftest"] = {"rc": rc, "threads": 148 * 128 * 4, "chain": 12, "mismatches": bad.value}
    emit()

    def rate(kind, ii, idf):
        v = C.c_double(0)
        rc = lib.kzgb_dfma_microbench(0, kind, ii, idf, C.byref(v))
        return v.value if rc == 0 else f"rc={rc}"

    out["dfma_per_s"] = rate(0, 0, 1024)
    emit()
    out["fqmul_int_per_s"] = rate(3, 256, 0)
    out["fqmul_dfma_per_s"] = rate(1, 0, 256)
    emit()
    out["hybrid"] = {}
    for ii, idf in ((256, 256), (256, 192), (256, 128), (256, 320), (256, 64)):
        out["hybrid"][f"int{ii}_dfma{idf}"] = rate(2, ii, idf)
        emit()


if __name__ == "__main__":
    main()
