"""Per-row measurement of SURVEY.md section 8(a): every function on the hot path through the C ABI on the GPU,
the CPU oracle (C++ restatement of the reference's algorithm) on the same inputs beside it, and the
results compared inline.  Writes one JSON line per row.

  python scripts/row_bench.py [--log-n 19] [--quick]

GPU times are wall-clock around the C-ABI call with HOST buffers (host<->device copies included), best of
`reps`; CPU times are one call of the oracle on all host threads (bounded samples where the reference's
algorithm would take minutes: stated per row).
"""
import argparse
import ctypes as C
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from bench import TAU, R_MOD, load_oracle_lib, make_blob  # noqa: E402

MONT = 1 << 256
P_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def best(fn, reps):
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append(time.perf_counter() - t0)
    return min(out)


def once(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=19)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="skip the slow CPU legs (a5/a6 at full size)")
    args = ap.parse_args()
    import numpy as np

    pkg = load_package()
    lib = pkg.lib
    eng = pkg.Engine(0)
    olib = load_oracle_lib()
    olib.ref_srs_decompress.restype = C.c_size_t
    th = olib.ref_hw_threads()
    logn = args.log_n
    n = 1 << logn
    rows = []

    def emit(row, what, gpu_s, cpu_s, parity, note=""):
        line = {"row": row, "what": what, "n": n, "gpu_ms": None if gpu_s is None else round(gpu_s * 1e3, 3),
                "cpu_ms": None if cpu_s is None else round(cpu_s * 1e3, 3), "cpu_threads": th, "parity": parity, "note": note}
        if gpu_s and cpu_s:
            line["cpu_over_gpu"] = round(cpu_s / gpu_s, 1)
        rows.append(line)
        print(json.dumps(line), flush=True)

    # ---- SRS: synthetic tau^i G on the GPU, read back for the CPU side ------------------------------------
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    t_pre = once(lambda: srs.precompute(n, 0))
    xy = C.create_string_buffer(64 * n)
    eng.check(lib.kzgb_srs_get_affine_mont(eng.h, 0, n, xy, None))
    srs_xy = xy.raw

    rnd = random.Random(7)
    seed = [rnd.randrange(R_MOD) for _ in range(512)]
    evals_mont = np.frombuffer(pkg.fr_to_mont_bytes(seed) * (n // 512), dtype=np.uint8).copy()
    for k in range(0, n, 4099):
        evals_mont[32 * k : 32 * k + 32] = np.frombuffer(pkg.fr_to_mont_bytes([rnd.randrange(R_MOD)]), dtype=np.uint8)
    evals_b = evals_mont.tobytes()
    out = C.create_string_buffer(64)
    inf = C.c_uint8(0)
    cout = C.create_string_buffer(64)

    # a3 commit_coeff_form = MSM over the monomial SRS
    eng.check(lib.kzgb_commit_coeff(eng.h, evals_b, n, out, C.byref(inf)))
    g = best(lambda: eng.check(lib.kzgb_commit_coeff(eng.h, evals_b, n, out, C.byref(inf))), args.reps)
    c = once(lambda: olib.ref_msm(srs_xy, evals_b, C.c_size_t(n), th, cout))
    emit("a3", "KZG::commit_coeff_form (G1 MSM over the SRS)", g, c, out.raw == cout.raw)

    # a10 Fr NTT (to_coeff_form / to_eval_form)
    buf = C.create_string_buffer(evals_b, 32 * n)
    g = best(lambda: eng.check(lib.kzgb_ntt_fr(eng.h, buf, n, 1)), args.reps)
    buf = C.create_string_buffer(evals_b, 32 * n)
    eng.check(lib.kzgb_ntt_fr(eng.h, buf, n, 1))
    cbuf = C.create_string_buffer(evals_b, 32 * n)
    c = once(lambda: olib.ref_ntt(cbuf, C.c_size_t(n), 1, th))
    ms = C.c_double(0)
    eng.check(lib.kzgb_bench_ntt(eng.h, logn, 1, 20, C.byref(ms)))
    emit("a10", "PolynomialEvalForm::to_coeff_form (Fr IFFT)", g, c, buf.raw == cbuf.raw,
         f"device-resident: {ms.value:.3f} ms per transform")

    # a9 roots of unity (the CPU leg is the reference's own algorithm: n serial multiplications, here Python big ints on one thread)
    rbuf = C.create_string_buffer(32 * n)
    rn = C.c_size_t(0)
    g = best(lambda: eng.check(lib.kzgb_roots_of_unity(eng.h, 32 * n, rbuf, n, C.byref(rn))), args.reps)
    from oracle import bn254 as o_

    t0 = time.perf_counter()
    roots_o = o_.calculate_roots_of_unity(32 * n)
    c9 = time.perf_counter() - t0
    sample = [0, 1, 2, n // 2, n - 1, 777 % n]
    par9 = all(pkg.fr_from_mont_bytes(rbuf.raw[32 * i : 32 * i + 32])[0] == roots_o[i] for i in sample)
    emit("a9", "helpers::calculate_roots_of_unity", g, c9, par9, "CPU = the reference's serial loop in Python big ints, 1 thread; parity on sampled indices")

    # a1 commit_eval_form
    eng.check(lib.kzgb_commit_eval(eng.h, evals_b, n, out, C.byref(inf)))
    g = best(lambda: eng.check(lib.kzgb_commit_eval(eng.h, evals_b, n, out, C.byref(inf))), args.reps)
    c = once(lambda: olib.ref_msm(srs_xy, cbuf.raw, C.c_size_t(n), th, cout)) + c
    emit("a1", "KZG::commit_eval_form (one MSM over the resident Lagrange table; CPU: Fr-IFFT + MSM)", g, c, out.raw == cout.raw,
         "CPU = oracle NTT + MSM; the reference's literal G1-IFFT form is ~200x more (bench.py cpu_baseline)")

    # a11 to_fr_array / a4 commit_blob / a8 challenge / a5 blob proof
    blob = make_blob(n, 99).tobytes()
    fr_gpu = C.create_string_buffer(32 * n)
    g = best(lambda: eng.check(lib.kzgb_to_fr_array(eng.h, blob, len(blob), fr_gpu)), args.reps)
    fr_cpu = C.create_string_buffer(32 * n)
    c = once(lambda: olib.ref_to_fr_array(blob, C.c_size_t(len(blob)), fr_cpu))
    emit("a11", "helpers::to_fr_array (blob bytes -> Fr)", g, c, fr_gpu.raw == fr_cpu.raw)

    cm = C.create_string_buffer(64)
    eng.check(lib.kzgb_commit_blob(eng.h, blob, len(blob), cm, C.byref(inf)))
    g = best(lambda: eng.check(lib.kzgb_commit_blob(eng.h, blob, len(blob), cm, C.byref(inf))), args.reps)
    ccm = C.create_string_buffer(64)
    c = once(lambda: olib.ref_commit_blob(blob, C.c_size_t(len(blob)), srs_xy, th, 0, ccm))
    emit("a4", "KZG::commit_blob", g, c, cm.raw == ccm.raw)

    z_gpu = C.create_string_buffer(32)
    g = best(lambda: eng.check(lib.kzgb_compute_challenge(eng.h, blob, len(blob), cm, 0, z_gpu)), args.reps)
    z_cpu = C.create_string_buffer(32)
    c = once(lambda: olib.ref_challenge(blob, C.c_size_t(len(blob)), cm, z_cpu))
    emit("a8", "helpers::compute_challenge (SHA-256 transcript, host side of the library)", g, c, z_gpu.raw == z_cpu.raw,
         "one sequential SHA-256 over 32n + 64 bytes on both sides")

    y_gpu = C.create_string_buffer(32)
    fr_b = fr_gpu.raw
    g = best(lambda: eng.check(lib.kzgb_evaluate_polynomial(eng.h, fr_b, n, z_gpu, y_gpu)), args.reps)
    n_s = n if not args.quick else min(n, 1 << 14)
    y_cpu = C.create_string_buffer(32)
    c = once(lambda: olib.ref_evaluate(fr_b, C.c_size_t(n_s), z_gpu, y_cpu))
    par = (y_gpu.raw == y_cpu.raw) if n_s == n else None
    emit("a7", "helpers::evaluate_polynomial_in_evaluation_form", g, c * (n / n_s), par,
         "CPU does n separate inversions as the reference does" + ("" if n_s == n else f"; timed at n = {n_s}, scaled linearly"))

    pf = C.create_string_buffer(64)
    eng.check(lib.kzgb_compute_blob_proof(eng.h, blob, len(blob), cm, 0, pf, C.byref(inf)))
    g = best(lambda: eng.check(lib.kzgb_compute_blob_proof(eng.h, blob, len(blob), cm, 0, pf, C.byref(inf))), args.reps)
    if args.quick:
        emit("a5", "KZG::compute_blob_proof", g, None, None, "CPU leg skipped (--quick)")
    else:
        cpf = C.create_string_buffer(64)
        c = once(lambda: olib.ref_blob_proof(blob, C.c_size_t(len(blob)), cm, srs_xy, th, 0, cpf))
        emit("a5", "KZG::compute_blob_proof", g, c, pf.raw == cpf.raw)

    # a6 compute_proof at a caller-supplied z, once outside and once INSIDE the domain (kzg.rs:237-260)
    w = pkg.get_primitive_root_of_unity(logn)
    for label, zval in (("z random", rnd.randrange(R_MOD)), ("z = w^5 (in the domain)", pow(w, 5, R_MOD))):
        zb = pkg.fr_to_mont_bytes([zval])
        yb = C.create_string_buffer(32)
        eng.check(lib.kzgb_compute_proof(eng.h, fr_b, n, zb, pf, C.byref(inf), yb))
        g = best(lambda: eng.check(lib.kzgb_compute_proof(eng.h, fr_b, n, zb, pf, C.byref(inf), yb)), args.reps)
        if args.quick:
            emit("a6", f"KZG::compute_proof, {label}", g, None, None, "CPU leg skipped (--quick)")
        else:
            cpf, cy = C.create_string_buffer(64), C.create_string_buffer(32)
            c = once(lambda: olib.ref_proof_at(fr_b, C.c_size_t(n), zb, srs_xy, th, cpf, cy))
            emit("a6", f"KZG::compute_proof, {label}", g, c, pf.raw == cpf.raw and yb.raw == cy.raw)

    # a2 g1_ifft (the reference's per-commit Lagrange SRS) at 2^12; GPU also at 2^16
    n12 = 1 << 12
    lag = C.create_string_buffer(64 * n12)
    linf = C.create_string_buffer(n12)
    eng.check(lib.kzgb_g1_ifft(eng.h, n12, lag, linf))
    g = best(lambda: eng.check(lib.kzgb_g1_ifft(eng.h, n12, lag, linf)), args.reps)
    clag = C.create_string_buffer(64 * n12)
    c = once(lambda: olib.ref_g1_ifft(srs_xy[: 64 * n12], C.c_size_t(n12), th, clag))
    line_n = n
    n = n12
    emit("a2", "KZG::g1_ifft (G1-point inverse NTT)", g, c, lag.raw == clag.raw)
    if logn >= 16:
        n16 = 1 << 16
        lag16, linf16 = C.create_string_buffer(64 * n16), C.create_string_buffer(n16)
        g16 = best(lambda: eng.check(lib.kzgb_g1_ifft(eng.h, n16, lag16, linf16)), 2)
        n = n16
        emit("a2", "KZG::g1_ifft (G1-point inverse NTT)", g16, c * (n16 * 16) / (n12 * 12), None, "CPU extrapolated by n log n from 2^12")
    n = line_n

    # a12 SRS ingest: gnark-BE compressed bytes -> affine points (decompression), then the window tables
    n_l = min(n, 1 << 17)
    rinv = pow(MONT, -1, P_MOD)
    half = (P_MOD - 1) // 2
    chunks = []
    for i in range(n_l):
        x = int.from_bytes(srs_xy[64 * i : 64 * i + 32], "little") * rinv % P_MOD
        y = int.from_bytes(srs_xy[64 * i + 32 : 64 * i + 64], "little") * rinv % P_MOD
        b = bytearray(x.to_bytes(32, "big"))
        b[0] |= 0xC0 if y > half else 0x80
        chunks.append(bytes(b))
    file_bytes = b"".join(chunks)
    eng2 = pkg.Engine(0)
    eng2.check(lib.kzgb_srs_load_gnark_be(eng2.h, file_bytes, n_l))
    g = best(lambda: eng2.check(lib.kzgb_srs_load_gnark_be(eng2.h, file_bytes, n_l)), args.reps)
    back = C.create_string_buffer(64 * n_l)
    eng2.check(lib.kzgb_srs_get_affine_mont(eng2.h, 0, n_l, back, None))
    cxy = C.create_string_buffer(64 * n_l)
    c = once(lambda: olib.ref_srs_decompress(file_bytes, C.c_size_t(n_l), th, cxy))
    n = n_l
    emit("a12", "SRS::new internals (read_g1_point_from_bytes_be x n)", g, c, back.raw == cxy.raw == srs_xy[: 64 * n_l],
         f"window-table precompute for 2^{logn} points (one-time, GPU only): {t_pre * 1e3:.1f} ms")
    eng2.close()
    n = line_n

    # a13 g1_lincomb (variable-base MSM), m = 4096; a16 validate_g1_point x 8192
    m = 4096
    n = m
    sc = pkg.fr_to_mont_bytes([rnd.randrange(R_MOD) for _ in range(m)])
    eng.check(lib.kzgb_msm_var(eng.h, srs_xy[: 64 * m], None, sc, m, out, C.byref(inf)))
    g = best(lambda: eng.check(lib.kzgb_msm_var(eng.h, srs_xy[: 64 * m], None, sc, m, out, C.byref(inf))), args.reps)
    c = once(lambda: olib.ref_msm(srs_xy[: 64 * m], sc, C.c_size_t(m), th, cout))
    emit("a13", "helpers::g1_lincomb (variable-base MSM)", g, c, out.raw == cout.raw)
    n = 2 * m
    g = best(lambda: eng.check(lib.kzgb_validate_g1_points(eng.h, srs_xy[: 64 * n], None, n)), args.reps)
    emit("a16", "helpers::validate_g1_point x n", g, None, True, "on-curve check; accepted all (negative cases in tests/)")

    # a14 + a15: verify_blob_kzg_proof_batch before the pairing, m pairs of 2^12-Fr blobs
    nb = 1 << 12
    mb = 4096 if not args.quick else 256
    eng3 = pkg.Engine(0)
    srs3 = pkg.SRS.synthetic(nb, TAU, engine=eng3)
    srs3.precompute(nb, 0)
    host = np.stack([make_blob(nb, 500 + i) for i in range(mb)])
    ptrs = (C.c_void_p * mb)(*[host[i].ctypes.data for i in range(mb)])
    lens = (C.c_size_t * mb)(*[nb * 32] * mb)
    cms, pfs = C.create_string_buffer(32 * mb), C.create_string_buffer(32 * mb)
    t_prove = once(lambda: eng3.check(lib.kzgb_commit_and_prove_blobs(eng3.h, ptrs, lens, mb, cms, pfs)))
    from oracle import bn254 as o

    cpts = [o.g1_deserialize_compressed(cms.raw[32 * i : 32 * i + 32]) for i in range(mb)]
    ppts = [o.g1_deserialize_compressed(pfs.raw[32 * i : 32 * i + 32]) for i in range(mb)]
    cxy_b, cinf_b = pkg.g1_to_abi(cpts)
    pxy_b, pinf_b = pkg.g1_to_abi(ppts)
    lhs, rhs = C.create_string_buffer(64), C.create_string_buffer(64)
    li, ri = C.c_uint8(0), C.c_uint8(0)

    def verify():
        eng3.check(lib.kzgb_verify_batch_rlc(eng3.h, ptrs, lens, mb, cxy_b, cinf_b, pxy_b, pinf_b, lhs, C.byref(li), rhs, C.byref(ri)))

    verify()
    g = best(verify, args.reps)
    lhs_pt = pkg.g1_from_abi(lhs.raw, bytes([li.value]))[0]
    rhs_pt = pkg.g1_from_abi(rhs.raw, bytes([ri.value]))[0]
    relation = lhs_pt is not None and o.g1_mul(lhs_pt, TAU) == rhs_pt
    # CPU: per-blob challenge + evaluation on a sample of 16 blobs (sequential in the reference, helpers.rs:635), scaled;
    # plus the three m-point MSMs of batch.rs:228,245,246
    xy3 = C.create_string_buffer(64 * nb)
    eng3.check(lib.kzgb_srs_get_affine_mont(eng3.h, 0, nb, xy3, None))
    sample = 16
    zc, yc = C.create_string_buffer(32), C.create_string_buffer(32)
    frs = C.create_string_buffer(32 * nb)

    def cpu_front():
        for i in range(sample):
            bb = host[i].tobytes()
            olib.ref_challenge(bb, C.c_size_t(len(bb)), cxy_b[64 * i : 64 * i + 64], zc)
            olib.ref_to_fr_array(bb, C.c_size_t(len(bb)), frs)
            olib.ref_evaluate(frs, C.c_size_t(nb), zc, yc)

    c_front = once(cpu_front) * mb / sample
    scm = pkg.fr_to_mont_bytes([rnd.randrange(R_MOD) for _ in range(mb)])
    c_msm = 3 * once(lambda: olib.ref_msm(pxy_b, scm, C.c_size_t(mb), 1, cout))
    n = mb
    emit("a14+a15", "verify_blob_kzg_proof_batch up to the pairing (m pairs of 2^12-Fr blobs)", g, c_front + c_msm, relation,
         f"parity = pairing relation rhs == tau * lhs on the synthetic SRS; CPU = {sample} blobs of challenge+evaluation "
         f"(single-threaded like the reference) scaled to m, + 3 single-threaded m-point MSMs; GPU prover made the {mb} proofs in {t_prove * 1e3:.0f} ms")
    print(json.dumps({"summary": "rows", "count": len(rows), "all_parity_ok": all(r["parity"] in (True, None) for r in rows)}))


if __name__ == "__main__":
    main()
