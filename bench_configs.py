"""Secondary bench configurations of BASELINE.json (the headline, configs[1], lives in bench.py):

  c3  batch of 1024 blobs x 2^16 Fr, commit + proof, sharded by blob across the ranks (strong scaling)
  c4  standalone 2^26-point G1 MSM sharded by point range, partial sums exchanged with one all_gather
  c5  verify_blob_kzg_proof_batch front half + RLC over 4096 (blob, commitment, proof) triples

Each prints one JSON line in bench.py's schema.  Parity of the timed configuration is checked in the
cpu_baseline leg against closed forms of the synthetic SRS (SRS_i = tau^i G):
  c4: scalars s_i = a^i  =>  MSM = ((a tau)^N - 1)/(a tau - 1) * G
  c5: the two outputs of the RLC satisfy  rhs = tau * lhs  (what the pairing check would verify).
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

from bench import TAU, R_MOD, ClockSampler, load_oracle_lib, make_blob

ROOT = os.path.dirname(os.path.abspath(__file__))
MONT = 1 << 256


def _dist_setup():
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    return torch, dist, world, rank, local_rank


def _timed(torch, dist, world, eng, lib, fn, steps):
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ms = C.c_double(0)
    barrier()
    w0 = time.perf_counter()
    eng.check(lib.kzgb_timer_begin(eng.h))
    for _ in range(steps):
        fn()
    eng.check(lib.kzgb_timer_end(eng.h, C.byref(ms)))
    wall = (time.perf_counter() - w0) * 1e3
    barrier()
    t = torch.tensor([ms.value, wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the timed region contains host work (SHA-256, exchange); the span that matters is the larger one
    return max(float(t[0]), float(t[1]))


def _fr_mont(v: int) -> bytes:
    return (v % R_MOD * MONT % R_MOD).to_bytes(32, "little")


def measure_c4(args, torch, dist, world, rank, local_rank, pkg, check_with_oracle=True):
    """BASELINE configs[3]: one 2^26-point fixed-base MSM cut by point range over the ranks (rank g holds SRS points
    [g N/G, (g+1) N/G) and their window table), the G partial sums exchanged with ONE all_gather (65 B per rank) and
    added on every rank.  Returns the JSON line as a dict on rank 0 (None elsewhere).  Shape of the reference's own
    large-commit bench: prover/benches/bench_kzg_commit_large_blobs.rs:17-37 (commit_coeff_form = one MSM, kzg.rs:107-125)."""
    sh = __import__("rust_kzg_bn254_b200.sharding", fromlist=["x"])
    lib = pkg.lib
    eng = pkg.Engine(local_rank)
    logn = args.log_n or 26
    N = 1 << logn
    first, count = sh.shard_range(N, rank, world)
    a = 0x1D5F3C29A7B4E6081122334455667788990AABBCCDDEEFF0123456789ABCDEF1 % R_MOD  # scalar base
    t0 = time.perf_counter()
    srs = pkg.SRS.synthetic(count, TAU, engine=eng, first=first)
    # fixed-base window tables over this rank's point range (W x count x 64 B: 48 GiB for 2^26 points on one GPU,
    # 6.5 GiB per GPU at 8); a failed allocation -> variable-base mode (one bucket set per window)
    tables = not getattr(args, "no_tables", False)
    if tables:
        try:
            srs.precompute(count, getattr(args, "window_bits", 0))
        except pkg.KzgError:
            tables = False
    if not tables:
        srs.precompute(0, -1)
    cbits, cwin, ctab = C.c_int(0), C.c_int(0), C.c_size_t(0)
    lib.kzgb_msm_config(eng.h, C.byref(cbits), C.byref(cwin), C.byref(ctab))
    scal = torch.empty(count * 32, dtype=torch.uint8, device="cuda")
    eng.check(lib.kzgb_fr_powers_dev(eng.h, _fr_mont(a), first, count, scal.data_ptr()))
    host_scal = scal.cpu().pin_memory()
    setup_s = time.perf_counter() - t0
    result = {}

    def step_dev():
        result["pt"] = sh.msm_srs_point_range_sharded(pkg, eng, scal.data_ptr(), first, count)

    def step_e2e():
        out = C.create_string_buffer(64)
        inf = C.c_uint8(0)
        eng.check(lib.kzgb_msm_srs_range(eng.h, host_scal.data_ptr(), 0, count, out, C.byref(inf)))
        result["pt_e2e"] = sh.reduce_g1_partials(pkg, sh.all_gather_bytes(out.raw + bytes([inf.value])))

    for _ in range(args.warmup):
        step_dev()
    lib.kzgb_stats(eng.h, None, None, None, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    ms = _timed(torch, dist, world, eng, lib, step_dev, args.steps)
    launches = eng.launch_count() - l0
    acc_ms, acc_n, acc_adds = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    lib.kzgb_stats(eng.h, C.byref(acc_ms), C.byref(acc_n), C.byref(acc_adds), 1)
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms = _timed(torch, dist, world, eng, lib, step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    line = None
    if rank == 0:
        imad = C.c_double(0)
        eng.check(lib.kzgb_microbench(eng.h, 0, C.byref(imad)))
        peak = imad.value / 136 / 1e9
        adds = acc_adds.value / max(1, acc_n.value)
        kms = acc_ms.value / max(1, acc_n.value)
        achieved = 10 * adds / (kms * 1e-3) / 1e9 if kms else 0.0
        # closed form of the synthetic inputs (SRS_i = tau^i G, s_i = a^i): MSM = ((a tau)^N - 1)/(a tau - 1) G.
        # Checked on every run through the library's own single-point multiplication, and against the CPU oracle's
        # g1_mul (checker only) unless told otherwise.
        at = a * TAU % R_MOD
        S = (pow(at, N, R_MOD) - 1) * pow(at - 1, -1, R_MOD) % R_MOD
        gen = (1, 2)
        own = pkg.g1_lincomb([gen], [S], eng)
        checks = {"closed_form_via_library_scalar_mul": result["pt"] == own and result["pt_e2e"] == own}
        if check_with_oracle:
            from oracle import bn254 as o

            expect = o.g1_mul(o.G1_GEN, S)
            checks["closed_form_via_cpu_oracle"] = result["pt"] == expect and result["pt_e2e"] == expect
        cpu_baseline = None
        if not args.skip_cpu_baseline:
            from oracle import bn254 as o

            olib = load_oracle_lib()
            threads = olib.ref_hw_threads()
            sample_n = 1 << 20
            small = o.synthetic_srs(4096)
            bases = b"".join((p[0] * MONT % o.P).to_bytes(32, "little") + (p[1] * MONT % o.P).to_bytes(32, "little") for p in small) * (sample_n // 4096)
            sc = bytes(host_scal[: sample_n * 32].numpy().tobytes())
            out = C.create_string_buffer(64)
            t1 = time.perf_counter()
            olib.ref_msm(bases, sc, C.c_size_t(sample_n), threads, out)
            dt = time.perf_counter() - t1
            cpu_baseline = {"value": sample_n / dt / 1e6, "unit": "Mpts/s", "cores": threads, "kind": "port",
                            "sample": "2^20-point MSM (arkworks window rule c=15, threads over windows), C++ restatement"}
        line = {
            "metric": "G1 MSM Mpts/s (2^%d points, sharded by point range)" % logn, "value": N * args.steps / (ms * 1e-3) / 1e6,
            "unit": "Mpts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 (8x32-bit Montgomery limbs, BN254 Fq/Fr)", "data": "synthetic",
            "config": {"workload": "standalone 2^%d-point G1 MSM over SRS_i = tau^i G, scalars a^i, sharded by point range; "
                                   "partial sums exchanged with one all_gather (65 B/rank) and added on every rank" % logn,
                       "points_per_rank": count, "fixed_base_tables": bool(ctab.value), "msm_window_bits": cbits.value,
                       "msm_windows": cwin.value, "table_GiB_per_rank": cwin.value * ctab.value * 64 / 2**30, "l2": "inputs larger than L2: %d MiB of points + %d MiB of scalars per rank" % (count * 64 >> 20, count * 32 >> 20)},
            "e2e": {"value": N * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "Mpts/s", "h2d_bytes_per_step": count * 32,
                    "d2h_bytes_per_step": 128 * 16, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(),
            "roofline": {"kernel": "k_accumulate (MSM bucket accumulation, XYZZ += affine)", "bound": "integer-pipe (IMAD)", "achieved": achieved, "peak": peak,
                         "unit": "GFqmul/s", "frac": achieved / peak if peak else None, "traffic": None,
                         "launch_ms": kms, "point_adds_per_launch": adds,
                         "frac_of_step": (10 * adds / (ms / args.steps * 1e-3) / 1e9) / peak if peak and ms else None,
                         "peak_source": "measured live: dependency-free IMAD chains / 136 (SURVEY.md 8d model)"},
            "checks": checks, "cpu_baseline": cpu_baseline, "setup_s": setup_s,
            "result_gnark_be": pkg.g1_to_gnark_be(result["pt"]).hex(),
        }
    del scal, host_scal, srs
    eng.close()
    torch.cuda.empty_cache()
    return line


def run_c4(args):
    torch, dist, world, rank, local_rank = _dist_setup()
    from __graft_entry__ import load_package

    pkg = load_package()
    line = measure_c4(args, torch, dist, world, rank, local_rank, pkg)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c3(args):
    torch, dist, world, rank, local_rank = _dist_setup()
    from __graft_entry__ import load_package

    pkg = load_package()
    sh = __import__("rust_kzg_bn254_b200.sharding", fromlist=["x"])
    lib = pkg.lib
    eng = pkg.Engine(local_rank)
    logn = args.log_n or 16
    n = 1 << logn
    total = 1024
    first, count = sh.shard_range(total, rank, world)
    t0 = time.perf_counter()
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    srs.precompute(n, 0)
    host = torch.empty((count, n * 32), dtype=torch.uint8).pin_memory()
    for i in range(count):
        host[i] = torch.from_numpy(make_blob(n, 7000 + first + i))
    dev = host.to("cuda")
    setup_s = time.perf_counter() - t0
    lens = (C.c_size_t * count)(*[n * 32] * count)
    hptr = (C.c_void_p * count)(*[host[i].data_ptr() for i in range(count)])
    dptr = (C.c_void_p * count)(*[dev[i].data_ptr() for i in range(count)])
    cm = C.create_string_buffer(32 * count)
    pf = C.create_string_buffer(32 * count)

    def step_dev():
        eng.check(lib.kzgb_commit_and_prove_blobs_dev(eng.h, dptr, hptr, lens, count, cm, pf))

    def step_e2e():
        eng.check(lib.kzgb_commit_and_prove_blobs(eng.h, hptr, lens, count, cm, pf))

    for _ in range(max(1, args.warmup // 2)):
        step_dev()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    ms = _timed(torch, dist, world, eng, lib, step_dev, args.steps)
    launches = eng.launch_count() - l0
    ref = (cm.raw, pf.raw)
    e2e_ms = _timed(torch, dist, world, eng, lib, step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    assert (cm.raw, pf.raw) == ref
    if rank == 0:
        cpu_baseline = None
        if not args.skip_cpu_baseline:
            olib = load_oracle_lib()
            threads = olib.ref_hw_threads()
            xy = C.create_string_buffer(64 * n)
            eng.check(lib.kzgb_srs_get_affine_mont(eng.h, 0, n, xy, None))
            b0 = bytes(host[0].numpy().tobytes())
            ocm = C.create_string_buffer(64)
            opf = C.create_string_buffer(64)
            t1 = time.perf_counter()
            reps = 4
            for _ in range(reps):
                olib.ref_commit_blob(b0, C.c_size_t(len(b0)), xy.raw, threads, 0, ocm)
                olib.ref_blob_proof(b0, C.c_size_t(len(b0)), ocm, xy.raw, threads, 0, opf)
            dt = (time.perf_counter() - t1) / reps
            c32 = C.create_string_buffer(32)
            p32 = C.create_string_buffer(32)
            lib.kzgb_g1_serialize_compressed(ocm, 0, c32)
            lib.kzgb_g1_serialize_compressed(opf, 0, p32)
            cpu_baseline = {"value": 1.0 / dt, "unit": "blobs/s", "cores": threads, "kind": "port",
                            "sample": "%d x 1 blob (2^%d Fr) commit+proof; C++ restatement of the arkworks algorithm" % (reps, logn),
                            "commitment_and_proof_match_gpu": c32.raw == ref[0][:32] and p32.raw == ref[1][:32]}
        line = {
            "metric": "blob commits+proofs/s (1024 blobs x 2^%d Fr, sharded by blob)" % logn, "value": total * args.steps / (ms * 1e-3),
            "unit": "blobs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 (8x32-bit Montgomery limbs, BN254 Fq/Fr)", "data": "synthetic",
            "config": {"workload": "batch of 1024 blobs x 2^%d Fr commit+proof, sharded by blob" % logn, "blobs_per_rank": count,
                       "l2": "inputs larger than L2: %d MiB of blobs per rank" % (count * n * 32 >> 20)},
            "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "blobs/s", "h2d_bytes_per_step": count * n * 32,
                    "d2h_bytes_per_step": count * 256, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": None, "cpu_baseline": cpu_baseline, "setup_s": setup_s,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    torch, dist, world, rank, local_rank = _dist_setup()
    from __graft_entry__ import load_package

    pkg = load_package()
    lib = pkg.lib
    eng = pkg.Engine(local_rank)
    logn = args.log_n or 12
    n = 1 << logn
    m = 4096
    t0 = time.perf_counter()
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    srs.precompute(n, 0)
    host = torch.empty((m, n * 32), dtype=torch.uint8).pin_memory()
    for i in range(m):
        host[i] = torch.from_numpy(make_blob(n, 90000 + 1000 * rank + i))
    lens = (C.c_size_t * m)(*[n * 32] * m)
    hptr = (C.c_void_p * m)(*[host[i].data_ptr() for i in range(m)])
    cm = C.create_string_buffer(32 * m)
    pf = C.create_string_buffer(32 * m)
    eng.check(lib.kzgb_commit_and_prove_blobs(eng.h, hptr, lens, m, cm, pf))  # the prover side, untimed

    def ark_to_gnark(b: bytes) -> bytes:  # arkworks LE + flags in byte 31  ->  gnark BE + flags in byte 0
        flags = b[31] & 0xC0
        body = bytes([b[31] & 0x3F]) + b[30::-1]
        top = 0x40 if flags & 0x40 else (0xC0 if flags & 0x80 else 0x80)
        return bytes([body[0] | top]) + body[1:]

    # the verifier's inputs are affine points: decompress the prover's outputs with the GPU decompressor (setup)
    tmp = pkg.Engine(local_rank)
    both = pkg.SRS.from_gnark_bytes(b"".join(ark_to_gnark(cm.raw[32 * i : 32 * i + 32]) for i in range(m))
                                    + b"".join(ark_to_gnark(pf.raw[32 * i : 32 * i + 32]) for i in range(m)), engine=tmp)
    xy = C.create_string_buffer(64 * 2 * m)
    inf = C.create_string_buffer(2 * m)
    tmp.check(lib.kzgb_srs_get_affine_mont(tmp.h, 0, 2 * m, xy, inf))
    cxy, cinf, pxy, pinf = xy.raw[: 64 * m], inf.raw[:m], xy.raw[64 * m :], inf.raw[m:]
    del both
    tmp.close()
    setup_s = time.perf_counter() - t0
    lhs, rhs = C.create_string_buffer(64), C.create_string_buffer(64)
    li, ri = C.c_uint8(0), C.c_uint8(0)

    def step():
        eng.check(lib.kzgb_verify_batch_rlc(eng.h, hptr, lens, m, cxy, cinf, pxy, pinf, lhs, C.byref(li), rhs, C.byref(ri)))

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    ms = _timed(torch, dist, world, eng, lib, step, args.steps)
    launches = eng.launch_count() - l0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if rank == 0:
        L = pkg.g1_from_abi(lhs.raw, bytes([li.value]))[0]
        Rr = pkg.g1_from_abi(rhs.raw, bytes([ri.value]))[0]
        ok = pkg.g1_lincomb([L], [TAU], eng) == Rr  # e(lhs, [tau]G2) == e(rhs, G2)  <=>  rhs = tau * lhs
        value = world * m * args.steps / (ms * 1e-3)
        line = {
            "metric": "batch-verify RLC pairs/s (4096 blob/proof pairs, 2^%d Fr each; pairing excluded)" % logn, "value": value,
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit Montgomery limbs, BN254 Fq/Fr)", "data": "synthetic",
            "config": {"workload": "verify_blob_kzg_proof_batch over 4096 pairs: per-blob challenge + evaluation, RLC scalars, "
                                   "G1 MSMs on GPU; final pairing stays in the reference code", "blob_fr": n,
                       "l2": "inputs larger than L2: %d MiB of blobs" % (m * n * 32 >> 20)},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": m * (n * 32 + 128), "d2h_bytes_per_step": m * 32 + 4096,
                    "ms_per_step": ms / args.steps},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": None,
            "cpu_baseline": {"rlc_outputs_satisfy_pairing_relation": ok}, "setup_s": setup_s,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
