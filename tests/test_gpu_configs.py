"""GPU parity at BASELINE-config sizes through size-independent properties of the synthetic SRS
(SRS_i = tau^i G with known tau): closed forms cost O(n) big-int work, no CPU MSM."""
import ctypes as C
import random

import pytest

import golden_data as g
from __graft_entry__ import load_package
from oracle import bn254 as o

pytestmark = pytest.mark.gpu
TAU = o.SYNTH_TAU


@pytest.fixture(scope="module")
def pkg():
    return load_package()


def _barycentric(evals, z):
    """p(z) for the interpolant of evals on the 2^k domain, O(n) with one batch inversion."""
    n = len(evals)
    w = o.PRIMITIVE_ROOTS_OF_UNITY[n.bit_length() - 1]
    roots = [1] * n
    for i in range(1, n):
        roots[i] = roots[i - 1] * w % o.R
    den = [(z - r) % o.R for r in roots]
    pref, acc = [], 1
    for d in den:
        pref.append(acc)
        acc = acc * d % o.R
    inv = pow(acc, -1, o.R)
    total = 0
    for i in range(n - 1, -1, -1):
        di = inv * pref[i] % o.R
        inv = inv * den[i] % o.R
        total = (total + evals[i] * roots[i] % o.R * di) % o.R
    return total * (pow(z, n, o.R) - 1) % o.R * pow(n, -1, o.R) % o.R


@pytest.mark.parametrize("logn", [16, 19])
def test_commit_and_proof_closed_form_at_bench_sizes(pkg, logn):
    """configs[1]/[2]: commitment = p(tau) G, proof = ((p(tau) - y)/(tau - z)) G, z from the real transcript."""
    n = 1 << logn
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    rnd = random.Random(logn)
    seed = [rnd.randrange(o.R) for _ in range(512)]
    evals = (seed * (n // 512))[:n]
    for k in range(0, n, 997):  # break the periodicity
        evals[k] = rnd.randrange(o.R)
    data = b"".join(e.to_bytes(32, "big") for e in evals)
    blob = pkg.Blob.from_unchecked(data)
    cs, ps = pkg.KZG.commit_and_prove_blobs([blob, blob], srs)
    ptau = _barycentric(evals, TAU)
    c = o.g1_mul(o.G1_GEN, ptau)
    assert cs[0] == cs[1] == o.g1_serialize_compressed(c)
    z = o.hash_to_field_element(
        o.FIAT_SHAMIR_PROTOCOL_DOMAIN + n.to_bytes(8, "big") + data + o.g1_serialize_compressed(c)
    )
    y = _barycentric(evals, z)
    pi = o.g1_mul(o.G1_GEN, (ptau - y) * pow((TAU - z) % o.R, -1, o.R) % o.R)
    assert ps[0] == ps[1] == o.g1_serialize_compressed(pi)
    # the single-call API agrees with the batch pipeline
    kzg = pkg.KZG()
    kzg.expanded_roots_of_unity = [0] * n  # only its length is checked (kzg.rs:135)
    assert kzg.commit_blob(blob, srs) == c
    assert kzg.compute_blob_proof(blob, c, srs) == pi


def test_config1_gpu_prover_verified_by_the_pairing(pkg):
    """BASELINE configs[0]: one 128 KiB blob (4096 Fr, the reference's blobs.txt) on a synthetic 2^12-point SRS.
    commit_blob + compute_blob_proof run on the GPU; verify_blob_kzg_proof (verifier/src/verify.rs:77-115: challenge,
    evaluation, e(C - y G1, G2) == e(proof, [tau - z] G2)) runs on the CPU oracle with [tau]G2 injected.  A proof for
    another blob fails, and the GPU's batch RLC outputs satisfy the final pairing of batch.rs:253-254."""
    n = 1 << 12
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    g2_tau = o.g2_mul(o.G2_GEN, TAU)
    data = g.blobs_txt()
    blob = pkg.Blob.new(data)
    kzg = pkg.KZG()
    kzg.calculate_and_store_roots_of_unity(len(blob))
    c = kzg.commit_blob(blob, srs)
    pi = kzg.compute_blob_proof(blob, c, srs)
    bo = o.Blob(data)
    assert o.verify_blob_kzg_proof(bo, c, pi, g2_tau)
    rnd = random.Random(41)
    other = b"".join(rnd.randrange(o.R).to_bytes(32, "big") for _ in range(n))
    blob2 = pkg.Blob.new(other)
    c2 = kzg.commit_blob(blob2, srs)
    pi2 = kzg.compute_blob_proof(blob2, c2, srs)
    assert o.verify_blob_kzg_proof(o.Blob(other), c2, pi2, g2_tau)
    assert not o.verify_blob_kzg_proof(bo, c, pi2, g2_tau)
    assert not o.verify_blob_kzg_proof(bo, c2, pi, g2_tau)
    # verify_blob_kzg_proof with its G1 side on the GPU (verify.rs:77-115): challenge, evaluation, C - [y] G1; the G2 side and
    # the pairing on the oracle: e(C - y G1, G2) == e(proof, [tau - z] G2)
    cmv, z, y = pkg.verify_blob_kzg_proof_g1(blob, c, pi, eng)
    assert z == o.compute_challenge(bo, c) and y == o.evaluate_polynomial_in_evaluation_form(bo.to_polynomial_eval_form(), z)
    x_minus_z = o.g2_add(g2_tau, o.g2_neg(o.g2_mul(o.G2_GEN, z)))
    assert o.pairings_verify(cmv, o.G2_GEN, pi, x_minus_z)
    cmv_bad, z2, _ = pkg.verify_blob_kzg_proof_g1(blob, c, pi2, eng)
    assert z2 == z and not o.pairings_verify(cmv_bad, o.G2_GEN, pi2, x_minus_z)
    assert pkg.verify_proof_g1(c, pi, y, eng) == cmv
    lhs, rhs = pkg.verify_blob_kzg_proof_batch_rlc([blob, blob2], [c, c2], [pi, pi2], eng)
    assert o.pairings_verify(lhs, g2_tau, rhs, o.G2_GEN)
    lhs_bad, rhs_bad = pkg.verify_blob_kzg_proof_batch_rlc([blob, blob2], [c, c2], [pi2, pi], eng)
    assert not o.pairings_verify(lhs_bad, g2_tau, rhs_bad, o.G2_GEN)


def test_point_range_sharded_msm_closed_form(pkg):
    """config 4 at 2^20 on one GPU in variable-base mode, as 1 range and as 3 ranges added on the host:
    scalars a^i => MSM = ((a tau)^N - 1)/(a tau - 1) G."""
    import torch

    sh = __import__("rust_kzg_bn254_b200.sharding", fromlist=["x"])
    N = 1 << 20
    a = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % o.R
    at = a * TAU % o.R
    expect = o.g1_mul(o.G1_GEN, (pow(at, N, o.R) - 1) * pow(at - 1, -1, o.R) % o.R)
    for world in (1, 3):
        partials = []
        for rank in range(world):
            first, count = sh.shard_range(N, rank, world)
            eng = pkg.Engine(0)
            srs = pkg.SRS.synthetic(count, TAU, engine=eng, first=first)
            srs.precompute(0, -1)
            scal = torch.empty(count * 32, dtype=torch.uint8, device="cuda")
            eng.check(pkg.lib.kzgb_fr_powers_dev(eng.h, pkg.fr_to_mont_bytes([a]), first, count, scal.data_ptr()))
            out = C.create_string_buffer(64)
            inf = C.c_uint8(0)
            eng.check(pkg.lib.kzgb_msm_srs_range_dev(eng.h, scal.data_ptr(), 0, count, out, C.byref(inf)))
            partials.append(out.raw + bytes([inf.value]))
            eng.close()
        assert sh.reduce_g1_partials(pkg, partials) == expect


def test_fixed_base_equals_variable_base(pkg):
    """The window-table MSM and the table-free MSM give the same point (2^16 points, random scalars)."""
    n = 1 << 16
    rnd = random.Random(5)
    sc = [rnd.randrange(o.R) for _ in range(n)]
    expect = o.tau_trick_msm(sc)
    e1 = pkg.Engine(0)
    s1 = pkg.SRS.synthetic(n, TAU, engine=e1)
    s1.precompute(n, 0)
    assert pkg.KZG().commit_coeff_form(pkg.PolynomialCoeffForm(sc), s1) == expect
    e2 = pkg.Engine(0)
    s2 = pkg.SRS.synthetic(n, TAU, engine=e2)
    s2.precompute(0, -1)
    assert pkg.KZG().commit_coeff_form(pkg.PolynomialCoeffForm(sc), s2) == expect
    # a different window width gives the same point too
    s1.precompute(n, 11)
    assert pkg.KZG().commit_coeff_form(pkg.PolynomialCoeffForm(sc), s1) == expect


def test_batch_verify_rlc_pairing_relation(pkg):
    """config 5 shape (m = 64 pairs of 2^12-Fr blobs): rhs = tau * lhs, and a tampered proof breaks it."""
    n = 1 << 12
    m = 64
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    rnd = random.Random(12)
    blobs = []
    for i in range(m):
        ev = [rnd.randrange(o.R) for _ in range(64)] * (n // 64)
        ev[i] = rnd.randrange(o.R)
        blobs.append(pkg.Blob.new(b"".join(e.to_bytes(32, "big") for e in ev)))
    blobs[3] = pkg.Blob.new(g.blobs_txt())
    cs, ps = pkg.KZG.commit_and_prove_blobs(blobs, srs)
    cpts = [o.g1_deserialize_compressed(c) for c in cs]
    ppts = [o.g1_deserialize_compressed(p) for p in ps]
    lhs, rhs = pkg.verify_blob_kzg_proof_batch_rlc(blobs, cpts, ppts, eng)
    assert lhs is not None and o.g1_mul(lhs, TAU) == rhs
    bad = list(ppts)
    bad[5] = o.g1_add(bad[5], o.G1_GEN)
    lhs2, rhs2 = pkg.verify_blob_kzg_proof_batch_rlc(blobs, cpts, bad, eng)
    assert o.g1_mul(lhs2, TAU) != rhs2
    # a small prefix against the full oracle restatement of batch.rs
    bo = [o.Blob(b.data()) for b in blobs[:3]]
    assert pkg.verify_blob_kzg_proof_batch_rlc(blobs[:3], cpts[:3], ppts[:3], eng) == o.verify_blob_kzg_proof_batch_rlc(bo, cpts[:3], ppts[:3])


def _geom(n):
    """sum_{i<n} tau^i mod r."""
    return (pow(TAU, n, o.R) - 1) * pow(TAU - 1, -1, o.R) % o.R


@pytest.mark.parametrize("waves", [4, 1])
def test_adversarial_scalar_distributions_2p16(pkg, waves):
    """SURVEY.md 8d 'D3' inputs at n = 2^16 on the fixed-base table, for two chunkings of the sorted list:
    all scalars equal (every point of a window in ONE bucket -> buckets spanning thousands
    of chunks), all r - 1 (all digits negative / carries through every window), a single non-zero
    scalar, all zero (identity), and two distinct values (two hot buckets per window).  Closed forms
    from SRS_i = tau^i G."""
    n = 1 << 16
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    srs.precompute(n, 0)
    kzg = pkg.KZG()
    G = o.G1_GEN
    s = 0x2F0E1D3C4B5A69788796A5B4C3D2E1F00112233445566778899AABBCCDDEEFF % o.R
    try:
        pkg.lib.kzgb_set_option(b"acc_waves", waves)
        assert kzg.commit_coeff_form(pkg.PolynomialCoeffForm([s] * n), srs) == o.g1_mul(G, s * _geom(n) % o.R)
        assert kzg.commit_coeff_form(pkg.PolynomialCoeffForm([o.R - 1] * n), srs) == o.g1_mul(G, (o.R - 1) * _geom(n) % o.R)
        one = [0] * n
        one[12345] = s
        assert kzg.commit_coeff_form(pkg.PolynomialCoeffForm(one), srs) == o.g1_mul(G, s * pow(TAU, 12345, o.R) % o.R)
        assert kzg.commit_coeff_form(pkg.PolynomialCoeffForm([0] * n), srs) is None
        t = (s * 7 + 1) % o.R
        even = (pow(TAU, n, o.R) - 1) * pow(TAU * TAU % o.R - 1, -1, o.R) % o.R  # sum of tau^(2i)
        two = [s, t] * (n // 2)
        assert kzg.commit_coeff_form(pkg.PolynomialCoeffForm(two), srs) == o.g1_mul(G, (s * even + t * TAU % o.R * even) % o.R)
    finally:
        pkg.lib.kzgb_set_option(b"acc_waves", 4)


def test_zero_and_constant_blobs_2p16(pkg):
    """verifier/tests/tests.rs:239-269 at size: the zero blob commits to the identity and proves to the
    identity; a constant polynomial c commits to c * G (the Lagrange basis sums to SRS_0) and its
    quotient is zero."""
    n = 1 << 16
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    zero = pkg.Blob.new(bytes(32 * n))
    cval = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF0123456789AB
    const = pkg.Blob.new(cval.to_bytes(32, "big") * n)
    cs, ps = pkg.KZG.commit_and_prove_blobs([zero, const, zero], srs)
    ident = o.g1_serialize_compressed(None)
    assert cs[0] == cs[2] == ident and ps[0] == ps[2] == ident
    assert cs[1] == o.g1_serialize_compressed(o.g1_mul(o.G1_GEN, cval))
    assert ps[1] == ident


def test_ntt_linearity_2p19(pkg):
    """ifft(a + 3 b) == ifft(a) + 3 ifft(b) element-wise at n = 2^19 (size-independent property)."""
    import numpy as np

    n = 1 << 19
    eng = pkg.Engine(0)
    rnd = random.Random(19)
    sa = [rnd.randrange(o.R) for _ in range(128)]
    sb = [rnd.randrange(o.R) for _ in range(128)]
    sc = [(x + 3 * y) % o.R for x, y in zip(sa, sb)]
    outs = []
    for seed in (sa, sb, sc):
        buf = C.create_string_buffer(pkg.fr_to_mont_bytes(seed) * (n // 128), 32 * n)
        eng.check(pkg.lib.kzgb_ntt_fr(eng.h, buf, n, 1))
        outs.append(buf.raw)
    for i in list(range(0, n, n // 64)) + [1, 2, 3, n - 1]:
        a, b, c = (pkg.fr_from_mont_bytes(o_[32 * i : 32 * i + 32])[0] for o_ in outs)
        assert c == (a + 3 * b) % o.R


# ---------------------------------------------------------------------------------------------------------
# BASELINE configs[2], [3] (one GPU's worth) and [4] at their FULL sizes, through size-independent properties
# ---------------------------------------------------------------------------------------------------------
def _raw_batch(lib, eng, host, count, nbytes):
    """kzgb_commit_and_prove_blobs on `count` blobs laid out back to back in a pinned torch tensor."""
    lens = (C.c_size_t * count)(*[nbytes] * count)
    ptrs = (C.c_void_p * count)(*[host[i].data_ptr() for i in range(count)])
    cm, pf = C.create_string_buffer(32 * count), C.create_string_buffer(32 * count)
    eng.check(lib.kzgb_commit_and_prove_blobs(eng.h, ptrs, lens, count, cm, pf))
    return cm.raw, pf.raw


def test_config3_full_size_1024_blobs_2p16(pkg):
    """configs[2] at full size: 1024 blobs x 2^16 Fr (2 GiB) through the grouped small-blob pipeline.  The second half
    of the batch repeats the first (equal inputs at different positions of different groups must give equal bytes),
    six blobs are checked against the closed forms commitment = p(tau) G and proof = ((p(tau) - y)/(tau - z)) G with
    z recomputed from the real transcript, and the whole batch is run twice (idempotence of the resident tables)."""
    import numpy as np
    import torch

    n, count = 1 << 16, 1024
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    rng = np.random.default_rng(2016)
    host = torch.empty((count, n * 32), dtype=torch.uint8).pin_memory()
    half = rng.integers(0, 256, size=(count // 2, n, 32), dtype=np.uint8)
    half[:, :, 0] = 0
    host[: count // 2] = torch.from_numpy(half.reshape(count // 2, -1))
    host[count // 2 :] = host[: count // 2]
    cm, pf = _raw_batch(pkg.lib, eng, host, count, n * 32)
    assert cm[: 32 * 512] == cm[32 * 512 :] and pf[: 32 * 512] == pf[32 * 512 :]
    assert len({cm[32 * i : 32 * i + 32] for i in range(512)}) == 512
    for i in (0, 1, 63, 64, 300, 511):
        data = bytes(host[i].numpy().tobytes())
        evals = [int.from_bytes(data[32 * k : 32 * k + 32], "big") for k in range(n)]
        ptau = _barycentric(evals, TAU)
        c = o.g1_mul(o.G1_GEN, ptau)
        assert cm[32 * i : 32 * i + 32] == o.g1_serialize_compressed(c)
        z = o.hash_to_field_element(o.FIAT_SHAMIR_PROTOCOL_DOMAIN + n.to_bytes(8, "big") + data + o.g1_serialize_compressed(c))
        y = _barycentric(evals, z)
        q = (ptau - y) * pow(TAU - z, -1, o.R) % o.R
        assert pf[32 * i : 32 * i + 32] == o.g1_serialize_compressed(o.g1_mul(o.G1_GEN, q))
    assert _raw_batch(pkg.lib, eng, host, count, n * 32) == (cm, pf)
    eng.close()


def test_config5_full_size_4096_pairs(pkg):
    """configs[4] at full size: verify_blob_kzg_proof_batch's front half + RLC over 4096 (blob, commitment, proof)
    triples of 2^12-Fr blobs.  The commitments and proofs come from the GPU prover; the two RLC outputs must satisfy
    what the final pairing checks, rhs = tau * lhs (batch.rs:253-254 with [tau]G2), the device-hashed and host-hashed
    challenge paths must agree, and one tampered proof among the 4096 must break the relation."""
    import numpy as np
    import torch

    n, m = 1 << 12, 4096
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    rng = np.random.default_rng(4096)
    arr = rng.integers(0, 256, size=(m, n, 32), dtype=np.uint8)
    arr[:, :, 0] = 0
    host = torch.from_numpy(arr.reshape(m, -1)).pin_memory()
    cm, pf = _raw_batch(pkg.lib, eng, host, m, n * 32)
    cpts = [o.g1_deserialize_compressed(cm[32 * i : 32 * i + 32]) for i in range(m)]
    ppts = [o.g1_deserialize_compressed(pf[32 * i : 32 * i + 32]) for i in range(m)]
    cxy, cinf = pkg.g1_to_abi(cpts)
    lens = (C.c_size_t * m)(*[n * 32] * m)
    ptrs = (C.c_void_p * m)(*[host[i].data_ptr() for i in range(m)])

    def rlc(proofs):
        pxy, pinf = pkg.g1_to_abi(proofs)
        lhs, rhs = C.create_string_buffer(64), C.create_string_buffer(64)
        li, ri = C.c_uint8(0), C.c_uint8(0)
        eng.check(pkg.lib.kzgb_verify_batch_rlc(eng.h, ptrs, lens, m, cxy, cinf, pxy, pinf, lhs, C.byref(li), rhs, C.byref(ri)))
        return pkg.g1_from_abi(lhs.raw, bytes([li.value]))[0], pkg.g1_from_abi(rhs.raw, bytes([ri.value]))[0]

    lhs, rhs = rlc(ppts)
    assert lhs is not None and o.g1_mul(lhs, TAU) == rhs
    try:
        pkg.lib.kzgb_set_option(b"fs_device", 0)  # host SHA-256 pool instead of the device kernel
        assert rlc(ppts) == (lhs, rhs)
    finally:
        pkg.lib.kzgb_set_option(b"fs_device", -1)
    bad = list(ppts)
    bad[2777] = o.g1_add(bad[2777], o.G1_GEN)
    lhs2, rhs2 = rlc(bad)
    assert o.g1_mul(lhs2, TAU) != rhs2
    eng.close()


def test_fixed_base_msm_2p24_closed_form(pkg):
    """One quarter of config 4's point count on one GPU with its window table (the library picks c = 22:
    12 windows x 2^24 points x 64 B = 12 GiB): scalars a^i => MSM = ((a tau)^N - 1)/(a tau - 1) G; the host-scalar
    entry point (chunked upload) must return the same point as the device-resident one."""
    import torch

    N = 1 << 24
    a = 0x1D5F3C29A7B4E6081122334455667788990AABBCCDDEEFF0123456789ABCDEF1 % o.R
    at = a * TAU % o.R
    expect = o.g1_mul(o.G1_GEN, (pow(at, N, o.R) - 1) * pow(at - 1, -1, o.R) % o.R)
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(N, TAU, engine=eng)
    srs.precompute(N, 0)
    cb, cw, ct = C.c_int(0), C.c_int(0), C.c_size_t(0)
    pkg.lib.kzgb_msm_config(eng.h, C.byref(cb), C.byref(cw), C.byref(ct))
    assert ct.value == N and cb.value >= 20
    scal = torch.empty(N * 32, dtype=torch.uint8, device="cuda")
    eng.check(pkg.lib.kzgb_fr_powers_dev(eng.h, pkg.fr_to_mont_bytes([a]), 0, N, scal.data_ptr()))
    out, inf = C.create_string_buffer(64), C.c_uint8(0)
    eng.check(pkg.lib.kzgb_msm_srs_range_dev(eng.h, scal.data_ptr(), 0, N, out, C.byref(inf)))
    assert pkg.g1_from_abi(out.raw, bytes([inf.value]))[0] == expect
    host = scal.cpu()
    eng.check(pkg.lib.kzgb_msm_srs_range(eng.h, host.data_ptr(), 0, N, out, C.byref(inf)))
    assert pkg.g1_from_abi(out.raw, bytes([inf.value]))[0] == expect
    eng.close()
