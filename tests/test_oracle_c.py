"""The C++ CPU baseline (oracle/cpu_ref.cpp) agrees with the big-int oracle and the reference
fixtures -- so the number bench.py reports beside the GPU is for the right computation."""
import ctypes as C
import os
import random
import subprocess

import pytest

import golden_data as g
from oracle import bn254 as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MONT = 1 << 256


@pytest.fixture(scope="module")
def olib():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return C.CDLL(os.path.join(ROOT, "oracle", "liboracle_cpu.so"))


def aff_bytes(pts):
    return b"".join(bytes(64) if p is None else (p[0] * MONT % o.P).to_bytes(32, "little") + (p[1] * MONT % o.P).to_bytes(32, "little") for p in pts)


def aff_from(buf):
    if not any(buf):
        return None
    ri = pow(MONT, -1, o.P)
    return (int.from_bytes(buf[:32], "little") * ri % o.P, int.from_bytes(buf[32:64], "little") * ri % o.P)


def fr_bytes(v):
    return b"".join((x % o.R * MONT % o.R).to_bytes(32, "little") for x in v)


def fr_from(buf):
    ri = pow(MONT, -1, o.R)
    return [int.from_bytes(buf[i : i + 32], "little") * ri % o.R for i in range(0, len(buf), 32)]


def test_sha256_is_the_oracles_own_and_matches_hashlib(olib):
    """The C++ oracle links nothing of the product: its SHA-256 is a plain FIPS 180-4 implementation of its own."""
    import hashlib

    rnd = random.Random(9)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 127, 128, 1000, 4096 + 17, 70001):
        data = bytes(rnd.getrandbits(8) for _ in range(n))
        out = C.create_string_buffer(32)
        olib.ref_sha256(data, C.c_size_t(n), out)
        assert out.raw == hashlib.sha256(data).digest(), n
    mk = open(os.path.join(ROOT, "oracle", "Makefile")).read()
    assert "rust-kzg-bn254_b200" not in mk.split("$(OUT):")[1]  # no product source in the oracle's build line


def test_inverse(olib):
    rnd = random.Random(1)
    for mod, fn in ((o.R, olib.ref_fr_inv), (o.P, olib.ref_fq_inv)):
        for x in [1, 2, mod - 1] + [rnd.randrange(1, mod) for _ in range(50)]:
            out = C.create_string_buffer(32)
            fn((x * MONT % mod).to_bytes(32, "little"), out)
            assert int.from_bytes(out.raw, "little") == pow(x, -1, mod) * MONT % mod


def test_ntt(olib):
    rnd = random.Random(2)
    for n in (1, 2, 8, 64, 1024):
        v = [rnd.randrange(o.R) for _ in range(n)]
        for inverse, ref in ((0, o.fft), (1, o.ifft)):
            buf = C.create_string_buffer(fr_bytes(v), 32 * n)
            olib.ref_ntt(buf, C.c_size_t(n), inverse, 4)
            assert fr_from(buf.raw) == ref(v)


def test_msm_and_g1_ifft(olib):
    rnd = random.Random(3)
    pts = g.srs_points_string()
    for n in (1, 5, 40, 700):
        sc = [rnd.randrange(o.R) for _ in range(n)]
        out = C.create_string_buffer(64)
        olib.ref_msm(aff_bytes(pts[:n]), fr_bytes(sc), C.c_size_t(n), 4, out)
        assert aff_from(out.raw) == o.msm(pts[:n], sc)
    out = C.create_string_buffer(64 * 64)
    olib.ref_g1_ifft(aff_bytes(pts[:64]), C.c_size_t(64), 4, out)
    assert [aff_from(out.raw[64 * i : 64 * i + 64]) for i in range(64)] == g.lagrange_srs_64()


def test_commit_and_proof(olib):
    pts = g.srs_points_string()
    srs = aff_bytes(pts[:256])
    rnd = random.Random(4)
    for raw in (g.gettysburg(), b"x", bytes(rnd.getrandbits(8) for _ in range(31 * 200))):
        bo = o.Blob.from_raw_data(raw)
        ko = o.KZG()
        ko.calculate_and_store_roots_of_unity(len(bo))
        c = ko.commit_blob(bo, pts[:256])
        pi = ko.compute_blob_proof(bo, c, pts[:256])
        for literal in (0, 1):
            if literal and len(bo) > 2048:
                continue
            cm = C.create_string_buffer(64)
            olib.ref_commit_blob(bo.data(), C.c_size_t(len(bo)), srs, 4, literal, cm)
            assert aff_from(cm.raw) == c
            pf = C.create_string_buffer(64)
            olib.ref_blob_proof(bo.data(), C.c_size_t(len(bo)), cm, srs, 4, literal, pf)
            assert aff_from(pf.raw) == pi
    # proof at a root of unity (z in the domain): the 40 reference KATs
    bo = o.Blob.from_raw_data(g.gettysburg())
    ev = bo.to_polynomial_eval_form().evaluations
    roots = o.calculate_roots_of_unity(len(bo))
    for idx, x, y in g.proof_eq_input()[:10]:
        pf = C.create_string_buffer(64)
        yo = C.create_string_buffer(32)
        olib.ref_proof_at(fr_bytes(ev), C.c_size_t(64), fr_bytes([roots[idx]]), aff_bytes(pts[:64]), 2, pf, yo)
        assert aff_from(pf.raw) == (x, y)
        assert fr_from(yo.raw)[0] == ev[idx]


def test_srs_decompress_evaluate_challenge_exports(olib):
    """The C++ oracle's SRS decompression (helpers.rs:175-226) against srs.g1.points.string, and its
    stand-alone evaluate / challenge exports against the big-int oracle."""
    olib.ref_srs_decompress.restype = C.c_size_t
    raw = g.g1_point_bytes()
    n = 3000
    out = C.create_string_buffer(64 * n)
    assert olib.ref_srs_decompress(raw, C.c_size_t(n), 4, out) == 0
    pts = g.srs_points_string()
    assert [aff_from(out.raw[64 * i : 64 * i + 64]) for i in range(n)] == pts
    inf = bytes([0x40]) + bytes(31)
    assert olib.ref_srs_decompress(inf + raw[:32], C.c_size_t(2), 1, out) == 0
    assert aff_from(out.raw[:64]) is None and aff_from(out.raw[64:128]) == pts[0]
    bo = o.Blob.from_raw_data(g.gettysburg())
    po = bo.to_polynomial_eval_form()
    rnd = random.Random(9)
    z = rnd.randrange(o.R)
    y = C.create_string_buffer(32)
    olib.ref_evaluate(fr_bytes(po.evaluations), C.c_size_t(len(po.evaluations)), fr_bytes([z]), y)
    assert fr_from(y.raw) == [o.evaluate_polynomial_in_evaluation_form(po, z)]
    c = o.KZG().commit_blob(bo, pts[:64])
    zc = C.create_string_buffer(32)
    olib.ref_challenge(bo.blob_data, C.c_size_t(len(bo.blob_data)), aff_bytes([c]), zc)
    assert fr_from(zc.raw) == [o.compute_challenge(bo, c)]
