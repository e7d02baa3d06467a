"""CPU checks of the FP64-pipe Fq arithmetic (csrc/field_dfma.cuh, csrc/ec_dfma.cuh): the same source the
kernels use, compiled for the host and run with the FPU in round-toward-zero mode (fma() == __fma_rz), against
Python big-int arithmetic and the oracle's group law.  No GPU involved."""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import bn254 as o

HERE = os.path.dirname(os.path.abspath(__file__))
P = o.P
W = 52
MASK = (1 << W) - 1
R260 = 1 << 260
R256 = 1 << 256


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "native", "dfma_test.cpp")
    out = os.path.join(HERE, "native", "libdfma_test.so")
    # -frounding-math / -ffp-contract=off: no compile-time folding or fusing of the floating-point operations
    subprocess.check_call(["g++", "-O1", "-frounding-math", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                           "-x", "c++", "-o", out, src])
    return C.CDLL(out)


def limbs(x):
    assert 0 <= x < R260
    return (C.c_uint64 * 5)(*[(x >> (W * i)) & MASK for i in range(5)])


def value(arr):
    assert all(v <= MASK for v in arr), "limb not normalised"
    return sum(int(v) << (W * i) for i, v in enumerate(arr))


def out5():
    return (C.c_uint64 * 5)()


EDGE = [0, 1, P - 1, P, P + 1, 2 * P, R260 - 1, (1 << 259) + 12345, MASK, MASK << 52, (1 << 260) - (1 << 208), 16 * P - 1]


def operands(rng, count, hi=R260):
    vals = list(EDGE)
    vals += [rng.randrange(hi) for _ in range(count)]
    vals += [sum(rng.choice([0, 1, MASK - 1, MASK]) << (W * i) for i in range(5)) for _ in range(count // 4)]
    return [v for v in vals if v < hi]


def test_generated_constants_are_current():
    import importlib.util
    path = os.path.join(os.path.dirname(HERE), "rust-kzg-bn254_b200", "csrc")
    spec = importlib.util.spec_from_file_location("gen_dfma", os.path.join(path, "gen_dfma.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(os.path.join(path, "field_dfma_consts.cuh")).read() == gen.render()  # (never rewritten: it is a build input)


def test_montgomery_product(lib):
    rng = random.Random(52)
    vals = operands(rng, 300)
    rinv = pow(R260, -1, P)
    for i, a in enumerate(vals):
        for b in (vals[(7 * i + 3) % len(vals)], vals[(i * i + 1) % len(vals)], a):
            if a * b // R260 + P >= R260:
                continue  # outside the domain: the result would not fit 5 limbs (operands of the curve code stay below 18 p)
            r = out5()
            lib.t5_mul(r, limbs(a), limbs(b))
            got = value(r)
            assert got % P == a * b * rinv % P
            assert got < a * b // R260 + P + 1  # the documented bound


def test_difference_of_products_in_one_reduction(lib):
    rng = random.Random(53)
    rinv = pow(R260, -1, P)
    for _ in range(400):
        a, b = rng.randrange(R260), rng.randrange(1 << 258)
        k = rng.randrange(1, 32)
        # c d <= k p 2^260
        c = rng.randrange(R260)
        d = rng.randrange(min(R260, k * P * R260 // max(c, 1) + 1))
        r = out5()
        lib.t5_mul2sub(r, limbs(a), limbs(b), limbs(c), limbs(d), k)
        got = value(r)
        assert got % P == (a * b - c * d) * rinv % P
        assert got < a * b // R260 + (k + 1) * P + 1
    # extremes: nothing subtracted / everything cancels
    for a, b, c, d, k in [(R260 - 1, 1 << 259, 0, 0, 0), (5, 7, 5, 7, 0), (P, P, P, P, 1), (0, 0, 16 * P, 2 * P, 1)]:
        r = out5()
        lib.t5_mul2sub(r, limbs(a), limbs(b), limbs(c), limbs(d), k)
        assert value(r) % P == (a * b - c * d) * rinv % P


def test_additions_and_subtractions(lib):
    rng = random.Random(54)
    vals = operands(rng, 200, hi=1 << 258)
    for i, a in enumerate(vals):
        b = vals[(5 * i + 2) % len(vals)]
        c = vals[(3 * i + 1) % len(vals)]
        r = out5()
        if a + b < R260:
            lib.t5_add(r, limbs(a), limbs(b))
            assert value(r) == a + b
        for k in (1, 6, 11, 16, 31):
            if b <= k * P and a - b + k * P < R260:
                lib.t5_sub(r, limbs(a), limbs(b), k)
                assert value(r) == a - b + k * P
            if b + c <= k * P and a - b - c + k * P < R260:
                lib.t5_sub2(r, limbs(a), limbs(b), limbs(c), k)
                assert value(r) == a - b - c + k * P


def test_zero_test_is_exact(lib):
    rng = random.Random(55)
    for k in range(32):
        assert lib.t5_is_zero_mod_p(limbs(k * P)) == 1
        for delta in (1, -1, 1 << 52, -(1 << 52), 1 << 208):
            v = k * P + delta
            if 0 <= v < 32 * P:
                assert lib.t5_is_zero_mod_p(limbs(v)) == 0
    for _ in range(2000):
        v = rng.randrange(32 * P)
        assert lib.t5_is_zero_mod_p(limbs(v)) == (1 if v % P == 0 else 0)
    # same low limb as a multiple of p, different upper limbs: the filter passes, the comparison must not
    for k in (1, 5, 17, 31):
        v = (k * P & MASK) | (((k * P >> 52) ^ 1) << 52)
        assert lib.t5_is_zero_mod_p(limbs(v)) == 0


def test_radix_conversions(lib):
    rng = random.Random(56)
    for x in [0, 1, P - 1, (1 << 254) - 1, (1 << 256) - 1] + [rng.randrange(1 << 256) for _ in range(300)]:
        r = out5()
        lib.t5_from_u32x8_times16(r, x.to_bytes(32, "little"))
        assert value(r) == 16 * x
    for a in operands(rng, 300):
        out = C.create_string_buffer(32)
        lib.t5_to_mont256(out, limbs(a))
        # a = x 2^260 -> canonical x 2^256
        assert int.from_bytes(out.raw, "little") == a * pow(16, -1, P) % P


def test_ranges_by_interval_arithmetic():
    """The fixed multiples of p in xyzz5_madd (16, 16, 6, 11, 1) against worst-case bounds, in units of p."""
    rho = P / R260
    mul = lambda a, b: a * b * rho + 1
    X1 = Y1 = 16.0
    ZZ = ZZZ = 1.0
    for _ in range(64):
        P_ = mul(16, ZZ) + 16
        R_ = mul(16, ZZZ) + 16
        PP, RR = mul(P_, P_), mul(R_, R_)
        PPP, Q = mul(P_, PP), mul(X1, PP)
        assert PPP + 2 * Q < 6  # X3 = RR - PPP - 2Q + 6p >= 0
        X3 = RR + 6
        assert X3 < 11  # Q - X3 + 11 p >= 0
        T = Q + 11
        assert Y1 * PPP * rho < 1  # Y1 PPP <= 1 p 2^260
        Y3 = R_ * T * rho + 2
        ZZ, ZZZ = max(ZZ, mul(ZZ, PP)), max(ZZZ, mul(ZZZ, PPP))
        X1, Y1 = max(X1, X3), max(Y1, Y3)
        assert max(X1, Y1) <= 16 and max(P_, R_, T) < 84  # representable below 2^260


def mont256_point(pt):
    return (pt[0] * R256 % P).to_bytes(32, "little") + (pt[1] * R256 % P).to_bytes(32, "little")


def run_chain(lib, pts):
    buf = b"".join(b"\0" * 64 if p is None else mont256_point(p) for p in pts)
    out5_, out8 = C.create_string_buffer(128), C.create_string_buffer(128)
    raw = (C.c_uint64 * 20)()
    lib.t5_accumulate(out5_, buf, len(pts), raw)
    lib.t8_accumulate(out8, buf, len(pts))
    aff5, aff8 = C.create_string_buffer(64), C.create_string_buffer(64)
    lib.t8_to_affine(aff5, out5_)
    lib.t8_to_affine(aff8, out8)
    words = [int.from_bytes(out5_.raw[i:i + 32], "little") for i in range(0, 128, 32)]
    assert all(w < P for w in words), "not canonical"
    return aff5.raw, aff8.raw, list(raw)


def test_xyzz_accumulation_matches_the_integer_formulas_and_the_oracle(lib):
    import golden_data as g

    rng = random.Random(57)
    base = g.srs_points_string()[:40]
    neg = lambda p: (p[0], (P - p[1]) % P)
    chains = [
        base[:1], base[:2], base[:25],
        [base[3], base[3]],                      # P + P: the doubling path
        [base[3], base[3], base[3], base[4]],
        [base[5], neg(base[5])],                 # P + (-P): identity
        [base[5], neg(base[5]), base[6], base[7]],
        [None, base[1], None, base[2]],          # identity entries of the padded sorted list
        [base[1], base[2], neg(base[2]), neg(base[1])],
        [],
    ]
    for _ in range(6):
        chains.append([rng.choice(base) if rng.random() < 0.8 else neg(rng.choice(base)) for _ in range(rng.randrange(2, 40))])
    rinv = pow(R256, -1, P)
    for pts in chains:
        a5, a8, raw = run_chain(lib, pts)
        assert a5 == a8
        expect = None
        for p_ in pts:
            if p_ is not None:
                expect = o.g1_add(expect, p_)
        got = None if a5 == b"\0" * 64 else (int.from_bytes(a5[:32], "little") * rinv % P, int.from_bytes(a5[32:], "little") * rinv % P)
        assert got == expect
        if len(pts) > 2 and got is not None and any(raw):
            x, y, zz, zzz = (value(raw[i:i + 5]) for i in range(0, 20, 5))
            assert x < 16 * P and y < 16 * P and zz < 2 * P and zzz < 2 * P  # the documented ranges
