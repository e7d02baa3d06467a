"""Big-integer CPU restatement of the rust-kzg-bn254 commitment path (TEST ORACLE).

Every function cites the reference file:line it follows (paths relative to the
reference repo root).  The heavy arithmetic of the reference lives in arkworks
0.5 (ark-bn254 / ark-ec / ark-poly / ark-ff / ark-serialize 0.5.0, Cargo.toml:57-62),
which is not vendored; its published algorithms/conventions are restated here and
anchored on the reference's own fixtures (see oracle/__init__.py).

Representation: Fq/Fr elements are Python ints in canonical form; G1 affine
points are ``(x, y)`` tuples, the identity is ``None``.
"""
from __future__ import annotations

import hashlib
from typing import Iterable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------
# Fields and curve (ark-bn254; constants corroborated by primitives/src/arith.rs:9
# and primitives/src/helpers.rs:161-164)
# ----------------------------------------------------------------------------
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr
B_COEFF = 3  # y^2 = x^3 + 3   (helpers.rs:202)
G1_GEN = (1, 2)
MONT_R = 1 << 256  # arkworks Fp<MontBackend<_,4>,4>: R = 2^256 (arith.rs:4-55)

BYTES_PER_FIELD_ELEMENT = 32  # consts.rs:4
SIZE_OF_G1_AFFINE_COMPRESSED = 32  # consts.rs:5
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"EIGENDA_FSBLOBVERIFY_V1_"  # consts.rs:8
RANDOM_CHALLENGE_KZG_BATCH_DOMAIN = b"EIGENDA_RCKZGBATCH___V1_"  # consts.rs:11
MAINNET_SRS_G1_SIZE = 268435456  # consts.rs:66
# arkworks Fr TWO_ADIC_ROOT_OF_UNITY == PRIMITIVE_ROOTS_OF_UNITY[28] (consts.rs:51)
TWO_ADIC_ROOT = 19103219067921713944291392827692070036145651957329286315305642004821462161904
TWO_ADICITY = 28


def _roots_table() -> List[int]:
    t = [0] * (TWO_ADICITY + 1)
    t[TWO_ADICITY] = TWO_ADIC_ROOT
    for k in range(TWO_ADICITY - 1, -1, -1):
        t[k] = t[k + 1] * t[k + 1] % R
    return t


# consts.rs:22-52 — table[k] is a primitive 2^k-th root of unity
PRIMITIVE_ROOTS_OF_UNITY = _roots_table()


class KzgError(Exception):
    """Mirror of primitives/src/errors.rs:32-86 (variant name + message)."""

    def __init__(self, variant: str, msg: str = ""):
        super().__init__(f"{variant}: {msg}" if msg else variant)
        self.variant = variant
        self.msg = msg


def fq_inv(a: int) -> int:
    return pow(a, P - 2, P)


def fr_inv(a: int) -> int:
    return pow(a, R - 2, R)


def fq_sqrt(a: int) -> Optional[int]:
    # p = 3 mod 4
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


# ----------------------------------------------------------------------------
# Montgomery reduction KAT helper (primitives/src/arith.rs:4-55)
# ----------------------------------------------------------------------------
def montgomery_reduce(z0: int, z1: int, z2: int, z3: int) -> Tuple[int, int, int, int]:
    """arith.rs:4-55: (z0..z3 as a 256-bit LE integer) * 2^-256 mod p, as 4 LE u64."""
    v = z0 | (z1 << 64) | (z2 << 128) | (z3 << 192)
    out = v * pow(MONT_R, -1, P) % P
    m = (1 << 64) - 1
    return (out & m, (out >> 64) & m, (out >> 128) & m, (out >> 192) & m)


# ----------------------------------------------------------------------------
# G1 arithmetic (ark-ec short Weierstrass; affine API, Jacobian inside)
# ----------------------------------------------------------------------------
Affine = Optional[Tuple[int, int]]
_JINF = (1, 1, 0)


def is_on_curve(pt: Affine) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - B_COEFF) % P == 0


def g1_neg(pt: Affine) -> Affine:
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % P)


def _jdbl(p):
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return _JINF
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def _jadd(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        if S1 == S2:
            return _jdbl(p)
        return _JINF
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def _jmadd(p, q: Tuple[int, int]):
    """Jacobian += affine (q not identity)."""
    X1, Y1, Z1 = p
    if Z1 == 0:
        return (q[0], q[1], 1)
    x2, y2 = q
    Z1Z1 = Z1 * Z1 % P
    U2 = x2 * Z1Z1 % P
    S2 = y2 * Z1 * Z1Z1 % P
    if U2 == X1:
        if S2 == Y1:
            return _jdbl(p)
        return _JINF
    H = (U2 - X1) % P
    HH = H * H % P
    I = 4 * HH % P
    J = H * I % P
    r = 2 * (S2 - Y1) % P
    V = X1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * Y1 * J) % P
    Z3 = ((Z1 + H) * (Z1 + H) - Z1Z1 - HH) % P
    return (X3, Y3, Z3)


def _to_jac(pt: Affine):
    return _JINF if pt is None else (pt[0], pt[1], 1)


def _to_aff(p) -> Affine:
    X, Y, Z = p
    if Z == 0:
        return None
    zi = fq_inv(Z)
    zi2 = zi * zi % P
    return (X * zi2 % P, Y * zi2 * zi % P)


def _batch_to_aff(ps) -> List[Affine]:
    """Montgomery-trick batch normalisation."""
    zs = [p[2] for p in ps]
    pref = []
    acc = 1
    for z in zs:
        pref.append(acc)
        if z:
            acc = acc * z % P
    inv = fq_inv(acc)
    out: List[Affine] = [None] * len(ps)
    for i in range(len(ps) - 1, -1, -1):
        z = zs[i]
        if not z:
            continue
        zi = inv * pref[i] % P
        inv = inv * z % P
        zi2 = zi * zi % P
        out[i] = (ps[i][0] * zi2 % P, ps[i][1] * zi2 * zi % P)
    return out


def g1_add(a: Affine, b: Affine) -> Affine:
    return _to_aff(_jadd(_to_jac(a), _to_jac(b)))


def _jmul(p, k: int):
    acc = _JINF
    if k == 0 or p[2] == 0:
        return acc
    for bit in bin(k)[2:]:
        acc = _jdbl(acc)
        if bit == "1":
            acc = _jadd(acc, p)
    return acc


def g1_mul(pt: Affine, k: int) -> Affine:
    return _to_aff(_jmul(_to_jac(pt), k % R))


def msm(bases: Sequence[Affine], scalars: Sequence[int]) -> Affine:
    """ark-ec 0.5 ``VariableBaseMSM::msm`` (call sites prover/src/kzg.rs:100,121,
    primitives/src/helpers.rs:332).  Result is a group element, so any correct
    bucket method gives the same affine point; this one is a plain unsigned
    Pippenger.  Length mismatch -> the reference maps arkworks' Err(min_len) to
    CommitError/MsmError."""
    if len(bases) != len(scalars):
        raise KzgError("MsmError", str(min(len(bases), len(scalars))))
    n = len(bases)
    if n == 0:
        return None
    c = 3 if n < 32 else max(3, min(16, n.bit_length() * 69 // 100 + 2))
    nwin = (254 + c - 1) // c
    mask = (1 << c) - 1
    sc = [s % R for s in scalars]
    win_sums = []
    for w in range(nwin):
        buckets = [_JINF] * (1 << c)
        sh = w * c
        for pt, s in zip(bases, sc):
            d = (s >> sh) & mask
            if d and pt is not None:
                buckets[d] = _jmadd(buckets[d], pt)
        run = _JINF
        tot = _JINF
        for d in range(mask, 0, -1):
            if buckets[d][2]:
                run = _jadd(run, buckets[d])
            tot = _jadd(tot, run)
        win_sums.append(tot)
    acc = _JINF
    for w in range(nwin - 1, -1, -1):
        for _ in range(c):
            acc = _jdbl(acc)
        acc = _jadd(acc, win_sums[w])
    return _to_aff(acc)


def g1_lincomb(points: Sequence[Affine], scalars: Sequence[int]) -> Affine:
    """primitives/src/helpers.rs:328-337."""
    return msm(points, scalars)


def validate_g1_point(pt: Affine) -> None:
    """primitives/src/helpers.rs:694-708 (subgroup check is trivially true on
    BN254 G1, cofactor 1; the identity passes ``is_on_curve``)."""
    if not is_on_curve(pt):
        raise KzgError("NotOnCurveError", "G1 point not on curve")


# ----------------------------------------------------------------------------
# Byte codecs
# ----------------------------------------------------------------------------
def lexicographically_largest(y: int) -> bool:
    """primitives/src/helpers.rs:151-173: y > (p-1)/2."""
    return y > (P - 1) // 2


def read_g1_point_from_bytes_be(b: bytes) -> Affine:
    """primitives/src/helpers.rs:175-226 (gnark compressed, big-endian)."""
    if len(b) != SIZE_OF_G1_AFFINE_COMPRESSED:
        raise KzgError("DeserializationError", "not enough bytes for g1 point")
    m_mask = 0b11 << 6
    m_inf, m_small, m_large = 0b01 << 6, 0b10 << 6, 0b11 << 6
    m_data = b[0] & m_mask
    if m_data == m_inf:
        if (b[0] & ~m_mask & 0xFF) != 0 or any(b[1:32]):
            raise KzgError("DeserializationError", "point at infinity not coded properly for g1")
        return None
    x = int.from_bytes(bytes([b[0] & ~m_mask & 0xFF]) + b[1:], "big") % P
    y2 = (x * x * x + B_COEFF) % P
    y = fq_sqrt(y2)
    if y is None:
        raise KzgError("NotOnCurveError", "compressed g1 point not on curve")
    if lexicographically_largest(y):
        if m_data == m_small:
            y = (-y) % P
    elif m_data == m_large:
        y = (-y) % P
    return (x, y)


def g1_to_gnark_be(pt: Affine) -> bytes:
    """Inverse of read_g1_point_from_bytes_be (the g1.point file format,
    helpers.rs:182-216): 0x40 = infinity, 0x80 = smaller y, 0xC0 = larger y."""
    if pt is None:
        return bytes([0x40]) + bytes(31)
    x, y = pt
    b = bytearray(x.to_bytes(32, "big"))
    b[0] |= 0xC0 if lexicographically_largest(y) else 0x80
    return bytes(b)


def g1_serialize_compressed(pt: Affine) -> bytes:
    """ark-serialize 0.5 ``CanonicalSerialize::serialize_compressed`` for
    ark-ec short-Weierstrass ``Affine`` (call sites helpers.rs:458-460,
    verifier/src/batch.rs:128,149-151): x little-endian, flags OR-ed into the
    top bits of byte 31: 0x80 iff y > -y, 0x40 for the identity (x written as 0).
    PARITY UNPINNED by the reference's fixtures."""
    if pt is None:
        b = bytearray(32)
        b[31] |= 0x40
        return bytes(b)
    x, y = pt
    b = bytearray(x.to_bytes(32, "little"))
    if y > (P - y) % P:
        b[31] |= 0x80
    return bytes(b)


def g1_deserialize_compressed(b: bytes) -> Affine:
    flags = b[31] & 0xC0
    if flags & 0x40:
        return None
    x = int.from_bytes(b[:31] + bytes([b[31] & 0x3F]), "little")
    y = fq_sqrt((x * x * x + B_COEFF) % P)
    if y is None:
        raise KzgError("NotOnCurveError", "not on curve")
    neg = y > (P - y) % P
    if bool(flags & 0x80) != neg:
        y = (-y) % P
    return (x, y)


def pad_payload(data: bytes) -> bytes:
    """primitives/src/helpers.rs:823-840."""
    out = bytearray()
    for i in range(0, len(data), 31):
        chunk = data[i : i + 31]
        out += b"\x00" + chunk + bytes(31 - len(chunk))
    return bytes(out)


def remove_internal_padding(padded: bytes) -> bytes:
    """primitives/src/helpers.rs:856-874."""
    if len(padded) % 32 != 0:
        raise KzgError("InvalidInputLength")
    return b"".join(padded[i + 1 : i + 32] for i in range(0, len(padded), 32))


def to_fr_array(data: bytes) -> List[int]:
    """primitives/src/helpers.rs:40-57: 32 B big-endian chunks -> Fr (mod r);
    the trailing partial chunk is right-padded with zeros."""
    out = []
    for i in range(0, len(data), 32):
        chunk = data[i : i + 32]
        if len(chunk) < 32:
            chunk = chunk + bytes(32 - len(chunk))
        out.append(int.from_bytes(chunk, "big") % R)
    return out


def to_byte_array(frs: Sequence[int], max_output_size: int) -> bytes:
    """primitives/src/helpers.rs:80-119."""
    n = len(frs)
    size = min(n * 32, max_output_size)
    data = bytearray(size)
    for i, e in enumerate(frs):
        v = (e % R).to_bytes(32, "big")
        start, end = i * 32, (i + 1) * 32
        if end > max_output_size:
            k = min(32, max_output_size - start)
            data[start : start + k] = v[:k]
            break
        data[start:end] = v
    return bytes(data)


def validate_blob_data_as_canonical_field_elements(data: bytes) -> None:
    """primitives/src/helpers.rs:784-811."""
    if len(data) % 32 != 0:
        raise KzgError("InvalidInputLength")
    for i in range(0, len(data), 32):
        if int.from_bytes(data[i : i + 32], "big") >= R:
            raise KzgError(
                "InvalidFieldElement",
                f"Field element at position {i // 32} is not canonical or invalid",
            )


def usize_to_be_bytes(n: int) -> bytes:
    """primitives/src/helpers.rs:769-779."""
    return n.to_bytes(8, "big")


def hash_to_field_element(msg: bytes) -> int:
    """primitives/src/helpers.rs:382-390."""
    return int.from_bytes(hashlib.sha256(msg).digest(), "big") % R


def next_power_of_two(n: int) -> int:
    """Rust usize::next_power_of_two (0 -> 1)."""
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


# ----------------------------------------------------------------------------
# Containers (primitives/src/blob.rs, primitives/src/polynomial.rs)
# ----------------------------------------------------------------------------
class PolynomialEvalForm:
    """primitives/src/polynomial.rs:13-140."""

    def __init__(self, evals: Sequence[int], len_underlying_blob_bytes: Optional[int] = None):
        if len(evals) > MAINNET_SRS_G1_SIZE:
            raise KzgError("GenericError", "Input size exceeds maximum polynomial size")
        self.len_underlying_blob_bytes = (
            len(evals) * 32 if len_underlying_blob_bytes is None else len_underlying_blob_bytes
        )
        n = next_power_of_two(len(evals))
        self.evaluations = [e % R for e in evals] + [0] * (n - len(evals))

    def __len__(self):
        return len(self.evaluations)

    def to_bytes_be(self) -> bytes:
        return to_byte_array(self.evaluations, self.len_underlying_blob_bytes)

    def to_coeff_form(self) -> "PolynomialCoeffForm":
        """polynomial.rs:130-140."""
        return PolynomialCoeffForm(ifft(self.evaluations), self.len_underlying_blob_bytes)


class PolynomialCoeffForm:
    """primitives/src/polynomial.rs:143-251."""

    def __init__(self, coeffs: Sequence[int], len_underlying_blob_bytes: Optional[int] = None):
        if len(coeffs) > MAINNET_SRS_G1_SIZE:
            raise KzgError("GenericError", "Input size exceeds maximum polynomial size")
        self.len_underlying_blob_bytes = (
            len(coeffs) * 32 if len_underlying_blob_bytes is None else len_underlying_blob_bytes
        )
        n = next_power_of_two(len(coeffs))
        self.coeffs = [c % R for c in coeffs] + [0] * (n - len(coeffs))

    def __len__(self):
        return len(self.coeffs)

    def to_bytes_be(self) -> bytes:
        return to_byte_array(self.coeffs, self.len_underlying_blob_bytes)

    def to_eval_form(self) -> PolynomialEvalForm:
        """polynomial.rs:241-251."""
        return PolynomialEvalForm(fft(self.coeffs), self.len_underlying_blob_bytes)


class Blob:
    """primitives/src/blob.rs:15-97."""

    def __init__(self, blob_data: bytes, _validate: bool = True):
        if _validate:
            validate_blob_data_as_canonical_field_elements(blob_data)
        self.blob_data = bytes(blob_data)

    @classmethod
    def from_raw_data(cls, raw: bytes) -> "Blob":
        return cls(pad_payload(raw), _validate=False)

    @classmethod
    def from_unchecked(cls, data: bytes) -> "Blob":
        """``impl From<Vec<u8>> for Blob`` (blob.rs:90-97): no validation."""
        return cls(data, _validate=False)

    def to_raw_data(self) -> bytes:
        return remove_internal_padding(self.blob_data)

    def data(self) -> bytes:
        return self.blob_data

    def __len__(self):
        return len(self.blob_data)

    def to_polynomial_eval_form(self) -> PolynomialEvalForm:
        return PolynomialEvalForm(to_fr_array(self.blob_data))

    def to_polynomial_coeff_form(self) -> PolynomialCoeffForm:
        return PolynomialCoeffForm(to_fr_array(self.blob_data))


# ----------------------------------------------------------------------------
# Roots of unity and Fr (I)FFT
# ----------------------------------------------------------------------------
def get_primitive_root_of_unity(power: int) -> int:
    """primitives/src/helpers.rs:365-370."""
    if power >= len(PRIMITIVE_ROOTS_OF_UNITY):
        raise KzgError("GenericError", "power must be <= 28")
    return PRIMITIVE_ROOTS_OF_UNITY[power]


def calculate_roots_of_unity(length_of_data_after_padding: int) -> List[int]:
    """primitives/src/helpers.rs:553-610: n = next_pow2(ceil(len/32)); returns
    [w^0 .. w^(n-1)], w = table[log2 n]."""
    if length_of_data_after_padding == 0:
        raise KzgError("GenericError", "Length of data after padding is 0")
    nelem = -(-length_of_data_after_padding // 32)
    if nelem > MAINNET_SRS_G1_SIZE:
        raise KzgError(
            "GenericError",
            "the length of data after padding is not valid with respect to the SRS",
        )
    n = next_power_of_two(nelem)
    w = get_primitive_root_of_unity(n.bit_length() - 1)
    # expand_root_of_unity (helpers.rs:592-610) then truncate the trailing 1
    roots = [1, w]
    while roots[-1] != 1:
        roots.append(roots[-1] * w % R)
    return roots[:-1]


def _domain_roots(n: int) -> List[int]:
    w = PRIMITIVE_ROOTS_OF_UNITY[n.bit_length() - 1]
    out = [1] * n
    for i in range(1, n):
        out[i] = out[i - 1] * w % R
    return out


def _bitrev_permute(a: list) -> list:
    n = len(a)
    bits = n.bit_length() - 1
    out = list(a)
    for i in range(n):
        j = int(format(i, f"0{bits}b")[::-1], 2) if bits else 0
        if j > i:
            out[i], out[j] = out[j], out[i]
    return out


def _ntt(vals: Sequence[int], w: int) -> List[int]:
    """Iterative radix-2 NTT, natural order in and out."""
    n = len(vals)
    a = _bitrev_permute([v % R for v in vals])
    length = 2
    while length <= n:
        wl = pow(w, n // length, R)
        half = length // 2
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * wl % R
        for s in range(0, n, length):
            for j in range(half):
                u = a[s + j]
                v = a[s + j + half] * tw[j] % R
                a[s + j] = (u + v) % R
                a[s + j + half] = (u - v) % R
        length *= 2
    return a


def fft(coeffs: Sequence[int]) -> List[int]:
    """ark-poly 0.5 ``GeneralEvaluationDomain::<Fr>::new(n).fft`` (call site
    primitives/src/polynomial.rs:242-246): natural-order coefficients ->
    evaluations p(w^i), w = TWO_ADIC_ROOT^(2^(28-k))."""
    n = len(coeffs)
    assert n & (n - 1) == 0 and n > 0
    return _ntt(coeffs, PRIMITIVE_ROOTS_OF_UNITY[n.bit_length() - 1])


def ifft(evals: Sequence[int]) -> List[int]:
    """ark-poly ``.ifft`` (call site polynomial.rs:131-135): inverse, with 1/n."""
    n = len(evals)
    assert n & (n - 1) == 0 and n > 0
    winv = fr_inv(PRIMITIVE_ROOTS_OF_UNITY[n.bit_length() - 1])
    ninv = fr_inv(n % R)
    return [v * ninv % R for v in _ntt(evals, winv)]


def g1_ifft(length: int, srs_g1: Sequence[Affine]) -> List[Affine]:
    """prover/src/kzg.rs:263-285: the same ark-poly ifft run over G1 points
    (twiddle multiplication = scalar multiplication), natural order, 1/n."""
    if length & (length - 1) != 0 or length == 0:
        raise KzgError("FFTError", "length provided is not a power of 2")
    n = length
    pts = _bitrev_permute([_to_jac(p) for p in srs_g1[:n]])
    winv = fr_inv(PRIMITIVE_ROOTS_OF_UNITY[n.bit_length() - 1])
    size = 2
    while size <= n:
        wl = pow(winv, n // size, R)
        half = size // 2
        tw = [1] * half
        for i in range(1, half):
            tw[i] = tw[i - 1] * wl % R
        for s in range(0, n, size):
            for j in range(half):
                u = pts[s + j]
                v = _jmul(pts[s + j + half], tw[j]) if tw[j] != 1 else pts[s + j + half]
                pts[s + j] = _jadd(u, v)
                pts[s + j + half] = _jadd(u, (v[0], (-v[1]) % P, v[2]))
        size *= 2
    ninv = fr_inv(n % R)
    pts = [_jmul(p, ninv) for p in pts]
    return _batch_to_aff(pts)


# ----------------------------------------------------------------------------
# Fiat-Shamir, evaluation, quotient (primitives/src/helpers.rs, prover/src/kzg.rs)
# ----------------------------------------------------------------------------
def compute_challenge(blob: Blob, commitment: Affine) -> int:
    """primitives/src/helpers.rs:411-472."""
    validate_g1_point(commitment)
    poly = blob.to_polynomial_eval_form()
    n = len(poly)
    buf = (
        FIAT_SHAMIR_PROTOCOL_DOMAIN
        + usize_to_be_bytes(n)
        + to_byte_array(poly.evaluations, n * 32)
        + g1_serialize_compressed(commitment)
    )
    assert len(buf) == 24 + 8 + 32 * n + 32
    return hash_to_field_element(buf)


def evaluate_polynomial_in_evaluation_form(poly: PolynomialEvalForm, z: int) -> int:
    """primitives/src/helpers.rs:475-535."""
    roots = calculate_roots_of_unity(poly.len_underlying_blob_bytes)
    if len(poly) != len(roots):
        raise KzgError("InvalidInputLength")
    width = len(poly)
    z %= R
    inv_width = fr_inv(width % R)
    for i, w in enumerate(roots):
        if w == z:
            return poly.evaluations[i]
    total = 0
    for f, w in zip(poly.evaluations, roots):
        total = (total + f * w % R * fr_inv((z - w) % R)) % R
    r = (pow(z, width, R) - 1) % R
    return total * r % R * inv_width % R


def compute_quotient_eval_on_domain(roots: Sequence[int], z: int, evals: Sequence[int], value: int) -> int:
    """prover/src/kzg.rs:237-260."""
    q = 0
    for i, w in enumerate(roots):
        if w == z:
            continue
        fi = (evals[i] - value) % R
        num = fi * w % R
        den = (z - w) % R * z % R
        q = (q + num * fr_inv(den)) % R
    return q


class KZG:
    """prover/src/kzg.rs:25-309, over ``srs`` = list of affine G1 points."""

    def __init__(self):
        self.expanded_roots_of_unity: List[int] = []

    def calculate_and_store_roots_of_unity(self, length_of_data_after_padding: int) -> None:
        self.expanded_roots_of_unity = calculate_roots_of_unity(length_of_data_after_padding)

    def get_nth_root_of_unity(self, i: int) -> Optional[int]:
        return self.expanded_roots_of_unity[i] if 0 <= i < len(self.expanded_roots_of_unity) else None

    def g1_ifft(self, length: int, srs: Sequence[Affine]) -> List[Affine]:
        return g1_ifft(length, srs)

    def commit_eval_form(self, poly: PolynomialEvalForm, srs: Sequence[Affine], literal: bool = False) -> Affine:
        """kzg.rs:84-104.  ``literal=True`` follows the reference literally
        (G1 IFFT of the SRS, then MSM over the Lagrange bases); the default uses
        the identity MSM(IFFT_G1(SRS), f) == MSM(SRS, IFFT_Fr(f)) (same group
        element; equality is asserted in tests/test_oracle.py)."""
        if len(poly) > len(srs):
            raise KzgError(
                "SrsCapacityExceeded", f"polynomial_len={len(poly)} srs_len={len(srs)}"
            )
        if literal:
            bases = g1_ifft(len(poly), srs)
            return msm(bases, poly.evaluations)
        return msm(srs[: len(poly)], ifft(poly.evaluations))

    def commit_coeff_form(self, poly: PolynomialCoeffForm, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:107-125."""
        if len(poly) > len(srs):
            raise KzgError("SerializationError", "polynomial length is not correct")
        return msm(srs[: len(poly)], poly.coeffs)

    def commit_blob(self, blob: Blob, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:182-185."""
        return self.commit_eval_form(blob.to_polynomial_eval_form(), srs)

    def compute_quotient(self, poly: PolynomialEvalForm, z: int) -> Tuple[int, List[int]]:
        """The y / quotient part of compute_proof_impl (kzg.rs:141-174)."""
        if len(poly) != len(self.expanded_roots_of_unity):
            raise KzgError("GenericError", "inconsistent length between blob and root of unities")
        z %= R
        evals = poly.evaluations
        y = evaluate_polynomial_in_evaluation_form(poly, z)
        roots = self.expanded_roots_of_unity
        q = []
        for i in range(len(roots)):
            den = (roots[i] - z) % R
            if den == 0:
                q.append(compute_quotient_eval_on_domain(roots, z, evals, y))
            else:
                q.append((evals[i] - y) % R * fr_inv(den) % R)
        return y, q

    def compute_proof_impl(self, poly: PolynomialEvalForm, z: int, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:128-178."""
        _, q = self.compute_quotient(poly, z)
        return self.commit_eval_form(PolynomialEvalForm(q), srs)

    def compute_proof(self, poly: PolynomialEvalForm, z: int, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:215-234."""
        if len(poly) != len(self.expanded_roots_of_unity):
            raise KzgError("GenericError", "inconsistent length between blob and root of unities")
        return self.compute_proof_impl(poly, z, srs)

    def compute_proof_with_known_z_fr_index(self, poly: PolynomialEvalForm, index: int, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:187-207."""
        z = self.get_nth_root_of_unity(index)
        if z is None:
            raise KzgError("GenericError", "Root of unity not found")
        return self.compute_proof(poly, z, srs)

    def compute_blob_proof(self, blob: Blob, commitment: Affine, srs: Sequence[Affine]) -> Affine:
        """kzg.rs:288-309."""
        validate_g1_point(commitment)
        poly = blob.to_polynomial_eval_form()
        z = compute_challenge(blob, commitment)
        return self.compute_proof_impl(poly, z, srs)


# ----------------------------------------------------------------------------
# Batch verification: everything up to (not including) the pairing
# (verifier/src/batch.rs, primitives/src/helpers.rs:613-662)
# ----------------------------------------------------------------------------
def compute_powers(base: int, count: int) -> List[int]:
    """primitives/src/helpers.rs:298-314."""
    out, cur = [], 1
    for _ in range(count):
        out.append(cur)
        cur = cur * base % R
    return out


def compute_challenges_and_evaluate_polynomial(blobs: Sequence[Blob], commitments: Sequence[Affine]):
    """primitives/src/helpers.rs:613-662."""
    if len(blobs) != len(commitments) and len(blobs) != 0:
        raise KzgError("GenericError", "length's of the input are not the same or is empty")
    zs, ys = [], []
    for blob, c in zip(blobs, commitments):
        poly = blob.to_polynomial_eval_form()
        z = compute_challenge(blob, c)
        ys.append(evaluate_polynomial_in_evaluation_form(poly, z))
        zs.append(z)
    return zs, ys


def compute_r_powers(commitments, zs, ys, proofs, blob_lengths: Sequence[int]) -> List[int]:
    """verifier/src/batch.rs:76-168.  Bytes 24..32 of the transcript are never
    written (stay zero); the count goes at 32..40."""
    n = len(commitments)
    buf = bytearray(40 + n * (32 * 4 + 8))
    buf[0:24] = RANDOM_CHALLENGE_KZG_BATCH_DOMAIN
    buf[32:40] = usize_to_be_bytes(n)
    off = 40
    for ln in blob_lengths:
        buf[off : off + 8] = int(ln).to_bytes(8, "big")
        off += 8
    for i in range(n):
        buf[off : off + 32] = g1_serialize_compressed(commitments[i]); off += 32
        buf[off : off + 32] = (zs[i] % R).to_bytes(32, "big"); off += 32
        buf[off : off + 32] = (ys[i] % R).to_bytes(32, "big"); off += 32
        buf[off : off + 32] = g1_serialize_compressed(proofs[i]); off += 32
    assert off == len(buf)
    r = hash_to_field_element(bytes(buf))
    return compute_powers(r, n)


def verify_kzg_proof_batch_rlc(commitments, zs, ys, proofs, blob_lengths) -> Tuple[Affine, Affine]:
    """verifier/src/batch.rs:185-249: returns (proof_lincomb, rhs_g1), the two
    G1 inputs of the final pairing check (batch.rs:253-254, out of scope)."""
    if not (len(commitments) == len(zs) == len(ys) == len(proofs)):
        raise KzgError("GenericError", "length's of the input are not the same")
    for c in commitments:
        validate_g1_point(c)
    for p in proofs:
        validate_g1_point(p)
    n = len(commitments)
    r_powers = compute_r_powers(commitments, zs, ys, proofs, blob_lengths)
    proof_lincomb = g1_lincomb(proofs, r_powers)
    c_minus_y, r_times_z = [], []
    for i in range(n):
        ys_enc = g1_mul(G1_GEN, ys[i])
        c_minus_y.append(g1_add(commitments[i], g1_neg(ys_enc)))
        r_times_z.append(r_powers[i] * zs[i] % R)
    proof_z_lincomb = g1_lincomb(proofs, r_times_z)
    c_minus_y_lincomb = g1_lincomb(c_minus_y, r_powers)
    rhs = g1_add(c_minus_y_lincomb, proof_z_lincomb)
    return proof_lincomb, rhs


def verify_blob_kzg_proof_batch_rlc(blobs: Sequence[Blob], commitments, proofs) -> Tuple[Affine, Affine]:
    """verifier/src/batch.rs:16-69 up to the pairing."""
    if not (len(commitments) == len(blobs) and len(proofs) == len(blobs)):
        raise KzgError("GenericError", "length's of the input are not the same")
    for c in commitments:
        validate_g1_point(c)
    for p in proofs:
        validate_g1_point(p)
    zs, ys = compute_challenges_and_evaluate_polynomial(blobs, commitments)
    lens = [len(b.to_polynomial_eval_form()) for b in blobs]
    return verify_kzg_proof_batch_rlc(commitments, zs, ys, proofs, lens)


# ----------------------------------------------------------------------------
# Synthetic SRS: SRS_i = tau^i * G with a known tau
# ----------------------------------------------------------------------------
SYNTH_TAU = int.from_bytes(hashlib.sha256(b"kzg-bn254-b200/tau/v1").digest(), "big") % R


def synthetic_srs(n: int, tau: int = SYNTH_TAU) -> List[Affine]:
    """tau^i * G for i < n (fixed-base comb over the generator, batch-normalised)."""
    # 4-bit fixed-base table of G: tbl[w][d] = d * 16^w * G
    tbl = []
    base = _to_jac(G1_GEN)
    for _ in range(64):
        row = [_JINF]
        for d in range(1, 16):
            row.append(_jadd(row[-1], base))
        rown = _batch_to_aff(row)
        tbl.append(rown)
        base = _jadd(row[15], base)
    out = []
    t = 1
    for _ in range(n):
        acc = _JINF
        k = t
        w = 0
        while k:
            d = k & 15
            if d:
                acc = _jmadd(acc, tbl[w][d])
            k >>= 4
            w += 1
        out.append(acc)
        t = t * tau % R
    return _batch_to_aff(out)


def tau_trick_msm(scalars: Sequence[int], tau: int = SYNTH_TAU) -> Affine:
    """MSM(SRS, s) for SRS_i = tau^i G equals (sum s_i tau^i) G."""
    acc = 0
    for s in reversed(scalars):
        acc = (acc * tau + s) % R
    return g1_mul(G1_GEN, acc)


# ----------------------------------------------------------------------------
# Pairing (TEST ORACLE ONLY; the product leaves the pairing in the reference's code).
# ark-bn254's optimal ate pairing restated on the polynomial tower
#   Fq2 = Fq[i]/(i^2 + 1),  Fq12 = Fq[w]/(w^12 - 18 w^6 + 82)   (w^6 = 9 + i),
# G2 on the sextic twist y^2 = x^3 + 3/(9 + i), mapped into Fq12 by (x, y) -> (x w^2, y w^3).
# Used by verify_proof / verify_blob_kzg_proof / verify_blob_kzg_proof_batch
# (verifier/src/verify.rs:10-115, verifier/src/batch.rs:16-69, helpers.rs:392-398) with an
# INJECTABLE [tau]G2 so config 1 verifies against the synthetic SRS; the default is the
# reference's mainnet constant G2_TAU (primitives/src/consts.rs:55-64).
# Anchors: G2_TAU and the generator lie on the twist and have order r; bilinearity in both
# arguments (tests/test_oracle.py).
# ----------------------------------------------------------------------------
Fq2 = Tuple[int, int]
G2Affine = Optional[Tuple[Fq2, Fq2]]

G2_GEN: G2Affine = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)
G2_TAU: G2Affine = (  # consts.rs:55-64
    (19394299006376106554626551996044114846855237028623244664226757033024550999552,
     10478571113809844268398751534081669357808742555529167819607714577862447855483),
    (9205262336805673656533560220225620941045451042642528799409071118332922267006,
     10552783866161062341197740743287753408530108186218052255509661543860392060676),
)
ATE_LOOP_COUNT = 29793968203157093288  # 6x + 2, x = 4965661367192848881


def fq2_add(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def fq2_sub(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def fq2_mul(a: Fq2, b: Fq2) -> Fq2:
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def fq2_inv(a: Fq2) -> Fq2:
    d = fq_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * d % P, -a[1] * d % P)


_TWIST_B: Fq2 = fq2_mul((3, 0), fq2_inv((9, 1)))  # 3 / (9 + i)


def g2_is_on_curve(q: G2Affine) -> bool:
    if q is None:
        return True
    x, y = q
    return fq2_mul(y, y) == fq2_add(fq2_mul(fq2_mul(x, x), x), _TWIST_B)


def g2_neg(q: G2Affine) -> G2Affine:
    return None if q is None else (q[0], (-q[1][0] % P, -q[1][1] % P))


def g2_add(a: G2Affine, b: G2Affine) -> G2Affine:
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if y1 != y2 or y1 == (0, 0):
            return None
        x1x1 = fq2_mul(x1, x1)
        lam = fq2_mul(fq2_add(fq2_add(x1x1, x1x1), x1x1), fq2_inv(fq2_add(y1, y1)))
    else:
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), x1), x2)
    return (x3, fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1))


def g2_mul(q: G2Affine, k: int) -> G2Affine:
    k %= R
    acc: G2Affine = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, q)
    return acc


# ---- Fq12 as coefficient lists of length 12 over Fq, modulus w^12 - 18 w^6 + 82
def _f12_mul(a: List[int], b: List[int]) -> List[int]:
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    for k in range(22, 11, -1):  # w^k = 18 w^(k-6) - 82 w^(k-12)
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


_F12_ONE = [1] + [0] * 11


def _f12_pow(a: List[int], e: int) -> List[int]:
    out = _F12_ONE
    for bit in bin(e)[2:]:
        out = _f12_mul(out, out)
        if bit == "1":
            out = _f12_mul(out, a)
    return out


def _poly_deg(p: List[int]) -> int:
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _f12_inv(a: List[int]) -> List[int]:
    """Extended Euclid in Fq[w] against the modulus polynomial."""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, -18 % P, 0, 0, 0, 0, 0, 1]
    while _poly_deg(low):
        # r = high // low (polynomial quotient)
        dl, dh = _poly_deg(low), _poly_deg(high)
        temp = list(high)
        q = [0] * 13
        inv_lead = fq_inv(low[dl])
        for i in range(dh - dl, -1, -1):
            q[i] = temp[dl + i] * inv_lead % P
            for c in range(dl + 1):
                temp[c + i] = (temp[c + i] - q[i] * low[c]) % P
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * q[j]) % P
                new[i + j] = (new[i + j] - low[i] * q[j]) % P
        lm, low, hm, high = nm, new, lm, low
    inv0 = fq_inv(low[0])
    return [x * inv0 % P for x in lm[:12]]


def _twist(q: Tuple[Fq2, Fq2]) -> Tuple[List[int], List[int]]:
    """G2 (on the twist, over Fq2) -> E(Fq12): a + b i  ->  (a - 9 b) + b w^6, then (x w^2, y w^3)."""
    (x0, x1), (y0, y1) = q
    nx = [0] * 12
    ny = [0] * 12
    nx[2], nx[8] = (x0 - 9 * x1) % P, x1
    ny[3], ny[9] = (y0 - 9 * y1) % P, y1
    return nx, ny


def _f12_sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def _f12_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def _line(p1, p2, t):
    """Line through p1, p2 (points of E(Fq12)) evaluated at t."""
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if x1 != x2:
        m = _f12_mul(_f12_sub(y2, y1), _f12_inv(_f12_sub(x2, x1)))
    elif y1 == y2:
        x1x1 = _f12_mul(x1, x1)
        m = _f12_mul([3 * c % P for c in x1x1], _f12_inv([2 * c % P for c in y1]))
    else:
        return _f12_sub(xt, x1)
    return _f12_sub(_f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))


def _e12_add(p1, p2):
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2 and y1 == y2:
        x1x1 = _f12_mul(x1, x1)
        m = _f12_mul([3 * c % P for c in x1x1], _f12_inv([2 * c % P for c in y1]))
    else:
        m = _f12_mul(_f12_sub(y2, y1), _f12_inv(_f12_sub(x2, x1)))
    x3 = _f12_sub(_f12_sub(_f12_mul(m, m), x1), x2)
    return x3, _f12_sub(_f12_mul(m, _f12_sub(x1, x3)), y1)


def miller_loop(q: G2Affine, p: Affine) -> List[int]:
    """Optimal ate Miller loop f_{6x+2,Q}(P) with the two Frobenius lines; no final exponentiation."""
    if q is None or p is None:
        return _F12_ONE
    Q = _twist(q)
    Pt = ([p[0]] + [0] * 11, [p[1]] + [0] * 11)
    Rp = Q
    f = _F12_ONE
    for i in range(ATE_LOOP_COUNT.bit_length() - 2, -1, -1):
        f = _f12_mul(_f12_mul(f, f), _line(Rp, Rp, Pt))
        Rp = _e12_add(Rp, Rp)
        if (ATE_LOOP_COUNT >> i) & 1:
            f = _f12_mul(f, _line(Rp, Q, Pt))
            Rp = _e12_add(Rp, Q)
    Q1 = (_f12_pow(Q[0], P), _f12_pow(Q[1], P))
    nQ2 = (_f12_pow(Q1[0], P), [-c % P for c in _f12_pow(Q1[1], P)])
    f = _f12_mul(f, _line(Rp, Q1, Pt))
    Rp = _e12_add(Rp, Q1)
    f = _f12_mul(f, _line(Rp, nQ2, Pt))
    return f


def final_exponentiation(f: List[int]) -> List[int]:
    return _f12_pow(f, (P**12 - 1) // R)


def pairing(q: G2Affine, p: Affine) -> List[int]:
    return final_exponentiation(miller_loop(q, p))


def pairings_verify(a1: Affine, a2: G2Affine, b1: Affine, b2: G2Affine) -> bool:
    """helpers.rs:392-398: e(a1, a2) == e(b1, b2), as one product e(a1, a2) * e(-b1, b2) == 1."""
    f = _f12_mul(miller_loop(a2, a1), miller_loop(b2, g1_neg(b1)))
    return final_exponentiation(f) == _F12_ONE


def verify_proof(commitment: Affine, proof: Affine, value_fr: int, z_fr: int, g2_tau: G2Affine = G2_TAU) -> bool:
    """verifier/src/verify.rs:10-73: e(C - y G1, G2) == e(proof, [tau - z] G2)."""
    validate_g1_point(commitment)
    validate_g1_point(proof)
    if not g2_is_on_curve(g2_tau):
        raise KzgError("NotOnCurveError", "Invalid trusted setup: G2_TAU not on curve")
    commit_minus_value = g1_add(commitment, g1_neg(g1_mul(G1_GEN, value_fr % R)))
    x_minus_z = g2_add(g2_tau, g2_neg(g2_mul(G2_GEN, z_fr % R)))
    if x_minus_z is None:
        raise KzgError("GenericError", "Evaluation point equals trusted setup secret")
    return pairings_verify(commit_minus_value, G2_GEN, proof, x_minus_z)


def verify_blob_kzg_proof(blob: Blob, commitment: Affine, proof: Affine, g2_tau: G2Affine = G2_TAU) -> bool:
    """verifier/src/verify.rs:77-115: challenge, evaluation, then verify_proof."""
    validate_g1_point(commitment)
    validate_g1_point(proof)
    poly = blob.to_polynomial_eval_form()
    z = compute_challenge(blob, commitment)
    y = evaluate_polynomial_in_evaluation_form(poly, z)
    return verify_proof(commitment, proof, y, z, g2_tau)


def verify_blob_kzg_proof_batch(blobs: Sequence[Blob], commitments, proofs, g2_tau: G2Affine = G2_TAU) -> bool:
    """verifier/src/batch.rs:16-69,170-255: the RLC front half, then e(lhs, [tau]G2) == e(rhs, G2)."""
    lhs, rhs = verify_blob_kzg_proof_batch_rlc(blobs, commitments, proofs)
    return pairings_verify(lhs, g2_tau, rhs, G2_GEN)
