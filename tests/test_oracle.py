"""Pins the CPU oracle against every fixture the reference holds for this path
 -- runs on CPU."""
import hashlib

import pytest

import golden_data as g
from oracle import bn254 as o

# primitives/src/consts.rs:22-52, first and a few middle entries (literal check, cf. helpers_test.rs:587-628)
ROOTS_LITERALS = {
    0: 1,
    1: 21888242871839275222246405745257275088548364400416034343698204186575808495616,
    2: 21888242871839275217838484774961031246007050428528088939761107053157389710902,
    6: 9088801421649573101014283686030284801466796108869023335878462724291607593530,
    12: 4158865282786404163413953114870269622875596290766033564087307867933865333818,
    19: 15549849457946371566896172786938980432421851627449396898353380550861104573629,
    28: 19103219067921713944291392827692070036145651957329286315305642004821462161904,
}


@pytest.fixture(scope="module")
def ref_srs():
    raw = g.g1_point_bytes()
    return [o.read_g1_point_from_bytes_be(raw[32 * i : 32 * i + 32]) for i in range(3000)]


def test_roots_table_literals():
    for k, v in ROOTS_LITERALS.items():
        assert o.PRIMITIVE_ROOTS_OF_UNITY[k] == v
    assert pow(o.PRIMITIVE_ROOTS_OF_UNITY[28], 1 << 27, o.R) != 1


def test_srs_decompression_kat(ref_srs):
    assert ref_srs == g.srs_points_string()
    raw = g.g1_point_bytes()
    assert all(o.g1_to_gnark_be(p) == raw[32 * i : 32 * i + 32] for i, p in enumerate(ref_srs))
    assert raw[:32] == bytes([0x80]) + bytes(30) + b"\x01"  # generator (1, 2)


def test_bytes_to_fr_kat():
    assert o.to_fr_array(g.blobs_txt()) == g.blobs_from_fr()
    o.Blob(g.blobs_txt())  # canonical: Blob::new accepts it


def test_pad_payload_literals():
    # primitives/tests/helpers_test.rs:468-476
    assert o.pad_payload(b"hi") == bytes([0, 104, 105] + [0] * 29)
    assert o.remove_internal_padding(o.pad_payload(b"hi")) == bytes([104, 105] + [0] * 29)
    padded = o.pad_payload(g.gettysburg())
    assert len(o.remove_internal_padding(padded)) == 1488
    assert len(padded) == 1536


def test_blob_validation():
    # primitives/tests/blob_test.rs:22-54
    with pytest.raises(o.KzgError):
        o.Blob(bytes(62))
    with pytest.raises(o.KzgError):
        o.Blob(b"\xff" * 32)
    o.Blob(bytes(64))


def test_g1_ifft_kat(ref_srs):
    assert o.g1_ifft(64, ref_srs) == g.lagrange_srs_64()
    with pytest.raises(o.KzgError) as e:
        o.g1_ifft(15, ref_srs)
    assert "length provided is not a power of 2" in str(e.value)  # prover/tests/kzg_test.rs:151-156


def test_proof_kat_and_commit_equivalence(ref_srs):
    blob = o.Blob.from_raw_data(g.gettysburg())
    poly = blob.to_polynomial_eval_form()
    assert len(poly) == 64
    k = o.KZG()
    k.calculate_and_store_roots_of_unity(len(blob))
    c = k.commit_blob(blob, ref_srs)
    assert c == k.commit_eval_form(poly, ref_srs, literal=True)  # MSM(IFFT_G1(SRS), f) == MSM(SRS, IFFT_Fr(f))
    assert c == k.commit_coeff_form(poly.to_coeff_form(), ref_srs)  # prover/tests/kzg_test.rs:57-89
    assert c == o.msm(g.lagrange_srs_64(), poly.evaluations)
    assert o.g1_to_gnark_be(c).hex() == "868bf472ebc0e26c297f8a9257c3f42a38af1e4612b60f6ac64d57dc272d50b1"
    for idx, x, y in g.proof_eq_input()[:12]:
        assert k.compute_proof_with_known_z_fr_index(poly, idx, ref_srs) == (x, y)


def test_eval_at_roots_and_fft_roundtrip():
    # prover/tests/kzg_test.rs:31-55, primitives/tests/polynomial_test.rs:47-64
    blob = o.Blob.from_raw_data(g.gettysburg())
    poly = blob.to_polynomial_eval_form()
    roots = o.calculate_roots_of_unity(len(blob))
    for i in (0, 1, 17, 63):
        assert o.evaluate_polynomial_in_evaluation_form(poly, roots[i]) == poly.evaluations[i]
    assert poly.to_coeff_form().to_eval_form().evaluations == poly.evaluations
    assert poly.to_bytes_be() == blob.data()


def test_synthetic_tau_flow():
    """SURVEY.md appendix D.2 cross-check vectors (tau = SHA-256("kzg-bn254-b200/tau/v1") mod r)."""
    assert o.SYNTH_TAU == 2480609854371098259468018140899271569021640719453669963486734696239309822386
    srs = o.synthetic_srs(64)
    blob = o.Blob.from_raw_data(g.gettysburg())
    k = o.KZG()
    k.calculate_and_store_roots_of_unity(len(blob))
    c = k.commit_blob(blob, srs)
    assert c == o.tau_trick_msm(o.ifft(blob.to_polynomial_eval_form().evaluations))
    assert o.g1_serialize_compressed(c).hex() == "92dc1c8eb53aad328198593c34535f8d2e74ca5357d41fbe2ab67b735a2b80a6"
    z = o.compute_challenge(blob, c)
    assert z == 16364512820228716287787019285793812063829111171541090311386264826634667274589
    pi = k.compute_blob_proof(blob, c, srs)
    assert o.g1_serialize_compressed(pi).hex() == "48d65224dd56b35a4374cded5cdf6341684b314a042cfce7a10e78ff09632520"
    # proof == ((p(tau) - y) / (tau - z)) * G
    y = o.evaluate_polynomial_in_evaluation_form(blob.to_polynomial_eval_form(), z)
    ptau = 0
    for cf in reversed(o.ifft(blob.to_polynomial_eval_form().evaluations)):
        ptau = (ptau * o.SYNTH_TAU + cf) % o.R
    assert pi == o.g1_mul(o.G1_GEN, (ptau - y) * o.fr_inv((o.SYNTH_TAU - z) % o.R) % o.R)


def test_rlc_transcript_layout():
    blob = o.Blob.from_raw_data(b"rlc")
    srs = o.synthetic_srs(4)
    k = o.KZG()
    k.calculate_and_store_roots_of_unity(len(blob))
    c = k.commit_blob(blob, srs)
    pi = k.compute_blob_proof(blob, c, srs)
    zs, ys = o.compute_challenges_and_evaluate_polynomial([blob], [c])
    powers = o.compute_r_powers([c], zs, ys, [pi], [1])
    buf = (
        b"EIGENDA_RCKZGBATCH___V1_" + bytes(8) + (1).to_bytes(8, "big") + (1).to_bytes(8, "big")
        + o.g1_serialize_compressed(c) + zs[0].to_bytes(32, "big") + ys[0].to_bytes(32, "big") + o.g1_serialize_compressed(pi)
    )
    assert powers == [1]
    assert len(buf) == 40 + 136
    r = int.from_bytes(hashlib.sha256(buf).digest(), "big") % o.R
    assert o.compute_r_powers([c, c], zs * 2, ys * 2, [pi, pi], [1, 1])[0] == 1
    assert r != 0


def test_pairing_constants_and_bilinearity():
    """The oracle's pairing (test infrastructure for config 1): the reference's G2_TAU (consts.rs:55-64) and
    the generator lie on the twist with order r; e is bilinear in both arguments and has order r."""
    for q in (o.G2_GEN, o.G2_TAU):
        assert o.g2_is_on_curve(q)
        assert o.g2_mul(q, o.R - 1) == o.g2_neg(q)
    assert not o.g2_is_on_curve(((1, 2), (3, 4)))
    e = o.pairing(o.G2_GEN, o.G1_GEN)
    assert e != o._F12_ONE and o._f12_pow(e, o.R) == o._F12_ONE
    a, b = 0x1234567890ABCDEF1234567, 0xFEDCBA09876543210FEDCBA0987
    assert o.pairing(o.G2_GEN, o.g1_mul(o.G1_GEN, a)) == o._f12_pow(e, a)
    assert o.pairing(o.g2_mul(o.G2_GEN, b), o.G1_GEN) == o._f12_pow(e, b)
    assert o.pairing(o.g2_mul(o.G2_GEN, b), o.g1_mul(o.G1_GEN, a)) == o._f12_pow(e, a * b % o.R)
    assert o.pairing(None, o.G1_GEN) == o._F12_ONE and o.pairing(o.G2_GEN, None) == o._F12_ONE
    x = o._f12_pow(e, 12345)
    assert o._f12_mul(x, o._f12_inv(x)) == o._F12_ONE


def test_config1_prove_verify_on_the_oracle():
    """BASELINE configs[0] restated on the CPU oracle with the synthetic SRS: commit_blob + compute_blob_proof +
    verify_blob_kzg_proof (verifier/src/verify.rs:10-115) with [tau]G2 injected; tampering fails
    (verifier/tests/tests.rs:28-192); batch verification (batch.rs:16-69) true / false."""
    srs = o.synthetic_srs(64)
    g2_tau = o.g2_mul(o.G2_GEN, o.SYNTH_TAU)
    blobs = [o.Blob.from_raw_data(g.gettysburg()), o.Blob.from_raw_data(b"second blob " * 20), o.Blob.from_raw_data(bytes(31 * 5))]
    cs, ps = [], []
    for blob in blobs:
        k = o.KZG()
        k.calculate_and_store_roots_of_unity(len(blob))
        c = k.commit_blob(blob, srs)
        cs.append(c)
        ps.append(k.compute_blob_proof(blob, c, srs))
    assert o.verify_blob_kzg_proof(blobs[0], cs[0], ps[0], g2_tau)
    assert not o.verify_blob_kzg_proof(blobs[0], cs[0], ps[0])  # the mainnet G2_TAU is a different tau
    assert not o.verify_blob_kzg_proof(blobs[0], cs[0], ps[1], g2_tau)
    assert not o.verify_blob_kzg_proof(blobs[1], cs[0], ps[0], g2_tau)
    assert o.verify_blob_kzg_proof(blobs[2], cs[2], ps[2], g2_tau)  # zero blob: identity commitment and proof
    assert cs[2] is None and ps[2] is None
    with pytest.raises(o.KzgError) as e:
        o.verify_blob_kzg_proof(blobs[0], (1, 3), ps[0], g2_tau)
    assert e.value.variant == "NotOnCurveError"
    z = 987654321
    y = o.evaluate_polynomial_in_evaluation_form(blobs[0].to_polynomial_eval_form(), z)
    k = o.KZG()
    k.calculate_and_store_roots_of_unity(len(blobs[0]))
    pi = k.compute_proof(blobs[0].to_polynomial_eval_form(), z, srs)
    assert o.verify_proof(cs[0], pi, y, z, g2_tau)
    assert not o.verify_proof(cs[0], pi, (y + 1) % o.R, z, g2_tau)
    with pytest.raises(o.KzgError) as e:
        o.verify_proof(cs[0], pi, y, o.SYNTH_TAU, g2_tau)  # z == tau (verify.rs:55-59)
    assert e.value.msg == "Evaluation point equals trusted setup secret"
    assert o.verify_blob_kzg_proof_batch(blobs, cs, ps, g2_tau)
    assert not o.verify_blob_kzg_proof_batch(blobs, cs, [ps[1], ps[0], ps[2]], g2_tau)
