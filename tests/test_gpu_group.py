"""kzgb_group (multi-GPU behind the C ABI): the sharded calls must return the bytes of the single-context calls.

The group paths are exercised on ONE GPU by listing device 0 twice or three times (independent contexts on one
GPU: the same code path -- SRS replication device to device, one host thread per member, shares by blob / by
point range, partial sums added on the host); with 2+ GPUs visible the same tests also run over distinct devices
(reference call sites: prover/src/kzg.rs:107-125 commit_coeff_form, :182-185 commit_blob, :288-309 compute_blob_proof)."""
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


def _device_sets():
    import torch

    sets = [[0, 0], [0, 0, 0]]
    n = torch.cuda.device_count()
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append(list(range(n)))
    return sets


@pytest.mark.parametrize("devices", _device_sets() if __import__("torch").cuda.is_available() else [[0, 0]])
def test_group_blob_batch_equals_single_context(pkg, devices):
    """A ragged batch (mixed sizes, a short last chunk, tiny blobs) cut across the members: commitment and proof
    bytes equal the single-context call and the CPU oracle on the reference's own SRS file."""
    rnd = random.Random(len(devices) * 7 + devices[-1])
    raws = [rnd.randbytes(k) for k in (31 * 64, 31 * 33, 31, 1, 31 * 200, 31 * 128, 500, 31 * 64, 31 * 7, 31 * 256)]
    blobs = [pkg.Blob.from_raw_data(r) for r in raws]
    eng = pkg.Engine(0)
    srs = pkg.SRS.from_gnark_bytes(g.g1_point_bytes(), engine=eng)
    cs1, ps1 = pkg.KZG.commit_and_prove_blobs(blobs, srs)
    grp = pkg.Group(devices)
    assert len(grp) == len(devices)
    grp.load_srs_gnark_bytes(g.g1_point_bytes())
    cs, ps = grp.commit_and_prove_blobs(blobs)
    assert cs == cs1 and ps == ps1
    pts = g.srs_points_string()
    okzg = o.KZG()
    for i in (3, 6, 8):  # small ones: the pure-Python oracle MSM takes a second per few hundred points
        ob = o.Blob.from_raw_data(raws[i])
        c = okzg.commit_blob(ob, pts)
        assert cs[i] == o.g1_serialize_compressed(c)
        okzg.calculate_and_store_roots_of_unity(len(ob))
        assert ps[i] == o.g1_serialize_compressed(okzg.compute_blob_proof(ob, c, pts))
    # fewer blobs than members, and none
    cs, ps = grp.commit_and_prove_blobs(blobs[:1])
    assert cs == cs1[:1] and ps == ps1[:1]
    assert grp.commit_and_prove_blobs([]) == ([], [])
    grp.close()


@pytest.mark.parametrize("devices", _device_sets() if __import__("torch").cuda.is_available() else [[0, 0]])
def test_group_msm_by_point_range_closed_form(pkg, devices):
    """2^16-point commit_coeff_form cut by point range, with and without the per-range window tables:
    = (sum s_i tau^i) G and = the single-context result; an n that does not divide evenly; n = 0."""
    n = 1 << 16
    grp = pkg.Group(devices)
    grp.load_srs_synthetic(n, TAU)
    rnd = random.Random(5)
    sc = [rnd.randrange(o.R) for _ in range(n)]

    def closed(v):
        acc = 0
        for s in reversed(v):
            acc = (acc * TAU + s) % o.R
        return o.g1_mul(o.G1_GEN, acc)

    want = closed(sc)
    assert grp.commit_coeff_form(pkg.PolynomialCoeffForm(sc)) == want  # tables built lazily per member ([0, share end))
    grp.precompute_ranges(n, 0)
    assert grp.commit_coeff_form(pkg.PolynomialCoeffForm(sc)) == want
    cb, cw, ct = C.c_int(0), C.c_int(0), C.c_size_t(0)
    pkg.lib.kzgb_msm_config(grp.member(len(devices) - 1).h, C.byref(cb), C.byref(cw), C.byref(ct))
    assert 0 < ct.value <= -(-n // len(devices)) + 32  # the last member's table covers its share only
    m = 40000 + 17
    out = C.create_string_buffer(64)
    inf = C.c_uint8(0)
    grp.check(pkg.lib.kzgb_group_msm_srs(grp.h, pkg.fr_to_mont_bytes(sc[:m]), m, out, C.byref(inf)))
    assert pkg.g1_from_abi(out.raw, bytes([inf.value]))[0] == closed(sc[:m])
    grp.check(pkg.lib.kzgb_group_msm_srs(grp.h, pkg.fr_to_mont_bytes(sc[:1]), 0, out, C.byref(inf)))
    assert inf.value == 1
    with pytest.raises(pkg.KzgError) as e:
        grp.check(pkg.lib.kzgb_group_msm_srs(grp.h, pkg.fr_to_mont_bytes(sc[:1]), n + 1, out, C.byref(inf)))
    assert e.value.variant == "SerializationError"
    grp.close()


def test_group_16mib_blobs_two_members(pkg):
    """BASELINE configs[1] through the group: four 2^19-Fr blobs over two members = the single-context bytes
    and the closed form p(tau) G of the first."""
    import numpy as np

    n = 1 << 19
    devices = [0, 1] if __import__("torch").cuda.device_count() >= 2 else [0, 0]
    rng = np.random.default_rng(7)
    datas = []
    for _ in range(4):
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 0] = 0
        datas.append(a.tobytes())
    blobs = [pkg.Blob.from_unchecked(d) for d in datas]
    eng = pkg.Engine(0)
    srs = pkg.SRS.synthetic(n, TAU, engine=eng)
    cs1, ps1 = pkg.KZG.commit_and_prove_blobs(blobs, srs)
    eng.close()
    grp = pkg.Group(devices)
    grp.load_srs_synthetic(n, TAU)
    cs, ps = grp.commit_and_prove_blobs(blobs)
    assert cs == cs1 and ps == ps1
    grp.close()
