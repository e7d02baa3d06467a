"""Loaders for the reference's own fixtures (copied verbatim under tests/golden/)."""
import functools
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _path(name):
    return os.path.join(GOLDEN, name)


@functools.lru_cache(None)
def g1_point_bytes() -> bytes:
    """prover/tests/test-files/g1.point: 3000 x 32 B gnark-BE compressed G1."""
    return open(_path("g1.point"), "rb").read()


@functools.lru_cache(None)
def srs_points_string():
    """prover/tests/test-files/srs.g1.points.string: affine coords of the same points."""
    out = []
    for line in open(_path("srs.g1.points.string")):
        line = line.strip()
        if line:
            x, y = line.split(",")
            out.append((int(x), int(y)))
    return out


@functools.lru_cache(None)
def lagrange_srs_64():
    """prover/tests/test-files/lagrangeG1SRS.txt: g1_ifft(64) of g1.point[..64]."""
    out = []
    for line in open(_path("lagrangeG1SRS.txt")):
        line = line.strip()
        if line:
            x, y = line.split(",")
            out.append((int(x), int(y)))
    return out


@functools.lru_cache(None)
def proof_eq_input():
    """prover/tests/test-files/kzg.proof.eq.input: rows (index, proof.x, proof.y)."""
    out = []
    for line in open(_path("kzg.proof.eq.input")):
        line = line.strip()
        if line:
            i, x, y = line.split(",")
            out.append((int(i), int(x), int(y)))
    return out


@functools.lru_cache(None)
def blobs_txt() -> bytes:
    return open(_path("blobs.txt"), "rb").read()


@functools.lru_cache(None)
def blobs_from_fr():
    return [int(l) for l in open(_path("blobs-from-fr.txt")) if l.strip()]


@functools.lru_cache(None)
def gettysburg() -> bytes:
    return open(_path("gettysburg.txt"), "rb").read()
