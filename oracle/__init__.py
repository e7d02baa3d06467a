"""CPU oracle for the rust-kzg-bn254 commitment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or
execute anything in this directory; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and there
only as the checker / the timed CPU baseline.

Parity status: pinned against the reference's own fixtures (copied under
``tests/golden/``): ``srs.g1.points.string`` (3000 decompressed points),
``lagrangeG1SRS.txt`` (g1_ifft(64)), ``kzg.proof.eq.input`` (40 proofs),
``blobs.txt``/``blobs-from-fr.txt`` (bytes -> Fr).  NOT pinned by any reference
fixture ("parity unpinned"): the arkworks ``serialize_compressed`` byte layout
that feeds the Fiat-Shamir transcripts (restated from the published arkworks 0.5
format), hence the challenge ``z`` of ``compute_blob_proof`` and the RLC scalar
``r``.  The verifier's pairing (``verify_proof`` / ``verify_blob_kzg_proof`` / the
final check of the batch verifier) is restated here too -- only so the tests can run
BASELINE config 1 end to end and check GPU-made proofs against the pairing equation
with an injected [tau]G2; it is anchored on the reference's ``G2_TAU`` constant lying
on the twist with order r and on bilinearity, and the product never computes a pairing.
The reference itself (Rust + un-vendored arkworks 0.5) cannot be built
in this image, so there is no ``oracle/_ref``.
"""
