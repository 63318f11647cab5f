"""CPU: the host-only part of the library that the verifier uses -- BN254 pairing and G2 arithmetic
(csrc/pairing.cuh) -- and the verifier's algebra (plonkathon_b200/verifier.py).

Pins: (1) the ceremony file itself: e([tau]_1, [1]_2) == e([1]_1, [tau]_2) with [tau]_1 and [tau]_2 = X2 read from
the shipped .ptau (the assertion the reference keeps as a comment, setup.py:61); (2) bilinearity and G2 scalar
multiplication against the restated py_ecc group law (oracle/shims); (3) the three golden proofs, which the
reference's completed verifier accepted when the fixtures were made (tests/golden/make_golden.py:190-199),
are accepted, and tampered ones rejected.  For (3) the G1 linear combinations are done by the oracle's naive
double-and-add here (no GPU in this suite); tests/test_gpu_parity.py repeats it through the GPU MSM."""
import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))

import py_ecc.bn128 as b  # noqa: E402  (restated shim)

import plonkathon_b200 as pb  # noqa: E402
from oracle import plonk_oracle as O  # noqa: E402
from plonkathon_b200 import _lib, verifier  # noqa: E402
from tests.golden_io import PTAU_HEAD, load_json, pt  # noqa: E402

R = O.R_MOD


def g2_of(q):
    """shim FQ2 pair -> product G2 point"""
    return (pb.FQ2([int(c) for c in q[0].coeffs]), pb.FQ2([int(c) for c in q[1].coeffs]))


def i1(p):
    return None if p is None else (int(p[0]), int(p[1]))


def test_ceremony_consistency_from_ptau():
    osetup = O.Setup.from_file(PTAU_HEAD)
    x2 = (pb.FQ2(osetup.X2[0]), pb.FQ2(osetup.X2[1]))
    tau1, tau2 = osetup.powers_of_x[1], osetup.powers_of_x[2]
    assert pb.pairing_product_is_one([(tau1, pb.G2), (O.g1_neg(O.G1), x2)])
    assert pb.pairing_product_is_one([(tau2, pb.G2), (O.g1_neg(tau1), x2)])
    assert not pb.pairing_product_is_one([(tau2, pb.G2), (O.g1_neg(O.G1), x2)])
    # and the product's own .ptau reader carries the same X2 (FQ2 objects, compared like test.py:47)
    assert g2_of(b.G2) == pb.G2


def test_bilinearity_and_g2_arithmetic():
    a, c = 0x1F3A9C55D2E4B7A1908877665544332211, 987654321987654321
    aP = i1(b.multiply(b.G1, a))
    acP = i1(b.multiply(b.G1, a * c % R))
    cQ = pb.g2_mul(pb.G2, c)
    assert cQ == g2_of(b.multiply(b.G2, c))
    assert pb.g2_mul(pb.G2, -c) == g2_of(b.neg(b.multiply(b.G2, c)))
    assert pb.g2_mul(pb.G2, 0) is None and pb.g2_mul(pb.G2, R) is None
    assert pb.g2_add(cQ, pb.G2) == g2_of(b.multiply(b.G2, c + 1))
    assert pb.g2_add(cQ, cQ) == g2_of(b.multiply(b.G2, 2 * c))
    assert pb.g2_add(cQ, pb.g2_mul(pb.G2, -c)) is None and pb.g2_add(None, cQ) == cQ
    # e(aP, cQ) == e(acP, Q)
    assert pb.pairing_product_is_one([(aP, cQ), (O.g1_neg(acP), pb.G2)])
    assert not pb.pairing_product_is_one([(aP, cQ), (O.g1_neg(i1(b.multiply(b.G1, a * c + 1))), pb.G2)])
    # identity factors drop out; a lone non-degenerate pairing is not 1
    assert pb.pairing_product_is_one([(None, pb.G2), (aP, None)])
    assert not pb.pairing_product_is_one([(aP, pb.G2)])
    # three-term product: e(P, Q)^(a+c) * e(-(a+c)P, Q) == 1
    assert pb.pairing_product_is_one([(aP, pb.G2), (i1(b.multiply(b.G1, c)), pb.G2),
                                      (O.g1_neg(i1(b.multiply(b.G1, a + c))), pb.G2)])


def test_rejects_points_off_curve():
    bad_g1 = (1, 3)
    with pytest.raises(_lib.PlonkB200Error, match="G1 point is not on the curve"):
        pb.pairing_product_is_one([(bad_g1, pb.G2)])
    bad_g2 = (pb.G2[0], pb.G2[1] + 1)
    with pytest.raises(_lib.PlonkB200Error, match="G2 point is not on the twist"):
        pb.pairing_product_is_one([(O.G1, bad_g2)])
    with pytest.raises(_lib.PlonkB200Error, match="not reduced"):
        out = ctypes.create_string_buffer(128)
        ident = ctypes.c_int()
        _lib.check(_lib.lib().pb200_g2_mul(b"\xff" * 128, bytes(32), out, ctypes.byref(ident)))


def test_fq2_value_object():
    x, y = pb.FQ2((3, 4)), pb.FQ2((5, 6))
    assert x * y == (3 * 5 - 4 * 6, 3 * 6 + 4 * 5) and x + y == [8, 10] and x - y == pb.FQ2((-2, -2))
    assert (x / y) * y == x and x * x.inv() == pb.FQ2.one() and -x + x == pb.FQ2.zero()
    assert x == pb.FQ2(x) and x != y and pb.FQ2((7, 0)) == 7


# ---- the verifier's algebra, with G1 combinations by the oracle (CPU suite has no GPU) --------------------------
def _oracle_lincomb(pairs, ctx=None):
    res = O.ec_lincomb_naive([(i1(p), int(n) % R) for p, n in pairs])
    return None if res is None else (pb.FQ(res[0]), pb.FQ(res[1]))


def _golden(name):
    entry = load_json("circuits.json")["circuits"][name]
    v = entry["vk"]
    x2 = (pb.FQ2([int(c) for c in v["X_2"][0]]), pb.FQ2([int(c) for c in v["X_2"][1]]))
    vk = pb.VerificationKey(entry["n"], *[pt(v[k]) for k in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3")],
                            x2, pb.Scalar(int(v["w"])))
    raw = O.proof_bytes({k: (pt(val) if isinstance(val, list) else int(val)) for k, val in entry["proof"].items()})
    return entry, vk, raw


@pytest.mark.parametrize("name", ["prover_test", "factorization", "poseidon"])
def test_verifier_accepts_golden_and_rejects_tampered(name, monkeypatch):
    monkeypatch.setattr(verifier, "ec_lincomb", _oracle_lincomb)
    entry, vk, raw = _golden(name)
    n, public = entry["n"], [int(x) for x in entry["public"]]
    proof = pb.Proof.from_bytes(raw)
    assert vk.verify_proof(n, proof, public)
    assert vk.verify_proof_unoptimized(n, proof, public)
    # wrong public input
    wrong = list(public)
    wrong[0] += 1
    assert not vk.verify_proof(n, proof, wrong) and not vk.verify_proof_unoptimized(n, proof, wrong)
    # one evaluation off by one
    k = 32 * 16  # c_eval
    bad = raw[:k] + ((int.from_bytes(raw[k:k + 32], "big") + 1) % R).to_bytes(32, "big") + raw[k + 32:]
    assert not vk.verify_proof(n, pb.Proof.from_bytes(bad), public)
    assert not vk.verify_proof_unoptimized(n, pb.Proof.from_bytes(bad), public)
    # an opening proof replaced by another curve point
    g = (1).to_bytes(32, "big") + (2).to_bytes(32, "big")
    bad = raw[:32 * 22] + g
    assert not vk.verify_proof(n, pb.Proof.from_bytes(bad), public)
    assert not vk.verify_proof_unoptimized(n, pb.Proof.from_bytes(bad), public)
    # z_shifted_eval only enters the second check of the unoptimized routine and the merged check of the other
    k = 32 * 19
    bad = raw[:k] + ((int.from_bytes(raw[k:k + 32], "big") + 5) % R).to_bytes(32, "big") + raw[k + 32:]
    assert not vk.verify_proof(n, pb.Proof.from_bytes(bad), public)
    assert not vk.verify_proof_unoptimized(n, pb.Proof.from_bytes(bad), public)
    # malformed proofs are rejected, not raised: a commitment that is not on the curve ...
    k = 32 * 1  # y of a_1
    bad = raw[:k] + ((int.from_bytes(raw[k:k + 32], "big") + 1) % O.Q_MOD).to_bytes(32, "big") + raw[k + 32:]
    assert not vk.verify_proof(n, pb.Proof.from_bytes(bad), public)
    assert not vk.verify_proof_unoptimized(n, pb.Proof.from_bytes(bad), public)
    # ... and the encoding is canonical: the same proof with x + q or an evaluation + r does not decode
    x0 = int.from_bytes(raw[:32], "big")
    if x0 + O.Q_MOD < 1 << 256:
        with pytest.raises(ValueError):
            pb.Proof.from_bytes((x0 + O.Q_MOD).to_bytes(32, "big") + raw[32:])
    k = 32 * 14
    e0 = int.from_bytes(raw[k:k + 32], "big")
    if e0 + R < 1 << 256:
        with pytest.raises(ValueError):
            pb.Proof.from_bytes(raw[:k] + (e0 + R).to_bytes(32, "big") + raw[k + 32:])


def test_pairing_entry_point_edge_cases():
    ok = ctypes.c_int(-1)
    L = _lib.lib()
    # empty product is 1
    _lib.check(L.pb200_pairing_check(None, None, None, None, 0, ctypes.byref(ok)))
    assert ok.value == 1
    # identity flags make the coordinate bytes irrelevant (they are not even range-checked for G1)
    g2 = b"".join(int(c).to_bytes(32, "little") for c in (*pb.G2[0].coeffs, *pb.G2[1].coeffs))
    _lib.check(L.pb200_pairing_check(b"\xff" * 64, bytes([1]), g2, bytes([0]), 1, ctypes.byref(ok)))
    assert ok.value == 1
    # e(P, Q) e(P, -Q) == 1 and the order of the factors does not matter
    nq = pb.g2_mul(pb.G2, -1)
    assert pb.pairing_product_is_one([(O.G1, pb.G2), (O.G1, nq)]) and pb.pairing_product_is_one([(O.G1, nq), (O.G1, pb.G2)])
    # scalars are reduced mod r: [r + 5] Q == [5] Q
    assert pb.g2_mul(pb.G2, R + 5) == pb.g2_mul(pb.G2, 5) == g2_of(b.multiply(b.G2, 5))
    with pytest.raises(AssertionError):
        pb.Proof.from_bytes(b"\x00" * 767)
