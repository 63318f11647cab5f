#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S
OWN, UNMODIFIED modules (``/root/reference``: curve.py, poly.py, transcript.py,
compiler/*, setup.py:from_file, TESTING_verifier_DO_NOT_OPEN.py) over the restated
third-party layer in ``oracle/shims`` (py_ecc 6.0.0 / merlin are not installable here).

Only runnable in the build container (needs /root/reference).  The fixtures it writes
are what the GPU-box tests compare against; nothing at test time reads /root/reference.

What is reference-produced vs oracle-produced is recorded per fixture:
  * ``kind: reference``      -- output of reference code (poly.fft, curve.ec_lincomb,
                                 compiler, Transcript) or a reference-published value
                                 (test.py KAT, vkey JSONs, proof.pickle);
  * ``kind: oracle-verified`` -- produced by oracle/plonk_oracle.py (the reference's
                                 prover rounds are stubs) and accepted by the reference's
                                 completed verifier ``TESTING_verifier_DO_NOT_OPEN.py``
                                 (both ``verify_proof`` and ``verify_proof_unoptimized``).
"""
import hashlib
import json
import os
import pickle
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
os.chdir(REF)  # reference modules open "test/..." relative paths

import numpy as np  # noqa: E402

import curve as rcurve  # noqa: E402  reference
import poly as rpoly  # noqa: E402  reference
import setup as rsetup  # noqa: E402  reference (from_file only; commit is a stub)
import prover as rprover  # noqa: E402  reference (Proof dataclass, Message types)
import transcript as rtranscript  # noqa: E402
from compiler.program import Program  # noqa: E402  reference compiler
from TESTING_verifier_DO_NOT_OPEN import TestingVerificationKey  # noqa: E402
import py_ecc.bn128 as b  # noqa: E402  (shim)
import test as rtest_pkg  # noqa: E402,F401
from test.mini_poseidon import rc, mds, poseidon_hash  # noqa: E402

from oracle import plonk_oracle as O  # noqa: E402

Scalar = rcurve.Scalar
R = O.R_MOD


def to_le(ints):
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in ints),
                         dtype=np.uint8).reshape(-1, 32).copy()


def pt_json(p):
    return None if p is None else [str(int(p[0])), str(int(p[1]))]


def ref_pt(p):
    return None if p is None else (int(p[0].n), int(p[1].n))


# ---------------------------------------------------------------- SRS
PTAU = os.path.join(REF, "test/powersOfTau28_hez_final_11.ptau")
ref_setup = rsetup.Setup.from_file(PTAU)
osetup = O.Setup.from_file(PTAU)
assert [ref_pt(p) for p in ref_setup.powers_of_x] == osetup.powers_of_x
assert (tuple(ref_setup.X2[0].coeffs), tuple(ref_setup.X2[1].coeffs)) == osetup.X2
# truncated copy of the data file: header + 2048 G1 powers + first two G2 powers
raw = open(PTAU, "rb").read()
# locate G2 start exactly as setup.py:44-51 does
factor = int.from_bytes(raw[80:112], "little")
target = (factor * b.G2[0].coeffs[0] % O.Q_MOD).to_bytes(32, "little")
g2pos = raw.find(target, 80 + 64 * 2048)
with open(os.path.join(HERE, "powersOfTau28_hez_final_11.head.ptau"), "wb") as f:
    f.write(raw[: g2pos + 32 * 8])
print("ptau head bytes", g2pos + 256)

# ---------------------------------------------------------------- transforms (reference poly.py)
rng = random.Random(20260924)
fft_cases = []
for logn in range(0, 11):
    n = 1 << logn
    vals = [rng.randrange(R) for _ in range(n)]
    if logn == 3:
        vals = [0, 1, R - 1, 2, R - 2, 0, 0, 5]
    P = rpoly.Polynomial([Scalar(v) for v in vals], rpoly.Basis.MONOMIAL)
    fwd = [x.n for x in P.fft().values]
    L = rpoly.Polynomial([Scalar(v) for v in vals], rpoly.Basis.LAGRANGE)
    inv = [x.n for x in L.ifft().values]
    case = {"n": n, "input": vals, "fft": fwd, "ifft": inv}
    assert O.fft(vals) == fwd and O.ifft(vals) == inv
    if 1 <= logn <= 8:
        off = rng.randrange(1, R)
        x = rng.randrange(R)
        case["offset"] = off
        case["coset_ext"] = [v.n for v in L.to_coset_extended_lagrange(Scalar(off)).values]
        case["coset_to_coeffs"] = [
            v.n for v in L.coset_extended_lagrange_to_coeffs(Scalar(off)).values]
        case["x"] = x
        case["bary"] = L.barycentric_eval(Scalar(x)).n
        # x on the domain: inv(0)=0 semantics of py_ecc
        w = Scalar.root_of_unity(n)
        case["x_root"] = (w ** (n // 2 + 0)).n if n > 1 else 1
        case["bary_root"] = L.barycentric_eval(Scalar(case["x_root"])).n
        assert O.to_coset_extended_lagrange(vals, off) == case["coset_ext"]
        assert O.coset_extended_lagrange_to_coeffs(vals, off) == case["coset_to_coeffs"]
        assert O.barycentric_eval(vals, x) == case["bary"]
        assert O.barycentric_eval(vals, case["x_root"]) == case["bary_root"]
    fft_cases.append(case)
json.dump({"kind": "reference", "source": "poly.py:113-195 run unmodified over oracle/shims",
           "cases": [{k: ([str(i) for i in v] if isinstance(v, list) else str(v))
                      for k, v in c.items()} for c in fft_cases]},
          open(os.path.join(HERE, "fft_vectors.json"), "w"))
print("fft vectors ok")

# ---------------------------------------------------------------- ec_lincomb (reference curve.py)
pts = ref_setup.powers_of_x
G = b.G1
lin_cases = []


def add_case(name, pairs):
    out = rcurve.ec_lincomb(pairs)
    naive = b.Z1
    for p, n in pairs:
        naive = b.add(naive, rcurve.ec_mul(p, n)) if p is not None else naive
    assert (out is None and naive is None) or out == naive, name
    ip = [(ref_pt(p), int(n.n) if hasattr(n, "n") else int(n)) for p, n in pairs]
    assert O.ec_lincomb(ip) == ref_pt(out), name
    lin_cases.append({"name": name, "points": [pt_json(p) for p, _ in ip],
                      "scalars": [str(n) for _, n in ip], "result": pt_json(ref_pt(out))})


add_case("single_one", [(G, 1)])
add_case("single_zero", [(G, 0)])
add_case("all_zero", [(pts[i], 0) for i in range(5)])
add_case("one_nonzero", [(pts[i], 7 if i == 3 else 0) for i in range(8)])
add_case("repeated_base", [(G, 3), (G, 5), (G, rng.randrange(R))])
add_case("p_plus_p", [(pts[1], 1), (pts[1], 1)])
add_case("p_minus_p", [(pts[2], 5), (b.neg(pts[2]), 5)])
add_case("cancel_to_identity", [(pts[2], 5), (pts[2], R - 5)])
add_case("none_points", [(None, 5), (pts[1], 9), (None, 0)])
add_case("negative_and_big", [(pts[1], -3), (pts[2], R + 11), (pts[3], Scalar(-1))])
add_case("r_minus_1", [(pts[i], R - 1) for i in range(4)])
add_case("small_scalars", [(pts[i], i % 3) for i in range(33)])
for n in (2, 7, 16, 63, 64, 128):
    add_case("random_%d" % n, [(pts[i], rng.randrange(R)) for i in range(n)])
add_case("skewed_96", [(pts[i], rng.choice([0, 0, 1, 1, 2, rng.randrange(R)]))
                       for i in range(96)])
json.dump({"kind": "reference", "source": "curve.py:38-111 run unmodified over oracle/shims",
           "cases": lin_cases}, open(os.path.join(HERE, "lincomb_vectors.json"), "w"))
print("lincomb vectors ok")


# ---------------------------------------------------------------- circuits (reference compiler)
def pk_arrays(program):
    pk = program.common_preprocessed_input()
    g = lambda p: [x.n for x in p.values]  # noqa: E731
    return O.Preprocessed(pk.group_order, g(pk.QM), g(pk.QL), g(pk.QR), g(pk.QO), g(pk.QC),
                          g(pk.S1), g(pk.S2), g(pk.S3))


def wire_arrays(program, witness):
    """prover.py:94-103: A/B/C[i] = witness[wires[i].L/R/O], None -> 0."""
    w = dict(witness)
    w[None] = 0
    A = [int(w[x.L]) % R for x in program.wires()]
    B = [int(w[x.R]) % R for x in program.wires()]
    C = [int(w[x.O]) % R for x in program.wires()]
    pub = [int(w[v]) for v in program.get_public_assignments()]
    return A, B, C, pub


def to_ref_proof(p):
    FQ = b.FQ
    pt = lambda q: (FQ(q[0]), FQ(q[1]))  # noqa: E731
    return rprover.Proof(
        rtranscript.Message1(pt(p["a_1"]), pt(p["b_1"]), pt(p["c_1"])),
        rtranscript.Message2(pt(p["z_1"])),
        rtranscript.Message3(pt(p["t_lo_1"]), pt(p["t_mid_1"]), pt(p["t_hi_1"])),
        rtranscript.Message4(*[Scalar(p[k]) for k in (
            "a_eval", "b_eval", "c_eval", "s1_eval", "s2_eval", "z_shifted_eval")]),
        rtranscript.Message5(pt(p["W_z_1"]), pt(p["W_zw_1"])))


def ref_verify(vk, n, proof, public):
    FQ, FQ2 = b.FQ, b.FQ2
    pt = lambda q: None if q is None else (FQ(q[0]), FQ(q[1]))  # noqa: E731
    tvk = TestingVerificationKey(
        group_order=n, Qm=pt(vk["Qm"]), Ql=pt(vk["Ql"]), Qr=pt(vk["Qr"]), Qo=pt(vk["Qo"]),
        Qc=pt(vk["Qc"]), S1=pt(vk["S1"]), S2=pt(vk["S2"]), S3=pt(vk["S3"]),
        X_2=(FQ2(list(vk["X_2"][0])), FQ2(list(vk["X_2"][1]))), w=Scalar(vk["w"]))
    rp = to_ref_proof(proof)
    assert tvk.verify_proof_unoptimized(n, rp, public)
    assert tvk.verify_proof(n, rp, public)


circuits = {}


def add_circuit(name, program, witness=None, vkey_json=None, expect_proof=None):
    pk = pk_arrays(program)
    n = pk.group_order
    entry = {"n": n}
    arrays = {"QM": pk.QM, "QL": pk.QL, "QR": pk.QR, "QO": pk.QO, "QC": pk.QC,
              "S1": pk.S1, "S2": pk.S2, "S3": pk.S3}
    vk = osetup.verification_key(pk)
    if vkey_json is not None:  # snarkjs-produced keys: reference-published (test.py:46-54)
        import utils as rutils
        theirs = json.load(open(os.path.join(REF, vkey_json)))
        for key in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3"):
            assert ref_pt(rutils.interpret_json_point(theirs[key])) == vk[key], (name, key)
        x2 = rutils.interpret_json_point(theirs["X_2"])
        assert (tuple(x2[0].coeffs), tuple(x2[1].coeffs)) == vk["X_2"]
        assert int(theirs["w"]) == vk["w"]
        entry["vk_kind"] = "reference (snarkjs %s)" % vkey_json
    else:
        entry["vk_kind"] = "oracle"
    entry["vk"] = {k: pt_json(vk[k]) for k in ("Qm", "Ql", "Qr", "Qo", "Qc", "S1", "S2", "S3")}
    entry["vk"]["w"] = str(vk["w"])
    entry["vk"]["X_2"] = [[str(c) for c in vk["X_2"][0]], [str(c) for c in vk["X_2"][1]]]
    if witness is not None:
        A, B, C, pub = wire_arrays(program, witness)
        arrays.update({"A": A, "B": B, "C": C})
        entry["public"] = [str(p) for p in pub]
        proof = O.Prover(osetup, pk).prove(A, B, C, pub)
        if expect_proof is not None:
            assert proof == expect_proof, name
            entry["proof_kind"] = "reference (test/proof.pickle)"
        else:
            ref_verify(vk, n, proof, pub)
            entry["proof_kind"] = "oracle-verified"
        entry["proof"] = {k: (pt_json(v) if isinstance(v, tuple) else str(v))
                          for k, v in proof.items()}
        entry["proof_sha256"] = hashlib.sha256(O.proof_bytes(proof)).hexdigest()
        print(name, "proof", entry["proof_kind"], entry["proof_sha256"])
    np.savez_compressed(os.path.join(HERE, "circuit_%s.npz" % name),
                        **{k: to_le(v) for k, v in arrays.items()})
    circuits[name] = entry


# test.py:14-34 commitment KAT
kat = osetup.commit(list(range(1, 9)))
assert kat == (16120260411117808045030798560855586501988622612038310041007562782458075125622,
               3125847109934958347271782137825877642397632921923926105820408033549219695465)
assert O.root_of_unity(8) == \
    19540430494807482326159819597004422086093766032135589407132600596362845576832

add_circuit("basic", Program(["c <== a * b"], 8), vkey_json="test/main.plonk.vkey.json")
add_circuit("ab_plus_a", Program(["ab === a - c", "-ab === a * b"], 8),
            vkey_json="test/main.plonk.vkey-58.json")
add_circuit("one_public", Program(["c public", "c === a * b"], 8),
            vkey_json="test/main.plonk.vkey-59.json")

# test.py:136-145 + test/proof.pickle (reference-published golden proof)
with open(os.path.join(REF, "test/proof.pickle"), "rb") as f:
    gp = pickle.load(f).flatten()
golden = {k: (ref_pt(v) if isinstance(v, tuple) else int(v.n)) for k, v in gp.items()}
add_circuit("prover_test", Program(["e public", "c <== a * b", "e <== c * d"], 8),
            witness={"a": 3, "b": 4, "c": 12, "d": 5, "e": 60}, expect_proof=golden)

# test.py:171-213
fact = Program.from_str(
    """n public
    pb0 === pb0 * pb0
    pb1 === pb1 * pb1
    pb2 === pb2 * pb2
    pb3 === pb3 * pb3
    qb0 === qb0 * qb0
    qb1 === qb1 * qb1
    qb2 === qb2 * qb2
    qb3 === qb3 * qb3
    pb01 <== pb0 + 2 * pb1
    pb012 <== pb01 + 4 * pb2
    p <== pb012 + 8 * pb3
    qb01 <== qb0 + 2 * qb1
    qb012 <== qb01 + 4 * qb2
    q <== qb012 + 8 * qb3
    n <== p * q""", 16)
fw = fact.fill_variable_assignments(
    {"pb3": 1, "pb2": 1, "pb1": 0, "pb0": 1, "qb3": 0, "qb2": 1, "qb1": 1, "qb0": 1})
add_circuit("factorization", fact, witness=fw)


# test.py:216-259 mini-Poseidon
def output_proof_lang():
    o = ["L0 public", "M0 public", "M64 public", "R0 <== 0"]
    for i in range(64):
        for j, pos in enumerate(("L", "M", "R")):
            f = {"x": i, "r": rc[i][j], "p": pos}
            if i < 4 or i >= 60 or pos == "L":
                o.append("{p}adj{x} <== {p}{x} + {r}".format(**f))
                o.append("{p}sq{x} <== {p}adj{x} * {p}adj{x}".format(**f))
                o.append("{p}qd{x} <== {p}sq{x} * {p}sq{x}".format(**f))
                o.append("{p}qn{x} <== {p}qd{x} * {p}adj{x}".format(**f))
            else:
                o.append("{p}qn{x} <== {p}{x} + {r}".format(**f))
        for j, pos in enumerate(("L", "M", "R")):
            o.append("{p}suma{x} <== Lqn{x} * {m}".format(x=i, p=pos, m=mds[j]))
            o.append("{p}sumb{x} <== {p}suma{x} + Mqn{x} * {m}".format(x=i, p=pos, m=mds[j + 1]))
            o.append("{p}{xp1} <== {p}sumb{x} + Rqn{x} * {m}".format(
                x=i, xp1=i + 1, p=pos, m=mds[j + 2]))
    return "\n".join(o)


if "--skip-poseidon" not in sys.argv:
    pos = Program.from_str(output_proof_lang(), 1024)
    pw = pos.fill_variable_assignments({"L0": 1, "M0": 2})
    assert pw["M64"] == poseidon_hash(1, 2).n
    add_circuit("poseidon", pos, witness=pw)

# ---------------------------------------------------------------- transcript vectors
tr = rtranscript.Transcript(b"plonk")
tr.append_point(b"a_1", (b.FQ(1), b.FQ(2)))
tr.append_scalar(b"a_eval", Scalar(12345))
ch = tr.get_and_append_challenge(b"beta")
ch2 = tr.get_and_append_challenge(b"gamma")
otr = O.Transcript(b"plonk")
otr.append_point(b"a_1", (1, 2))
otr.append_scalar(b"a_eval", 12345)
assert (otr.get_and_append_challenge(b"beta"), otr.get_and_append_challenge(b"gamma")) \
    == (ch.n, ch2.n)

json.dump({
    "commit_kat": {"kind": "reference (test.py:23-28)", "lagrange": [str(i) for i in range(1, 9)],
                   "point": pt_json(kat)},
    "omega_8": str(O.root_of_unity(8)),
    "transcript": {"kind": "reference (transcript.py over merlin shim)",
                   "ops": "plonk|point a_1 (1,2)|scalar a_eval 12345|challenge beta|challenge gamma",
                   "beta": str(ch.n), "gamma": str(ch2.n)},
    "merlin_vector": "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615",
    "circuits": circuits,
}, open(os.path.join(HERE, "circuits.json"), "w"), indent=1)
print("done")
