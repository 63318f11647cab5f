"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement (plain Python ints) of the PLONK proving hot path of
0xPARC/plonkathon, function by function, each citing the reference file:line it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may import this module, and only as the checker / the reported
baseline -- never as something the product calls.

Pinning (see tests/test_oracle_pins.py, tests/golden/):
  * the ``test.py:23-28`` commitment KAT and ``test.py:30-33`` omega_8;
  * the three snarkjs verification keys ``test/main.plonk.vkey*.json`` (24 G1 points);
  * ``test/proof.pickle`` (all 15 proof fields, i.e. rounds 1-5 + transcript);
  * Merlin's published conformance vector;
  * in the build container, element-for-element agreement with the reference's own,
    unmodified ``poly.py`` / ``curve.py`` / ``transcript.py`` / ``compiler`` imported
    from /root/reference over the ``oracle/shims`` packages (tests/test_oracle_vs_reference.py).

The arithmetic below the reference (``py_ecc 6.0.0`` field and curve ops, ``merlin``)
is third-party and absent from /root/reference; it is restated from the published
algorithms (see oracle/shims/).  The completed ``Setup.commit`` /
``Setup.verification_key`` / ``Prover.round_1..5`` bodies are stubs in the mounted
branch (``setup.py:66-77``, ``prover.py:86-306``); they are restated from the in-tree
comments, the sanity asserts, and the completed test verifier
(``TESTING_verifier_DO_NOT_OPEN.py:186-262``), and are pinned by ``test/proof.pickle``.
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.join(_HERE, "shims") not in sys.path:
    sys.path.insert(0, os.path.join(_HERE, "shims"))

from merlin import MerlinTranscript  # noqa: E402  (oracle's own restatement)

# ----------------------------------------------------------------------------
# constants (py_ecc.bn128: curve_order, field_modulus; curve.py:5,11)
# ----------------------------------------------------------------------------
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
PRIMITIVE_ROOT = 5  # curve.py:5
G1 = (1, 2)

G1Point = Optional[tuple]  # (x, y) ints mod Q_MOD, or None for the identity (py_ecc Z1)


def inv0(a: int, m: int) -> int:
    """py_ecc ``prime_field_inv``: modular inverse with inv(0) == 0."""
    a %= m
    return pow(a, -1, m) if a else 0


# ----------------------------------------------------------------------------
# Scalar helpers (curve.py:10-24)
# ----------------------------------------------------------------------------
def root_of_unity(group_order: int) -> int:
    """curve.py:15-16 -- 5 ** ((r-1) // group_order)."""
    return pow(PRIMITIVE_ROOT, (R_MOD - 1) // group_order, R_MOD)


def roots_of_unity(group_order: int) -> list[int]:
    """curve.py:20-24 -- [1, w, w^2, ...] by repeated multiplication."""
    o = [1, root_of_unity(group_order)]
    while len(o) < group_order:
        o.append(o[-1] * o[1] % R_MOD)
    return o[:group_order] if group_order >= 2 else [1]


# ----------------------------------------------------------------------------
# Polynomial transforms (poly.py)
# ----------------------------------------------------------------------------
def _fft(vals: list[int], roots: list[int]) -> list[int]:
    """poly.py:117-127 -- recursive radix-2 DIT, natural order in and out."""
    if len(vals) == 1:
        return vals
    L = _fft(vals[::2], roots[::2])
    R = _fft(vals[1::2], roots[::2])
    o = [0] * len(vals)
    h = len(L)
    for i in range(h):
        t = R[i] * roots[i]
        o[i] = (L[i] + t) % R_MOD
        o[i + h] = (L[i] - t) % R_MOD
    return o


def fft(vals: Sequence[int], inv: bool = False) -> list[int]:
    """poly.py:113-145 -- forward: coefficients -> evaluations at 1,w,..; inverse:
    reversed roots then multiply by 1/n (poly.py:132-139)."""
    n = len(vals)
    roots = roots_of_unity(n)
    nvals = [v % R_MOD for v in vals]
    if inv:
        invlen = inv0(n, R_MOD)
        rev = [roots[0]] + roots[1:][::-1]
        return [x * invlen % R_MOD for x in _fft(nvals, rev)]
    return _fft(nvals, roots)


def ifft(vals: Sequence[int]) -> list[int]:
    """poly.py:147-148."""
    return fft(vals, True)


def to_coset_extended_lagrange(vals: Sequence[int], offset: int) -> list[int]:
    """poly.py:156-163 -- ifft(n), coefficient i times offset^i, zero-pad to 4n, fft(4n)."""
    n = len(vals)
    coeffs = ifft(vals)
    xp = [pow(offset, i, R_MOD) * c % R_MOD for i, c in enumerate(coeffs)] + [0] * (3 * n)
    return fft(xp)


def coset_extended_lagrange_to_coeffs(vals: Sequence[int], offset: int) -> list[int]:
    """poly.py:169-177 -- ifft(4n) then coefficient i times offset^-i."""
    shifted = ifft(vals)
    io = inv0(offset, R_MOD)
    return [v * pow(io, i, R_MOD) % R_MOD for i, v in enumerate(shifted)]


def barycentric_eval(vals: Sequence[int], x: int) -> int:
    """poly.py:181-195 -- (x^n - 1)/n * sum(v_i w^i / (x - w^i)), inv(0)=0 semantics."""
    n = len(vals)
    roots = roots_of_unity(n)
    x %= R_MOD
    s = 0
    for v, w in zip(vals, roots):
        s += v * w % R_MOD * inv0(x - w, R_MOD)
    return (pow(x, n, R_MOD) - 1) * inv0(n, R_MOD) % R_MOD * (s % R_MOD) % R_MOD


# ----------------------------------------------------------------------------
# G1 affine group law (py_ecc.bn128 add/double/multiply; one inversion per op)
# ----------------------------------------------------------------------------
def g1_is_on_curve(p: G1Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - 3) % Q_MOD == 0


def g1_double(p: G1Point) -> G1Point:
    if p is None:
        return None
    x, y = p
    m = 3 * x * x * inv0(2 * y, Q_MOD) % Q_MOD
    nx = (m * m - 2 * x) % Q_MOD
    ny = (-m * nx + m * x - y) % Q_MOD
    return (nx, ny)


def g1_add(p1: G1Point, p2: G1Point) -> G1Point:
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2 and y1 == y2:
        return g1_double(p1)
    if x1 == x2:
        return None
    m = (y2 - y1) * inv0(x2 - x1, Q_MOD) % Q_MOD
    nx = (m * m - x1 - x2) % Q_MOD
    ny = (-m * nx + m * x1 - y1) % Q_MOD
    return (nx, ny)


def g1_neg(p: G1Point) -> G1Point:
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_multiply(p: G1Point, n: int) -> G1Point:
    """py_ecc ``multiply`` (double-and-add), as used by curve.py:30-33 ``ec_mul``."""
    n %= R_MOD
    acc = None
    add = p
    while n:
        if n & 1:
            acc = g1_add(acc, add)
        add = g1_double(add)
        n >>= 1
    return acc


ec_mul = g1_multiply


# ----------------------------------------------------------------------------
# ec_lincomb -> lincomb -> multisubset (curve.py:38-111)
# ----------------------------------------------------------------------------
def multisubset(numbers, subsets, adder, zero):
    """curve.py:59-86 -- power-set tables over partitions of 1+int(ln(#subsets+1))
    numbers, one table lookup per (subset, partition)."""
    psize = 1 + int(math.log(len(subsets) + 1))
    numbers = list(numbers)
    while len(numbers) % psize:
        numbers.append(zero)
    tables = []
    for start in range(0, len(numbers), psize):
        tbl = [zero]
        for value in numbers[start:start + psize]:
            tbl += [adder(t, value) for t in tbl]
        tables.append(tbl)
    sums = []
    for subset in subsets:
        acc = zero
        for k, tbl in enumerate(tables):
            idx = 0
            for j in range(psize):
                if k * psize + j in subset:
                    idx += 1 << j
            acc = adder(acc, tbl[idx])
        sums.append(acc)
    return sums


def lincomb(numbers, factors, adder, zero):
    """curve.py:91-111 -- bit-slice the factors into subsets, sum each subset with
    ``multisubset``, then Horner over the bits (top bit first)."""
    maxbitlen = max(len(bin(f)) - 2 for f in factors)  # ValueError on empty input (curve.py:93)
    subsets = [
        {i for i in range(len(numbers)) if factors[i] & (1 << j)} for j in range(maxbitlen + 1)
    ]
    sums = multisubset(numbers, subsets, adder, zero)
    acc = zero
    for i in range(len(subsets) - 1, -1, -1):
        acc = adder(adder(acc, acc), sums[i])
    return acc


def ec_lincomb(pairs) -> G1Point:
    """curve.py:38-44 -- sum_i n_i * P_i; scalars reduced ``int(n) % curve_order``."""
    pairs = list(pairs)
    return lincomb(
        [pt for pt, _ in pairs], [int(n) % R_MOD for _, n in pairs], g1_add, None
    )


def ec_lincomb_naive(pairs) -> G1Point:
    """curve.py:45-49 (the 'Equivalent to' comment): sum of ec_mul."""
    o = None
    for pt, n in pairs:
        o = g1_add(o, g1_multiply(pt, int(n) % R_MOD))
    return o


# ----------------------------------------------------------------------------
# Setup (setup.py)
# ----------------------------------------------------------------------------
SETUP_FILE_G1_STARTPOS = 80  # setup.py:11
SETUP_FILE_POWERS_POS = 60  # setup.py:12


@dataclass
class Setup:
    powers_of_x: list  # list[(x, y)] ints
    X2: object  # ((c0, c1), (c0, c1)) ints -- G2 point [x]_2

    @classmethod
    def from_file(cls, filename: str) -> "Setup":
        """setup.py:24-63 -- snarkjs .ptau: byte 60 = log2(#powers); G1 points from
        byte 80, 32-byte little-endian coordinates multiplied by a constant factor
        recovered from the first point (== generator)."""
        contents = open(filename, "rb").read()
        powers = 2 ** contents[SETUP_FILE_POWERS_POS]
        values = [
            int.from_bytes(contents[i:i + 32], "little")
            for i in range(SETUP_FILE_G1_STARTPOS, SETUP_FILE_G1_STARTPOS + 32 * powers * 2, 32)
        ]
        assert max(values) < Q_MOD
        factor = values[0] * inv0(G1[0], Q_MOD) % Q_MOD
        finv = inv0(factor, Q_MOD)
        values = [v * finv % Q_MOD for v in values]
        powers_of_x = [(values[2 * i], values[2 * i + 1]) for i in range(powers)]
        # G2 side (setup.py:44-59): scan for the generator's x.c0 in the same encoding
        g2_x_c0 = 10857046999023057135944570762232829481370756359578518086990519993285655852781
        target = (factor * g2_x_c0 % Q_MOD).to_bytes(32, "little")
        pos = contents.find(target, SETUP_FILE_G1_STARTPOS + 32 * powers * 2)
        assert pos >= 0
        enc = contents[pos + 32 * 4: pos + 32 * 8]
        xv = [int.from_bytes(enc[i:i + 32], "little") * finv % Q_MOD for i in range(0, 128, 32)]
        X2 = ((xv[0], xv[1]), (xv[2], xv[3]))
        return cls(powers_of_x, X2)

    def commit(self, lagrange_values: Sequence[int]) -> G1Point:
        """setup.py:66-72 -- ifft to the monomial basis, then ec_lincomb with the SRS."""
        coeffs = ifft(lagrange_values)
        if len(coeffs) > len(self.powers_of_x):
            raise Exception("Not enough powers in setup")
        return ec_lincomb([(self.powers_of_x[i], c) for i, c in enumerate(coeffs)])

    def commit_coeffs(self, coeffs: Sequence[int]) -> G1Point:
        return ec_lincomb([(self.powers_of_x[i], c) for i, c in enumerate(coeffs)])

    def verification_key(self, pk: "Preprocessed") -> dict:
        """setup.py:75-77 + verifier.py:10-37 -- eight commitments, X_2 and w."""
        return {
            "group_order": pk.group_order,
            "Qm": self.commit(pk.QM), "Ql": self.commit(pk.QL), "Qr": self.commit(pk.QR),
            "Qo": self.commit(pk.QO), "Qc": self.commit(pk.QC),
            "S1": self.commit(pk.S1), "S2": self.commit(pk.S2), "S3": self.commit(pk.S3),
            "X_2": self.X2, "w": root_of_unity(pk.group_order),
        }


# ----------------------------------------------------------------------------
# array-level circuit description (compiler/program.py:10-30 CommonPreprocessedInput)
# ----------------------------------------------------------------------------
@dataclass
class Preprocessed:
    group_order: int
    QM: list
    QL: list
    QR: list
    QO: list
    QC: list
    S1: list
    S2: list
    S3: list


# ----------------------------------------------------------------------------
# Transcript (transcript.py:58-123)
# ----------------------------------------------------------------------------
class Transcript(MerlinTranscript):
    def append_scalar(self, label: bytes, item: int):
        self.append_message(label, (item % R_MOD).to_bytes(32, "big"))  # transcript.py:62-63

    def append_point(self, label: bytes, item):
        # transcript.py:65-67: x then y under the same label; identity is unsupported
        self.append_message(label, item[0].to_bytes(32, "big"))
        self.append_message(label, item[1].to_bytes(32, "big"))

    def get_and_append_challenge(self, label: bytes) -> int:
        """transcript.py:69-75 -- 255 squeezed bytes, big-endian, mod r; redraw on 0."""
        while True:
            cb = self.challenge_bytes(label, 255)
            f = int.from_bytes(cb, "big") % R_MOD
            if f != 0:
                self.append_message(label, cb)
                return f

    def round_1(self, a_1, b_1, c_1):
        self.append_point(b"a_1", a_1)
        self.append_point(b"b_1", b_1)
        self.append_point(b"c_1", c_1)
        return self.get_and_append_challenge(b"beta"), self.get_and_append_challenge(b"gamma")

    def round_2(self, z_1):
        self.append_point(b"z_1", z_1)
        return (self.get_and_append_challenge(b"alpha"),
                self.get_and_append_challenge(b"fft_cofactor"))

    def round_3(self, t_lo_1, t_mid_1, t_hi_1):
        self.append_point(b"t_lo_1", t_lo_1)
        self.append_point(b"t_mid_1", t_mid_1)
        self.append_point(b"t_hi_1", t_hi_1)
        return self.get_and_append_challenge(b"zeta")

    def round_4(self, a_eval, b_eval, c_eval, s1_eval, s2_eval, z_shifted_eval):
        for lbl, v in ((b"a_eval", a_eval), (b"b_eval", b_eval), (b"c_eval", c_eval),
                       (b"s1_eval", s1_eval), (b"s2_eval", s2_eval),
                       (b"z_shifted_eval", z_shifted_eval)):
            self.append_scalar(lbl, v)
        return self.get_and_append_challenge(b"v")

    def round_5(self, W_z_1, W_zw_1):
        self.append_point(b"W_z_1", W_z_1)
        self.append_point(b"W_zw_1", W_zw_1)
        return self.get_and_append_challenge(b"u")


# ----------------------------------------------------------------------------
# Proof (prover.py:11-35) and its canonical 768-byte serialisation
# ----------------------------------------------------------------------------
PROOF_FIELDS = (  # prover.py:18-35 ``flatten`` order
    "a_1", "b_1", "c_1", "z_1", "t_lo_1", "t_mid_1", "t_hi_1",
    "a_eval", "b_eval", "c_eval", "s1_eval", "s2_eval", "z_shifted_eval",
    "W_z_1", "W_zw_1",
)


def proof_bytes(proof: dict) -> bytes:
    """The 15 ``Proof.flatten()`` fields in order, G1 as x||y, every integer 32-byte
    big-endian exactly as the transcript absorbs them (transcript.py:62-67)."""
    out = bytearray()
    for k in PROOF_FIELDS:
        v = proof[k]
        if isinstance(v, tuple):
            out += int(v[0]).to_bytes(32, "big") + int(v[1]).to_bytes(32, "big")
        else:
            out += int(v).to_bytes(32, "big")
    assert len(out) == 768
    return bytes(out)


# ----------------------------------------------------------------------------
# Prover (prover.py:39-315), completed per the in-tree comments / SURVEY App. D
# ----------------------------------------------------------------------------
def _padd(a, b):
    return [(x + y) % R_MOD for x, y in zip(a, b)]


def _psub(a, b):
    return [(x - y) % R_MOD for x, y in zip(a, b)]


def _pmul(a, b):
    return [x * y % R_MOD for x, y in zip(a, b)]


def _pscale(a, k):
    return [x * k % R_MOD for x in a]


def _padds(a, k):  # LAGRANGE-basis scalar add (poly.py:32-37)
    return [(x + k) % R_MOD for x in a]


class Prover:
    """prover.py:39-84.  ``pk`` is a :class:`Preprocessed`; the witness is given at
    array level: A/B/C wire values per row (prover.py:97-103) and the list of public
    input values (prover.py:57-62)."""

    def __init__(self, setup: Setup, pk: Preprocessed, check: bool = True):
        self.group_order = pk.group_order
        self.setup = setup
        self.pk = pk
        self.check = check

    def rlc(self, t1, t2):
        """prover.py:314-315."""
        return (t1 + t2 * self.beta + self.gamma) % R_MOD

    def fft_expand(self, x):
        """prover.py:308-309."""
        return to_coset_extended_lagrange(x, self.fft_cofactor)

    def expanded_evals_to_coeffs(self, x):
        """prover.py:311-312."""
        return coset_extended_lagrange_to_coeffs(x, self.fft_cofactor)

    def prove(self, A, B, C, public_inputs) -> dict:
        n = self.group_order
        tr = Transcript(b"plonk")  # prover.py:53
        self.PI = [(-int(v)) % R_MOD for v in public_inputs] + [0] * (n - len(public_inputs))
        a_1, b_1, c_1 = self.round_1(A, B, C)
        self.beta, self.gamma = tr.round_1(a_1, b_1, c_1)
        z_1 = self.round_2()
        self.alpha, self.fft_cofactor = tr.round_2(z_1)
        t_lo_1, t_mid_1, t_hi_1 = self.round_3()
        self.zeta = tr.round_3(t_lo_1, t_mid_1, t_hi_1)
        evals = self.round_4()
        self.v = tr.round_4(*evals)
        W_z_1, W_zw_1 = self.round_5()
        vals = (a_1, b_1, c_1, z_1, t_lo_1, t_mid_1, t_hi_1) + tuple(evals) + (W_z_1, W_zw_1)
        return dict(zip(PROOF_FIELDS, vals))

    def round_1(self, A, B, C):
        """prover.py:86-119."""
        n, pk = self.group_order, self.pk
        self.A = [int(v) % R_MOD for v in A] + [0] * (n - len(A))
        self.B = [int(v) % R_MOD for v in B] + [0] * (n - len(B))
        self.C = [int(v) % R_MOD for v in C] + [0] * (n - len(C))
        if self.check:  # prover.py:108-116
            for i in range(n):
                assert (self.A[i] * pk.QL[i] + self.B[i] * pk.QR[i]
                        + self.A[i] * self.B[i] * pk.QM[i] + self.C[i] * pk.QO[i]
                        + self.PI[i] + pk.QC[i]) % R_MOD == 0, "gate %d unsatisfied" % i
        return (self.setup.commit(self.A), self.setup.commit(self.B), self.setup.commit(self.C))

    def round_2(self):
        """prover.py:121-152 -- permutation grand product."""
        n, pk = self.group_order, self.pk
        roots = roots_of_unity(n)
        Z = [1]
        for i in range(n):
            num = (self.rlc(self.A[i], roots[i]) * self.rlc(self.B[i], 2 * roots[i])
                   * self.rlc(self.C[i], 3 * roots[i])) % R_MOD
            den = (self.rlc(self.A[i], pk.S1[i]) * self.rlc(self.B[i], pk.S2[i])
                   * self.rlc(self.C[i], pk.S3[i])) % R_MOD
            Z.append(Z[-1] * num % R_MOD * inv0(den, R_MOD) % R_MOD)
        assert Z.pop() == 1  # prover.py:132
        self.Z = Z
        return self.setup.commit(Z)

    def round_3(self):
        """prover.py:154-226 -- quotient polynomial on the 4n coset."""
        n, pk = self.group_order, self.pk
        k = self.fft_cofactor
        quarter = roots_of_unity(4 * n)
        xs = [k * m % R_MOD for m in quarter]  # coset points
        A_b, B_b, C_b = (self.fft_expand(v) for v in (self.A, self.B, self.C))
        PI_b = self.fft_expand(self.PI)
        QL_b, QR_b, QM_b, QO_b, QC_b = (
            self.fft_expand(v) for v in (pk.QL, pk.QR, pk.QM, pk.QO, pk.QC))
        Z_b = self.fft_expand(self.Z)
        Zw_b = Z_b[4:] + Z_b[:4]  # Z(wX): shift by 4 on the 4x-finer domain
        S1_b, S2_b, S3_b = (self.fft_expand(v) for v in (pk.S1, pk.S2, pk.S3))
        ZH_b = [(pow(x, n, R_MOD) - 1) % R_MOD for x in xs]
        L0_b = self.fft_expand([1] + [0] * (n - 1))  # prover.py:181-183
        al, be, ga = self.alpha, self.beta, self.gamma
        Q = []
        for j in range(4 * n):
            a, b, c, x = A_b[j], B_b[j], C_b[j], xs[j]
            gate = (a * QL_b[j] + b * QR_b[j] + a * b % R_MOD * QM_b[j] + c * QO_b[j]
                    + PI_b[j] + QC_b[j])
            p1 = (a + be * x + ga) * (b + 2 * be * x + ga) % R_MOD * (c + 3 * be * x + ga) % R_MOD
            p2 = ((a + be * S1_b[j] + ga) * (b + be * S2_b[j] + ga) % R_MOD
                  * (c + be * S3_b[j] + ga) % R_MOD)
            num = (gate + al * (p1 * Z_b[j] - p2 * Zw_b[j])
                   + al * al % R_MOD * (Z_b[j] - 1) * L0_b[j]) % R_MOD
            Q.append(num * inv0(ZH_b[j], R_MOD) % R_MOD)
        T = self.expanded_evals_to_coeffs(Q)
        assert T[-n:] == [0] * n  # prover.py:205-208
        self.T1c, self.T2c, self.T3c = T[:n], T[n:2 * n], T[2 * n:3 * n]
        self.T1, self.T2, self.T3 = fft(self.T1c), fft(self.T2c), fft(self.T3c)
        if self.check:  # prover.py:215-219
            assert (barycentric_eval(self.T1, k)
                    + barycentric_eval(self.T2, k) * pow(k, n, R_MOD)
                    + barycentric_eval(self.T3, k) * pow(k, 2 * n, R_MOD)) % R_MOD == Q[0]
        return (self.setup.commit(self.T1), self.setup.commit(self.T2),
                self.setup.commit(self.T3))

    def round_4(self):
        """prover.py:228-239."""
        z = self.zeta
        w = root_of_unity(self.group_order)
        self.a_eval = barycentric_eval(self.A, z)
        self.b_eval = barycentric_eval(self.B, z)
        self.c_eval = barycentric_eval(self.C, z)
        self.s1_eval = barycentric_eval(self.pk.S1, z)
        self.s2_eval = barycentric_eval(self.pk.S2, z)
        self.z_shifted_eval = barycentric_eval(self.Z, z * w % R_MOD)
        return (self.a_eval, self.b_eval, self.c_eval, self.s1_eval, self.s2_eval,
                self.z_shifted_eval)

    def round_5(self):
        """prover.py:241-306 -- linearisation polynomial and the two opening proofs,
        built in the coset extended Lagrange basis as the comments prescribe."""
        n, pk = self.group_order, self.pk
        k, zeta, v = self.fft_cofactor, self.zeta, self.v
        al, be, ga = self.alpha, self.beta, self.gamma
        w = root_of_unity(n)
        L0_ev = (pow(zeta, n, R_MOD) - 1) * inv0(n * (zeta - 1), R_MOD) % R_MOD
        ZH_ev = (pow(zeta, n, R_MOD) - 1) % R_MOD
        T1_b, T2_b, T3_b = (self.fft_expand(t) for t in (self.T1, self.T2, self.T3))
        QL_b, QR_b, QM_b, QO_b, QC_b = (
            self.fft_expand(p) for p in (pk.QL, pk.QR, pk.QM, pk.QO, pk.QC))
        Z_b = self.fft_expand(self.Z)
        S3_b = self.fft_expand(pk.S3)
        PI_ev = barycentric_eval(self.PI, zeta)
        a, b, c = self.a_eval, self.b_eval, self.c_eval
        s1, s2, zw = self.s1_eval, self.s2_eval, self.z_shifted_eval
        c1 = (a + be * zeta + ga) * (b + 2 * be * zeta + ga) % R_MOD \
            * (c + 3 * be * zeta + ga) % R_MOD * al % R_MOD
        c2 = (a + be * s1 + ga) * (b + be * s2 + ga) % R_MOD * al % R_MOD * zw % R_MOD
        zn, z2n = pow(zeta, n, R_MOD), pow(zeta, 2 * n, R_MOD)
        R_b = []
        for j in range(4 * n):
            r = (a * QL_b[j] + b * QR_b[j] + a * b % R_MOD * QM_b[j] + c * QO_b[j]
                 + PI_ev + QC_b[j]
                 + c1 * Z_b[j]
                 - c2 * ((c + be * S3_b[j] + ga) % R_MOD)
                 + al * al % R_MOD * L0_ev % R_MOD * (Z_b[j] - 1)
                 - ZH_ev * ((T1_b[j] + zn * T2_b[j] + z2n * T3_b[j]) % R_MOD))
            R_b.append(r % R_MOD)
        R_coeffs = self.expanded_evals_to_coeffs(R_b)
        assert R_coeffs[n:] == [0] * (3 * n)
        R = fft(R_coeffs[:n])
        assert barycentric_eval(R, zeta) == 0  # prover.py:267
        A_b, B_b, C_b = (self.fft_expand(p) for p in (self.A, self.B, self.C))
        S1_b, S2_b = self.fft_expand(pk.S1), self.fft_expand(pk.S2)
        quarter = roots_of_unity(4 * n)
        v2, v3, v4, v5 = (pow(v, e, R_MOD) for e in (2, 3, 4, 5))
        Wz_b, Wzw_b = [], []
        for j in range(4 * n):
            x = k * quarter[j] % R_MOD
            num = (R_b[j] + v * (A_b[j] - a) + v2 * (B_b[j] - b) + v3 * (C_b[j] - c)
                   + v4 * (S1_b[j] - s1) + v5 * (S2_b[j] - s2)) % R_MOD
            Wz_b.append(num * inv0(x - zeta, R_MOD) % R_MOD)
            Wzw_b.append((Z_b[j] - zw) * inv0(x - zeta * w, R_MOD) % R_MOD)
        Wz = self.expanded_evals_to_coeffs(Wz_b)
        assert Wz[n:] == [0] * (3 * n)  # prover.py:288
        Wzw = self.expanded_evals_to_coeffs(Wzw_b)
        assert Wzw[n:] == [0] * (3 * n)  # prover.py:299
        return (self.setup.commit(fft(Wz[:n])), self.setup.commit(fft(Wzw[:n])))


# ----------------------------------------------------------------------------
# Verification (TESTING_verifier_DO_NOT_OPEN.py:39-163 `verify_proof`), for structured test SRSs whose
# toxic value tau is known: the single pairing equation e(X_2, A) == e(G2, B) holds iff tau * A == B in G1,
# so the check needs no pairing.  Used as an independent end-to-end check of proofs at sizes (2^20 gates)
# the reference itself cannot produce.  Validated against the pairing verifier on the golden circuits'
# family in tests/test_oracle_pins.py.
# ----------------------------------------------------------------------------
def eval_lagrange_at(vals: Sequence[int], x: int) -> int:
    """Value at x of the polynomial given by Lagrange values on the n-th roots of unity (barycentric form with
    one batched inversion; x must not be a root of unity)."""
    n = len(vals)
    w = root_of_unity(n)
    diffs, cur = [], 1
    for _ in range(n):
        diffs.append((x - cur) % R_MOD)
        cur = cur * w % R_MOD
    pref, run = [], 1
    for d in diffs:
        pref.append(run)
        run = run * d % R_MOD
    inv = pow(run, -1, R_MOD)
    acc, cur_w = 0, pow(w, n - 1, R_MOD)
    w_inv = pow(w, -1, R_MOD)
    for i in range(n - 1, -1, -1):
        di = inv * pref[i] % R_MOD
        inv = inv * diffs[i] % R_MOD
        if vals[i]:
            acc += vals[i] * cur_w % R_MOD * di
        cur_w = cur_w * w_inv % R_MOD
    return (pow(x, n, R_MOD) - 1) * inv0(n, R_MOD) % R_MOD * (acc % R_MOD) % R_MOD


def compute_challenges(proof: dict):
    """TESTING_verifier_DO_NOT_OPEN.py:267-277."""
    tr = Transcript(b"plonk")
    beta, gamma = tr.round_1(proof["a_1"], proof["b_1"], proof["c_1"])
    alpha, _cof = tr.round_2(proof["z_1"])
    zeta = tr.round_3(proof["t_lo_1"], proof["t_mid_1"], proof["t_hi_1"])
    v = tr.round_4(*[proof[k] for k in ("a_eval", "b_eval", "c_eval", "s1_eval", "s2_eval", "z_shifted_eval")])
    u = tr.round_5(proof["W_z_1"], proof["W_zw_1"])
    return beta, gamma, alpha, zeta, v, u


def verify_proof_trapdoor(group_order: int, vk: dict, proof: dict, public: Sequence[int], tau: int) -> bool:
    """TESTING_verifier_DO_NOT_OPEN.py:39-163 with the final pairing equation checked through tau.
    vk: dict with G1 points Qm Ql Qr Qo Qc S1 S2 S3."""
    n = group_order
    beta, gamma, alpha, zeta, v, u = compute_challenges(proof)
    w = root_of_unity(n)
    ZH_ev = (pow(zeta, n, R_MOD) - 1) % R_MOD
    L0_ev = ZH_ev * inv0(n * (zeta - 1), R_MOD) % R_MOD
    # PI(zeta) = sum_i (-public_i) L_i(zeta),  L_i(zeta) = w^i (zeta^n - 1) / (n (zeta - w^i))
    PI_ev = 0
    for i, p in enumerate(public):
        wi = pow(w, i, R_MOD)
        PI_ev += (-p) * wi % R_MOD * ZH_ev % R_MOD * inv0(n * (zeta - wi), R_MOD)
    PI_ev %= R_MOD
    a, b, c = proof["a_eval"], proof["b_eval"], proof["c_eval"]
    s1, s2, zw = proof["s1_eval"], proof["s2_eval"], proof["z_shifted_eval"]
    r0 = (PI_ev - L0_ev * alpha * alpha
          - alpha * (a + beta * s1 + gamma) * (b + beta * s2 + gamma) % R_MOD * (c + gamma) % R_MOD * zw) % R_MOD
    D = ec_lincomb_naive([
        (vk["Qm"], a * b), (vk["Ql"], a), (vk["Qr"], b), (vk["Qo"], c), (vk["Qc"], 1),
        (proof["z_1"], ((a + beta * zeta + gamma) * (b + 2 * beta * zeta + gamma) % R_MOD
                        * (c + 3 * beta * zeta + gamma) % R_MOD * alpha + L0_ev * alpha * alpha + u)),
        (vk["S3"], -(a + beta * s1 + gamma) * (b + beta * s2 + gamma) % R_MOD * alpha * beta % R_MOD * zw),
        (proof["t_lo_1"], -ZH_ev), (proof["t_mid_1"], -ZH_ev * pow(zeta, n, R_MOD)),
        (proof["t_hi_1"], -ZH_ev * pow(zeta, 2 * n, R_MOD)),
    ])
    F = ec_lincomb_naive([(D, 1), (proof["a_1"], v), (proof["b_1"], pow(v, 2, R_MOD)),
                          (proof["c_1"], pow(v, 3, R_MOD)), (vk["S1"], pow(v, 4, R_MOD)),
                          (vk["S2"], pow(v, 5, R_MOD))])
    E = g1_multiply(G1, (-r0 + v * a + v * v * b + pow(v, 3, R_MOD) * c + pow(v, 4, R_MOD) * s1
                         + pow(v, 5, R_MOD) * s2 + u * zw) % R_MOD)
    lhs = g1_multiply(ec_lincomb_naive([(proof["W_z_1"], 1), (proof["W_zw_1"], u)]), tau)
    rhs = ec_lincomb_naive([(proof["W_z_1"], zeta), (proof["W_zw_1"], u * zeta % R_MOD * w),
                            (F, 1), (E, -1)])
    return lhs == rhs


def proof_from_bytes(raw: bytes) -> dict:
    w = [int.from_bytes(raw[i:i + 32], "big") for i in range(0, 768, 32)]
    vals = [(w[0], w[1]), (w[2], w[3]), (w[4], w[5]), (w[6], w[7]), (w[8], w[9]), (w[10], w[11]), (w[12], w[13])]
    vals += w[14:20]
    vals += [(w[20], w[21]), (w[22], w[23])]
    return dict(zip(PROOF_FIELDS, vals))
