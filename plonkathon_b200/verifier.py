"""Drop-in for the reference's ``verifier.py``: ``VerificationKey`` (verifier.py:10-37) with the two
verification routines the reference leaves as stubs (verifier.py:40-92; their completed form is what
``test.py:115-133`` runs as ``TestingVerificationKey``) and ``compute_challenges`` (verifier.py:95-106).

The G1 linear combinations go through ``ec_lincomb`` (GPU MSM), G2 arithmetic and the pairing through the
library's host code (csrc/pairing.cuh).  Each pairing equation  e(L, [1]_2) == e(W, Q)  is checked as
e(L, [1]_2) * e(-W, Q) == 1 -- a product of two Miller loops and one final exponentiation.

Notation follows the PLONK paper's verifier: zeta evaluation point, v batching challenge, u the challenge
that merges the two opening checks, bars for the prover-supplied evaluations."""
from __future__ import annotations

from dataclasses import dataclass

from .curve import G1, G2, Scalar, ec_lincomb, g1_neg, g2_add, g2_mul, pairing_product_is_one
from . import _lib
from .field import CURVE_ORDER, FIELD_MODULUS
from .transcript import Transcript


def _lagrange_terms_at(group_order: int, values, x: Scalar) -> Scalar:
    """sum_i values[i] * L_i(x) for the first len(values) Lagrange polynomials of the domain:
    L_i(x) = w^i (x^n - 1) / (n (x - w^i)).  Equal to Polynomial(values + zeros, LAGRANGE).barycentric_eval(x)
    (poly.py:181-195; division by zero yields zero there as here) without materialising n - len(values) zeros."""
    w = Scalar.root_of_unity(group_order)
    zh = x ** group_order - 1
    acc, wi = Scalar(0), Scalar(1)
    for val in values:
        acc += Scalar(val) * wi * zh / ((x - wi) * group_order)
        wi *= w
    return acc


@dataclass
class VerificationKey:
    """verifier.py:10-37."""
    group_order: int
    Qm: object
    Ql: object
    Qr: object
    Qo: object
    Qc: object
    S1: object
    S2: object
    S3: object
    X_2: object
    w: Scalar

    # ---- shared by both routines: steps 4-7 of the paper's verifier
    def _common(self, group_order: int, pf, public):
        beta, gamma, alpha, zeta, v, u = self.compute_challenges(pf)
        proof = pf.flatten()
        zh_ev = zeta ** group_order - 1
        l0_ev = zh_ev / ((zeta - 1) * group_order)
        pi_ev = _lagrange_terms_at(group_order, [-int(x) % CURVE_ORDER for x in public], zeta)
        return beta, gamma, alpha, zeta, v, u, proof, zh_ev, l0_ev, pi_ev

    @staticmethod
    def _well_formed(pf) -> bool:
        """every commitment of the proof is a point of the curve y^2 = x^3 + 3 with reduced coordinates and every
        evaluation is a reduced scalar: a malformed proof is rejected (False), it must not reach the MSM"""
        for v in pf.flatten().values():
            if isinstance(v, tuple):
                x, y = int(v[0]), int(v[1])
                if not (0 <= x < FIELD_MODULUS and 0 <= y < FIELD_MODULUS and (y * y - x * x * x - 3) % FIELD_MODULUS == 0):
                    return False
            elif v is None or not 0 <= int(v) < CURVE_ORDER:
                return False
        return True

    def verify_proof(self, group_order: int, pf, public=[]) -> bool:
        """verifier.py:40-73: the batched form -- one pairing equation, the linearisation commitment never
        formed on its own.  Malformed proofs (points off the curve, the identity) are rejected, not raised."""
        if not self._well_formed(pf):
            return False
        try:
            return self._verify_batched(group_order, pf, public)
        except _lib.PlonkB200Error:
            return False

    def _verify_batched(self, group_order: int, pf, public) -> bool:
        beta, gamma, alpha, zeta, v, u, proof, zh_ev, l0_ev, pi_ev = self._common(group_order, pf, public)
        a, b, c = proof["a_eval"], proof["b_eval"], proof["c_eval"]
        s1, s2, zw = proof["s1_eval"], proof["s2_eval"], proof["z_shifted_eval"]
        root = Scalar.root_of_unity(group_order)

        perm_bar = (a + beta * s1 + gamma) * (b + beta * s2 + gamma) * alpha * zw
        # the part of r(zeta) that needs no commitment
        r0 = pi_ev - l0_ev * alpha * alpha - perm_bar * (c + gamma)
        zeta_n = zeta ** group_order
        # [D] = [r] - r0 + u [z]
        d_pt = ec_lincomb([
            (self.Qm, a * b), (self.Ql, a), (self.Qr, b), (self.Qo, c), (self.Qc, 1),
            (proof["z_1"], (a + beta * zeta + gamma) * (b + beta * 2 * zeta + gamma) * (c + beta * 3 * zeta + gamma)
             * alpha + l0_ev * alpha * alpha + u),
            (self.S3, -perm_bar * beta),
            (proof["t_lo_1"], -zh_ev), (proof["t_mid_1"], -zh_ev * zeta_n), (proof["t_hi_1"], -zh_ev * zeta_n * zeta_n),
        ])
        f_pt = ec_lincomb([(d_pt, 1), (proof["a_1"], v), (proof["b_1"], v ** 2), (proof["c_1"], v ** 3),
                           (self.S1, v ** 4), (self.S2, v ** 5)])
        e_scalar = -r0 + v * a + v ** 2 * b + v ** 3 * c + v ** 4 * s1 + v ** 5 * s2 + u * zw
        # e(W_z + u W_zw, [x]_2) == e(zeta W_z + u zeta w W_zw + F - E, [1]_2)
        lhs = ec_lincomb([(proof["W_z_1"], 1), (proof["W_zw_1"], u)])
        rhs = ec_lincomb([(proof["W_z_1"], zeta), (proof["W_zw_1"], u * zeta * root), (f_pt, 1), (G1, -e_scalar)])
        return pairing_product_is_one([(lhs, self.X_2), (g1_neg(rhs), G2)])

    def verify_proof_unoptimized(self, group_order: int, pf, public=[]) -> bool:
        """verifier.py:76-92: rebuild the commitment to the prover's linearisation polynomial R (R(zeta) == 0),
        then check the opening at zeta and the opening of Z at zeta*w separately."""
        if not self._well_formed(pf):
            return False
        try:
            return self._verify_unoptimized(group_order, pf, public)
        except _lib.PlonkB200Error:
            return False

    def _verify_unoptimized(self, group_order: int, pf, public) -> bool:
        beta, gamma, alpha, zeta, v, u, proof, zh_ev, l0_ev, pi_ev = self._common(group_order, pf, public)
        a, b, c = proof["a_eval"], proof["b_eval"], proof["c_eval"]
        s1, s2, zw = proof["s1_eval"], proof["s2_eval"], proof["z_shifted_eval"]
        root = Scalar.root_of_unity(group_order)
        zeta_n = zeta ** group_order
        sigma_bar = (a + beta * s1 + gamma) * (b + beta * s2 + gamma) * zw

        r_pt = ec_lincomb([
            # gate constraint with the wire values fixed to their evaluations
            (self.Qm, a * b), (self.Ql, a), (self.Qr, b), (self.Qo, c), (self.Qc, 1), (G1, pi_ev),
            # permutation argument: Z(X) keeps its commitment, S3(X) too, everything else is a number
            (proof["z_1"], (a + beta * zeta + gamma) * (b + beta * 2 * zeta + gamma) * (c + beta * 3 * zeta + gamma) * alpha),
            (self.S3, -sigma_bar * alpha * beta), (G1, -sigma_bar * alpha * (c + gamma)),
            # (Z(X) - 1) L0(zeta)
            (proof["z_1"], l0_ev * alpha * alpha), (G1, -l0_ev * alpha * alpha),
            # - Z_H(zeta) (T1 + zeta^n T2 + zeta^2n T3)
            (proof["t_lo_1"], -zh_ev), (proof["t_mid_1"], -zh_ev * zeta_n), (proof["t_hi_1"], -zh_ev * zeta_n * zeta_n),
        ])
        # opening at zeta of  R + v(A - a) + v^2(B - b) + v^3(C - c) + v^4(S1 - s1) + v^5(S2 - s2)
        batch = ec_lincomb([
            (r_pt, 1), (proof["a_1"], v), (proof["b_1"], v ** 2), (proof["c_1"], v ** 3), (self.S1, v ** 4),
            (self.S2, v ** 5), (G1, -(v * a + v ** 2 * b + v ** 3 * c + v ** 4 * s1 + v ** 5 * s2)),
        ])
        x_minus_zeta = g2_add(self.X_2, g2_mul(G2, -zeta))
        if not pairing_product_is_one([(batch, G2), (g1_neg(proof["W_z_1"]), x_minus_zeta)]):
            return False
        # opening of Z at zeta * w
        z_open = ec_lincomb([(proof["z_1"], 1), (G1, -zw)])
        x_minus_zeta_w = g2_add(self.X_2, g2_mul(G2, -(zeta * root)))
        return pairing_product_is_one([(z_open, G2), (g1_neg(proof["W_zw_1"]), x_minus_zeta_w)])

    def compute_challenges(self, proof):
        """verifier.py:95-106: replay the prover's transcript over the proof's five messages."""
        transcript = Transcript(b"plonk")
        beta, gamma = transcript.round_1(proof.msg_1)
        alpha, _fft_cofactor = transcript.round_2(proof.msg_2)
        zeta = transcript.round_3(proof.msg_3)
        v = transcript.round_4(proof.msg_4)
        u = transcript.round_5(proof.msg_5)
        return beta, gamma, alpha, zeta, v, u
