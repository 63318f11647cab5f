"""ORACLE / TEST INFRASTRUCTURE ONLY -- the oracle prover at sizes the pure-Python one cannot reach.

Same code as plonk_oracle.Prover (the reference's round structure: 29 coset extensions, per-row Python loops,
barycentric evaluations), with its two library kernels swapped for the plain-C restatement in oracle/c/:
`fft` (poly.py:113-149) -> oc_fr_fft and `Setup.commit`'s `ec_lincomb` (curve.py:38-44) -> oc_g1_lincomb, and the
SRS [tau^i]G produced by oc_g1_powers.  The C functions are themselves pinned against the Python oracle and the
reference's golden vectors (tests/test_oracle_c.py); tests/test_oracle_fast.py checks that this composite prover
returns the same proofs as the pure-Python oracle.  Used by tests/golden/make_proof_2p20.py to produce the golden
2^20-gate proof (34 minutes on one core)."""
import contextlib

import numpy as np

from . import c_oracle as CO
from . import plonk_oracle as O


def _to_np(vals) -> np.ndarray:
    raw = b"".join((int(v) % O.R_MOD).to_bytes(32, "little") for v in vals)
    return np.frombuffer(raw, dtype=np.uint8).reshape(-1, 32)


def _from_np(arr: np.ndarray) -> list:
    raw = arr.tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def fft(vals, inv: bool = False) -> list:
    if len(vals) == 1:
        return [int(vals[0]) % O.R_MOD]
    return _from_np(CO.fft(_to_np(vals), inv))


def ifft(vals) -> list:
    return fft(vals, True)


class Setup:
    """Structured test SRS [tau^i]G, i < n, kept as an (n, 64) byte array; duck-types plonk_oracle.Setup for the
    prover (``commit``) and exposes ``points(i)`` for spot checks."""

    def __init__(self, tau: int, n: int):
        self.tau = tau
        self.pts = CO.g1_powers(tau, n)
        self.X2 = None

    def point(self, i: int):
        raw = self.pts[i].tobytes()
        return int.from_bytes(raw[:32], "little"), int.from_bytes(raw[32:], "little")

    def commit(self, lagrange_values):
        """setup.py:66-72"""
        coeffs = ifft(lagrange_values)
        if len(coeffs) > self.pts.shape[0]:
            raise Exception("Not enough powers in setup")
        return CO.g1_lincomb(self.pts[:len(coeffs)], _to_np(coeffs))


@contextlib.contextmanager
def c_kernels():
    """plonk_oracle's module-level fft / ifft (used by its coset helpers and rounds) answered by the C restatement"""
    saved = O.fft, O.ifft
    O.fft, O.ifft = fft, ifft
    try:
        yield
    finally:
        O.fft, O.ifft = saved


def prove(setup: Setup, pk: "O.Preprocessed", A, B, C, public_inputs, check: bool = True) -> dict:
    with c_kernels():
        return O.Prover(setup, pk, check=check).prove(A, B, C, public_inputs)
