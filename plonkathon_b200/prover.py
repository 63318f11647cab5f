"""Drop-in for the reference's ``prover.py``: ``Proof`` (prover.py:11-35) and ``Prover`` with
``prove`` and ``round_1..5`` (prover.py:39-306).  The rounds run on the GPU through the C ABI
(csrc/prover.cu); vectors stay device-resident between rounds and only the 9 points + 6 scalars of the
proof become Python objects.

``Prover(setup, program)`` accepts the reference's ``Program`` (duck-typed: ``group_order``,
``common_preprocessed_input()``, ``wires()``, ``get_public_assignments()``).  For circuits that never
existed as ``Program`` objects (synthetic 2^k-gate circuits) use ``Prover.from_arrays`` and
``prove_arrays`` with numpy buffers."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib
from .curve import Scalar
from .field import CURVE_ORDER, FIELD_MODULUS, FQ
from .poly import Basis, _log2_exact, scalars_to_bytes
from .transcript import Message1, Message2, Message3, Message4, Message5, Transcript

PK_ORDER = ("QM", "QL", "QR", "QO", "QC", "S1", "S2", "S3")  # compiler/program.py:10-30
PROOF_FIELDS = ("a_1", "b_1", "c_1", "z_1", "t_lo_1", "t_mid_1", "t_hi_1", "a_eval", "b_eval", "c_eval",
                "s1_eval", "s2_eval", "z_shifted_eval", "W_z_1", "W_zw_1")


@dataclass
class Proof:
    msg_1: Message1
    msg_2: Message2
    msg_3: Message3
    msg_4: Message4
    msg_5: Message5

    def flatten(self):
        """prover.py:18-35."""
        m1, m2, m3, m4, m5 = self.msg_1, self.msg_2, self.msg_3, self.msg_4, self.msg_5
        vals = (m1.a_1, m1.b_1, m1.c_1, m2.z_1, m3.t_lo_1, m3.t_mid_1, m3.t_hi_1, m4.a_eval, m4.b_eval,
                m4.c_eval, m4.s1_eval, m4.s2_eval, m4.z_shifted_eval, m5.W_z_1, m5.W_zw_1)
        return dict(zip(PROOF_FIELDS, vals))

    def to_bytes(self) -> bytes:
        """Canonical 768-byte form: flatten() order, G1 as x||y, 32-byte big-endian integers."""
        out = bytearray()
        for v in self.flatten().values():
            if isinstance(v, tuple):
                out += v[0].n.to_bytes(32, "big") + v[1].n.to_bytes(32, "big")
            else:
                out += v.n.to_bytes(32, "big")
        return bytes(out)

    @classmethod
    def from_bytes(cls, raw: bytes) -> "Proof":
        """Inverse of to_bytes.  The encoding is canonical: coordinates must be below q and evaluations below r
        (ValueError otherwise) -- a second byte string for the same proof would make proofs malleable."""
        assert len(raw) == 768
        w = [int.from_bytes(raw[i:i + 32], "big") for i in range(0, 768, 32)]
        for k, x in enumerate(w):
            if x >= (CURVE_ORDER if 14 <= k < 20 else FIELD_MODULUS):
                raise ValueError("non-canonical proof encoding (word %d is not reduced)" % k)
        pt = lambda k: (FQ(w[k]), FQ(w[k + 1]))  # noqa: E731
        return cls(Message1(pt(0), pt(2), pt(4)), Message2(pt(6)), Message3(pt(8), pt(10), pt(12)),
                   Message4(*[Scalar(x) for x in w[14:20]]), Message5(pt(20), pt(22)))


def _as_le_rows(values, n) -> np.ndarray:
    """list of ints / Scalars, or an (m,32) uint8 / (m,8) uint32 array -> contiguous (n,32) uint8, zero padded."""
    if isinstance(values, np.ndarray):
        arr = np.ascontiguousarray(values).view(np.uint8).reshape(-1, 32)
    else:
        raw = b"".join((v.n if hasattr(v, "n") else int(v) % CURVE_ORDER).to_bytes(32, "little") for v in values)
        arr = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 32)
    if arr.shape[0] < n:
        arr = np.concatenate([arr, np.zeros((n - arr.shape[0], 32), dtype=np.uint8)])
    assert arr.shape[0] == n
    return np.ascontiguousarray(arr)


def _pts(raw: bytes, count: int):
    return [(FQ(int.from_bytes(raw[64 * k:64 * k + 32], "little")),
             FQ(int.from_bytes(raw[64 * k + 32:64 * k + 64], "little"))) for k in range(count)]


def _raise(err: _lib.PlonkB200Error):
    if str(err).startswith("AssertionError"):
        raise AssertionError(str(err)) from None
    raise err


class Prover:
    _CREATE = "pb200_prover_create"

    def __init__(self, setup, program):
        """prover.py:45-49."""
        self.group_order = program.group_order
        self.setup = setup
        self.program = program
        self.pk = program.common_preprocessed_input()
        cols = {k: scalars_to_bytes(getattr(self.pk, k).values) for k in PK_ORDER}
        self._create(setup, self.group_order, cols)

    @classmethod
    def from_arrays(cls, setup, group_order: int, pk_arrays: dict, ctx=None):
        """pk_arrays: QM QL QR QO QC S1 S2 S3 -> list of ints or (n,32) uint8 little-endian arrays.
        ``ctx``: run this prover on another context (stream + scratch) of the same device than the setup's; the SRS
        is shared read-only, so several provers can be driven concurrently from different host threads."""
        self = cls.__new__(cls)
        self.group_order = group_order
        self.setup = setup
        self.program = None
        self.pk = None
        cols = {k: _as_le_rows(pk_arrays[k], group_order) for k in PK_ORDER}
        self._create(setup, group_order, cols, ctx)
        return self

    def _create(self, setup, n, cols, ctx=None):
        self.ctx = ctx or setup.ctx
        self._log_n = _log2_exact(n)
        keep = [c if isinstance(c, bytes) else c.tobytes() for c in (cols[k] for k in PK_ORDER)]
        arr = (ctypes.c_char_p * 8)(*keep)
        h = ctypes.c_void_p()
        create = getattr(_lib.lib(), self._CREATE)  # the multi-GPU prover creates its sharded counterpart
        _lib.check(create(self.ctx.handle, setup._srs, self._log_n, ctypes.cast(arr, ctypes.c_void_p), ctypes.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().pb200_prover_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ array-level fast path
    def prove_arrays(self, A, B, C, public) -> bytes:
        """One C-ABI call for the whole proof (rounds 1-5 + transcript); returns the canonical 768 bytes."""
        n = self.group_order
        a, b, c = (_as_le_rows(v, n) for v in (A, B, C))
        pub = _as_le_rows(public, len(public)) if len(public) else np.zeros((0, 32), dtype=np.uint8)
        out = ctypes.create_string_buffer(768)
        try:
            _lib.check(_lib.lib().pb200_prover_prove(
                self._h, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                c.ctypes.data_as(ctypes.c_void_p), pub.ctypes.data_as(ctypes.c_void_p), pub.shape[0], out))
        except _lib.PlonkB200Error as e:
            _raise(e)
        return out.raw

    # ------------------------------------------------------------------ the reference's surface
    def prove(self, witness) -> Proof:
        """prover.py:51-84."""
        transcript = Transcript(b"plonk")
        msg_1 = self.round_1(witness)  # also collects the public inputs (prover.py:57-62)
        self.beta, self.gamma = transcript.round_1(msg_1)
        msg_2 = self.round_2()
        self.alpha, self.fft_cofactor = transcript.round_2(msg_2)
        msg_3 = self.round_3()
        self.zeta = transcript.round_3(msg_3)
        msg_4 = self.round_4()
        self.v = transcript.round_4(msg_4)
        msg_5 = self.round_5()
        return Proof(msg_1, msg_2, msg_3, msg_4, msg_5)

    def round_1(self, witness) -> Message1:
        """prover.py:86-119."""
        if None not in witness:
            witness[None] = 0
        wires = self.program.wires()
        n = self.group_order
        A = [int(witness[w.L]) % CURVE_ORDER for w in wires]
        B = [int(witness[w.R]) % CURVE_ORDER for w in wires]
        C = [int(witness[w.O]) % CURVE_ORDER for w in wires]
        self._public = [int(witness[v]) % CURVE_ORDER for v in self.program.get_public_assignments()]
        return self.round_1_arrays(A, B, C, self._public)

    def round_1_arrays(self, A, B, C, public) -> Message1:
        """round 1 from wire-value columns (lists of ints or (n,32) uint8 arrays) instead of a witness dict"""
        n = self.group_order
        self._public = [int(x) % CURVE_ORDER for x in public]
        a, b, c = (_as_le_rows(v, n) for v in (A, B, C))
        pub = _as_le_rows(self._public, len(self._public)) if self._public else np.zeros((0, 32), np.uint8)
        out = ctypes.create_string_buffer(192)
        try:
            _lib.check(_lib.lib().pb200_prover_round1(
                self._h, a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p),
                c.ctypes.data_as(ctypes.c_void_p), pub.ctypes.data_as(ctypes.c_void_p), pub.shape[0], out))
        except _lib.PlonkB200Error as e:
            _raise(e)
        return Message1(*self._commitments(0, 3, out.raw))

    @staticmethod
    def _le(x) -> bytes:
        return (int(x) % CURVE_ORDER).to_bytes(32, "little")

    def round_2(self) -> Message2:
        """prover.py:121-152."""
        out = ctypes.create_string_buffer(64)
        try:
            _lib.check(_lib.lib().pb200_prover_round2(self._h, self._le(self.beta), self._le(self.gamma), out))
        except _lib.PlonkB200Error as e:
            _raise(e)
        return Message2(*self._commitments(3, 1, out.raw))

    def round_3(self) -> Message3:
        """prover.py:154-226."""
        out = ctypes.create_string_buffer(192)
        try:
            _lib.check(_lib.lib().pb200_prover_round3(self._h, self._le(self.alpha), self._le(self.fft_cofactor), out))
        except _lib.PlonkB200Error as e:
            _raise(e)
        return Message3(*self._commitments(4, 3, out.raw))

    def round_4(self) -> Message4:
        """prover.py:228-239."""
        out = ctypes.create_string_buffer(192)
        _lib.check(_lib.lib().pb200_prover_round4(self._h, self._le(self.zeta), out))
        return Message4(*[Scalar(int.from_bytes(out.raw[32 * k:32 * k + 32], "little")) for k in range(6)])

    def round_5(self) -> Message5:
        """prover.py:241-306."""
        out = ctypes.create_string_buffer(128)
        try:
            _lib.check(_lib.lib().pb200_prover_round5(self._h, self._le(self.v), out))
        except _lib.PlonkB200Error as e:
            _raise(e)
        return Message5(*self._commitments(7, 2, out.raw))

    # ------------------------------------------------------------------ round state (prover.py: self.A .. self.T3)
    def _state(self, which: int, basis):
        """a vector of the device-resident round state as a lazily materialised Polynomial (stays in HBM)"""
        import torch
        from .poly import Polynomial
        n = self.group_order
        t = torch.empty((n, 32), dtype=torch.uint8, device=torch.device("cuda", self.ctx.device))
        _lib.check(_lib.lib().pb200_prover_read_vector(self._h, which, ctypes.c_void_p(t.data_ptr())))
        return Polynomial(None, basis, _dev=t)

    A = property(lambda self: self._state(0, Basis.LAGRANGE), doc="prover.py:97-103 (after round_1)")
    B = property(lambda self: self._state(1, Basis.LAGRANGE))
    C = property(lambda self: self._state(2, Basis.LAGRANGE))
    Z = property(lambda self: self._state(3, Basis.LAGRANGE), doc="prover.py:147 (after round_2)")
    PI = property(lambda self: self._state(4, Basis.LAGRANGE), doc="prover.py:57-63 (after round_1)")
    # T1, T2, T3 are Lagrange-basis Polynomials in the reference (prover.py:209-219): the forward transform of the
    # three coefficient thirds the library keeps (after round_3)
    T1 = property(lambda self: self._state(5, Basis.MONOMIAL).fft(ctx=self.ctx))
    T2 = property(lambda self: self._state(6, Basis.MONOMIAL).fft(ctx=self.ctx))
    T3 = property(lambda self: self._state(7, Basis.MONOMIAL).fft(ctx=self.ctx))

    def _commitments(self, first_slot: int, count: int, raw: bytes):
        """commitments a round produced"""
        return _pts(raw, count)

    def fft_expand(self, x):
        """prover.py:308-309 -- x.to_coset_extended_lagrange(self.fft_cofactor)."""
        return x.to_coset_extended_lagrange(self.fft_cofactor, ctx=self.ctx)

    def expanded_evals_to_coeffs(self, x):
        """prover.py:311-312 -- x.coset_extended_lagrange_to_coeffs(self.fft_cofactor)."""
        return x.coset_extended_lagrange_to_coeffs(self.fft_cofactor, ctx=self.ctx)

    def rlc(self, term_1, term_2):
        """prover.py:314-315."""
        return term_1 + term_2 * self.beta + self.gamma
