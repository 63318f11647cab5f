"""Drop-in for the reference's ``poly.py``: ``Basis`` and ``Polynomial`` with the same methods,
assertions and list-of-Scalar ``.values`` (poly.py:10-195).  The transforms run on the GPU
(csrc/ntt.cu) through the host-buffer C ABI; element-wise ring ops stay list-based like the
reference (they are not on the accelerated path -- the prover uses device-resident vectors)."""
from __future__ import annotations

import ctypes
from enum import Enum
from typing import Optional

from . import _lib
from .curve import Scalar
from .field import CURVE_ORDER


class Basis(Enum):
    LAGRANGE = 1
    MONOMIAL = 2


def _log2_exact(n: int) -> int:
    assert n >= 1 and n & (n - 1) == 0, "length must be a power of two"
    return n.bit_length() - 1


def scalars_to_bytes(values) -> bytes:
    return b"".join(v.n.to_bytes(32, "little") for v in values)


def bytes_to_scalars(raw: bytes):
    new = Scalar.__new__
    out = []
    for i in range(0, len(raw), 32):
        s = new(Scalar)
        s.n = int.from_bytes(raw[i:i + 32], "little")
        out.append(s)
    return out


class Polynomial:
    """``values`` is the reference's list[Scalar].  Results of the GPU transforms stay resident in HBM (a
    canonical [n, 32]-byte tensor) and only materialise ``values`` when somebody reads them, so chains such as
    ``p.ifft().fft()``, ``setup.commit(p.fft())`` or ``p.to_coset_extended_lagrange(k)
    .coset_extended_lagrange_to_coeffs(k)`` cross the Python-object boundary once."""

    def __init__(self, values, basis: Basis, _dev=None):
        assert isinstance(basis, Basis)
        if _dev is None:
            assert all(isinstance(x, Scalar) for x in values)
        self._values = values
        self._dev = _dev
        self.basis = basis

    @property
    def values(self):
        if self._values is None:
            self._values = bytes_to_scalars(self._dev.cpu().numpy().tobytes())
        return self._values

    @values.setter
    def values(self, v):
        self._values = v
        self._dev = None

    def __len__(self):
        return len(self._values) if self._values is not None else int(self._dev.shape[0])

    def _device(self, ctx):
        """canonical [n, 32] uint8 CUDA tensor holding the values (uploaded once)"""
        import torch
        if self._dev is None:
            raw = bytearray(scalars_to_bytes(self._values))
            t = torch.frombuffer(raw, dtype=torch.uint8).reshape(-1, 32) if raw else torch.empty((0, 32), dtype=torch.uint8)
            self._dev = t.to(torch.device("cuda", ctx.device))
            torch.cuda.current_stream(self._dev.device).synchronize()  # the library runs on its own stream
        return self._dev

    def __eq__(self, other):
        return (self.basis == other.basis) and (self.values == other.values)

    # ---- ring operations (poly.py:23-100)
    def _zip(self, other, op, lagrange_only=False):
        assert len(self.values) == len(other.values)
        assert self.basis == other.basis
        if lagrange_only:
            assert self.basis == Basis.LAGRANGE
        return Polynomial([op(x, y) for x, y in zip(self.values, other.values)], self.basis)

    def __add__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x + y)
        assert isinstance(other, Scalar)
        if self.basis == Basis.LAGRANGE:
            return Polynomial([x + other for x in self.values], self.basis)
        return Polynomial([self.values[0] + other] + self.values[1:], self.basis)

    def __sub__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x - y)
        assert isinstance(other, Scalar)
        if self.basis == Basis.LAGRANGE:
            return Polynomial([x - other for x in self.values], self.basis)
        return Polynomial([self.values[0] - other] + self.values[1:], self.basis)

    def __mul__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x * y, lagrange_only=True)
        assert isinstance(other, Scalar)
        return Polynomial([x * other for x in self.values], self.basis)

    def __truediv__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x / y, lagrange_only=True)
        assert isinstance(other, Scalar)
        return Polynomial([x / other for x in self.values], self.basis)

    def shift(self, shift: int):
        assert self.basis == Basis.LAGRANGE
        assert shift < len(self.values)
        return Polynomial(self.values[shift:] + self.values[:shift], self.basis)

    # ---- transforms (GPU)
    def _ctx(self, ctx):
        return ctx or _lib.default_context()

    def _run(self, ctx, out_rows, call):
        import torch
        ctx = self._ctx(ctx)
        d_in = self._device(ctx)
        d_out = torch.empty((out_rows, 32), dtype=torch.uint8, device=d_in.device)
        torch.cuda.current_stream(d_in.device).synchronize()
        call(ctx, ctypes.c_void_p(d_in.data_ptr()), ctypes.c_void_p(d_out.data_ptr()))
        ctx.sync()
        return d_out

    def fft(self, inv=False, ctx: Optional[_lib.Context] = None):
        """poly.py:113-145."""
        if inv:
            assert self.basis == Basis.LAGRANGE
        else:
            assert self.basis == Basis.MONOMIAL
        n = len(self)
        log_n = _log2_exact(n)
        out = self._run(ctx, n, lambda c, i, o: _lib.check(
            _lib.lib().pb200_fr_ntt(c.handle, i, o, log_n, 1 if inv else 0)))
        return Polynomial(None, Basis.MONOMIAL if inv else Basis.LAGRANGE, _dev=out)

    def ifft(self, ctx=None):
        """poly.py:147-148."""
        return self.fft(True, ctx)

    def to_coset_extended_lagrange(self, offset, ctx=None):
        """poly.py:156-163."""
        assert self.basis == Basis.LAGRANGE
        n = len(self)
        log_n = _log2_exact(n)
        off = (int(offset) % CURVE_ORDER).to_bytes(32, "little")
        out = self._run(ctx, 4 * n, lambda c, i, o: _lib.check(
            _lib.lib().pb200_fr_coset_extend(c.handle, i, o, log_n, off)))
        return Polynomial(None, Basis.LAGRANGE, _dev=out)

    def coset_extended_lagrange_to_coeffs(self, offset, ctx=None):
        """poly.py:169-177."""
        assert self.basis == Basis.LAGRANGE
        n = len(self)
        log_n = _log2_exact(n)
        off = (int(offset) % CURVE_ORDER).to_bytes(32, "little")
        out = self._run(ctx, n, lambda c, i, o: _lib.check(
            _lib.lib().pb200_fr_coset_to_coeffs(c.handle, i, o, log_n, off)))
        return Polynomial(None, Basis.MONOMIAL, _dev=out)

    def barycentric_eval(self, x, ctx=None):
        """poly.py:181-195."""
        assert self.basis == Basis.LAGRANGE
        import torch
        ctx = self._ctx(ctx)
        d_in = self._device(ctx)
        torch.cuda.current_stream(d_in.device).synchronize()
        out = ctypes.create_string_buffer(32)
        xb = (int(x) % CURVE_ORDER).to_bytes(32, "little")
        _lib.check(_lib.lib().pb200_fr_barycentric_eval(ctx.handle, ctypes.c_void_p(d_in.data_ptr()),
                                                        _log2_exact(len(self)), xb, out))
        return Scalar(int.from_bytes(out.raw, "little"))
