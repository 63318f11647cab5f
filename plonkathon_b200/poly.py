"""Drop-in for the reference's ``poly.py``: ``Basis`` and ``Polynomial`` with the same methods,
assertions and list-of-Scalar ``.values`` (poly.py:10-195).  The transforms run on the GPU
(csrc/ntt.cu); the element-wise ring operations run on the GPU too (csrc/poly_ops.cu, pb200_fr_vec_op) whenever an
operand is device-resident -- the result of a transform, a round's state read off the prover -- so a
reference-style round written with Polynomial arithmetic (prover.py:188-202) never materialises 2^k Scalars; plain
list operands keep the reference's list arithmetic."""
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
        """The reference's list[Scalar].  Reading it makes the list the source of truth, as in the reference: the HBM
        copy is dropped, so in-place edits of the list are seen by every later transform / commitment."""
        if self._values is None:
            self._values = bytes_to_scalars(self._dev.cpu().numpy().tobytes())
        self._dev = None
        return self._values

    @values.setter
    def values(self, v):
        self._values = v
        self._dev = None

    def __len__(self):
        return len(self._values) if self._values is not None else int(self._dev.shape[0])

    @property
    def on_device(self) -> bool:
        return self._dev is not None

    def _device(self, ctx):
        """canonical [n, 32] uint8 CUDA tensor holding the values, on the context's device.  A device-born polynomial
        hands out its HBM copy; a list-backed one is uploaded on every call -- the list is the source of truth and
        may have been edited in place since the last call (the reference's semantics)."""
        import torch
        dev = torch.device("cuda", ctx.device)
        if self._dev is not None:
            if self._dev.device != dev:
                self._dev = self._dev.to(dev)
            return self._dev
        raw = bytearray(scalars_to_bytes(self._values))
        t = torch.frombuffer(raw, dtype=torch.uint8).reshape(-1, 32) if raw else torch.empty((0, 32), dtype=torch.uint8)
        t = t.to(dev)
        torch.cuda.current_stream(dev).synchronize()  # the library runs on its own stream
        return t

    def __eq__(self, other):
        if self.basis != other.basis:
            return False
        if self.on_device and other.on_device and self._dev.device == other._dev.device:
            import torch
            return self._dev.shape == other._dev.shape and bool(torch.equal(self._dev, other._dev))
        if self.on_device or other.on_device:  # compare bytes without building 2^k Scalars on the resident side
            if len(self) != len(other):
                return False
            raw = lambda p: p._dev.cpu().numpy().tobytes() if p.on_device else scalars_to_bytes(p._values)  # noqa: E731
            return raw(self) == raw(other)
        return self.values == other.values

    # ---- ring operations (poly.py:23-109)
    # op codes of pb200_fr_vec_op
    _ADD, _SUB, _MUL, _DIV, _ADD_S, _SUB_S, _MUL_S, _ADD_S0, _SUB_S0, _SHIFT = range(10)

    def _vec(self, op, other=None, scalar=None, shift=0, ctx=None):
        """one element-wise kernel over device-resident canonical vectors; the result stays in HBM"""
        import torch
        ctx = self._ctx(ctx)
        a = self._device(ctx)
        b = other._device(ctx) if other is not None else None
        out = torch.empty_like(a)
        torch.cuda.current_stream(a.device).synchronize()
        sb = (int(scalar) % CURVE_ORDER).to_bytes(32, "little") if scalar is not None else None
        _lib.check(_lib.lib().pb200_fr_vec_op(ctx.handle, op, ctypes.c_void_p(a.data_ptr()),
                                              ctypes.c_void_p(b.data_ptr()) if b is not None else None, sb,
                                              ctypes.c_void_p(out.data_ptr()), len(self), shift))
        ctx.sync()
        return Polynomial(None, self.basis, _dev=out)

    def _zip(self, other, op, dev_op, lagrange_only=False):
        assert len(self) == len(other)
        assert self.basis == other.basis
        if lagrange_only:
            assert self.basis == Basis.LAGRANGE
        if self.on_device or other.on_device:
            return self._vec(dev_op, other)
        return Polynomial([op(x, y) for x, y in zip(self.values, other.values)], self.basis)

    def __add__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x + y, self._ADD)
        assert isinstance(other, Scalar)
        if self.on_device:
            return self._vec(self._ADD_S if self.basis == Basis.LAGRANGE else self._ADD_S0, scalar=other.n)
        if self.basis == Basis.LAGRANGE:
            return Polynomial([x + other for x in self.values], self.basis)
        return Polynomial([self.values[0] + other] + self.values[1:], self.basis)

    def __sub__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x - y, self._SUB)
        assert isinstance(other, Scalar)
        if self.on_device:
            return self._vec(self._SUB_S if self.basis == Basis.LAGRANGE else self._SUB_S0, scalar=other.n)
        if self.basis == Basis.LAGRANGE:
            return Polynomial([x - other for x in self.values], self.basis)
        return Polynomial([self.values[0] - other] + self.values[1:], self.basis)

    def __mul__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x * y, self._MUL, lagrange_only=True)
        assert isinstance(other, Scalar)
        if self.on_device:
            return self._vec(self._MUL_S, scalar=other.n)
        return Polynomial([x * other for x in self.values], self.basis)

    def __truediv__(self, other):
        if isinstance(other, Polynomial):
            return self._zip(other, lambda x, y: x / y, self._DIV, lagrange_only=True)
        assert isinstance(other, Scalar)
        if self.on_device:
            return self._vec(self._MUL_S, scalar=(Scalar(1) / other).n)  # py_ecc: x / 0 == 0
        return Polynomial([x / other for x in self.values], self.basis)

    def shift(self, shift: int):
        assert self.basis == Basis.LAGRANGE
        assert shift < len(self)
        if self.on_device:
            return self._vec(self._SHIFT, shift=shift)
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
