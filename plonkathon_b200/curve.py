"""Drop-in for the reference's ``curve.py``: ``Scalar`` (curve.py:10-24), ``ec_mul`` (curve.py:30-33)
and ``ec_lincomb`` (curve.py:38-44), the latter backed by the GPU Pippenger MSM in csrc/msm.cu."""
from __future__ import annotations

import ctypes
from typing import NewType, Optional

from . import _lib
from .field import CURVE_ORDER, FIELD_MODULUS, FQ, PrimeFieldElement

primitive_root = 5  # curve.py:5
G1Point = NewType("G1Point", tuple)
curve_order = CURVE_ORDER
field_modulus = FIELD_MODULUS
G1 = (FQ(1), FQ(2))
Z1 = None


class Scalar(PrimeFieldElement):
    """Fr element (curve.py:10-24)."""
    field_modulus = CURVE_ORDER
    __slots__ = ()

    @classmethod
    def root_of_unity(cls, group_order: int):
        return Scalar(5) ** ((cls.field_modulus - 1) // group_order)

    @classmethod
    def roots_of_unity(cls, group_order: int):
        w = cls.root_of_unity(group_order).n
        out, cur = [], 1
        for _ in range(max(group_order, 2)):
            out.append(Scalar(cur))
            cur = cur * w % CURVE_ORDER
        return out


def _coord(c) -> int:
    return c.n if hasattr(c, "n") else int(c)


def _pt_bytes(pt) -> bytes:
    return _coord(pt[0]).to_bytes(32, "little") + _coord(pt[1]).to_bytes(32, "little")


def _pt_from(buf: bytes, is_identity: int):
    if is_identity:
        return Z1
    return (FQ(int.from_bytes(buf[:32], "little")), FQ(int.from_bytes(buf[32:64], "little")))


def ec_lincomb(pairs, ctx: Optional[_lib.Context] = None):
    """curve.py:38-44.  ``pairs``: iterable of (G1 point | None, Scalar | int).  Scalars are reduced
    ``int(n) % curve_order`` (negative ints legal); ``None`` points contribute nothing.  Empty input
    raises ValueError like the reference (``max()`` of an empty list, curve.py:93)."""
    pairs = list(pairs)
    if not pairs:
        raise ValueError("max() arg is an empty sequence")
    ctx = ctx or _lib.default_context()
    live = [(p, int(n) % CURVE_ORDER) for p, n in pairs if p is not None]
    live = [(p, n) for p, n in live if n != 0]
    if not live:
        return Z1
    pts = b"".join(_pt_bytes(p) for p, _ in live)
    sc = b"".join(n.to_bytes(32, "little") for _, n in live)
    out = ctypes.create_string_buffer(64)
    ident = ctypes.c_int(0)
    _lib.check(_lib.lib().pb200_g1_msm_host(ctx.handle, pts, sc, len(live), out, ctypes.byref(ident)))
    return _pt_from(out.raw, ident.value)


def ec_mul(pt, coeff, ctx: Optional[_lib.Context] = None):
    """curve.py:30-33 -- single scalar multiplication (an MSM of one term)."""
    if hasattr(coeff, "n"):
        coeff = coeff.n
    if pt is None:
        return Z1
    return ec_lincomb([(pt, coeff % CURVE_ORDER)], ctx)
