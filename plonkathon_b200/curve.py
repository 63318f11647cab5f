"""Drop-in for the reference's ``curve.py``: ``Scalar`` (curve.py:10-24), ``ec_mul`` (curve.py:30-33)
and ``ec_lincomb`` (curve.py:38-44), the latter backed by the GPU Pippenger MSM in csrc/msm.cu."""
from __future__ import annotations

import ctypes
from typing import NewType, Optional

from . import _lib
from .field import CURVE_ORDER, FIELD_MODULUS, FQ, FQ2, PrimeFieldElement

primitive_root = 5  # curve.py:5
G1Point = NewType("G1Point", tuple)
curve_order = CURVE_ORDER
field_modulus = FIELD_MODULUS
G1 = (FQ(1), FQ(2))
Z1 = None
G2Point = NewType("G2Point", tuple)  # curve.py:7 -- (FQ2, FQ2) on the twist y^2 = x^3 + 3/(9+u); identity = None
G2 = (FQ2((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634)),
      FQ2((8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531)))
Z2 = None


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


# ---------------------------------------------------------------------------------------------------------
# G2 and the pairing (host code in the library, csrc/pairing.cuh) -- what the reference's verifier takes from
# py_ecc: b.add / b.multiply / b.neg on G2 and b.pairing (TESTING_verifier_DO_NOT_OPEN.py:148-151, 237-262)
# ---------------------------------------------------------------------------------------------------------
def _g2_bytes(pt) -> bytes:
    if pt is None:
        return bytes(128)
    return b"".join(_coord(c).to_bytes(32, "little") for c in (*FQ2(pt[0]).coeffs, *FQ2(pt[1]).coeffs))


def _g2_from(buf: bytes, is_identity: int):
    if is_identity:
        return Z2
    w = [int.from_bytes(buf[i:i + 32], "little") for i in range(0, 128, 32)]
    return (FQ2(w[0:2]), FQ2(w[2:4]))


def g2_mul(pt, coeff):
    """b.multiply(pt, coeff) on G2; the scalar is reduced mod the curve order (negative ints legal)."""
    if pt is None:
        return Z2
    k = (coeff.n if hasattr(coeff, "n") else int(coeff)) % CURVE_ORDER
    out = ctypes.create_string_buffer(128)
    ident = ctypes.c_int(0)
    _lib.check(_lib.lib().pb200_g2_mul(_g2_bytes(pt), k.to_bytes(32, "little"), out, ctypes.byref(ident)))
    return _g2_from(out.raw, ident.value)


def g2_add(p, q):
    """b.add(p, q) on G2."""
    out = ctypes.create_string_buffer(128)
    ident = ctypes.c_int(0)
    _lib.check(_lib.lib().pb200_g2_add(_g2_bytes(p), 1 if p is None else 0, _g2_bytes(q), 1 if q is None else 0,
                                       out, ctypes.byref(ident)))
    return _g2_from(out.raw, ident.value)


def g2_neg(pt):
    return None if pt is None else (FQ2(pt[0]), -FQ2(pt[1]))


def g1_neg(pt):
    return None if pt is None else (FQ(pt[0]), -FQ(pt[1]))


def pairing_product_is_one(pairs) -> bool:
    """prod e(P, Q) == 1 over ``pairs`` of (G1 point | None, G2 point | None): one product of Miller loops and
    a single final exponentiation.  ``b.pairing(Q1, P1) == b.pairing(Q2, P2)`` is the case
    ``[(P1, Q1), (neg(P2), Q2)]``."""
    pairs = list(pairs)
    g1 = b"".join(bytes(64) if p is None else _pt_bytes(p) for p, _ in pairs)
    g1i = bytes(1 if p is None else 0 for p, _ in pairs)
    g2 = b"".join(_g2_bytes(q) for _, q in pairs)
    g2i = bytes(1 if q is None else 0 for _, q in pairs)
    ok = ctypes.c_int(0)
    _lib.check(_lib.lib().pb200_pairing_check(g1, g1i, g2, g2i, len(pairs), ctypes.byref(ok)))
    return bool(ok.value)
