"""Host-side field element objects at the drop-in boundary.

The reference's callers hold py_ecc ``FQ`` objects (``curve.py:10-24`` ``Scalar``; G1 coordinates
``b.FQ``).  The product keeps that object model at the edge -- ``.n``, ``+ - * / **``, comparison
with ints (``test.py:23-28``), ``inv(0) == 0`` -- while all bulk arithmetic happens on the GPU."""
from __future__ import annotations

CURVE_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FIELD_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583


def _inv0(a: int, m: int) -> int:
    a %= m
    return pow(a, -1, m) if a else 0


class PrimeFieldElement:
    field_modulus = None
    __slots__ = ("n",)

    def __init__(self, val=0):
        if isinstance(val, PrimeFieldElement):
            self.n = val.n
        elif isinstance(val, int):
            self.n = val % self.field_modulus
        elif hasattr(val, "n") and isinstance(val.n, int):  # foreign FQ-like object
            self.n = val.n % self.field_modulus
        else:
            raise TypeError("Expected an int or field element, got {}".format(type(val)))

    @staticmethod
    def _v(o):
        if isinstance(o, int):
            return o
        if hasattr(o, "n"):
            return o.n
        raise TypeError("Expected an int or field element, got {}".format(type(o)))

    def __add__(self, o): return type(self)(self.n + self._v(o))
    __radd__ = __add__
    def __mul__(self, o): return type(self)(self.n * self._v(o))
    __rmul__ = __mul__
    def __sub__(self, o): return type(self)(self.n - self._v(o))
    def __rsub__(self, o): return type(self)(self._v(o) - self.n)
    def __truediv__(self, o): return type(self)(self.n * _inv0(self._v(o), self.field_modulus))
    def __rtruediv__(self, o): return type(self)(self._v(o) * _inv0(self.n, self.field_modulus))
    def __neg__(self): return type(self)(-self.n)

    def __pow__(self, e: int):
        if e < 0:
            raise ValueError("negative exponent")
        return type(self)(pow(self.n, e, self.field_modulus))

    def __eq__(self, o):
        if isinstance(o, int):
            return self.n == o
        if hasattr(o, "n"):
            return self.n == o.n
        raise TypeError("Expected an int or field element, got {}".format(type(o)))

    def __ne__(self, o): return not self == o
    def __hash__(self): return hash(self.n)
    def __int__(self): return self.n
    def __index__(self): return self.n
    def __repr__(self): return repr(self.n)

    def __getstate__(self): return {"n": self.n}
    def __setstate__(self, st): self.n = st["n"]

    @classmethod
    def one(cls): return cls(1)

    @classmethod
    def zero(cls): return cls(0)


class FQ(PrimeFieldElement):
    """BN254 base-field element (py_ecc ``bn128.FQ``)."""
    field_modulus = FIELD_MODULUS
    __slots__ = ()
