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


class FQ2:
    """Element c0 + c1*u of Fq[u]/(u^2 + 1) (py_ecc ``bn128.FQ2``): the coordinates of G2 points such as the
    SRS element ``X2`` (setup.py:58) and ``VerificationKey.X_2`` (verifier.py:32, compared at test.py:47).
    Host-side value object only; G2 arithmetic and the pairing run inside the library (csrc/pairing.cuh)."""
    __slots__ = ("coeffs",)

    def __init__(self, coeffs):
        c = tuple(coeffs.coeffs) if isinstance(coeffs, FQ2) else tuple(coeffs)
        if len(c) != 2:
            raise ValueError("FQ2 takes two coefficients")
        self.coeffs = (FQ(c[0]), FQ(c[1]))

    @staticmethod
    def _c(o):
        if isinstance(o, FQ2):
            return o.coeffs[0].n, o.coeffs[1].n
        if isinstance(o, int) or hasattr(o, "n"):
            return (o.n if hasattr(o, "n") else o) % FIELD_MODULUS, 0
        a, b = o
        return FQ(a).n, FQ(b).n

    def __add__(self, o):
        a, b = self._c(o)
        return FQ2((self.coeffs[0].n + a, self.coeffs[1].n + b))
    __radd__ = __add__

    def __sub__(self, o):
        a, b = self._c(o)
        return FQ2((self.coeffs[0].n - a, self.coeffs[1].n - b))

    def __neg__(self):
        return FQ2((-self.coeffs[0].n, -self.coeffs[1].n))

    def __mul__(self, o):
        a, b = self._c(o)
        x, y = self.coeffs[0].n, self.coeffs[1].n
        return FQ2((x * a - y * b, x * b + y * a))
    __rmul__ = __mul__

    def inv(self):
        x, y = self.coeffs[0].n, self.coeffs[1].n
        d = _inv0(x * x + y * y, FIELD_MODULUS)
        return FQ2((x * d, -y * d))

    def __truediv__(self, o):
        return self * FQ2(self._c(o)).inv()

    def __eq__(self, o):
        try:
            return (self.coeffs[0].n, self.coeffs[1].n) == self._c(o)
        except (TypeError, ValueError):
            return NotImplemented

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    def __hash__(self):
        return hash((self.coeffs[0].n, self.coeffs[1].n))

    def __iter__(self):
        return iter(self.coeffs)

    def __getitem__(self, i):
        return self.coeffs[i]

    def __repr__(self):
        return repr(tuple(self.coeffs))

    @classmethod
    def one(cls):
        return cls((1, 0))

    @classmethod
    def zero(cls):
        return cls((0, 0))
