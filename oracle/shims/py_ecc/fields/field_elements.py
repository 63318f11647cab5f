"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Restatement of the prime-field / extension-field element semantics of the
third-party package ``py-ecc 6.0.0`` (pinned by the reference at
``poetry.lock:362-363``; the package itself is absent from /root/reference and
cannot be installed here).  Only the behaviour the reference's call sites rely
on is restated (``curve.py:1-2,11,16``, ``setup.py:35-59``, ``utils.py:4-21``,
``poly.py`` via ``Scalar``):

* ``FQ(int | FQ)`` reduces mod ``field_modulus`` (negatives allowed);
* ``+ - * /`` accept FQ or int on either side; ``/`` multiplies by the modular
  inverse with the *inv0* convention: ``inverse(0) == 0`` (no exception);
* ``**`` takes a non-negative int exponent; ``==`` accepts FQ or int;
* ``FQP`` is a polynomial extension ``Fp[w]/(w^deg + sum c_i w^i)`` used for
  FQ2 (the G2 point in the SRS file) and FQ12 (pairing, verifier only).
"""


def prime_field_inv(a: int, n: int) -> int:
    """Modular inverse with inverse(0) == 0 (py_ecc's convention)."""
    a %= n
    if a == 0:
        return 0
    return pow(a, -1, n)


class FQ:
    field_modulus = None

    def __init__(self, val=0):
        if self.field_modulus is None:
            raise AttributeError("Field Modulus hasn't been specified")
        if isinstance(val, FQ):
            self.n = val.n
        elif isinstance(val, int):
            self.n = val % self.field_modulus
        else:
            raise TypeError(
                "Expected an int or FQ object, but got object of type {}".format(type(val))
            )

    @staticmethod
    def _val(other):
        if isinstance(other, FQ):
            return other.n
        if isinstance(other, int):
            return other
        raise TypeError("Expected an int or FQ object, but got {}".format(type(other)))

    def __add__(self, other):
        return type(self)(self.n + self._val(other))

    __radd__ = __add__

    def __mul__(self, other):
        return type(self)(self.n * self._val(other))

    __rmul__ = __mul__

    def __sub__(self, other):
        return type(self)(self.n - self._val(other))

    def __rsub__(self, other):
        return type(self)(self._val(other) - self.n)

    def __truediv__(self, other):
        return type(self)(self.n * prime_field_inv(self._val(other), self.field_modulus))

    __div__ = __truediv__

    def __rtruediv__(self, other):
        return type(self)(prime_field_inv(self.n, self.field_modulus) * self._val(other))

    __rdiv__ = __rtruediv__

    def __pow__(self, other: int):
        # square-and-multiply in py_ecc; the value is the same as pow()
        if other < 0:
            raise ValueError("negative exponent")
        return type(self)(pow(self.n, other, self.field_modulus))

    def __eq__(self, other):
        if isinstance(other, FQ):
            return self.n == other.n
        if isinstance(other, int):
            return self.n == other
        raise TypeError("Expected an int or FQ object, but got {}".format(type(other)))

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self.n)

    def __neg__(self):
        return type(self)(-self.n)

    def __repr__(self):
        return repr(self.n)

    def __int__(self):
        return self.n

    @classmethod
    def one(cls):
        return cls(1)

    @classmethod
    def zero(cls):
        return cls(0)


def _deg(p):
    d = len(p) - 1
    while d and p[d] == 0:
        d -= 1
    return d


def _poly_rounded_div(a, b, mod):
    dega, degb = _deg(a), _deg(b)
    temp = list(a)
    o = [0] * len(a)
    for i in range(dega - degb, -1, -1):
        q = temp[degb + i] * prime_field_inv(b[degb], mod) % mod
        o[i] = (o[i] + q) % mod
        for c in range(degb + 1):
            temp[c + i] = (temp[c + i] - q * b[c]) % mod
    return o[: _deg(o) + 1]


class FQP:
    """Element of Fp[w] / (w^degree + modulus_coeffs . (1, w, w^2, ...))."""

    degree = 0
    field_modulus = None
    modulus_coeffs = ()

    def __init__(self, coeffs):
        if len(coeffs) != self.degree:
            raise ValueError("wrong number of coefficients")
        p = self.field_modulus
        self.coeffs = tuple((c.n if isinstance(c, FQ) else int(c)) % p for c in coeffs)

    def _wrap(self, ints):
        return type(self)(ints)

    def __add__(self, other):
        return self._wrap([a + b for a, b in zip(self.coeffs, other.coeffs)])

    def __sub__(self, other):
        return self._wrap([a - b for a, b in zip(self.coeffs, other.coeffs)])

    def __mul__(self, other):
        p = self.field_modulus
        if isinstance(other, (int, FQ)):
            k = other.n if isinstance(other, FQ) else other
            return self._wrap([c * k for c in self.coeffs])
        d = self.degree
        b = [0] * (2 * d - 1)
        for i, x in enumerate(self.coeffs):
            if x:
                for j, y in enumerate(other.coeffs):
                    b[i + j] += x * y
        # reduce by w^d = -sum(modulus_coeffs[i] w^i)
        for exp in range(d - 2, -1, -1):
            top = b.pop() % p
            if top:
                for i, c in enumerate(self.modulus_coeffs):
                    if c:
                        b[exp + i] -= top * c
        return self._wrap(b)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, (int, FQ)):
            k = other.n if isinstance(other, FQ) else other
            return self * prime_field_inv(k, self.field_modulus)
        return self * other.inv()

    __div__ = __truediv__

    def __pow__(self, other: int):
        o = type(self).one()
        t = self
        while other > 0:
            if other & 1:
                o = o * t
            other >>= 1
            t = t * t
        return o

    def inv(self):
        # extended Euclid over Fp[w] against the field polynomial
        p = self.field_modulus
        d = self.degree
        lm, hm = [1] + [0] * d, [0] * (d + 1)
        low, high = list(self.coeffs) + [0], list(self.modulus_coeffs) + [1]
        while _deg(low):
            r = _poly_rounded_div(high, low, p)
            r += [0] * (d + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(d + 1):
                for j in range(d + 1 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm = [x % p for x in nm]
            new = [x % p for x in new]
            lm, low, hm, high = nm, new, lm, low
        return self._wrap(lm[:d]) / low[0]

    def __eq__(self, other):
        return self.coeffs == other.coeffs

    def __ne__(self, other):
        return not self == other

    def __neg__(self):
        return self._wrap([-c for c in self.coeffs])

    def __repr__(self):
        return repr(self.coeffs)

    @classmethod
    def one(cls):
        return cls([1] + [0] * (cls.degree - 1))

    @classmethod
    def zero(cls):
        return cls([0] * cls.degree)
