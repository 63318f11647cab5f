"""ORACLE / TEST INFRASTRUCTURE ONLY.  Restated subset of ``py_ecc.bn128``
(py-ecc 6.0.0, reference pin ``poetry.lock:362-363``): alt_bn128 constants,
affine group law (one field inversion per operation, identity = None) and the
optimal-ate pairing (verifier only).  Call sites in the reference:
``curve.py:2,11,33,39-44``, ``setup.py:2,35-59``, ``utils.py:1-21``,
``TESTING_verifier_DO_NOT_OPEN.py:120-160,237-262``."""
from ..fields import bn128_FQ as FQ, bn128_FQ2 as FQ2, bn128_FQ12 as FQ12, bn128_FQP as FQP
from ..fields import BN128_FIELD_MODULUS as field_modulus

curve_order = 21888242871839275222246405745257275088548364400416034343698204186575808495617

# y^2 = x^3 + 3 over Fq; the sextic twist over Fq2; the curve over Fq12
b = FQ(3)
b2 = FQ2([3, 0]) / FQ2([9, 1])
b12 = FQ12([3] + [0] * 11)

G1 = (FQ(1), FQ(2))
G2 = (
    FQ2([
        10857046999023057135944570762232829481370756359578518086990519993285655852781,
        11559732032986387107991004021392285783925812861821192530917403151452391805634,
    ]),
    FQ2([
        8495653923123431417604973247489272438418190587263600148770280649306958101930,
        4082367875863433681332203403145435568316851327593401208105741076214120093531,
    ]),
)
Z1 = None
Z2 = None


def is_inf(pt):
    return pt is None


def is_on_curve(pt, b_):
    if pt is None:
        return True
    x, y = pt
    return y * y - x * x * x == b_


def double(pt):
    if pt is None:
        return None
    x, y = pt
    m = 3 * x * x / (2 * y)
    nx = m * m - 2 * x
    ny = -m * nx + m * x - y
    return (nx, ny)


def add(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    x1, y1 = p1
    x2, y2 = p2
    if x2 == x1 and y2 == y1:
        return double(p1)
    if x2 == x1:
        return None
    m = (y2 - y1) / (x2 - x1)
    nx = m * m - x1 - x2
    ny = -m * nx + m * x1 - y1
    return (nx, ny)


def multiply(pt, n):
    if n == 0:
        return None
    if n == 1:
        return pt
    if not n % 2:
        return multiply(double(pt), n // 2)
    return add(multiply(double(pt), n // 2), pt)


def eq(p1, p2):
    return p1 == p2


def neg(pt):
    if pt is None:
        return None
    x, y = pt
    return (x, -y)


# ---------------------------------------------------------------- pairing
ate_loop_count = 29793968203157093288
log_ate_loop_count = 63
_w = FQ12([0, 1] + [0] * 10)


def twist(pt):
    if pt is None:
        return None
    x, y = pt
    xc = [x.coeffs[0] - x.coeffs[1] * 9, x.coeffs[1]]
    yc = [y.coeffs[0] - y.coeffs[1] * 9, y.coeffs[1]]
    nx = FQ12([xc[0]] + [0] * 5 + [xc[1]] + [0] * 5)
    ny = FQ12([yc[0]] + [0] * 5 + [yc[1]] + [0] * 5)
    return (nx * _w ** 2, ny * _w ** 3)


def cast_point_to_fq12(pt):
    if pt is None:
        return None
    x, y = pt
    return (FQ12([x.n] + [0] * 11), FQ12([y.n] + [0] * 11))


def linefunc(P1, P2, T):
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = (y2 - y1) / (x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = 3 * x1 * x1 / (2 * y1)
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(Q, P):
    if Q is None or P is None:
        return FQ12.one()
    R = Q
    f = FQ12.one()
    for i in range(log_ate_loop_count, -1, -1):
        f = f * f * linefunc(R, R, P)
        R = double(R)
        if ate_loop_count & (2 ** i):
            f = f * linefunc(R, Q, P)
            R = add(R, Q)
    Q1 = (Q[0] ** field_modulus, Q[1] ** field_modulus)
    nQ2 = (Q1[0] ** field_modulus, -(Q1[1] ** field_modulus))
    f = f * linefunc(R, Q1, P)
    R = add(R, Q1)
    f = f * linefunc(R, nQ2, P)
    return f ** ((field_modulus ** 12 - 1) // curve_order)


def pairing(Q, P):
    assert is_on_curve(Q, b2)
    assert is_on_curve(P, b)
    return miller_loop(twist(Q), cast_point_to_fq12(P))
