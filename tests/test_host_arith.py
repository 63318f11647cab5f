"""CPU: the limb-level field / curve arithmetic of csrc/field.cuh + csrc/curve.cuh, compiled for
the host (PTX carry primitives replaced by their emulation) and checked against Python ints and
the oracle's affine group law.  This is the same C++ the CUDA kernels instantiate."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import plonk_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "plonkathon_b200", "csrc")
R256 = 1 << 256


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(ROOT, "build", "host_selftest.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(CSRC, "host_selftest.cpp")
    deps = [src] + [os.path.join(CSRC, h) for h in ("field.cuh", "curve.cuh", "fieldd.cuh", "msm_digits.cuh", "msm_affine.cuh", "modinv.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src,
                               "-I", CSRC, "-o", out])
    return ctypes.CDLL(out)


def limbs(x, n=1):
    return (ctypes.c_uint32 * (8 * n))(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8 * n)])


def unlimbs(buf, k=0):
    return sum(int(buf[8 * k + i]) << (32 * i) for i in range(8))


def fop(lib, field, op, a, b=0):
    out = (ctypes.c_uint32 * 8)()
    assert lib.hs_field_op(field, op, limbs(a), limbs(b), out) == 0
    return unlimbs(out)


EDGE = [0, 1, 2, 3, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1, (1 << 128) + 1, (1 << 224) - 1]


@pytest.mark.parametrize("field,p", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_field_ops(lib, field, p):
    rng = random.Random(field)
    vals = EDGE + [p - 1, p - 2, (p - 1) // 2, R256 % p, (R256 * R256) % p] + \
        [rng.randrange(p) for _ in range(200)]
    rinv = pow(R256, -1, p)
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert fop(lib, field, 0, a, b) == (a + b) % p
        assert fop(lib, field, 1, a, b) == (a - b) % p
        assert fop(lib, field, 2, a, b) == a * b * rinv % p
        assert fop(lib, field, 8, a) == a * a * rinv % p
        assert fop(lib, field, 3, a) == (-a) % p
        assert fop(lib, field, 7, a) == 2 * a % p
        assert fop(lib, field, 5, a) == a * R256 % p
        assert fop(lib, field, 6, a) == a * rinv % p
    for a in vals[:40]:
        am = a * R256 % p
        inv = fop(lib, field, 4, am)
        assert inv == (pow(a, -1, p) * R256 % p if a else 0)


def mont(x):
    return x * R256 % O.Q_MOD


def unmont(x):
    return x * pow(R256, -1, O.Q_MOD) % O.Q_MOD


def xyzz(lib, pt):
    """affine int point (or None) -> XYZZ limbs buffer"""
    if pt is None:
        return limbs(0, 4)
    v = mont(pt[0]) | (mont(pt[1]) << 256) | (mont(1) << 512) | (mont(1) << 768)
    return limbs(v, 4)


def to_affine(lib, acc):
    out = (ctypes.c_uint32 * 32)()
    inf = lib.hs_curve_op(3, acc, limbs(0, 4), 0, out)
    return None if inf else (unmont(unlimbs(out, 0)), unmont(unlimbs(out, 1)))


def test_curve_ops(lib):
    rng = random.Random(5)
    G = O.G1
    pts = [O.g1_multiply(G, rng.randrange(1, O.R_MOD)) for _ in range(12)]
    for i, p in enumerate(pts):
        q = pts[(i + 1) % len(pts)]
        for a, b in ((p, q), (p, p), (p, O.g1_neg(p)), (None, q), (p, None)):
            acc = xyzz(lib, a)
            out = (ctypes.c_uint32 * 32)()
            if b is not None:  # mixed add
                bb = limbs(mont(b[0]) | (mont(b[1]) << 256), 2)
                lib.hs_curve_op(0, acc, bb, 0, out)
                assert to_affine(lib, out) == O.g1_add(a, b)
                out_u = (ctypes.c_uint32 * 32)()
                lib.hs_curve_op(4, acc, bb, 0, out_u)  # SIMT-uniform variant
                assert to_affine(lib, out_u) == O.g1_add(a, b)
            lib.hs_curve_op(1, acc, xyzz(lib, b), 0, out)  # full add
            assert to_affine(lib, out) == O.g1_add(a, b)
            out_u2 = (ctypes.c_uint32 * 32)()
            lib.hs_curve_op(5, acc, xyzz(lib, b), 0, out_u2)  # select-based full add
            assert to_affine(lib, out_u2) == O.g1_add(a, b)
        # chains with non-trivial ZZ: ((p+q)+q)+(p+q) etc.
        acc = xyzz(lib, p)
        bb = limbs(mont(q[0]) | (mont(q[1]) << 256), 2)
        o1 = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(0, acc, bb, 0, o1)
        o2 = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(0, o1, bb, 0, o2)
        exp = O.g1_add(O.g1_add(p, q), q)
        assert to_affine(lib, o2) == exp
        o3 = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(1, o2, o1, 0, o3)
        assert to_affine(lib, o3) == O.g1_add(exp, O.g1_add(p, q))
        o3u = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(5, o2, o1, 0, o3u)
        assert to_affine(lib, o3u) == to_affine(lib, o3)
        o4 = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(1, o3, o3, 0, o4)  # projective doubling through add
        assert to_affine(lib, o4) == O.g1_double(O.g1_add(exp, O.g1_add(p, q)))
        o5 = (ctypes.c_uint32 * 32)()
        lib.hs_curve_op(2, o3, limbs(0, 4), 0, o5)
        assert to_affine(lib, o5) == to_affine(lib, o4)


@pytest.mark.parametrize("field,p", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_fp64_pipe_multiplier(lib, field, p):
    """csrc/fieldd.cuh: the DFMA-based Montgomery product (52-bit limbs in doubles, radix 2^260), host-emulated"""
    rng = random.Random(10 + field)
    r260_inv = pow(1 << 260, -1, p)
    vals = EDGE + [p - 1, p - 2, (1 << 52) - 1, (1 << 104) - 1, ((1 << 52) - 1) << 52, (1 << 253) + 12345] + \
        [rng.randrange(p) for _ in range(300)]
    for i, a in enumerate(vals):
        b = vals[(i * 5 + 1) % len(vals)]
        out = (ctypes.c_uint32 * 8)()
        assert lib.hs_fieldd_mul(field, limbs(a), limbs(b), out) == 0
        assert unlimbs(out) == a * b * r260_inv % p, (hex(a), hex(b))


def test_msm_signed_digit_slicing(lib):
    """csrc/msm_digits.cuh: for every window size the MSM can pick, the signed digits reconstruct the scalar,
    stay within [-2^(c-1), 2^(c-1)] and leave no carry -- including the field's edge values"""
    rng = random.Random(21)
    scalars = [0, 1, 2, O.R_MOD - 1, O.R_MOD - 2, (1 << 253), (1 << 254) - 1 - ((1 << 254) - O.R_MOD) - 1,
               (1 << 200) - 1, int("55" * 31, 16), int("aa" * 31, 16) % O.R_MOD] + [rng.randrange(O.R_MOD) for _ in range(60)]
    for c in range(4, 23):
        for s_ in scalars:
            digits = (ctypes.c_int32 * 64)()
            nw = ctypes.c_uint32(0)
            carry = lib.hs_msm_digits(limbs(s_), c, digits, ctypes.byref(nw))
            assert carry == 0, (c, hex(s_))
            assert nw.value == (256 + c - 1) // c
            ds = [digits[w] for w in range(nw.value)]
            assert all(abs(d) <= 1 << (c - 1) for d in ds), (c, hex(s_))
            assert sum(d << (c * w) for w, d in enumerate(ds)) == s_, (c, hex(s_))


def test_msm_affine_rounds_on_host(lib):
    """csrc/msm_affine.cuh: the per-thread bodies of the batched-affine bucket accumulation, run round by round on
    the CPU, against the oracle's affine group law -- bucket shapes 0..37, doubled points, opposite points,
    identities that travel through later rounds, and every (B, F) blocking incl. ones that split buckets"""
    rng = random.Random(99)
    base = [O.g1_multiply(O.G1, rng.randrange(1, O.R_MOD)) for _ in range(24)]
    table = (ctypes.c_uint32 * (16 * len(base)))()
    for i, (x, y) in enumerate(base):
        for k in range(8):
            table[16 * i + k] = (x >> (32 * k)) & 0xFFFFFFFF
            table[16 * i + 8 + k] = (y >> (32 * k)) & 0xFFFFFFFF
    P, N = (lambda i: (i, 0)), (lambda i: (i, 1))  # entry = (table index, negated?)
    buckets = [
        [], [P(0)], [P(1), P(2)], [P(3), P(3)], [P(4), N(4)], [P(5), N(5), P(6)], [P(6), P(7), P(8), N(8)],
        [P(9), N(9), P(10), N(10)], [P(11)] * 7, [N(12)] * 8, [P(13), P(13), N(13), N(13), P(14)],
        [], [], [P(1), N(1), P(1), N(1), P(1), P(2), N(2), P(3), P(3)],
    ]
    for size in (3, 5, 16, 31, 37):
        buckets.append([(rng.randrange(len(base)), rng.randrange(2)) for _ in range(size)])
    buckets += [[], [N(23)]]
    entries, offsets = [], [0]
    for bk in buckets:
        entries += [i | (s << 31) for i, s in bk]
        offsets.append(len(entries))
    expect = []
    for bk in buckets:
        acc = None
        for i, s in bk:
            acc = O.g1_add(acc, O.g1_neg(base[i]) if s else base[i])
        expect.append(acc)
    nb = len(buckets)
    sorted_arr = (ctypes.c_uint32 * max(1, len(entries)))(*entries)
    off_arr = (ctypes.c_uint32 * (nb + 1))(*offsets)
    for B, F in ((1, 1), (2, 3), (3, 64), (5, 2), (8, 8), (32, 32), (1000, 1)):
        out = (ctypes.c_uint32 * (16 * nb))()
        inf = (ctypes.c_uint8 * nb)()
        rounds = lib.hs_msm_affine_rounds(table, len(base), sorted_arr, off_arr, nb, B, F, out, inf)
        assert rounds == 6  # ceil(log2(37))
        for b in range(nb):
            got = None if inf[b] else (unlimbs(out, 2 * b), unlimbs(out, 2 * b + 1))
            assert got == expect[b], (B, F, b)
    # nothing to add at all: zero rounds, buckets read straight from the table with their signs
    off1 = (ctypes.c_uint32 * 4)(0, 1, 1, 2)
    ent1 = (ctypes.c_uint32 * 2)(5, 7 | (1 << 31))
    out = (ctypes.c_uint32 * 48)()
    inf = (ctypes.c_uint8 * 3)()
    assert lib.hs_msm_affine_rounds(table, len(base), ent1, off1, 3, 4, 4, out, inf) == 0
    assert (unlimbs(out, 0), unlimbs(out, 1)) == base[5] and inf[1] == 1
    assert (unlimbs(out, 4), unlimbs(out, 5)) == O.g1_neg(base[7])


@pytest.mark.parametrize("field,p", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_safegcd_inverse(lib, field, p):
    """csrc/modinv.cuh: inversion by batches of 30 Bernstein-Yang division steps on signed 30-bit limbs -- plain
    integers against pow(x, -1, p), and the Montgomery-form wrapper against the Fermat fp_inv it is meant to replace"""
    rng = random.Random(40 + field)
    vals = EDGE + [p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, R256 % p, 1 << 253, (1 << 253) + 1, 3 << 252 if (3 << 252) < p else 5,
                   (1 << 30) - 1, 1 << 30, (1 << 60) + 1, p - (1 << 30), p - (1 << 200)]
    vals += [rng.randrange(p) for _ in range(3000)]
    vals += [rng.randrange(1 << k) for k in range(1, 254, 7) for _ in range(4)]
    for a in vals:
        a %= p
        assert fop(lib, field, 10, a) == (pow(a, -1, p) if a else 0), a
    for a in vals[:400]:
        a %= p
        assert fop(lib, field, 9, a) == fop(lib, field, 4, a), a
