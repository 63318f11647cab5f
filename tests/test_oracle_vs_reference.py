"""CPU, build container only: the oracle restatement vs the reference's own unmodified
modules imported from /root/reference over oracle/shims.  Skipped where the reference tree
does not exist (the GPU box)."""
import os
import random
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

_SCRIPT = r'''
import sys, random
sys.path.insert(0, %(shims)r); sys.path.insert(0, %(ref)r); sys.path.insert(0, %(root)r)
import curve, poly
from oracle import plonk_oracle as O
rng = random.Random(7)
S = curve.Scalar
for logn in (0, 1, 2, 5, 9):
    n = 1 << logn
    v = [rng.randrange(O.R_MOD) for _ in range(n)]
    assert [x.n for x in poly.Polynomial([S(a) for a in v], poly.Basis.MONOMIAL).fft().values] == O.fft(v)
    assert [x.n for x in poly.Polynomial([S(a) for a in v], poly.Basis.LAGRANGE).ifft().values] == O.ifft(v)
assert [x.n for x in S.roots_of_unity(16)] == O.roots_of_unity(16)
import py_ecc.bn128 as b
pts = [b.multiply(b.G1, rng.randrange(1, 1000)) for _ in range(20)]
sc = [rng.randrange(O.R_MOD) for _ in range(20)]
ref = curve.ec_lincomb(list(zip(pts, sc)))
ora = O.ec_lincomb([((p[0].n, p[1].n), s) for p, s in zip(pts, sc)])
assert (ref[0].n, ref[1].n) == ora
# the mock-adder self-test of curve.py:115-142 against the restated lincomb
nums = [rng.randrange(10**20) for _ in range(40)]
fac = [rng.randrange(2**256) for _ in range(40)]
assert O.lincomb(nums, fac, lambda x, y: x + y, 0) == curve.lincomb(nums, fac) == sum(a*f for a, f in zip(nums, fac))
print("OK")
'''


def test_restatement_matches_reference_modules():
    code = _SCRIPT % {"shims": os.path.join(ROOT, "oracle", "shims"), "ref": REF, "root": ROOT}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-2000:]
