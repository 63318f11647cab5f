"""CPU, build container only: the array-level circuit builder (plonkathon_b200/synthetic.py) against the
reference's own compiler on the same wiring -- selectors, permutation polynomials, wire values and
public inputs must agree element for element."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")

_SCRIPT = r'''
import sys
sys.path.insert(0, %(shims)r); sys.path.insert(0, %(ref)r); sys.path.insert(0, %(root)r)
from compiler.program import Program
from plonkathon_b200 import synthetic as syn
for log_n, fill in ((3, 1.0), (5, 1.0), (6, 0.8)):
    c = syn.build_circuit(log_n, seed=log_n, n_public=2, fill=fill, with_text=True)
    prog = Program(c.text, 1 << log_n)
    pk = prog.common_preprocessed_input()
    S1, S2, S3 = syn.permutation_polys(c.wire_L, c.wire_R, c.wire_O, c.group_order, c.n_constraints)
    g = lambda p: [x.n for x in p.values]
    assert g(pk.QL) == c.QL and g(pk.QR) == c.QR and g(pk.QM) == c.QM and g(pk.QO) == c.QO and g(pk.QC) == c.QC
    assert g(pk.S1) == S1 and g(pk.S2) == S2 and g(pk.S3) == S3
    names = {("v%%d" %% i): v for i, v in enumerate(c.values)}
    names[None] = 0
    A, B, C = c.wires_values()
    m = c.n_constraints
    assert [names[w.L] for w in prog.wires()] == A[:m]
    assert [names[w.R] for w in prog.wires()] == B[:m]
    assert [names[w.O] for w in prog.wires()] == C[:m]
    assert [names[v] for v in prog.get_public_assignments()] == c.public_values()
print("OK")
'''


def test_builder_matches_reference_compiler():
    code = _SCRIPT % {"shims": os.path.join(ROOT, "oracle", "shims"), "ref": REF, "root": ROOT}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stderr[-3000:]
