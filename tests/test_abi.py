"""CPU: the C-ABI library loads and exports every symbol include/plonk_b200.h declares (no compute calls
without a GPU), the ctypes binding covers them all, and the product path refuses to run without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "plonk_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from plonkathon_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: __graft_entry__.build()"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(raw, s), "missing export: " + s
    _lib.lib()
    assert set(_lib.EXPORTS) == set(syms), set(_lib.EXPORTS) ^ set(syms)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from plonkathon_b200 import _lib
    with pytest.raises(_lib.PlonkB200Error, match="no CUDA device"):
        _lib.Context(0)
    import plonkathon_b200 as pb
    with pytest.raises(_lib.PlonkB200Error):
        pb.Polynomial([pb.Scalar(1), pb.Scalar(2)], pb.Basis.MONOMIAL).fft()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "plonkathon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert not re.search(r'#include\s*[<"][^>"]*oracle', src), f


def test_host_side_objects():
    """Scalar / FQ semantics at the boundary (py_ecc FQ conventions the reference relies on)."""
    import plonkathon_b200 as pb
    S = pb.Scalar
    assert S(-1).n == pb.CURVE_ORDER - 1 and (S(3) / S(0)).n == 0 and (1 / S(2)) * 2 == 1
    assert S.root_of_unity(8) == 19540430494807482326159819597004422086093766032135589407132600596362845576832
    assert [r.n for r in S.roots_of_unity(4)][2] == pb.CURVE_ORDER - 1
    assert (pb.FQ(5), pb.FQ(7)) == (5, 7)
    p = pb.Polynomial([S(1), S(2)], pb.Basis.LAGRANGE) + pb.Polynomial([S(3), S(4)], pb.Basis.LAGRANGE)
    assert [v.n for v in p.values] == [4, 6]
    assert [v.n for v in (pb.Polynomial([S(1), S(2)], pb.Basis.MONOMIAL) + S(5)).values] == [6, 2]
    assert [v.n for v in pb.Polynomial([S(1), S(2), S(3), S(4)], pb.Basis.LAGRANGE).shift(1).values] == [2, 3, 4, 1]
    with pytest.raises(ValueError):
        pb.ec_lincomb([])


def test_transcript_is_host_code_and_matches_vectors():
    """The Merlin transcript inside the .so needs no GPU: check it here against Merlin's conformance vector,
    the reference-generated challenge fixture and the oracle's restatement on random schedules."""
    import json
    import random
    import plonkathon_b200 as pb
    from oracle import plonk_oracle as O
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "circuits.json")))
    t = pb.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == g["merlin_vector"]
    tr = pb.Transcript(b"plonk")
    tr.append_point(b"a_1", (pb.FQ(1), pb.FQ(2)))
    tr.append_scalar(b"a_eval", pb.Scalar(12345))
    assert tr.get_and_append_challenge(b"beta") == int(g["transcript"]["beta"])
    assert tr.get_and_append_challenge(b"gamma") == int(g["transcript"]["gamma"])
    rng = random.Random(4)
    a, b = pb.Transcript(b"plonk"), O.Transcript(b"plonk")
    for i in range(40):
        v = rng.randrange(O.R_MOD)
        a.append_scalar(b"s%d" % i, pb.Scalar(v))
        b.append_scalar(b"s%d" % i, v)
        assert a.get_and_append_challenge(b"c").n == b.get_and_append_challenge(b"c")
