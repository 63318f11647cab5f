"""CPU: oracle/fast.py (the oracle prover with its fft and ec_lincomb answered by the C restatement, and a
C-generated structured SRS) returns byte-for-byte the proofs of the pure-Python oracle -- the link that lets the
golden 2^20-gate proof (tests/golden/proof_2p20.json) stand for the oracle at a size it cannot reach itself."""
import pytest

from oracle import fast as F
from oracle import plonk_oracle as O
from plonkathon_b200 import synthetic as syn

TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF


@pytest.mark.parametrize("log_n,seed,n_public", [(3, 1, 1), (5, 2, 2), (6, 7, 0)])
def test_fast_oracle_equals_python_oracle(log_n, seed, n_public):
    c = syn.build_circuit(log_n, seed=seed, n_public=n_public)
    n = c.group_order
    S1, S2, S3 = syn.permutation_polys(c.wire_L, c.wire_R, c.wire_O, n, c.n_constraints)
    pk = O.Preprocessed(n, c.QM, c.QL, c.QR, c.QO, c.QC, S1, S2, S3)
    A, B, C = c.wires_values()
    fsetup = F.Setup(TAU, n)
    pts, cur = [], O.G1
    for i in range(n):
        assert fsetup.point(i) == cur
        pts.append(cur)
        cur = O.g1_multiply(cur, TAU)
    slow = O.Prover(O.Setup(pts, None), pk).prove(A, B, C, c.public_values())
    fast = F.prove(fsetup, pk, A, B, C, c.public_values())
    assert O.proof_bytes(fast) == O.proof_bytes(slow)
    assert O.fft is not F.fft  # the patch is undone


@pytest.mark.parametrize("name", ["proof_2p20.json", "proof_2p20_seed20260924.json", "proof_2p22.json"])
def test_golden_2p20_record_is_consistent(name):
    """the committed golden proofs (the GPU test's circuit, seed 7, and bench.py's, seed 20260924): hash matches the bytes, and the proof verifies under the reference's verification
    equation through the known tau (cheap: a handful of scalar multiplications)"""
    import hashlib
    import json
    import os
    from tests.golden_io import GOLDEN
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("golden 2^20 proof not generated yet")
    rec = json.load(open(path))
    raw = bytes.fromhex(rec["proof_hex"])
    assert len(raw) == 768 and hashlib.sha256(raw).hexdigest() == rec["sha256"]
    log_n = rec["log_n"]
    assert log_n == int(name.split("_")[1][2:].split(".")[0]) and int(rec["tau"], 16) == TAU
    vk = {k: tuple(int(x) for x in v) for k, v in rec["vk"].items()}
    assert O.verify_proof_trapdoor(1 << log_n, vk, O.proof_from_bytes(raw), [int(x) for x in rec["public"]], TAU)
