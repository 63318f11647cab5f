"""Generates tests/golden/proof_2p20.json: the oracle's proof of the 2^20-gate synthetic circuit that
tests/test_gpu_parity.py::test_prove_2p20_gates_verifies proves on the GPU (seed 7, two public inputs, structured SRS
with the test tau), so that the GPU proof can be compared BYTE FOR BYTE at BASELINE.json's headline size.

The pure-Python oracle would need hours; this uses oracle/fast.py (same prover code, fft / ec_lincomb / SRS from the
C restatement; equivalence checked in tests/test_oracle_fast.py).  One core, 34 minutes, ~7 GB:

    python tests/golden/make_proof_2p20.py
    GOLDEN_SEED=20260924 python tests/golden/make_proof_2p20.py     # bench.py's circuit -> proof_2p20_seed20260924.json
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fast as F  # noqa: E402
from oracle import plonk_oracle as O  # noqa: E402
from plonkathon_b200 import synthetic as syn  # noqa: E402

TAU = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF
LOG_N, SEED, N_PUBLIC = int(os.environ.get("GOLDEN_LOG_N", "20")), int(os.environ.get("GOLDEN_SEED", "7")), 2
t0 = time.time()


def log(msg):
    print("[%7.1f s] %s" % (time.time() - t0, msg), flush=True)


c = syn.build_circuit(LOG_N, seed=SEED, n_public=N_PUBLIC)
n = c.group_order
S1, S2, S3 = syn.permutation_polys(c.wire_L, c.wire_R, c.wire_O, n, c.n_constraints)
pk = O.Preprocessed(n, c.QM, c.QL, c.QR, c.QO, c.QC, S1, S2, S3)
A, B, C = c.wires_values()
log("circuit built")
setup = F.Setup(TAU, n)
log("SRS generated")
proof = F.prove(setup, pk, A, B, C, c.public_values())
raw = O.proof_bytes(proof)
log("proof done")
with F.c_kernels():
    vk = {name: setup.commit(col) for name, col in (("Qm", c.QM), ("Ql", c.QL), ("Qr", c.QR), ("Qo", c.QO), ("Qc", c.QC),
                                                     ("S1", S1), ("S2", S2), ("S3", S3))}
log("verification key done")
assert O.verify_proof_trapdoor(n, vk, O.proof_from_bytes(raw), c.public_values(), TAU)
rec = {"log_n": LOG_N, "seed": SEED, "n_public": N_PUBLIC, "tau": hex(TAU), "public": [str(x) for x in c.public_values()],
       "sha256": hashlib.sha256(raw).hexdigest(), "proof_hex": raw.hex(),
       "vk": {k: [str(v[0]), str(v[1])] for k, v in vk.items()},
       "generator": "tests/golden/make_proof_2p20.py (oracle/fast.py: plonk_oracle.Prover over the C restatement)",
       "seconds": round(time.time() - t0, 1)}
name = "proof_2p%d" % LOG_N + ("" if SEED == 7 else "_seed%d" % SEED)
out = os.path.join(HERE, name + ".json")
json.dump(rec, open(out, "w"), indent=1)
log("wrote " + out + " sha256 " + rec["sha256"])
