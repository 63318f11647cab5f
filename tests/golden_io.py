"""Helpers to read the committed golden fixtures (tests/golden/, written by make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PTAU_HEAD = os.path.join(GOLDEN, "powersOfTau28_hez_final_11.head.ptau")


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def ints(strs):
    return [int(s) for s in strs]


def pt(p):
    return None if p is None else (int(p[0]), int(p[1]))


def le_to_ints(arr):
    raw = np.ascontiguousarray(arr).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def load_circuit(name):
    """-> (entry dict from circuits.json, {array name: list[int]})."""
    entry = load_json("circuits.json")["circuits"][name]
    z = np.load(os.path.join(GOLDEN, "circuit_%s.npz" % name))
    return entry, {k: le_to_ints(z[k]) for k in z.files}


def proof_from_entry(entry):
    return {k: (pt(v) if isinstance(v, list) else int(v)) for k, v in entry["proof"].items()}
