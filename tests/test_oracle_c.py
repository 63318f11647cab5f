"""CPU: the C restatement (oracle/c/plonk_oracle.c) against the pinned Python oracle and the golden vectors."""
import random

import numpy as np

from oracle import c_oracle as C
from oracle import plonk_oracle as O
from tests.golden_io import PTAU_HEAD, ints, load_json, pt


def le(ints_):
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in ints_), dtype=np.uint8).reshape(-1, 32).copy()


def un(arr):
    raw = arr.tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def test_c_fft_matches_golden_and_python_oracle():
    for c in load_json("fft_vectors.json")["cases"]:
        v = ints(c["input"])
        assert un(C.fft(le(v))) == ints(c["fft"])
        assert un(C.fft(le(v), True)) == ints(c["ifft"])
    rng = random.Random(2)
    v = [rng.randrange(O.R_MOD) for _ in range(1 << 12)]
    assert un(C.fft(le(v))) == O.fft(v)


def test_c_lincomb_matches_golden_and_python_oracle():
    for c in load_json("lincomb_vectors.json")["cases"]:
        pairs = [(pt(p), int(s) % O.R_MOD) for p, s in zip(c["points"], c["scalars"]) if p is not None]
        if not pairs:
            continue
        pts = np.frombuffer(b"".join(p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little") for p, _ in pairs),
                            dtype=np.uint8).reshape(-1, 64)
        assert C.g1_lincomb(pts, le([s for _, s in pairs])) == O.ec_lincomb_naive(pairs), c["name"]
    s = O.Setup.from_file(PTAU_HEAD)
    rng = random.Random(3)
    n = 700
    sc = [rng.randrange(O.R_MOD) for _ in range(n)]
    pts = np.frombuffer(b"".join(p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little") for p in s.powers_of_x[:n]),
                        dtype=np.uint8).reshape(-1, 64)
    assert C.g1_lincomb(pts, le(sc)) == O.ec_lincomb(list(zip(s.powers_of_x[:n], sc)))


def test_c_eval_lagrange():
    rng = random.Random(4)
    v = [rng.randrange(O.R_MOD) for _ in range(256)]
    v[3] = 0
    x = rng.randrange(O.R_MOD)
    assert C.eval_lagrange(le(v), x) == O.eval_lagrange_at(v, x) == O.barycentric_eval(v, x)
