"""Array-level circuit construction for sizes the reference's front end cannot reach.

The reference compiler (compiler/program.py, compiler/utils.py:45-47) is O(n^2) and string-based; the
2^20 / 2^22-gate configurations are therefore built directly as arrays with the same conventions:
gate selectors as ``AssemblyEqn.gate()`` produces them (compiler/assembly.py:43-68), the copy-constraint
permutation exactly as ``Program.make_s_polynomials`` builds it (compiler/program.py:70-113: for every
variable, sort its cells by (row, column), and store at each cell the label w^row * column of the previous
cell of the cycle; all unused cells form one more cycle under the ``None`` variable).
tests/test_synthetic_vs_reference.py checks this against the reference compiler on the same wiring."""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np

from .field import CURVE_ORDER

R = CURVE_ORDER


@dataclass
class ArrayCircuit:
    group_order: int
    n_constraints: int
    # per-row wire variable ids (-1 == None), rows >= n_constraints are unused
    wire_L: np.ndarray
    wire_R: np.ndarray
    wire_O: np.ndarray
    # selector values per row (Python ints mod r)
    QL: list
    QR: list
    QM: list
    QO: list
    QC: list
    n_public: int
    values: list  # value of every variable id
    text: list  # the same circuit in the reference's constraint language (small sizes only)

    def wires_values(self):
        val = self.values
        get = lambda ids: [val[i] if i >= 0 else 0 for i in ids.tolist()]  # noqa: E731
        n, m = self.group_order, self.n_constraints
        pad = [0] * (n - m)
        return get(self.wire_L[:m]) + pad, get(self.wire_R[:m]) + pad, get(self.wire_O[:m]) + pad

    def public_values(self):
        return [self.values[i] for i in self.wire_L[:self.n_public].tolist()]


def roots_of_unity(n: int) -> list:
    w = pow(5, (R - 1) // n, R)
    out, cur = [], 1
    for _ in range(n):
        out.append(cur)
        cur = cur * w % R
    return out


def permutation_polys(wire_L, wire_R, wire_O, group_order: int, n_constraints: int):
    """S1, S2, S3 Lagrange values (lists of ints) -- compiler/program.py:70-113."""
    n = group_order
    roots = roots_of_unity(n)
    ids = np.full((n, 3), -1, dtype=np.int64)
    ids[:n_constraints, 0] = wire_L[:n_constraints]
    ids[:n_constraints, 1] = wire_R[:n_constraints]
    ids[:n_constraints, 2] = wire_O[:n_constraints]
    flat = ids.reshape(-1)  # cell index = row * 3 + (column - 1): already sorted by (row, column)
    order = np.argsort(flat, kind="stable")  # groups cells by variable, keeping (row, column) order
    sorted_ids = flat[order]
    # previous cell within each group (cyclically): S[cell] = label(previous cell of the same variable)
    start = np.ones(len(order), dtype=bool)
    start[1:] = sorted_ids[1:] != sorted_ids[:-1]
    prev = np.empty(len(order), dtype=np.int64)
    prev[1:] = order[:-1]
    group_start_pos = np.flatnonzero(start)
    group_end_pos = np.append(group_start_pos[1:], len(order)) - 1
    prev[group_start_pos] = order[group_end_pos]  # first cell of a cycle points at the last
    S = [[0] * n, [0] * n, [0] * n]
    for cell, pc in zip(order.tolist(), prev.tolist()):
        prow, pcol = divmod(pc, 3)
        row, col = divmod(cell, 3)
        S[col][row] = roots[prow] * (pcol + 1) % R
    return S[0], S[1], S[2]


def build_circuit(log_n: int, seed: int = 20260924, n_public: int = 2, fill: float = 1.0,
                  with_text: bool = False) -> ArrayCircuit:
    """Deterministic synthetic circuit with 2^log_n rows: ``n_public`` public-input rows, then a chain of
    multiplication / addition / add-constant gates whose operands are drawn from recently produced
    variables (so the permutation is non-trivial and the witness values are pseudo-random field elements)."""
    n = 1 << log_n
    rng = random.Random(seed)
    m = max(n_public + 1, int(n * fill))
    m = min(m, n)
    values = []
    wL = np.full(n, -1, dtype=np.int64)
    wR = np.full(n, -1, dtype=np.int64)
    wO = np.full(n, -1, dtype=np.int64)
    QL, QR, QM, QO, QC = ([0] * n for _ in range(5))
    text = []

    def name(i):
        return "v%d" % i

    # public rows: "x public" -> L = 1, O = 0 (compiler/assembly.py:160-164, 43-68)
    for i in range(n_public):
        values.append(rng.randrange(1, R))
        wL[i] = i
        QL[i] = 1
        if with_text:
            text.append("%s public" % name(i))
    # two private seeds
    for _ in range(2):
        values.append(rng.randrange(1, R))
    window = 64
    row = n_public
    first = True
    while row < m:
        nv = len(values)
        lo = max(0, nv - window)
        ia = rng.randrange(lo, nv)
        ib = rng.randrange(lo, nv)
        if first:  # make sure the private seeds are used so every variable appears in some cell
            ia, ib = n_public, n_public + 1
            first = False
        kind = rng.randrange(3)
        out = nv
        if kind == 0 or ia == ib:  # c <== a * b : M = -1, O = 1
            values.append(values[ia] * values[ib] % R)
            wL[row], wR[row], wO[row] = ia, ib, out
            QM[row], QO[row] = R - 1, 1
            if with_text:
                text.append("%s <== %s * %s" % (name(out), name(ia), name(ib)))
        elif kind == 1:  # c <== a + b : L = R = -1, O = 1
            values.append((values[ia] + values[ib]) % R)
            wL[row], wR[row], wO[row] = ia, ib, out
            QL[row], QR[row], QO[row] = R - 1, R - 1, 1
            if with_text:
                text.append("%s <== %s + %s" % (name(out), name(ia), name(ib)))
        else:  # c <== a + k : L = -1, C = -k, O = 1
            k = rng.randrange(1, 1 << 30)
            values.append((values[ia] + k) % R)
            # a single-variable expression puts the variable on both input wires (compiler/assembly.py:146-148)
            wL[row], wR[row], wO[row] = ia, ia, out
            QL[row], QC[row], QO[row] = R - 1, (R - k) % R, 1
            if with_text:
                text.append("%s <== %s + %d" % (name(out), name(ia), k))
        row += 1
    return ArrayCircuit(n, m, wL, wR, wO, QL, QR, QM, QO, QC, n_public, values, text)


def circuit_arrays(c: ArrayCircuit):
    """-> (pk dict of (n,32) uint8 arrays, A, B, C arrays, public list) ready for Prover.from_arrays /
    prove_arrays."""
    S1, S2, S3 = permutation_polys(c.wire_L, c.wire_R, c.wire_O, c.group_order, c.n_constraints)
    to_le = lambda ints: np.frombuffer(  # noqa: E731
        b"".join(int(x).to_bytes(32, "little") for x in ints), dtype=np.uint8).reshape(-1, 32).copy()
    pk = {"QM": to_le(c.QM), "QL": to_le(c.QL), "QR": to_le(c.QR), "QO": to_le(c.QO), "QC": to_le(c.QC),
          "S1": to_le(S1), "S2": to_le(S2), "S3": to_le(S3)}
    A, B, C = c.wires_values()
    return pk, to_le(A), to_le(B), to_le(C), c.public_values()
