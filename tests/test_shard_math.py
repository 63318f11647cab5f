"""CPU: the decompositions behind the one-proof-across-GPUs path (csrc/prover.cu with world > 1, csrc/ntt_shard.cuh),
restated with the oracle's transforms -- every identity the sharded prover relies on, for 2, 4 and 8 ranks."""
import random

import pytest

from oracle import plonk_oracle as O

P = O.R_MOD


def _poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


@pytest.mark.parametrize("world", [2, 4, 8])
def test_coset_slice_extension_is_local(world):
    """Rank r's slice of the 4n coset g<mu> is {g mu^(world k + r)}: a coset of the subgroup of order 4n / world.
    Evaluating an n-coefficient polynomial there = multiply coefficient i by (g mu^r)^i, wrap around the slice length
    (world = 8: the slice is shorter than the polynomial), one forward transform of the slice length
    (prover.cu coset_extend / ntt.cu fold)."""
    rng = random.Random(world)
    n, g = 16, 5
    mu = O.root_of_unity(4 * n)
    ne = 4 * n // world
    coeffs = [rng.randrange(P) for _ in range(n)]
    full = [_poly_eval(coeffs, g * pow(mu, j, P) % P) for j in range(4 * n)]
    for r in range(world):
        shift = g * pow(mu, r, P) % P
        scaled = [c * pow(shift, i, P) % P for i, c in enumerate(coeffs)]
        folded = [0] * ne
        for i, v in enumerate(scaled):
            folded[i % ne] = (folded[i % ne] + v) % P
        assert O.fft(folded) == full[r::world]
        # Z(w x) on the slice: the same coefficients on the slice shifted by w = mu^4 -- local index + 4 / world when
        # world divides 4, otherwise (world = 8) the slice of rank r + 4, extended separately with shift * w
        zw = [_poly_eval(coeffs, g * pow(mu, j + 4, P) % P) for j in range(4 * n)]
        if 4 % world == 0:
            sh = 4 // world
            assert [full[r::world][(k + sh) % ne] for k in range(ne)] == zw[r::world]
        else:
            shift_w = shift * pow(mu, 4, P) % P
            folded = [0] * ne
            for i, c in enumerate(coeffs):
                folded[i % ne] = (folded[i % ne] + c * pow(shift_w, i, P)) % P
            assert O.fft(folded) == zw[r::world]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_quotient_inverse_transform_join(world):
    """Round 3's way back (prover.py:205-226 across ranks): every rank holds T on its slice; local inverse transform,
    store multiplier mu^(-r j0) / world, ONE allgather, then per j0 a world-point inverse DFT over the rank index and
    g^-j on the way out give the 4n coefficients (the top n are zero for a valid quotient)."""
    rng = random.Random(7 * world)
    n, g = 8, 5
    N = 4 * n
    mu = O.root_of_unity(N)
    ne = N // world
    t = [rng.randrange(P) for _ in range(3 * n)] + [0] * n
    evals = [_poly_eval(t, g * pow(mu, j, P) % P) for j in range(N)]
    mu_inv, inv_world = pow(mu, -1, P), pow(world, -1, P)
    U = []
    for r in range(world):
        y = O.fft(evals[r::world], inv=True)
        U.append([v * pow(mu_inv, r * j0, P) * inv_world % P for j0, v in enumerate(y)])
    wg_inv = pow(mu_inv, ne, P)
    got = [0] * N
    for j0 in range(ne):
        for j1 in range(world):
            v = sum(U[r][j0] * pow(wg_inv, r * j1, P) for r in range(world)) % P
            j = j0 + ne * j1
            got[j] = v * pow(g, -j, P) % P
    assert got == t


@pytest.mark.parametrize("world", [2, 4, 8])
def test_small_dft_butterflies_match_definition(world):
    """ntt_shard.cuh small_dft: bit-reversed input, radix-2 DIT stages with w^(j G / 2 half)"""
    rng = random.Random(world)
    w = pow(O.root_of_unity(64), 64 // world, P)
    x = [rng.randrange(P) for _ in range(world)]
    lg = world.bit_length() - 1
    y = [x[int(format(i, "0%db" % lg)[::-1], 2)] for i in range(world)]
    for s in range(lg):
        half = 1 << s
        for i in range(0, world, 2 * half):
            for j in range(half):
                v = y[i + j + half] * pow(w, j * (world // (2 * half)), P) % P
                u = y[i + j]
                y[i + j], y[i + j + half] = (u + v) % P, (u - v) % P
    assert y == [sum(x[r] * pow(w, r * k, P) for r in range(world)) % P for k in range(world)]
