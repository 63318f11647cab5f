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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_sharded_division_and_grand_product(world):
    """Rounds 2 and 5 across ranks (csrc/prover.cu, world > 1).  Division by (X - z) in coefficient space is a weighted
    suffix sum q_(k-1) = z^-k sum_(m >= k) N_m z^m: a rank handles a slab of m, its carry is the sum of the slabs
    above, the top element of its output slab is carry * z^-(hi).  The grand product Z_(i+1) = Z_i f_i is an exclusive
    prefix product: a rank's carry is the product of the slabs below."""
    rng = random.Random(31 * world)
    n = 32
    z = rng.randrange(2, P)
    q = [rng.randrange(P) for _ in range(n - 1)]           # quotient, degree n - 2
    N = [0] * n                                             # N = q * (X - z): exactly divisible
    for i, c in enumerate(q):
        N[i + 1] = (N[i + 1] + c) % P
        N[i] = (N[i] - c * z) % P
    zinv = pow(z, -1, P)
    ns = n // world
    u = [N[m] * pow(z, m, P) % P for m in range(n)]
    totals = [sum(u[r * ns:(r + 1) * ns]) % P for r in range(world)]
    assert sum(totals) % P == 0                             # the remainder N(z)
    out = [None] * n
    for r in range(world):
        lo, hi = r * ns, (r + 1) * ns
        carry = sum(totals[r + 1:]) % P
        run = carry
        for m in range(hi - 1, lo - 1, -1):                 # k_sufsum_apply on the slab, from the top
            run = (run + u[m]) % P
            if m - lo >= 1:
                out[m - 1] = run * pow(zinv, m, P) % P
        out[hi - 1] = carry * pow(zinv, hi, P) % P
    assert out == q + [0]
    f = [rng.randrange(1, P) for _ in range(n)]
    Z = [1]
    for x in f[:-1]:
        Z.append(Z[-1] * x % P)
    prods = []
    for r in range(world):
        t = 1
        for x in f[r * ns:(r + 1) * ns]:
            t = t * x % P
        prods.append(t)
    got = []
    for r in range(world):
        run = 1
        for t in prods[:r]:
            run = run * t % P                               # k_prod_carry
        for x in f[r * ns:(r + 1) * ns]:                    # k_prod_apply: exclusive prefix inside the slab
            got.append(run)
            run = run * x % P
    assert got == Z
