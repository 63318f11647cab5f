// Modular inversion by the Bernstein-Yang "safegcd" division steps in batches of 30 on signed 30-bit limbs: the
// per-thread inversion of the batched-affine bucket accumulation (msm_bucket.cuh), where a Fermat chain (fp_inv:
// 380 dependent Montgomery products) would cost more than the 6 products per addition it amortises.
// One batch = 30 "half-delta" division steps on the low words (branch-free, ~17 ALU instructions per step) that
// produce a 2x2 transition matrix with entries below 2^30 in magnitude, which is then applied to the full-width
// (f, g) exactly and to (d, e) modulo p.  At most 20 batches for a 256-bit modulus (590-step bound); the loop leaves
// early once g == 0 (uniform enough across a warp: 16-18 batches for almost every BN254 input).  No chain of
// dependent 256-bit multiplications, no table, same instruction stream on every lane.
// Host/device code like field.cuh: unit-tested on the CPU against Python's pow(x, -1, p) (tests/test_host_arith.py).
#pragma once
#include "field.cuh"

namespace pb200 {

struct S30 {
  int32_t v[9];  // value = sum v[i] * 2^(30 i); v[0..7] in [0, 2^30) when normalised, v[8] carries the sign
};

#define PB_M30 0x3fffffff

// bits [30 i, 30 i + 30) of a 256-bit little-endian word array
PB_HD uint32_t limb30_of(const uint32_t* w, int i) {
  const int bit = 30 * i, word = bit >> 5, sh = bit & 31;
  uint64_t lo = w[word];
  uint64_t hi = word + 1 < 8 ? w[word + 1] : 0;
  return (uint32_t)(((lo | (hi << 32)) >> sh) & PB_M30);
}

template <class P>
struct ModInv30 {
  static PB_HD int32_t p30(int i) {
    uint32_t w[8];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = P::p(k);
    return (int32_t)limb30_of(w, i);
  }
  static PB_HD uint32_t pinv30() {  // p^-1 mod 2^30 (Newton on the low word; p is odd)
    const uint32_t p0 = P::p(0);
    uint32_t x = 1;
    for (int k = 0; k < 5; k++) x *= 2u - p0 * x;
    return x & PB_M30;
  }
};

template <class P>
PB_HD S30 s30_from_fp(const Fp<P>& a) {
  S30 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.v[i] = (int32_t)limb30_of(a.v, i);
  return r;
}
// normalised, non-negative, < 2^256
template <class P>
PB_HD Fp<P> s30_to_fp(const S30& a) {
  Fp<P> r;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // bits [32k, 32k+32): limb i = 32k / 30 and possibly the next one
    const int bit = 32 * k, i = bit / 30, sh = bit - 30 * i;
    uint64_t x = (uint64_t)(uint32_t)a.v[i] >> sh;
    x |= (uint64_t)(uint32_t)a.v[i + 1] << (30 - sh);
    if (i + 2 < 9) x |= (uint64_t)(uint32_t)a.v[i + 2] << (60 - sh);
    r.v[k] = (uint32_t)x;
  }
  return r;
}

struct Trans30 {
  int32_t u, v, q, r;  // (f, g) <- (u f + v g, q f + r g) / 2^30
};

// 30 division steps on the low words ("half-delta" variant: zeta = -(delta + 1/2), start zeta = -1; 590 steps
// suffice for any 256-bit modulus, i.e. 20 batches).  Branch-free, ~17 ALU instructions per step:
//   g odd and zeta < 0 : (f, g) <- (g, (g - f) / 2), zeta <- -zeta - 2
//   g odd              : g <- (g + f) / 2,           zeta <- zeta - 1
//   g even             : g <- g / 2,                 zeta <- zeta - 1
// The transition matrix t accumulates the same operations scaled by 2^30 (entries below 2^30 in magnitude).
PB_HD int32_t divsteps30(int32_t zeta, uint32_t f0, uint32_t g0, Trans30& t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;  // two's complement; read back as int32
  uint32_t f = f0, g = g0;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t m1 = (uint32_t)(zeta >> 31);  // all ones when zeta < 0
    const uint32_t m2 = 0u - (g & 1u);     // all ones when g is odd
    const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;  // +-f, +-u, +-v
    g += x & m2;
    q += y & m2;
    r += z & m2;
    m1 &= m2;
    zeta = (int32_t)(((uint32_t)zeta ^ m1) - 1u);
    f += g & m1;
    u += q & m1;
    v += r & m1;
    g >>= 1;  // only the low 30 - i bits of f and g are meaningful, the lost top bit is not one of them
    u <<= 1;
    v <<= 1;
  }
  t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
  return zeta;
}

// (f, g) <- t (f, g) / 2^30, exact
PB_HD void update_fg30(S30& f, S30& g, const Trans30& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  int64_t cf = u * f.v[0] + v * g.v[0];
  int64_t cg = q * f.v[0] + r * g.v[0];
  cf >>= 30;  // the low 30 bits are zero by construction
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cf += u * f.v[i] + v * g.v[i];
    cg += q * f.v[i] + r * g.v[i];
    f.v[i - 1] = (int32_t)(cf & PB_M30);
    g.v[i - 1] = (int32_t)(cg & PB_M30);
    cf >>= 30;
    cg >>= 30;
  }
  f.v[8] = (int32_t)cf;
  g.v[8] = (int32_t)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p, keeping both in (-2p, p)
template <class P>
PB_HD void update_de30(S30& d, S30& e, const Trans30& t) {
  const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
  const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;  // all ones when negative
  int32_t md = (t.u & sd) + (t.v & se);
  int32_t me = (t.q & sd) + (t.r & se);
  int64_t cd = u * d.v[0] + v * e.v[0];
  int64_t ce = q * d.v[0] + r * e.v[0];
  // multiples of p that clear the low 30 bits
  const uint32_t pinv = ModInv30<P>::pinv30();
  md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & PB_M30);
  me -= (int32_t)((pinv * (uint32_t)ce + (uint32_t)me) & PB_M30);
  cd += (int64_t)ModInv30<P>::p30(0) * md;
  ce += (int64_t)ModInv30<P>::p30(0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 9; i++) {
    cd += u * d.v[i] + v * e.v[i] + (int64_t)ModInv30<P>::p30(i) * md;
    ce += q * d.v[i] + r * e.v[i] + (int64_t)ModInv30<P>::p30(i) * me;
    d.v[i - 1] = (int32_t)(cd & PB_M30);
    e.v[i - 1] = (int32_t)(ce & PB_M30);
    cd >>= 30;
    ce >>= 30;
  }
  d.v[8] = (int32_t)cd;
  e.v[8] = (int32_t)ce;
}

PB_HD void s30_propagate(S30& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int32_t carry = a.v[i] >> 30;  // arithmetic: -1 for a negative limb
    a.v[i] &= PB_M30;
    a.v[i + 1] += carry;
  }
}

// x^-1 mod p for a plain integer x < p (NOT a Montgomery operation); 0 -> 0
template <class P>
PB_HD Fp<P> fp_inv_plain_gcd(const Fp<P>& x) {
  S30 f, g = s30_from_fp(x), d, e;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    f.v[i] = ModInv30<P>::p30(i);
    d.v[i] = 0;
    e.v[i] = 0;
  }
  e.v[0] = 1;
  int32_t eta = -1;
  for (int batch = 0; batch < 20; batch++) {  // 20 * 30 >= 590 half-delta division steps: enough for any 256-bit input
    int32_t gz = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) gz |= g.v[i];
    if (gz == 0) break;
    Trans30 t;
    const uint32_t f0 = (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30);
    const uint32_t g0 = (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30);
    eta = divsteps30(eta, f0, g0, t);
    update_de30<P>(d, e, t);
    update_fg30(f, g, t);
  }
  // f = +-1 (or p when x == 0, in which case d == 0): result = sign(f) * d, brought into [0, p)
  const int32_t fneg = f.v[8] >> 31;
  int32_t m = d.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] += ModInv30<P>::p30(i) & m;
  s30_propagate(d);
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] = (d.v[i] ^ fneg) - fneg;
  s30_propagate(d);
  m = d.v[8] >> 31;
#pragma unroll
  for (int i = 0; i < 9; i++) d.v[i] += ModInv30<P>::p30(i) & m;
  s30_propagate(d);
  return s30_to_fp<P>(d);
}

// Montgomery-form inverse with the same contract as fp_inv: a R -> a^-1 R, inv(0) == 0
template <class P>
PB_HD Fp<P> fp_inv_gcd(const Fp<P>& a) {
  return fp_mul(fp_inv_plain_gcd(a), Fp<P>::r3());  // (a R)^-1 R^3 R^-1 = a^-1 R
}

}  // namespace pb200
