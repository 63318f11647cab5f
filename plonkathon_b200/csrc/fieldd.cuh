// PROTOTYPE -- not used by the shipped kernels (see profiles/r01_pipe_microbench.md for why).
// FP64-pipe field arithmetic for BN254 Fq/Fr: a second Montgomery multiplier that runs on the DFMA pipe, which
// is idle next to the integer multiplier and issues 64 DFMA/clk/SM concurrently with it.  Correct on host
// (tests/test_host_arith.py) and device (tools/ubench/modmul_d.cu: 0 mismatches in 2M products), but at 400 SASS
// instructions per product it only matches the IMAD multiplier (60 vs 66 G products/s) and mixing the two gains
// 2 %: it needs chained-FMA accumulation (far fewer integer adds) before it pays.
//
// Representation: 5 limbs of 52 bits, each held as an exact integer in a double; Montgomery radix R' = 2^260.
// Limb products use the classic two-FMA exact split (round-toward-zero):
//     h = fma_rz(a, b, 2^104)              -> 2^104 + floor(ab / 2^52) * 2^52
//     l = fma_rz(a, b, (2^104 + 2^52) - h) -> 2^52 + (ab mod 2^52)
// so the mantissa fields of h and l are the high and low 52-bit halves of the 104-bit product.  The halves are
// summed per column as raw 64-bit integers (exponent fields included; their totals are compile-time constants
// folded into the initial column values).  Reduction is word-serial Montgomery in base 2^52.
// Host builds emulate fma_rz with fesetround so the same code is unit-tested on the CPU.
#pragma once
#include <stdint.h>
#include "field.cuh"
#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#include <cstring>
#endif

namespace pb200 {

#if defined(__CUDA_ARCH__)
PB_D double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
PB_D int64_t d2bits(double x) { return __double_as_longlong(x); }
PB_D double bits2d(int64_t x) { return __longlong_as_double(x); }
#else
inline double fma_rz(double a, double b, double c) {
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile double r = std::fma(a, b, c);
  fesetround(old);
  return r;
}
inline int64_t d2bits(double x) { int64_t r; memcpy(&r, &x, 8); return r; }
inline double bits2d(int64_t x) { double r; memcpy(&r, &x, 8); return r; }
#endif

#define PB_D52 4503599627370496.0                      /* 2^52 */
#define PB_C1 20282409603651670423947251286016.0       /* 2^104 */
#define PB_C2 20282409603651674927546878656512.0       /* 2^104 + 2^52 */
#define PB_MASK52 0xFFFFFFFFFFFFFULL
#define PB_EXP_LO 0x4330000000000000LL                 /* bit pattern of 2^52  (exponent of the low halves) */
#define PB_EXP_HI 0x4670000000000000LL                 /* bit pattern of 2^104 (exponent of the high halves) */

// exact conversion of an integer < 2^52 to double and back
PB_HD double u52_to_d(uint64_t x) { return bits2d((int64_t)(x | (uint64_t)PB_EXP_LO)) - PB_D52; }
PB_HD uint64_t d_to_u52(double d) { return (uint64_t)d2bits(d + PB_D52) & PB_MASK52; }

template <class P>
struct FpDConst {
  // limbs of p in base 2^52, -p^-1 mod 2^52, and 2^256 mod p (to move between the two Montgomery radices)
  static PB_HD uint64_t p52(int i) {
    // limbs recomputed from the 32-bit constants (constant-folded by the compiler)
    uint64_t w0 = (uint64_t)P::p(0) | ((uint64_t)P::p(1) << 32), w1 = (uint64_t)P::p(2) | ((uint64_t)P::p(3) << 32);
    uint64_t w2 = (uint64_t)P::p(4) | ((uint64_t)P::p(5) << 32), w3 = (uint64_t)P::p(6) | ((uint64_t)P::p(7) << 32);
    return i == 0 ? (w0 & PB_MASK52)
         : i == 1 ? (((w0 >> 52) | (w1 << 12)) & PB_MASK52)
         : i == 2 ? (((w1 >> 40) | (w2 << 24)) & PB_MASK52)
         : i == 3 ? (((w2 >> 28) | (w3 << 36)) & PB_MASK52)
                  : (w3 >> 16);
  }
  static PB_HD uint64_t np52() {  // -p^-1 mod 2^52 by Newton iteration on the low limb
    uint64_t p0 = p52(0), x = 1;
    for (int k = 0; k < 6; k++) x *= 2 - p0 * x;
    return ((uint64_t)0 - x) & PB_MASK52;
  }
};

template <class P>
struct FpD {
  double v[5];
};

// 8 x u32 limbs -> 5 x 52-bit limbs as doubles (pure re-slicing, the VALUE is unchanged)
template <class P>
PB_HD FpD<P> fpd_from_u32(const Fp<P>& a) {
  uint64_t w0 = (uint64_t)a.v[0] | ((uint64_t)a.v[1] << 32), w1 = (uint64_t)a.v[2] | ((uint64_t)a.v[3] << 32);
  uint64_t w2 = (uint64_t)a.v[4] | ((uint64_t)a.v[5] << 32), w3 = (uint64_t)a.v[6] | ((uint64_t)a.v[7] << 32);
  FpD<P> r;
  r.v[0] = u52_to_d(w0 & PB_MASK52);
  r.v[1] = u52_to_d(((w0 >> 52) | (w1 << 12)) & PB_MASK52);
  r.v[2] = u52_to_d(((w1 >> 40) | (w2 << 24)) & PB_MASK52);
  r.v[3] = u52_to_d(((w2 >> 28) | (w3 << 36)) & PB_MASK52);
  r.v[4] = u52_to_d(w3 >> 16);
  return r;
}
template <class P>
PB_HD Fp<P> fpd_to_u32(const FpD<P>& a) {
  uint64_t l0 = d_to_u52(a.v[0]), l1 = d_to_u52(a.v[1]), l2 = d_to_u52(a.v[2]), l3 = d_to_u52(a.v[3]), l4 = d_to_u52(a.v[4]);
  uint64_t w0 = l0 | (l1 << 52), w1 = (l1 >> 12) | (l2 << 40), w2 = (l2 >> 24) | (l3 << 28), w3 = (l3 >> 36) | (l4 << 16);
  Fp<P> r;
  r.v[0] = (uint32_t)w0; r.v[1] = (uint32_t)(w0 >> 32); r.v[2] = (uint32_t)w1; r.v[3] = (uint32_t)(w1 >> 32);
  r.v[4] = (uint32_t)w2; r.v[5] = (uint32_t)(w2 >> 32); r.v[6] = (uint32_t)w3; r.v[7] = (uint32_t)(w3 >> 32);
  return r;
}

// one exact 52x52 -> (hi, lo) product, accumulated as raw bit patterns into two 64-bit column sums
#define PB_DMAC(HI_COL, LO_COL, A, B)                         \
  {                                                           \
    double h__ = fma_rz((A), (B), PB_C1);                     \
    double l__ = fma_rz((A), (B), PB_C2 - h__);               \
    (HI_COL) += (uint64_t)d2bits(h__);                        \
    (LO_COL) += (uint64_t)d2bits(l__);                        \
  }

// Montgomery product in base 2^52: returns a * b * 2^-260 mod p, fully reduced, for a, b < p
template <class P>
PB_HD FpD<P> fpd_mul(const FpD<P>& a, const FpD<P>& b) {
  typedef FpDConst<P> K;
  // column sums V[0..9]; every added bit pattern carries its exponent field, so start each column at minus the
  // total of the exponent fields it is going to receive (computed mod 2^64):
  //   column k receives lo halves of the products with i + j == k      (a*b: nlo_ab(k); reduction: nlo_red(k))
  //   and hi halves of the products with i + j == k - 1.
  uint64_t V[10];
#pragma unroll
  for (int k = 0; k < 10; k++) {
    int nlo_ab = (k <= 8) ? ((k < 5 ? k : 8 - k) + 1) : 0;                    // products of a*b in column k
    int nhi_ab = (k >= 1) ? (((k - 1) < 5 ? (k - 1) : 8 - (k - 1)) + 1) : 0;  // ... in column k-1
    if (k - 1 > 8) nhi_ab = 0;
    int nlo_red = (k <= 8) ? ((k < 5 ? k : 8 - k) + 1) : 0;                   // q_i * p_j with i + j == k: same counts
    int nhi_red = nhi_ab;
    V[k] = (uint64_t)0 - ((uint64_t)(nlo_ab + nlo_red) * (uint64_t)PB_EXP_LO + (uint64_t)(nhi_ab + nhi_red) * (uint64_t)PB_EXP_HI);
  }
#pragma unroll
  for (int i = 0; i < 5; i++) {
#pragma unroll
    for (int j = 0; j < 5; j++) PB_DMAC(V[i + j + 1], V[i + j], a.v[i], b.v[j]);
  }
  const double npd = u52_to_d(K::np52());
  uint64_t carry = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    // Column i now holds every contribution except the low half of q_i * p_0 (about to be computed), whose
    // exponent field was pre-subtracted above: add it back to read the true partial column value.
    uint64_t t = V[i] + carry + (uint64_t)PB_EXP_LO;
    double td = u52_to_d(t & PB_MASK52);
    double qh = fma_rz(td, npd, PB_C1);
    double ql = fma_rz(td, npd, PB_C2 - qh);
    double qd = ql - PB_D52;  // q = (t * np) mod 2^52, exact integer in a double
#pragma unroll
    for (int j = 0; j < 5; j++) PB_DMAC(V[i + j + 1], V[i + j], qd, u52_to_d(K::p52(j)));
    // column i is now complete and divisible by 2^52
    uint64_t done = V[i] + carry;
    carry = done >> 52;
  }
  FpD<P> r;
  uint64_t limb[5];
#pragma unroll
  for (int k = 0; k < 5; k++) {
    uint64_t s = V[5 + k] + carry;
    limb[k] = s & PB_MASK52;
    carry = s >> 52;
  }
  // conditional subtraction of p (result < 2p)
  uint64_t d[5];
  int64_t borrow = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    int64_t x = (int64_t)limb[k] - (int64_t)K::p52(k) + borrow;
    d[k] = (uint64_t)x & PB_MASK52;
    borrow = x >> 63;  // -1 when negative
  }
#pragma unroll
  for (int k = 0; k < 5; k++) r.v[k] = u52_to_d(borrow ? limb[k] : d[k]);
  return r;
}

}  // namespace pb200
