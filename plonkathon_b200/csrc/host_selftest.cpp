// Host build of the limb-level arithmetic in field.cuh / curve.cuh (same code path as the device,
// PTX carry-chain primitives replaced by their emulation).  TEST INFRASTRUCTURE: loaded only by
// tests/test_host_arith.py through ctypes; never linked into libplonk_b200.so.
#include "field.cuh"
#include "curve.cuh"
#include "fieldd.cuh"
#include "msm_digits.cuh"
#include <cstring>
using namespace pb200;

template <class F> static F ld(const uint32_t* p) { F r; memcpy(r.v, p, 32); return r; }
template <class F> static void st(uint32_t* p, const F& a) { memcpy(p, a.v, 32); }

extern "C" {
// op: 0 add, 1 sub, 2 mul, 3 neg, 4 inv, 5 to_mont, 6 from_mont, 7 dbl, 8 sqr ; field: 0 Fr, 1 Fq
int hs_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
#define RUN(F)                                                   \
  {                                                              \
    F x = ld<F>(a), y = ld<F>(b), r;                             \
    switch (op) {                                                \
      case 0: r = fp_add(x, y); break;                           \
      case 1: r = fp_sub(x, y); break;                           \
      case 2: r = fp_mul(x, y); break;                           \
      case 3: r = fp_neg(x); break;                              \
      case 4: r = fp_inv(x); break;                              \
      case 5: r = fp_to_mont(x); break;                          \
      case 6: r = fp_from_mont(x); break;                        \
      case 7: r = fp_dbl(x); break;                              \
      case 8: r = fp_sqr(x); break;                              \
      default: return -1;                                        \
    }                                                            \
    st(out, r);                                                  \
  }
  if (field == 0) RUN(Fr) else RUN(Fq)
  return 0;
}

// signed digits of a canonical scalar for window size c: writes W = ceil(256 / c) digits (sign * magnitude) and
// returns the carry left after the last window (must be 0 for scalars < r)
int hs_msm_digits(const uint32_t* scalar, uint32_t c, int32_t* digits, uint32_t* n_windows) {
  MsmGeom g;
  g.c = c;
  g.W = (256 + c - 1) / c;
  g.half = 1u << (c - 1);
  g.bucket_stride = g.half;
  g.point_stride = 0;
  g.nb = g.half * g.W;
  g.batch = 1;
  Fr s = ld<Fr>(scalar);
  DigitWalk dw(&s, 0, 0);
  for (uint32_t w = 0; w < g.W; w++) {
    uint32_t neg, d = dw.next(w, g, neg);
    digits[w] = neg ? -(int32_t)d : (int32_t)d;
  }
  *n_windows = g.W;
  return (int)dw.carry;
}

// FP64-pipe multiplier: out = a * b * 2^-260 mod p on plain integers a, b < p (8 x u32 limbs in and out)
int hs_fieldd_mul(int field, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (field == 0) { Fr r = fpd_to_u32(fpd_mul(fpd_from_u32(ld<Fr>(a)), fpd_from_u32(ld<Fr>(b)))); st(out, r); }
  else { Fq r = fpd_to_u32(fpd_mul(fpd_from_u32(ld<Fq>(a)), fpd_from_u32(ld<Fq>(b)))); st(out, r); }
  return 0;
}

// G1 ops on Montgomery-form coordinates.  xyzz: 4x8 limbs (X, Y, ZZ, ZZZ); affine: 2x8 limbs + inf flag.
// op: 0 xyzz += affine, 1 xyzz += xyzz, 2 double, 3 to_affine
int hs_curve_op(int op, const uint32_t* acc_in, const uint32_t* other, int other_inf, uint32_t* out) {
  G1XYZZ acc;
  memcpy(&acc, acc_in, sizeof(acc));
  switch (op) {
    case 0: {
      G1Affine p;
      memcpy(&p.x, other, 32);
      memcpy(&p.y, other + 8, 32);
      if (!other_inf) g1_add_mixed(acc, p);
      break;
    }
    case 4: {
      G1Affine p;
      memcpy(&p.x, other, 32);
      memcpy(&p.y, other + 8, 32);
      if (!other_inf) g1_add_mixed_uniform(acc, p);
      break;
    }
    case 1: {
      G1XYZZ q;
      memcpy(&q, other, sizeof(q));
      g1_add(acc, q);
      break;
    }
    case 2: g1_double(acc); break;
    case 5: {
      G1XYZZ q;
      memcpy(&q, other, sizeof(q));
      g1_add_uniform(acc, q);
      break;
    }
    case 3: {
      G1Affine p;
      bool inf = g1_to_affine(acc, p);
      memcpy(out, &p.x, 32);
      memcpy(out + 8, &p.y, 32);
      return inf ? 1 : 0;
    }
    default: return -1;
  }
  memcpy(out, &acc, sizeof(acc));
  return 0;
}
}
