// Host build of the limb-level arithmetic in field.cuh / curve.cuh (same code path as the device,
// PTX carry-chain primitives replaced by their emulation).  TEST INFRASTRUCTURE: loaded only by
// tests/test_host_arith.py through ctypes; never linked into libplonk_b200.so.
#include "field.cuh"
#include "curve.cuh"
#include "fieldd.cuh"
#include "msm_digits.cuh"
#include "msm_affine.cuh"
#include "modinv.cuh"
#include <vector>
#include <cstring>
using namespace pb200;

template <class F> static F ld(const uint32_t* p) { F r; memcpy(r.v, p, 32); return r; }
template <class F> static void st(uint32_t* p, const F& a) { memcpy(p, a.v, 32); }

extern "C" {
// op: 0 add, 1 sub, 2 mul, 3 neg, 4 inv, 5 to_mont, 6 from_mont, 7 dbl, 8 sqr, 9 inv by safegcd (Montgomery
// contract of op 4), 10 plain-integer inverse by safegcd ; field: 0 Fr, 1 Fq
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
      case 9: r = fp_inv_gcd(x); break;                          \
      case 10: r = fp_inv_plain_gcd(x); break;                   \
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

// Bucket accumulation by rounds of batched affine additions (msm_affine.cuh), every thread body run in a loop.
// table: n_pts canonical affine points (16 words each); sorted: entries (index | sign << 31) grouped by bucket;
// offsets: nb + 1.  out: nb canonical affine points + out_inf flags.  Returns the number of rounds, -1 on bad input.
int hs_msm_affine_rounds(const uint32_t* table, uint32_t n_pts, const uint32_t* sorted, const uint32_t* offsets,
                         uint32_t nb, uint32_t B, uint32_t F, uint32_t* out, uint8_t* out_inf) {
  if (B == 0 || F == 0 || F > 64) return -1;
  std::vector<G1Affine> tab(n_pts);
  for (uint32_t i = 0; i < n_pts; i++) {
    tab[i].x = fp_to_mont(ld<Fq>(table + 16 * i));
    tab[i].y = fp_to_mont(ld<Fq>(table + 16 * i + 8));
  }
  std::vector<uint32_t> off_in(offsets, offsets + nb + 1), off_out(nb + 1);
  std::vector<G1Affine> cur, nxt;
  int rounds = 0;
  bool first = true;
  for (;;) {
    uint32_t maxc = 0;
    for (uint32_t b = 0; b < nb; b++) maxc = std::max(maxc, off_in[b + 1] - off_in[b]);
    AffineRound a;
    a.table = tab.data();
    a.sorted = first ? sorted : nullptr;
    a.in = first ? nullptr : cur.data();
    a.off_in = off_in.data();
    a.nb = nb;
    a.B = B;
    if (maxc <= 1) {
      for (uint32_t b = 0; b < nb; b++) {
        G1XYZZ r = affine_round_bucket(a, b);
        G1Affine p;
        bool inf = g1_to_affine(r, p);
        out_inf[b] = inf ? 1 : 0;
        Fq x = fp_from_mont(p.x), y = fp_from_mont(p.y);
        st(out + 16 * b, x);
        st(out + 16 * b + 8, y);
      }
      return rounds;
    }
    off_out[0] = 0;
    for (uint32_t b = 0; b < nb; b++) off_out[b + 1] = off_out[b] + ((off_in[b + 1] - off_in[b] + 1) >> 1);
    const uint32_t S = off_out[nb], T = affine_round_threads(S, B);
    nxt.assign(S, G1Affine());
    std::vector<Fq> prefix(S), prod(T + 40);  // a spare warp and a bit: those threads must do nothing
    std::vector<uint32_t> desc(S);
    a.off_out = off_out.data();
    a.out = nxt.data();
    a.prefix = prefix.data();
    a.desc = desc.data();
    a.thread_prod = prod.data();
    for (uint32_t t = 0; t < T + 40; t++) affine_round_forward(a, t);
    for (uint32_t u = 0; u < affine_invert_threads(T, F) + 2; u++) affine_round_invert(prod.data(), T, F, u);
    for (uint32_t t = T + 40; t-- > 0;) affine_round_backward(a, t);
    cur.swap(nxt);
    off_in = off_out;
    first = false;
    rounds++;
  }
}
}
