// Bucket accumulation by rounds of batched AFFINE additions (the alternative to the XYZZ segment accumulation
// in msm.cu; same inputs -- the bucket-sorted entry list -- same output -- one point per bucket).
//
// An affine chord addition costs 1 inversion + 2 mul + 1 sqr; sharing the inversion over many independent
// additions with Montgomery's trick (3 mul each) makes it 6 products per addition against 10 for XYZZ += affine.
// Independence comes from the shape of the work: in one round every bucket with k points does floor(k/2)
// additions of neighbouring pairs (an odd last point is carried over), so the bucket sizes halve and
// ceil(log2(max bucket size)) rounds leave one point per bucket.  Output slot s of a round belongs to bucket b
// (off_out[b] <= s < off_out[b+1]); with j = s - off_out[b] it is the sum of input elements off_in[b] + 2j and
// off_in[b] + 2j + 1, or a copy of the former when that is the bucket's odd last element.
//
// One round is three launches:
//   forward  : thread t owns B output slots (lane-interleaved within its warp's 32 * B consecutive slots); it
//              multiplies the denominators of its additions into a running product, storing the product BEFORE each
//              slot (prefix[s]), and classifies each slot (desc[s]);
//   invert   : the per-thread products are inverted in place, again with Montgomery's trick (F per thread, one
//              Fermat inversion each) -- two levels, so an inversion is shared by B * F additions;
//   backward : thread t walks its slots in reverse, peeling 1/denominator off the inverted product, and writes the
//              sums.
// Exceptional cases are exact, as the reference's group law has them (curve.py:38-44 over py_ecc `add`): equal
// points are doubled (tangent slope, denominator 2y), opposite points give the identity, an identity operand
// returns the other one.  The identity is encoded in an affine slot by x.v[7] == 0xffffffff (not a reduced field
// element).  Points of the first round are gathered from the point table through the sorted entry list (index |
// sign << 31), later rounds read the previous round's output.
//
// The thread bodies are host/device functions so that tests/test_host_arith.py can run whole rounds on the CPU
// (csrc/host_selftest.cpp) against the oracle's group law; the __global__ wrappers live in msm.cu.
#pragma once
#include "curve.cuh"

namespace pb200 {

#define PB_AFF_COPY 0u
#define PB_AFF_ADD 1u
#define PB_AFF_DBL 2u
#define PB_AFF_INF 3u
#define PB_AFF_TAKE_Q 4u
#define PB_AFF_TAKE_P 5u
#define PB_AFF_INDEX_BITS 29
#define PB_AFF_INDEX_MASK ((1u << PB_AFF_INDEX_BITS) - 1)

struct AffineRound {
  const G1Affine* table;    // first round: point table (Montgomery affine), addressed through `sorted`
  const uint32_t* sorted;   // first round: point index | sign << 31 per entry; nullptr in later rounds
  const G1Affine* in;       // later rounds: the previous round's output
  const uint32_t* off_in;   // nb + 1 offsets of the input layout
  const uint32_t* off_out;  // nb + 1 offsets of the output layout: off_out[b+1] - off_out[b] = ceil(count_in / 2)
  uint32_t nb;              // buckets
  uint32_t B;               // output slots per thread
  G1Affine* out;            // one point per output slot
  Fq* prefix;               // one per output slot
  uint32_t* desc;           // one per output slot: first input element | kind << 29
  Fq* thread_prod;          // one per thread: product of its denominators, later its inverse
};

PB_HD bool aff_is_identity_x(const Fq& x) { return x.v[7] == 0xffffffffu; }
PB_HD G1Affine aff_identity() {
  G1Affine r;
  for (int i = 0; i < 8; i++) { r.x.v[i] = 0xffffffffu; r.y.v[i] = 0; }
  return r;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ Fq aff_ld_fq(const Fq* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fq r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ uint32_t aff_ld_u32(const uint32_t* p) { return __ldg(p); }
#else
inline Fq aff_ld_fq(const Fq* p) { return *p; }
inline uint32_t aff_ld_u32(const uint32_t* p) { return *p; }
#endif

// x coordinate of input element e
PB_HD Fq aff_load_x(const AffineRound& a, uint32_t e) {
  if (a.sorted) return aff_ld_fq(&a.table[aff_ld_u32(a.sorted + e) & 0x7fffffffu].x);
  return aff_ld_fq(&a.in[e].x);
}
// y coordinate of input element e (sign applied in the first round)
PB_HD Fq aff_load_y(const AffineRound& a, uint32_t e) {
  if (a.sorted) {
    uint32_t v = aff_ld_u32(a.sorted + e);
    Fq y = aff_ld_fq(&a.table[v & 0x7fffffffu].y);
    return (v >> 31) ? fp_neg(y) : y;
  }
  return aff_ld_fq(&a.in[e].y);
}
PB_HD G1Affine aff_load(const AffineRound& a, uint32_t e) {
  G1Affine p;
  p.x = aff_load_x(a, e);
  p.y = aff_load_y(a, e);
  return p;
}

// first index i in [lo, hi) with a[i] > key (hi if none)
PB_HD uint32_t aff_upper_bound(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t key) {
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (aff_ld_u32(a + mid) > key) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Slot assignment: a warp owns 32 * B consecutive output slots and lane l takes slots w0 + l, w0 + 32 + l, ...,
// so that at every step the 32 lanes touch 32 consecutive slots (coalesced prefix / desc / out traffic, and
// consecutive input pairs).  The bucket of a slot is found by bisecting off_out between the buckets of the warp's
// first and last slot.
struct AffineSpan {
  uint32_t w0, w1;      // the warp's slot range [w0, w1)
  uint32_t lane;
  uint32_t b_lo, b_hi;  // buckets of slots w0 and w1 - 1
};
PB_HD bool affine_span(const AffineRound& a, uint32_t t, AffineSpan& sp) {
  const uint32_t S = a.off_out[a.nb];
  const uint64_t w0 = (uint64_t)(t >> 5) * 32 * a.B;
  if (w0 >= S) return false;
  sp.w0 = (uint32_t)w0;
  sp.w1 = (uint32_t)(w0 + 32ull * a.B < S ? w0 + 32ull * a.B : S);
  sp.lane = t & 31;
  sp.b_lo = aff_upper_bound(a.off_out, 0, a.nb + 1, sp.w0) - 1;
  sp.b_hi = aff_upper_bound(a.off_out, sp.b_lo, a.nb + 1, sp.w1 - 1) - 1;
  return true;
}
// number of threads that own a product this round (all lanes of every warp with at least one slot)
PB_HD uint32_t affine_round_threads(uint32_t S, uint32_t B) {
  return (uint32_t)(((uint64_t)S + 32ull * B - 1) / (32ull * B)) * 32;
}

// ---- forward pass of thread t -------------------------------------------------------------------------------
PB_HD void affine_round_forward(const AffineRound& a, uint32_t t) {
  AffineSpan sp;
  if (!affine_span(a, t, sp)) return;
  Fq acc = Fq::one();
  uint32_t b = sp.b_lo;
  for (uint32_t s = sp.w0 + sp.lane; s < sp.w1; s += 32) {
    b = aff_upper_bound(a.off_out, b + 1, sp.b_hi + 1, s) - 1;  // off_out[b] <= s < off_out[b+1]
    const uint32_t j = s - aff_ld_u32(a.off_out + b);
    const uint32_t in_lo = aff_ld_u32(a.off_in + b), in_cnt = aff_ld_u32(a.off_in + b + 1) - in_lo;
    const uint32_t e0 = in_lo + 2 * j;
    uint32_t kind = PB_AFF_COPY;
    if (2 * j + 1 < in_cnt) {
      Fq x1 = aff_load_x(a, e0), x2 = aff_load_x(a, e0 + 1);
      Fq d = Fq::one();
      if (aff_is_identity_x(x1)) {
        kind = PB_AFF_TAKE_Q;
      } else if (aff_is_identity_x(x2)) {
        kind = PB_AFF_TAKE_P;
      } else if (x1 != x2) {
        kind = PB_AFF_ADD;
        d = fp_sub(x2, x1);
      } else {
        Fq y1 = aff_load_y(a, e0), y2 = aff_load_y(a, e0 + 1);
        if (y1 == y2) {  // y != 0: the group has odd order
          kind = PB_AFF_DBL;
          d = fp_dbl(y1);
        } else {
          kind = PB_AFF_INF;
        }
      }
      if (kind == PB_AFF_ADD || kind == PB_AFF_DBL) {
        a.prefix[s] = acc;
        acc = fp_mul(acc, d);
      }
    }
    a.desc[s] = e0 | (kind << PB_AFF_INDEX_BITS);
  }
  a.thread_prod[t] = acc;
}

// ---- in-place inversion of the per-thread products ------------------------------------------------------------------
// U = ceil(n / F) threads; thread u owns entries u, u + U, u + 2U, ... (at most F <= 64 of them: neighbouring
// threads touch neighbouring entries).  All of a thread's entries are fetched before the product chain starts, so
// the loads overlap instead of each waiting in front of a multiplication.
PB_HD uint32_t affine_invert_threads(uint32_t n, uint32_t F) { return (n + F - 1) / F; }
PB_HD void affine_round_invert(Fq* prod, uint32_t n, uint32_t F, uint32_t u) {
  const uint32_t U = affine_invert_threads(n, F);
  if (u >= U) return;
  Fq val[64], pref[64];
  uint32_t cnt = 0;
  for (uint32_t k = 0; k < F; k++) {
    const uint64_t i = (uint64_t)k * U + u;
    if (i >= n) break;
    val[k] = prod[i];
    cnt = k + 1;
  }
  Fq run = Fq::one();
  for (uint32_t k = 0; k < cnt; k++) {
    pref[k] = run;
    run = fp_mul(run, val[k]);
  }
  Fq inv = fp_inv(run);
  for (uint32_t k = cnt; k-- > 0;) {
    prod[(uint64_t)k * U + u] = fp_mul(inv, pref[k]);
    inv = fp_mul(inv, val[k]);
  }
}

// ---- backward pass of thread t ------------------------------------------------------------------------------
PB_HD void affine_round_backward(const AffineRound& a, uint32_t t) {
  AffineSpan sp;
  if (!affine_span(a, t, sp)) return;
  if (sp.w0 + sp.lane >= sp.w1) return;
  Fq inv = a.thread_prod[t];  // 1 / (product of this thread's denominators)
  const uint32_t last = sp.w0 + sp.lane + ((sp.w1 - 1 - sp.w0 - sp.lane) & ~31u);  // this lane's last slot
  for (uint32_t s = last;; s -= 32) {
    const uint32_t dsc = a.desc[s];
    const uint32_t e0 = dsc & PB_AFF_INDEX_MASK, kind = dsc >> PB_AFF_INDEX_BITS;
    G1Affine r;
    if (kind == PB_AFF_ADD || kind == PB_AFF_DBL) {
      const G1Affine p = aff_load(a, e0), q = aff_load(a, e0 + 1);
      Fq d, num;
      if (kind == PB_AFF_ADD) {
        d = fp_sub(q.x, p.x);
        num = fp_sub(q.y, p.y);
      } else {
        d = fp_dbl(p.y);
        Fq xx = fp_sqr(p.x);
        num = fp_add(fp_dbl(xx), xx);
      }
      const Fq dinv = fp_mul(inv, a.prefix[s]);
      inv = fp_mul(inv, d);
      const Fq lam = fp_mul(num, dinv);
      r.x = fp_sub(fp_sub(fp_sqr(lam), p.x), q.x);
      r.y = fp_sub(fp_mul(lam, fp_sub(p.x, r.x)), p.y);
    } else if (kind == PB_AFF_INF) {
      r = aff_identity();
    } else if (kind == PB_AFF_TAKE_Q) {
      r = aff_load(a, e0 + 1);
    } else {  // COPY, TAKE_P
      r = aff_load(a, e0);
    }
    a.out[s] = r;
    if (s < sp.w0 + 32 + sp.lane) break;
  }
}

// ---- after the last round: bucket b holds 0 or 1 element ----------------------------------------------------------
PB_HD G1XYZZ affine_round_bucket(const AffineRound& a, uint32_t b) {
  const uint32_t lo = a.off_in[b], cnt = a.off_in[b + 1] - lo;
  if (cnt == 0) return G1XYZZ::identity();
  const G1Affine p = aff_load(a, lo);
  if (aff_is_identity_x(p.x)) return G1XYZZ::identity();
  return g1_from_affine(p);
}

}  // namespace pb200
