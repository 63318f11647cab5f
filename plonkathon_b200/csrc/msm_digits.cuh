// Signed-digit window slicing of MSM scalars (shared by the histogram and scatter kernels of msm.cu and
// unit-tested on the host through csrc/host_selftest.cpp).
#pragma once
#include "field.cuh"

namespace pb200 {

struct MsmGeom {
  uint32_t c, W;            // window bits, number of windows
  uint32_t half;            // 2^(c-1) buckets per window
  uint32_t fixed_base;      // 1: one bucket set per batched scalar vector, shared by all windows; 0: one per window
  uint64_t point_stride;    // generic: 0 ; fixed-base: n (index of window w's copy of point i = w*n + i)
  uint32_t batch;           // fixed-base only: number of MSMs sharing the points
  uint32_t lo, nloc;        // this launch owns bucket magnitudes d with lo <= d - 1 < lo + nloc (bucket-range shard) ...
  uint32_t own_log, own_rank;  // ... or, when own_log > 0, every 2^own_log-th bucket: (d - 1) mod 2^own_log == own_rank,
                               // local index (d - 1) >> own_log, nloc = half >> own_log (strided shard: the buckets
                               // that skewed digits concentrate on -- the short top window -- spread over all ranks)
  uint32_t sets;            // bucket sets: fixed-base: batch ; generic: W
  uint32_t nb;              // sets * nloc local buckets
};

// local bucket of digit magnitude d >= 1 of window w of scalar vector k, or 0xffffffff when another rank owns it
PB_HD uint32_t msm_bucket_key(const MsmGeom& g, uint32_t k, uint32_t w, uint32_t d) {
  uint32_t j;
  if (g.own_log) {
    if (((d - 1) & ((1u << g.own_log) - 1)) != g.own_rank) return 0xffffffffu;
    j = (d - 1) >> g.own_log;
  } else {
    j = d - 1 - g.lo;  // wraps for d - 1 < lo
    if (j >= g.nloc) return 0xffffffffu;
  }
  return (g.fixed_base ? k : w) * g.nloc + j;
}

struct ScalarBatch { const Fr* p[4]; };

// ---- signed-digit walk shared by the histogram and scatter passes
struct DigitWalk {
  Fr s;
  uint32_t carry;
  PB_HD DigitWalk(const Fr* scalars, uint64_t i, int from_mont) : carry(0) {
    s = scalars[i];
    if (from_mont) s = fp_from_mont(s);
  }
  // digit of window w as (magnitude d in [0, 2^(c-1)], sign)
  PB_HD uint32_t next(uint32_t w, const MsmGeom& g, uint32_t& neg) {
    uint32_t bit = w * g.c;
    uint32_t limb = bit >> 5, off = bit & 31;
    uint64_t two = limb < 8 ? s.v[limb] : 0;
    if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
    uint32_t raw = (uint32_t)(two >> off) & ((1u << g.c) - 1);
    uint32_t d = raw + carry;
    neg = 0;
    if (d > g.half) { d = (1u << g.c) - d; neg = 1; carry = 1; } else carry = 0;
    return d;
  }
};


}  // namespace pb200
