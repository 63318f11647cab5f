// Signed-digit window slicing of MSM scalars (shared by the histogram and scatter kernels of msm.cu and
// unit-tested on the host through csrc/host_selftest.cpp).
#pragma once
#include "field.cuh"

namespace pb200 {

struct MsmGeom {
  uint32_t c, W;            // window bits, number of windows
  uint32_t half;            // 2^(c-1) buckets per window
  uint32_t bucket_stride;   // generic: half ; fixed-base: 0
  uint64_t point_stride;    // generic: 0 ; fixed-base: n (index of window w's copy of point i = w*n + i)
  uint32_t nb;              // total buckets
  uint32_t batch;           // fixed-base only: number of MSMs sharing the points (bucket set k at k * half)
};

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
