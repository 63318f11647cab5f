// The join of the slab-sharded NTT (poly.py:113-149 across G = 2, 4 or 8 GPUs): a length-G DFT per element.
//
// N = G * M.  Rank r transforms the decimated sequence x[r::G] (M points) locally and multiplies by w_N^(r k0) on
// the store:  u_r[k0] = w_N^(r k0) * sum_j1 x[G j1 + r] w_M^(j1 k0).  After ONE allgather of the G sub-spectra,
//     X[k0 + M k1] = sum_r u_r[k0] w_G^(r k1),        w_G = w_N^M,
// i.e. for every k0 a G-point DFT over the rank index, done here in registers by radix-2 butterflies (5 products
// for G = 8) instead of G - 1 products per output.  The inverse transform is the same with inverse roots and 1/G
// folded into the store multiplier; the M outputs of a DFT are M apart, so both the G loads and the G stores of a
// warp are coalesced.  Host/device code (unit-tested on the CPU through csrc/host_selftest.cpp).
#pragma once
#include "field.cuh"

namespace pb200 {

struct DftTw {
  Fr w[4];  // w_G^k, k < G/2
};

template <int LG>
PB_HD void small_dft(Fr (&x)[1 << LG], const DftTw& tw) {
  constexpr int G = 1 << LG;
  // bit reversal of the input index (compile-time unrolled)
#pragma unroll
  for (int i = 0; i < G; i++) {
    int j = 0;
#pragma unroll
    for (int b = 0; b < LG; b++) j |= ((i >> b) & 1) << (LG - 1 - b);
    if (j > i) { Fr t = x[i]; x[i] = x[j]; x[j] = t; }
  }
#pragma unroll
  for (int s = 0; s < LG; s++) {
    const int half = 1 << s;
#pragma unroll
    for (int i = 0; i < G; i += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; j++) {
        Fr v = x[i + j + half];
        if (j) v = fp_mul(v, tw.w[j * (G / (2 * half))]);
        const Fr u = x[i + j];
        x[i + j] = fp_add(u, v);
        x[i + j + half] = fp_sub(u, v);
      }
    }
  }
}

}  // namespace pb200
