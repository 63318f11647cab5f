// Prover state shared between prover.cu (the rounds) and capi.cu (the C ABI accessors).
#pragma once
#include "common.cuh"

namespace pb200 {

struct Srs;
void srs_msm(Context* ctx, Srs* srs, const Fr* d_scalars, uint64_t m, bool scalars_mont, uint8_t* out_xy, int* is_identity);
void srs_msm_batch(Context* ctx, Srs* srs, const Fr* const* d_scalars, uint32_t batch, uint64_t m, bool scalars_mont,
                   uint8_t* out_xy, int* is_identity);
void srs_msm_batch_partial(Context* ctx, Srs* srs, const Fr* const* d_scalars, uint32_t batch, uint64_t first,
                           uint64_t count, uint32_t bucket_lo, uint32_t bucket_hi, bool scalars_mont, G1XYZZ* out);
void srs_msm_batch_sharded(Context* ctx, Srs* srs, const Fr* const* d_scalars, uint32_t batch, uint64_t m,
                           bool scalars_mont, uint8_t* out_xy, int* is_identity);
uint32_t srs_bucket_count(Srs* s);

struct Proof {
  uint8_t pts[9][64];    // a_1 b_1 c_1 z_1 t_lo t_mid t_hi W_z W_zw  (canonical LE x||y)
  uint8_t evals[6][32];  // a b c s1 s2 z_shifted (canonical LE)
};

struct Prover {
  Context* ctx;
  Srs* srs;
  int log_n;
  uint64_t n;
  // One proof across the GPUs of a box (world > 1; the context carries the communicator): rank r owns every
  // world-th point of the 4n coset -- x_j = g mu^(world j + r), j < n_ext = 4n / world, itself a coset of the subgroup
  // of order n_ext -- so the coset extensions, the cached selector extensions and the quotient are local and divide by
  // world; inverse transforms are slab-sharded with one allgather at the join (ntt_shard.cuh); commitments split the
  // buckets (msm.cu).  world == 1 is the same code with n_ext = 4n.
  int world = 1, rank = 0, log_world = 0;
  int log_ext = 0;       // log2(n_ext)
  uint64_t n_ext = 0;
  uint32_t fold = 1;     // n / n_ext when the coset slice is shorter than a coefficient vector (world = 8), else 1
  uint64_t zw_shift = 4; // Z(w x_j) = Z-extension at local index j + 4 / world ...
  bool zw_separate = false;  // ... or, when 4 % world != 0, a separately extended vector (ext[5])
  // per-circuit (all Montgomery)
  DevBuf sel_coeff[8];   // QM QL QR QO QC S1 S2 S3, coefficient form
  DevBuf sel_lag[8];     // same, Lagrange values (QM..QC for the gate check, S1..S3 for round 2)
  DevBuf sel_ext[8];     // same, on this rank's slice of the fixed 4n coset
  DevBuf roots;          // w^i, i < n
  DevBuf gpow;           // (g mu^rank)^i, i < n     (coset shift on load)
  DevBuf gpow_w;         // (g mu^(rank+4))^i, i < n (only when zw_separate)
  DevBuf ginv_pow;       // g^-i, i < 4n            (undo the shift on store)
  DevBuf xs;             // x_j, j < n_ext
  DevBuf l0_ext;         // L0 on the slice
  Fr g, g_inv, zh_inv[4];
  // per-proof state
  DevBuf lag[4];         // A B C Z Lagrange
  DevBuf coeff[5];       // a b c z pi coefficients
  DevBuf pi_lag;
  DevBuf ext[6];         // A B C Z PI Z(wX) on the slice
  DevBuf tq;             // quotient evaluations (n_ext) / coefficients (4n; sharded: 3n coefficients)
  DevBuf tq_loc;         // sharded: quotient evaluations on the slice
  DevBuf tmp[5];
  DevBuf aux_tmp;        // pass buffer of the side-stream coset transforms (n_ext)
  bool overlap = true;   // run the round-3 coset extensions of A, B, C (and Z) beside the round-1/2 MSMs
  DevBuf flags;
  Fr beta, gamma, alpha, fft_cofactor, zeta, v;   // Montgomery
  Fr ev[6];                                      // Montgomery evaluations (round 4)
  Fr pi_ev;
  // public inputs: when there are at most 8, PI is a combination of cached Lagrange-basis coset vectors
  uint64_t n_public = 0;
  bool pi_sparse = false;
  std::vector<DevBuf> pi_basis;   // L_i on the slice (n_ext each), i < 8
  std::vector<Fr> pub_neg;        // -public_i, Montgomery (host)
  Proof proof;

  enum { QM = 0, QL, QR, QO, QC, S1, S2, S3 };

  // several commitments in one pass over the SRS (out: count * 64 bytes, contiguous)
  void commit_batch(const Fr* const* d_coeffs, uint32_t count, uint64_t m, uint8_t* out_xy) {
    int ident[4] = {0, 0, 0, 0};
    if (world > 1) srs_msm_batch_sharded(ctx, srs, d_coeffs, count, m, true, out_xy, ident);
    else srs_msm_batch(ctx, srs, d_coeffs, count, m, true, out_xy, ident);
    for (uint32_t k = 0; k < count; k++)
      PB_CHECK(!ident[k], "commitment is the point at infinity (unsupported by the reference transcript)");
  }
  void commit(const Fr* d_coeffs, uint64_t m, uint8_t* out_xy) { commit_batch(&d_coeffs, 1, m, out_xy); }
};


}  // namespace pb200
