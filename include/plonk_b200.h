/* plonk_b200.h -- C ABI of libplonk_b200.so, the B200-native PLONK proving hot path.
 *
 * The reference (0xPARC/plonkathon) is pure Python and has no FFI of its own; these entry points are
 * what a binding for its hot path would call, one per reference callable (file:line cited per entry,
 * relative to the reference tree).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - Fr element  : 32 bytes, little-endian integer in [0, r), r = BN254 group order (curve.py:11).
 *   - G1 affine   : 64 bytes, x || y, each a 32-byte little-endian integer in [0, q) (py_ecc FQ.n).
 *                   The identity (py_ecc Z1 == None) is reported through an `is_identity` flag.
 *   - "d_" pointers are device pointers on the context's device; "h_" pointers are host memory.
 *   - Vectors at the ABI are in canonical (non-Montgomery) form unless a parameter says otherwise.
 *   - Every call returns 0 on success, non-zero on error; pb200_last_error() describes the failure
 *     (thread-local).  The Python facade turns errors into exceptions/AssertionErrors matching the
 *     reference's behaviour.
 *   - Work is issued on the context's CUDA stream; calls that return host data synchronise it.
 *   - A context (and the SRS / prover objects created on it) is not thread-safe: one context per host thread
 *     and per GPU, like the reference's single-threaded call sequence.  Objects are freed by their _destroy
 *     call; the caller owns every h_ / d_ buffer it passes in.
 */
#ifndef PLONK_B200_H
#define PLONK_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb200_ctx pb200_ctx;
typedef struct pb200_srs pb200_srs;
typedef struct pb200_prover pb200_prover;
typedef struct pb200_transcript pb200_transcript;

/* ---- context ------------------------------------------------------------------------------- */
const char* pb200_last_error(void);
const char* pb200_version(void);
/* cuda_stream may be NULL (the library creates its own non-blocking stream). */
int pb200_ctx_create(int device, void* cuda_stream, pb200_ctx** out);
void pb200_ctx_destroy(pb200_ctx* ctx);
int pb200_ctx_sync(pb200_ctx* ctx);
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t pb200_ctx_launches(pb200_ctx* ctx);
/* per-kernel device timing with CUDA events on the launching stream (bench.py's roofline object).
 * enable != 0 clears the records and starts recording.  category 0: MSM bucket accumulation kernel,
 * 1: NTT pass kernel.  Returns the summed duration and the number of launches recorded. */
int pb200_ctx_timing(pb200_ctx* ctx, int enable);
int pb200_ctx_timing_read(pb200_ctx* ctx, int category, double* total_ms, uint64_t* count);
/* the context's CUDA stream (cudaStream_t) so callers can time with events on it */
void* pb200_ctx_stream(pb200_ctx* ctx);

/* ---- Fr vectors ---------------------------------------------------------------------------- */
/* canonical <-> Montgomery form, in place allowed */
int pb200_fr_to_mont(pb200_ctx* ctx, const void* d_in, void* d_out, uint64_t n);
int pb200_fr_from_mont(pb200_ctx* ctx, const void* d_in, void* d_out, uint64_t n);

/* poly.py:23-109  the ring operations of Polynomial on device-resident canonical vectors of n elements:
 *   op 0 a + b, 1 a - b, 2 a * b, 3 a / b (element-wise, py_ecc's inv(0) == 0);
 *   op 4 a + s, 5 a - s, 6 a * s for a Scalar s (h_scalar, canonical) on every element (LAGRANGE basis) --
 *   Polynomial / Scalar is op 6 with 1/s; op 7 / 8: + s / - s on element 0 only (MONOMIAL basis, poly.py:32-37);
 *   op 9 shift: out[i] = a[(i + shift) mod n] (poly.py:102-107; not in place).  d_out may alias d_a otherwise. */
int pb200_fr_vec_op(pb200_ctx* ctx, int op, const void* d_a, const void* d_b, const uint8_t* h_scalar, void* d_out,
                    uint64_t n, uint64_t shift);

/* poly.py:113-149  Polynomial.fft(inv) / ifft: natural order in and out, n = 2^log_n <= 2^28.
 * inverse != 0 uses the reversed roots and multiplies by n^-1.  d_out may alias d_in. */
int pb200_fr_ntt(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, int inverse);
int pb200_fr_ntt_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n, int inverse);

/* The local building block of the slab-sharded transform: the 2^log_m-point NTT of the strided sub-sequence
 * d_in[offset + stride * i] (rank h of G transforms x[h::G]). */
int pb200_fr_ntt_decimated(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_m, int inverse, uint64_t stride,
                           uint64_t offset);

/* poly.py:156-163  to_coset_extended_lagrange(offset): n Lagrange values -> 4n evaluations on
 * offset * <w_4n>.  h_offset: 32-byte canonical Fr. */
int pb200_fr_coset_extend(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, const uint8_t* h_offset);
int pb200_fr_coset_extend_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n,
                               const uint8_t* h_offset);
/* poly.py:169-177  coset_extended_lagrange_to_coeffs(offset): N values (N = 2^log_n, the extended size)
 * -> N coefficients. */
int pb200_fr_coset_to_coeffs(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, const uint8_t* h_offset);
int pb200_fr_coset_to_coeffs_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n,
                                  const uint8_t* h_offset);
/* poly.py:181-195  barycentric_eval(x) with py_ecc's inv(0) == 0 convention when x is on the domain. */
int pb200_fr_barycentric_eval(pb200_ctx* ctx, const void* d_vals, unsigned log_n, const uint8_t* h_x, uint8_t* h_out);
int pb200_fr_barycentric_eval_host(pb200_ctx* ctx, const uint8_t* h_vals, unsigned log_n, const uint8_t* h_x,
                                   uint8_t* h_out);

/* ---- G1 MSM -------------------------------------------------------------------------------- */
/* curve.py:38-44  ec_lincomb(pairs): sum_i scalars[i] * points[i].  Points must be on the curve and not
 * the identity (the facade drops None points); scalars already reduced mod r (curve.py:41).
 * n == 0 is an error (the reference raises ValueError from max(), curve.py:93). */
int pb200_g1_msm(pb200_ctx* ctx, const void* d_points, const void* d_scalars, uint64_t n, uint8_t* h_out_xy,
                 int* is_identity);
int pb200_g1_msm_host(pb200_ctx* ctx, const uint8_t* h_points, const uint8_t* h_scalars, uint64_t n,
                      uint8_t* h_out_xy, int* is_identity);

/* ---- SRS / Setup --------------------------------------------------------------------------- */
/* setup.py:16-22  Setup.powers_of_x.  h_points: n affine points (canonical).  precompute != 0 builds the
 * fixed-base window table in HBM (size ceil(256/c) * n * 64 bytes). */
int pb200_srs_create(pb200_ctx* ctx, const uint8_t* h_points, uint64_t n, int precompute, pb200_srs** out);
/* Structured test SRS generated on the device: points [tau^i]G, i < n, for a known (toxic) tau --
 * the 2^20 / 2^22 configurations need more powers than the reference's shipped .ptau holds
 * (setup.py:27 reads 2^11).  h_tau: canonical 32-byte Fr. */
int pb200_srs_generate(pb200_ctx* ctx, const uint8_t* h_tau, uint64_t n, int precompute, pb200_srs** out);
/* the same for the Lagrange basis of the size-n domain (n a power of two): points
 * [L_i(tau)]G, i < n -- what section 12 of a snarkjs .ptau holds for the ceremony's tau.  Against such an SRS
 * Setup.commit(values) (setup.py:66-72) is ONE MSM over the values, with no inverse transform:
 * pb200_srs_commit_coeffs / _host with the LAGRANGE values in place of coefficients. */
int pb200_srs_generate_lagrange(pb200_ctx* ctx, const uint8_t* h_tau, uint64_t n, int precompute, pb200_srs** out);
/* copy `count` points starting at `first` back to the host (canonical x||y) */
int pb200_srs_export(pb200_ctx* ctx, pb200_srs* srs, uint8_t* h_points, uint64_t first, uint64_t count);
void pb200_srs_destroy(pb200_srs* srs);
uint64_t pb200_srs_size(pb200_srs* srs);
/* setup.py:66-72  Setup.commit(values): values in the LAGRANGE basis -> ifft -> MSM with powers_of_x.
 * n = 2^log_n must be <= srs size. */
int pb200_srs_commit_lagrange(pb200_ctx* ctx, pb200_srs* srs, const void* d_values, unsigned log_n,
                              uint8_t* h_out_xy, int* is_identity);
int pb200_srs_commit_lagrange_host(pb200_ctx* ctx, pb200_srs* srs, const uint8_t* h_values, unsigned log_n,
                                   uint8_t* h_out_xy, int* is_identity);
/* MSM of m coefficients (monomial basis) with the first m powers. */
int pb200_srs_commit_coeffs(pb200_ctx* ctx, pb200_srs* srs, const void* d_coeffs, uint64_t m, int coeffs_montgomery,
                            uint8_t* h_out_xy, int* is_identity);
/* same from host memory (canonical scalars) */
int pb200_srs_commit_coeffs_host(pb200_ctx* ctx, pb200_srs* srs, const uint8_t* h_coeffs, uint64_t m,
                                 uint8_t* h_out_xy, int* is_identity);

/* ---- Prover (prover.py:39-306) ---------------------------------------------------------------- */
/* prover.py:45-49  Prover(setup, program): h_pk = 8 pointers, in the order of CommonPreprocessedInput
 * (compiler/program.py:10-30): QM QL QR QO QC S1 S2 S3, each 2^log_n Lagrange values (canonical).
 * Converts them to coefficients and to a cached 4n coset extension in HBM. */
int pb200_prover_create(pb200_ctx* ctx, pb200_srs* srs, unsigned log_n, const uint8_t* const* h_pk,
                        pb200_prover** out);
void pb200_prover_destroy(pb200_prover* p);
/* prover.py:51-84  prove(witness): h_A/h_B/h_C = wire values per row (prover.py:97-103), h_public = the
 * public input values in order (prover.py:57-62; the library negates them).  Writes the canonical 768-byte
 * proof: Proof.flatten() order (prover.py:18-35), G1 as x||y, 32-byte big-endian integers.
 * Fails (error string starts with "AssertionError") where the reference's asserts would. */
int pb200_prover_prove(pb200_prover* p, const uint8_t* h_A, const uint8_t* h_B, const uint8_t* h_C,
                       const uint8_t* h_public, uint64_t n_public, uint8_t* h_proof768);
/* same, with the wire values already resident in HBM (canonical form, n x 32 bytes each) */
int pb200_prover_prove_device(pb200_prover* p, const void* d_A, const void* d_B, const void* d_C,
                              const uint8_t* h_public, uint64_t n_public, uint8_t* h_proof768);
/* the individual rounds, challenges supplied by the caller's transcript; outputs little-endian */
int pb200_prover_round1(pb200_prover* p, const uint8_t* h_A, const uint8_t* h_B, const uint8_t* h_C,
                        const uint8_t* h_public, uint64_t n_public, uint8_t* h_abc_xy /*3*64*/);   /* prover.py:86 */
int pb200_prover_round2(pb200_prover* p, const uint8_t* beta, const uint8_t* gamma, uint8_t* h_z_xy);  /* :121 */
int pb200_prover_round3(pb200_prover* p, const uint8_t* alpha, const uint8_t* fft_cofactor,
                        uint8_t* h_t_xy /*3*64*/);                                                   /* :154 */
int pb200_prover_round4(pb200_prover* p, const uint8_t* zeta, uint8_t* h_evals /*6*32*/);             /* :228 */
int pb200_prover_round5(pb200_prover* p, const uint8_t* v, uint8_t* h_w_xy /*2*64*/);                 /* :241 */

/* The round state the reference keeps on `self` (read by its own sanity asserts, prover.py:108-116, 137-145,
 * 215-219), copied out in canonical form to a device buffer of 2^log_n elements: which = 0 A, 1 B, 2 C, 3 Z, 4 PI
 * (Lagrange values), 5 T1, 6 T2, 7 T3 (coefficients).  Valid after the round that produces them. */
int pb200_prover_read_vector(pb200_prover* p, int which, void* d_out);
/* canonical 768-byte proof of the last rounds run on this prover (Proof.flatten() order, prover.py:18-35) */
int pb200_prover_serialize(pb200_prover* p, uint8_t* h_proof768);

/* ---- multi-GPU: one process per GPU, one communicator per context -------------------------
 * The library issues its data-path collectives itself, on the context's stream, through NCCL (bound at run time
 * from the libnccl.so.2 the process has loaded; the single-GPU entry points work without it).  Rendezvous stays
 * with the caller: one rank draws pb200_comm_unique_id, distributes the 128 bytes (torch.distributed in
 * plonkathon_b200/parallel.py), every rank calls pb200_comm_init.  2, 4 or 8 ranks of one box. */
int pb200_comm_unique_id(uint8_t* out128);
int pb200_comm_init(pb200_ctx* ctx, const uint8_t* id128, int rank, int world);
/* rank / world of the context (0 / 1 without a communicator), data-path collectives issued so far and the bytes
 * this rank received in them */
int pb200_comm_info(pb200_ctx* ctx, int* rank, int* world, uint64_t* collectives, uint64_t* bytes_received);
/* poly.py:113-149 slab-sharded across the ranks: d_in (2^log_n elements, present on every rank; rank h reads only
 * x[h::G]) -> d_out (the full transform, on every rank).  Local 2^log_n / G-point transforms with the join twiddle
 * fused into the store, ONE allgather, then a G-point DFT per element (csrc/ntt_shard.cuh). */
int pb200_fr_ntt_sharded(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, int inverse);
/* setup.py:66-72's MSM with the 2^(c-1) signed-digit buckets split evenly over the ranks (every rank holds the
 * coefficients and an SRS replica, walks all digits, but sorts / accumulates / reduces only its buckets); ONE
 * allgather of 256 bytes per rank at the join; the same affine result on every rank. */
int pb200_srs_commit_coeffs_sharded(pb200_ctx* ctx, pb200_srs* srs, const void* d_coeffs, uint64_t m,
                                    int coeffs_montgomery, uint8_t* h_out_xy, int* is_identity);
/* prover.py:45-49 for one proof across the ranks: as pb200_prover_create, but rank r caches and works on every
 * G-th point of the 4n coset (coset extensions, selector cache and quotient divide by G), interpolations are
 * slab-sharded, commitments bucket-sharded.  Every rank calls the prove / round entry points with the same
 * inputs and gets the same 768 bytes.  A failing check (the reference's asserts) fails on every rank alike. */
int pb200_prover_create_sharded(pb200_ctx* ctx, pb200_srs* srs, unsigned log_n, const uint8_t* const* h_pk,
                                pb200_prover** out);
/* Operator-level shards with a caller-side join (tests, other transports): the partial sum over the SRS powers
 * [first, first+count) and the bucket magnitudes [bucket_lo, bucket_hi) of pb200_srs_bucket_count's range, as one
 * XYZZ point (128 bytes, Montgomery limbs); pb200_g1_combine_partials_host adds such partials. */
int pb200_srs_commit_partial(pb200_ctx* ctx, pb200_srs* srs, const void* d_coeffs, uint64_t first, uint64_t count,
                             uint32_t bucket_lo, uint32_t bucket_hi, int coeffs_montgomery, uint8_t* h_xyzz128);
int pb200_srs_bucket_count(pb200_srs* srs, uint32_t* out);
int pb200_g1_combine_partials_host(const uint8_t* h_xyzz, unsigned count, uint8_t* h_out_xy, int* is_identity);
/* the host half of the sharded commitment's join: h_sr = [world][sets] pairs (S = sum of the rank's buckets,
 * R = sum_j (j+1) B_j over them, j the rank's local bucket index; 2 x 128 bytes XYZZ).  nloc > 0: rank rho owns the
 * contiguous buckets [rho * nloc, (rho+1) * nloc); nloc == 0: strided ownership, rank rho owns the buckets
 * world * k + rho (what pb200_srs_commit_coeffs_sharded uses: skewed digits spread over all ranks). */
int pb200_g1_join_bucket_shards_host(const uint8_t* h_sr, unsigned world, unsigned sets, uint32_t nloc, uint8_t* h_out_xy,
                                     int* is_identity);

/* ---- Transcript (transcript.py:58-123; host code) ---------------------------------------------- */
int pb200_transcript_create(const uint8_t* label, size_t label_len, pb200_transcript** out);
void pb200_transcript_destroy(pb200_transcript* t);
int pb200_transcript_append_message(pb200_transcript* t, const uint8_t* label, size_t label_len,
                                    const uint8_t* msg, size_t msg_len);
int pb200_transcript_challenge_bytes(pb200_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out,
                                     size_t n);
/* transcript.py:69-75: challenge as a canonical little-endian Fr (never zero) */
int pb200_transcript_get_and_append_challenge(pb200_transcript* t, const uint8_t* label, size_t label_len,
                                              uint8_t* out_le32);

/* ---- Pairing and G2 (host code) ------------------------------------------------
 * Replaces py_ecc's `b.pairing`, `b.add`/`b.multiply` on G2 as the reference's verifier calls them
 * (TESTING_verifier_DO_NOT_OPEN.py:148-151, 237-262) and `b.is_on_curve(X2, b.b2)` (setup.py:59).
 * G1 points: 64 B x||y, G2 points: 128 B x.c0||x.c1||y.c0||y.c1 (the .ptau order, setup.py:53-58), every
 * coordinate 32-byte little-endian canonical; identity flags are one byte per point (NULL = none).
 * Points off their curve are an error; membership of the order-r subgroup of G2 is NOT checked. */
/* *ok = 1 iff  prod_i e(g1_i, g2_i) == 1  in GT */
int pb200_pairing_check(const uint8_t* h_g1, const uint8_t* h_g1_identity, const uint8_t* h_g2,
                        const uint8_t* h_g2_identity, unsigned count, int* ok);
int pb200_g2_mul(const uint8_t* h_point, const uint8_t* h_scalar_le32, uint8_t* h_out, int* is_identity);
int pb200_g2_add(const uint8_t* h_p, int p_identity, const uint8_t* h_q, int q_identity, uint8_t* h_out,
                 int* is_identity);

/* ---- micro-benchmarks (bench.py / profiles only) --------------------------------------------- */
/* runs `iters` dependent Montgomery products per thread over `threads` threads; returns elapsed ms */
int pb200_bench_modmul(pb200_ctx* ctx, int field /*0 Fr, 1 Fq*/, uint64_t threads, uint32_t iters, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PLONK_B200_H */
