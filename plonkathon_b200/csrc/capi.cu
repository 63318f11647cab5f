// extern "C" surface of libplonk_b200.so (declared in include/plonk_b200.h).
#include "../../include/plonk_b200.h"

#include <cstring>

#include "common.cuh"
#include "prover.cuh"
#include "pairing.cuh"
#include "transcript.cuh"

namespace pb200 {
// ntt.cu
void ntt_run(Context* ctx, const Fr* in, Fr* out, int log_n, bool inverse, uint64_t n_in, const Fr* in_scale,
             const Fr* out_scale);
void launch_powers(Context* ctx, Fr* out, uint64_t n, const Fr& base, const Fr& scale);
Fr fr_from_u64(uint64_t x);
void ntt_run_strided(Context* ctx, const Fr* in, Fr* out, int log_n, bool inverse, uint64_t n_in,
                     const Fr* in_scale, const Fr* out_scale, uint64_t in_mul, uint64_t in_add);
void ntt_slab_combine(Context* ctx, const Fr* sub, Fr* out, int log_m, int log_g, uint64_t slab, bool inverse);
// poly_ops.cu
void fr_to_mont(Context* ctx, const Fr* in, Fr* out, uint64_t n);
void fr_from_mont(Context* ctx, const Fr* in, Fr* out, uint64_t n);
void barycentric_eval(Context* ctx, const Fr* d_vals, int log_n, const Fr& x_mont, Fr* h_out);
float bench_modmul(Context* ctx, int field, uint64_t threads, uint32_t iters);
// msm.cu
struct Srs;
Srs* srs_create(Context* ctx, const uint8_t* h_points, uint64_t n, int precompute);
void srs_destroy(Srs* s);
Srs* srs_generate(Context* ctx, const Fr& tau_canonical, uint64_t n, int precompute);
Srs* srs_generate_lagrange(Context* ctx, const Fr& tau_canonical, uint64_t n, int precompute);
void srs_export(Context* ctx, Srs* srs, uint8_t* h_points, uint64_t first, uint64_t count);
void srs_msm(Context* ctx, Srs* srs, const Fr* d_scalars, uint64_t m, bool scalars_mont, uint8_t* out_xy, int* is_identity);
uint64_t srs_size(Srs* s);
uint32_t msm_default_window(uint64_t n, bool fixed_base);
void msm_run(Context* ctx, const G1Affine* points, uint64_t n, const Fr* scalars, bool scalars_mont, uint32_t c,
             bool fixed_base, uint64_t point_stride, uint8_t* out_xy, int* is_identity);
void affine_to_mont(Context* ctx, const G1Affine* in, G1Affine* out, uint64_t n);
// prover.cu
Prover* prover_create(Context* ctx, Srs* srs, int log_n, const uint8_t* const* h_pk);
void prover_destroy(Prover* p);
void prover_prove(Prover* P, const uint8_t* hA, const uint8_t* hB, const uint8_t* hC, const uint8_t* h_public,
                  uint64_t n_public, uint8_t* out768, bool wires_on_device);
void prover_round1(Prover* P, const uint8_t* hA, const uint8_t* hB, const uint8_t* hC, const uint8_t* h_public,
                   uint64_t n_public, bool wires_on_device);
void prover_round2(Prover* P, const Fr& beta_c, const Fr& gamma_c);
void prover_round3(Prover* P, const Fr& alpha_c, const Fr& cofactor_c);
void prover_round4(Prover* P, const Fr& zeta_c);
void prover_round5(Prover* P, const Fr& v_c);
void prover_set_shard(Prover* P, uint64_t first, uint64_t count, bool enable);
void prover_serialize(const Prover* P, uint8_t* out768);
void g1_combine_partials_host(const G1XYZZ* parts, uint32_t count, uint8_t* out_xy, int* is_identity);
}  // namespace pb200

using namespace pb200;

static thread_local std::string g_err;

#define PB_API_BEGIN try {
#define PB_API_END                        \
  return 0;                               \
  }                                       \
  catch (const std::exception& e) {       \
    g_err = e.what();                     \
    return 1;                             \
  }                                       \
  catch (...) {                           \
    g_err = "unknown error";              \
    return 1;                             \
  }

static Context* C(pb200_ctx* c) { return reinterpret_cast<Context*>(c); }

static Fr load_fr_canonical(const uint8_t* h) {
  Fr a;
  memcpy(a.v, h, 32);
  // reject non-canonical input
  Fr m = Fr::modulus();
  bool lt = false;
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] != m.v[i]) { lt = a.v[i] < m.v[i]; break; }
  }
  PB_CHECK(lt, "Fr value not reduced below the modulus");
  return a;
}

extern "C" {

const char* pb200_last_error(void) { return g_err.c_str(); }
const char* pb200_version(void) { return "plonk_b200 0.1 (sm_100a)"; }

int pb200_ctx_create(int device, void* cuda_stream, pb200_ctx** out) {
  PB_API_BEGIN
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  PB_CHECK(e == cudaSuccess && count > 0, "no CUDA device: libplonk_b200 has no CPU fallback");
  PB_CHECK(device >= 0 && device < count, "bad device ordinal");
  PB_CUDA(cudaSetDevice(device));
  auto ctx = std::make_unique<Context>();
  ctx->device = device;
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (cuda_stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  } else {
    PB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  *out = reinterpret_cast<pb200_ctx*>(ctx.release());
  PB_API_END
}

void pb200_ctx_destroy(pb200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(C(ctx)->device);
  cudaStreamSynchronize(C(ctx)->stream);
  delete C(ctx);
}

int pb200_ctx_sync(pb200_ctx* ctx) {
  PB_API_BEGIN
  PB_CUDA(cudaStreamSynchronize(C(ctx)->stream));
  PB_API_END
}

uint64_t pb200_ctx_launches(pb200_ctx* ctx) { return C(ctx)->launches; }

int pb200_ctx_timing(pb200_ctx* ctx, int enable) {
  PB_API_BEGIN
  Context* c = C(ctx);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& v : c->timed) {
    for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    v.clear();
  }
  c->timing = enable != 0;
  PB_API_END
}
int pb200_ctx_timing_read(pb200_ctx* ctx, int category, double* total_ms, uint64_t* count) {
  PB_API_BEGIN
  Context* c = C(ctx);
  PB_CHECK(category >= 0 && category < 4, "bad timing category");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  double t = 0;
  for (auto& pr : c->timed[category]) {
    float ms = 0;
    PB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    t += ms;
  }
  *total_ms = t;
  *count = c->timed[category].size();
  PB_API_END
}
void* pb200_ctx_stream(pb200_ctx* ctx) { return (void*)C(ctx)->stream; }

int pb200_fr_to_mont(pb200_ctx* ctx, const void* d_in, void* d_out, uint64_t n) {
  PB_API_BEGIN
  fr_to_mont(C(ctx), (const Fr*)d_in, (Fr*)d_out, n);
  PB_API_END
}
int pb200_fr_from_mont(pb200_ctx* ctx, const void* d_in, void* d_out, uint64_t n) {
  PB_API_BEGIN
  fr_from_mont(C(ctx), (const Fr*)d_in, (Fr*)d_out, n);
  PB_API_END
}

int pb200_fr_ntt(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, int inverse) {
  PB_API_BEGIN
  ntt_run(C(ctx), (const Fr*)d_in, (Fr*)d_out, (int)log_n, inverse != 0, (uint64_t)1 << log_n, nullptr, nullptr);
  PB_API_END
}

int pb200_fr_ntt_decimated(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_m, int inverse, uint64_t stride,
                           uint64_t offset) {
  PB_API_BEGIN
  PB_CHECK(stride >= 1, "bad stride");
  ntt_run_strided(C(ctx), (const Fr*)d_in, (Fr*)d_out, (int)log_m, inverse != 0, (uint64_t)1 << log_m, nullptr, nullptr,
                  stride, offset);
  PB_API_END
}
int pb200_fr_ntt_slab_combine(pb200_ctx* ctx, const void* d_sub, void* d_out, unsigned log_m, unsigned log_g,
                              uint64_t slab, int inverse) {
  PB_API_BEGIN
  PB_CHECK(log_g >= 1 && log_g <= 3 && slab < ((uint64_t)1 << log_g), "slab NTT supports 2, 4 or 8 ranks");
  ntt_slab_combine(C(ctx), (const Fr*)d_sub, (Fr*)d_out, (int)log_m, (int)log_g, slab, inverse != 0);
  PB_API_END
}

// staging helpers for the host-buffer entry points
struct HostStage {
  Context* ctx;
  DevBuf in, out;
  HostStage(Context* c, const uint8_t* h_in, size_t in_bytes, size_t out_bytes) : ctx(c), in(in_bytes), out(out_bytes) {
    if (in_bytes) PB_CUDA(cudaMemcpyAsync(in.p, h_in, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  void finish(uint8_t* h_out, size_t bytes) {
    PB_CUDA(cudaMemcpyAsync(h_out, out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
};

int pb200_fr_ntt_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n, int inverse) {
  PB_API_BEGIN
  size_t bytes = ((size_t)1 << log_n) * 32;
  HostStage st(C(ctx), h_in, bytes, bytes);
  ntt_run(C(ctx), st.in.as<Fr>(), st.out.as<Fr>(), (int)log_n, inverse != 0, (uint64_t)1 << log_n, nullptr, nullptr);
  st.finish(h_out, bytes);
  PB_API_END
}

// poly.py:156-163: ifft(n) ; coefficient i *= offset^i ; zero-pad to 4n ; fft(4n)
static void coset_extend(Context* ctx, const Fr* d_in, Fr* d_out, int log_n, const uint8_t* h_offset) {
  uint64_t n = (uint64_t)1 << log_n;
  Fr off = fp_to_mont(load_fr_canonical(h_offset));
  DevBuf coeffs(n * 32), powers(n * 32);
  ntt_run(ctx, d_in, coeffs.as<Fr>(), log_n, true, n, nullptr, nullptr);
  launch_powers(ctx, powers.as<Fr>(), n, off, Fr::one());
  ntt_run(ctx, coeffs.as<Fr>(), d_out, log_n + 2, false, n, powers.as<Fr>(), nullptr);
  PB_CUDA(cudaStreamSynchronize(ctx->stream));  // temporaries die here
}

// poly.py:169-177: ifft(N) ; coefficient i *= offset^-i
static void coset_to_coeffs(Context* ctx, const Fr* d_in, Fr* d_out, int log_n, const uint8_t* h_offset) {
  uint64_t n = (uint64_t)1 << log_n;
  Fr off = fp_to_mont(load_fr_canonical(h_offset));
  Fr inv = fp_inv(off);  // inv(0) == 0 like py_ecc
  DevBuf powers(n * 32);
  launch_powers(ctx, powers.as<Fr>(), n, inv, Fr::one());
  ntt_run(ctx, d_in, d_out, log_n, true, n, nullptr, powers.as<Fr>());
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
}

int pb200_fr_coset_extend(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, const uint8_t* h_offset) {
  PB_API_BEGIN
  coset_extend(C(ctx), (const Fr*)d_in, (Fr*)d_out, (int)log_n, h_offset);
  PB_API_END
}
int pb200_fr_coset_extend_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n,
                               const uint8_t* h_offset) {
  PB_API_BEGIN
  size_t bytes = ((size_t)1 << log_n) * 32;
  HostStage st(C(ctx), h_in, bytes, bytes * 4);
  coset_extend(C(ctx), st.in.as<Fr>(), st.out.as<Fr>(), (int)log_n, h_offset);
  st.finish(h_out, bytes * 4);
  PB_API_END
}
int pb200_fr_coset_to_coeffs(pb200_ctx* ctx, const void* d_in, void* d_out, unsigned log_n, const uint8_t* h_offset) {
  PB_API_BEGIN
  coset_to_coeffs(C(ctx), (const Fr*)d_in, (Fr*)d_out, (int)log_n, h_offset);
  PB_API_END
}
int pb200_fr_coset_to_coeffs_host(pb200_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, unsigned log_n,
                                  const uint8_t* h_offset) {
  PB_API_BEGIN
  size_t bytes = ((size_t)1 << log_n) * 32;
  HostStage st(C(ctx), h_in, bytes, bytes);
  coset_to_coeffs(C(ctx), st.in.as<Fr>(), st.out.as<Fr>(), (int)log_n, h_offset);
  st.finish(h_out, bytes);
  PB_API_END
}

int pb200_fr_barycentric_eval(pb200_ctx* ctx, const void* d_vals, unsigned log_n, const uint8_t* h_x, uint8_t* h_out) {
  PB_API_BEGIN
  Fr x = fp_to_mont(load_fr_canonical(h_x));
  Fr r;
  barycentric_eval(C(ctx), (const Fr*)d_vals, (int)log_n, x, &r);
  memcpy(h_out, r.v, 32);
  PB_API_END
}
int pb200_fr_barycentric_eval_host(pb200_ctx* ctx, const uint8_t* h_vals, unsigned log_n, const uint8_t* h_x,
                                   uint8_t* h_out) {
  PB_API_BEGIN
  size_t bytes = ((size_t)1 << log_n) * 32;
  HostStage st(C(ctx), h_vals, bytes, 0);
  Fr x = fp_to_mont(load_fr_canonical(h_x));
  Fr r;
  barycentric_eval(C(ctx), st.in.as<Fr>(), (int)log_n, x, &r);
  memcpy(h_out, r.v, 32);
  PB_API_END
}

int pb200_g1_msm(pb200_ctx* ctx, const void* d_points, const void* d_scalars, uint64_t n, uint8_t* h_out_xy,
                 int* is_identity) {
  PB_API_BEGIN
  PB_CHECK(n > 0, "ec_lincomb of an empty list (the reference raises ValueError, curve.py:93)");
  Context* c = C(ctx);
  DevBuf mont(n * sizeof(G1Affine));
  affine_to_mont(c, (const G1Affine*)d_points, mont.as<G1Affine>(), n);
  msm_run(c, mont.as<G1Affine>(), n, (const Fr*)d_scalars, false, msm_default_window(n, false), false, 0, h_out_xy,
          is_identity);
  PB_API_END
}

int pb200_g1_msm_host(pb200_ctx* ctx, const uint8_t* h_points, const uint8_t* h_scalars, uint64_t n,
                      uint8_t* h_out_xy, int* is_identity) {
  PB_API_BEGIN
  PB_CHECK(n > 0, "ec_lincomb of an empty list (the reference raises ValueError, curve.py:93)");
  Context* c = C(ctx);
  DevBuf pts(n * 64), sc(n * 32);
  PB_CUDA(cudaMemcpyAsync(pts.p, h_points, n * 64, cudaMemcpyHostToDevice, c->stream));
  PB_CUDA(cudaMemcpyAsync(sc.p, h_scalars, n * 32, cudaMemcpyHostToDevice, c->stream));
  int rc = pb200_g1_msm(ctx, pts.p, sc.p, n, h_out_xy, is_identity);
  if (rc) throw Error(g_err);
  PB_API_END
}

int pb200_srs_create(pb200_ctx* ctx, const uint8_t* h_points, uint64_t n, int precompute, pb200_srs** out) {
  PB_API_BEGIN
  PB_CHECK(n > 0, "empty SRS");
  *out = reinterpret_cast<pb200_srs*>(srs_create(C(ctx), h_points, n, precompute));
  PB_API_END
}
int pb200_srs_generate(pb200_ctx* ctx, const uint8_t* h_tau, uint64_t n, int precompute, pb200_srs** out) {
  PB_API_BEGIN
  PB_CHECK(n > 0, "empty SRS");
  *out = reinterpret_cast<pb200_srs*>(srs_generate(C(ctx), load_fr_canonical(h_tau), n, precompute));
  PB_API_END
}
int pb200_srs_generate_lagrange(pb200_ctx* ctx, const uint8_t* h_tau, uint64_t n, int precompute, pb200_srs** out) {
  PB_API_BEGIN
  PB_CHECK(n > 0, "empty SRS");
  *out = reinterpret_cast<pb200_srs*>(srs_generate_lagrange(C(ctx), load_fr_canonical(h_tau), n, precompute));
  PB_API_END
}
int pb200_srs_commit_coeffs_host(pb200_ctx* ctx, pb200_srs* srs, const uint8_t* h_coeffs, uint64_t m,
                                 uint8_t* h_out_xy, int* is_identity) {
  PB_API_BEGIN
  HostStage st(C(ctx), h_coeffs, (size_t)m * 32, 0);
  srs_msm(C(ctx), reinterpret_cast<Srs*>(srs), (const Fr*)st.in.p, m, false, h_out_xy, is_identity);
  PB_API_END
}
int pb200_srs_export(pb200_ctx* ctx, pb200_srs* srs, uint8_t* h_points, uint64_t first, uint64_t count) {
  PB_API_BEGIN
  srs_export(C(ctx), reinterpret_cast<Srs*>(srs), h_points, first, count);
  PB_API_END
}
void pb200_srs_destroy(pb200_srs* srs) { srs_destroy(reinterpret_cast<Srs*>(srs)); }
uint64_t pb200_srs_size(pb200_srs* srs) { return srs_size(reinterpret_cast<Srs*>(srs)); }

int pb200_srs_commit_lagrange(pb200_ctx* ctx, pb200_srs* srs, const void* d_values, unsigned log_n, uint8_t* h_out_xy,
                              int* is_identity) {
  PB_API_BEGIN
  Context* c = C(ctx);
  uint64_t n = (uint64_t)1 << log_n;
  PB_CHECK(n <= srs_size(reinterpret_cast<Srs*>(srs)), "Not enough powers in setup");
  DevBuf coeffs(n * 32);
  ntt_run(c, (const Fr*)d_values, coeffs.as<Fr>(), (int)log_n, true, n, nullptr, nullptr);
  srs_msm(c, reinterpret_cast<Srs*>(srs), coeffs.as<Fr>(), n, false, h_out_xy, is_identity);
  PB_API_END
}
int pb200_srs_commit_lagrange_host(pb200_ctx* ctx, pb200_srs* srs, const uint8_t* h_values, unsigned log_n,
                                   uint8_t* h_out_xy, int* is_identity) {
  PB_API_BEGIN
  size_t bytes = ((size_t)1 << log_n) * 32;
  HostStage st(C(ctx), h_values, bytes, 0);
  int rc = pb200_srs_commit_lagrange(ctx, srs, st.in.p, log_n, h_out_xy, is_identity);
  if (rc) throw Error(g_err);
  PB_API_END
}
int pb200_srs_commit_coeffs(pb200_ctx* ctx, pb200_srs* srs, const void* d_coeffs, uint64_t m, int coeffs_montgomery,
                            uint8_t* h_out_xy, int* is_identity) {
  PB_API_BEGIN
  srs_msm(C(ctx), reinterpret_cast<Srs*>(srs), (const Fr*)d_coeffs, m, coeffs_montgomery != 0, h_out_xy, is_identity);
  PB_API_END
}

int pb200_prover_create(pb200_ctx* ctx, pb200_srs* srs, unsigned log_n, const uint8_t* const* h_pk,
                        pb200_prover** out) {
  PB_API_BEGIN
  *out = reinterpret_cast<pb200_prover*>(prover_create(C(ctx), reinterpret_cast<Srs*>(srs), (int)log_n, h_pk));
  PB_API_END
}
void pb200_prover_destroy(pb200_prover* p) { prover_destroy(reinterpret_cast<Prover*>(p)); }

int pb200_prover_prove(pb200_prover* p, const uint8_t* h_A, const uint8_t* h_B, const uint8_t* h_C,
                       const uint8_t* h_public, uint64_t n_public, uint8_t* h_proof768) {
  PB_API_BEGIN
  prover_prove(reinterpret_cast<Prover*>(p), h_A, h_B, h_C, h_public, n_public, h_proof768, false);
  PB_API_END
}
int pb200_prover_prove_device(pb200_prover* p, const void* d_A, const void* d_B, const void* d_C,
                              const uint8_t* h_public, uint64_t n_public, uint8_t* h_proof768) {
  PB_API_BEGIN
  prover_prove(reinterpret_cast<Prover*>(p), (const uint8_t*)d_A, (const uint8_t*)d_B, (const uint8_t*)d_C, h_public,
               n_public, h_proof768, true);
  PB_API_END
}
int pb200_prover_round1(pb200_prover* p, const uint8_t* h_A, const uint8_t* h_B, const uint8_t* h_C,
                        const uint8_t* h_public, uint64_t n_public, uint8_t* h_abc_xy) {
  PB_API_BEGIN
  Prover* P = reinterpret_cast<Prover*>(p);
  prover_round1(P, h_A, h_B, h_C, h_public, n_public, false);
  memcpy(h_abc_xy, P->proof.pts[0], 3 * 64);
  PB_API_END
}
int pb200_prover_round2(pb200_prover* p, const uint8_t* beta, const uint8_t* gamma, uint8_t* h_z_xy) {
  PB_API_BEGIN
  Prover* P = reinterpret_cast<Prover*>(p);
  prover_round2(P, load_fr_canonical(beta), load_fr_canonical(gamma));
  memcpy(h_z_xy, P->proof.pts[3], 64);
  PB_API_END
}
int pb200_prover_round3(pb200_prover* p, const uint8_t* alpha, const uint8_t* fft_cofactor, uint8_t* h_t_xy) {
  PB_API_BEGIN
  Prover* P = reinterpret_cast<Prover*>(p);
  prover_round3(P, load_fr_canonical(alpha), load_fr_canonical(fft_cofactor));
  memcpy(h_t_xy, P->proof.pts[4], 3 * 64);
  PB_API_END
}
int pb200_prover_round4(pb200_prover* p, const uint8_t* zeta, uint8_t* h_evals) {
  PB_API_BEGIN
  Prover* P = reinterpret_cast<Prover*>(p);
  prover_round4(P, load_fr_canonical(zeta));
  memcpy(h_evals, P->proof.evals[0], 6 * 32);
  PB_API_END
}
int pb200_prover_round5(pb200_prover* p, const uint8_t* v, uint8_t* h_w_xy) {
  PB_API_BEGIN
  Prover* P = reinterpret_cast<Prover*>(p);
  prover_round5(P, load_fr_canonical(v));
  memcpy(h_w_xy, P->proof.pts[7], 2 * 64);
  PB_API_END
}

int pb200_prover_set_shard(pb200_prover* p, uint64_t first, uint64_t count, int enable) {
  PB_API_BEGIN
  prover_set_shard(reinterpret_cast<Prover*>(p), first, count, enable != 0);
  PB_API_END
}
int pb200_prover_read_partials(pb200_prover* p, unsigned first_slot, unsigned count, uint8_t* h_xyzz) {
  PB_API_BEGIN
  PB_CHECK(first_slot + count <= 9, "proof has 9 commitment slots");
  memcpy(h_xyzz, reinterpret_cast<Prover*>(p)->partials + first_slot, (size_t)count * sizeof(G1XYZZ));
  PB_API_END
}
int pb200_prover_set_points(pb200_prover* p, unsigned first_slot, unsigned count, const uint8_t* h_xy) {
  PB_API_BEGIN
  PB_CHECK(first_slot + count <= 9, "proof has 9 commitment slots");
  memcpy(reinterpret_cast<Prover*>(p)->proof.pts[first_slot], h_xy, (size_t)count * 64);
  PB_API_END
}
int pb200_prover_serialize(pb200_prover* p, uint8_t* h_proof768) {
  PB_API_BEGIN
  prover_serialize(reinterpret_cast<Prover*>(p), h_proof768);
  PB_API_END
}
int pb200_g1_combine_partials_host(const uint8_t* h_xyzz, unsigned count, uint8_t* h_out_xy, int* is_identity) {
  PB_API_BEGIN
  std::vector<G1XYZZ> parts(count);
  memcpy(parts.data(), h_xyzz, (size_t)count * sizeof(G1XYZZ));
  g1_combine_partials_host(parts.data(), count, h_out_xy, is_identity);
  PB_API_END
}

int pb200_transcript_create(const uint8_t* label, size_t label_len, pb200_transcript** out) {
  PB_API_BEGIN
  *out = reinterpret_cast<pb200_transcript*>(new Transcript(std::string((const char*)label, label_len)));
  PB_API_END
}
void pb200_transcript_destroy(pb200_transcript* t) { delete reinterpret_cast<Transcript*>(t); }
int pb200_transcript_append_message(pb200_transcript* t, const uint8_t* label, size_t label_len, const uint8_t* msg,
                                    size_t msg_len) {
  PB_API_BEGIN
  reinterpret_cast<Transcript*>(t)->append_message(std::string((const char*)label, label_len), msg, msg_len);
  PB_API_END
}
int pb200_transcript_challenge_bytes(pb200_transcript* t, const uint8_t* label, size_t label_len, uint8_t* out,
                                     size_t n) {
  PB_API_BEGIN
  reinterpret_cast<Transcript*>(t)->challenge_bytes(std::string((const char*)label, label_len), out, n);
  PB_API_END
}
int pb200_transcript_get_and_append_challenge(pb200_transcript* t, const uint8_t* label, size_t label_len,
                                              uint8_t* out_le32) {
  PB_API_BEGIN
  Fr f = reinterpret_cast<Transcript*>(t)->get_and_append_challenge(std::string((const char*)label, label_len));
  memcpy(out_le32, f.v, 32);
  PB_API_END
}

// ---- pairing / G2 (host code, pairing.cuh) ----
static Fq load_fq_canonical(const uint8_t* h) {
  Fq a;
  memcpy(a.v, h, 32);
  Fq m = Fq::modulus();
  bool lt = false;
  for (int i = 7; i >= 0; i--) {
    if (a.v[i] != m.v[i]) { lt = a.v[i] < m.v[i]; break; }
  }
  PB_CHECK(lt, "Fq value not reduced below the modulus");
  return fp_to_mont(a);
}
static G2Affine load_g2(const uint8_t* h, bool inf) {
  G2Affine q;
  q.inf = inf;
  q.x = {load_fq_canonical(h), load_fq_canonical(h + 32)};
  q.y = {load_fq_canonical(h + 64), load_fq_canonical(h + 96)};
  PB_CHECK(inf || g2_on_curve(q), "G2 point is not on the twist curve");
  return q;
}
static void store_g2(const G2Affine& q, uint8_t* out, int* is_identity) {
  *is_identity = q.inf ? 1 : 0;
  memset(out, 0, 128);
  if (q.inf) return;
  const Fq c[4] = {fp_from_mont(q.x.a), fp_from_mont(q.x.b), fp_from_mont(q.y.a), fp_from_mont(q.y.b)};
  for (int k = 0; k < 4; k++) memcpy(out + 32 * k, c[k].v, 32);
}
static const Bn254Pairing& pairing_engine() {
  static const Bn254Pairing engine;  // constants derived once (thread-safe static initialisation)
  return engine;
}

int pb200_pairing_check(const uint8_t* h_g1, const uint8_t* h_g1_identity, const uint8_t* h_g2,
                        const uint8_t* h_g2_identity, unsigned count, int* ok) {
  PB_API_BEGIN
  std::vector<G1Host> ps(count);
  std::vector<G2Affine> qs(count);
  const Fq three = fq_small(3);
  for (unsigned i = 0; i < count; i++) {
    ps[i].inf = h_g1_identity && h_g1_identity[i];
    if (!ps[i].inf) {
      ps[i].x = load_fq_canonical(h_g1 + 64 * i);
      ps[i].y = load_fq_canonical(h_g1 + 64 * i + 32);
      PB_CHECK(fp_sqr(ps[i].y) == fp_add(fp_mul(fp_sqr(ps[i].x), ps[i].x), three), "G1 point is not on the curve");
    }
    qs[i] = load_g2(h_g2 + 128 * i, h_g2_identity && h_g2_identity[i]);
  }
  *ok = pairing_engine().product_is_one(ps, qs) ? 1 : 0;
  PB_API_END
}
int pb200_g2_mul(const uint8_t* h_point, const uint8_t* h_scalar_le32, uint8_t* h_out, int* is_identity) {
  PB_API_BEGIN
  G2Affine p = load_g2(h_point, false);
  uint32_t k[8];
  memcpy(k, h_scalar_le32, 32);
  store_g2(g2_mul(p, k), h_out, is_identity);
  PB_API_END
}
int pb200_g2_add(const uint8_t* h_p, int p_identity, const uint8_t* h_q, int q_identity, uint8_t* h_out,
                 int* is_identity) {
  PB_API_BEGIN
  store_g2(g2_add(load_g2(h_p, p_identity != 0), load_g2(h_q, q_identity != 0)), h_out, is_identity);
  PB_API_END
}

int pb200_bench_modmul(pb200_ctx* ctx, int field, uint64_t threads, uint32_t iters, float* ms_out) {
  PB_API_BEGIN
  *ms_out = bench_modmul(C(ctx), field, threads, iters);
  PB_API_END
}

}  // extern "C"
