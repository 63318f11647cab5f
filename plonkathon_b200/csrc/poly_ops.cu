// Element-wise and reduction kernels over Fr vectors used around the NTT / MSM cores:
// canonical <-> Montgomery conversion, barycentric evaluation (poly.py:181-195), and the
// modmul throughput micro-benchmark that gives the integer-pipe ceiling the rooflines are read against.
#include "common.cuh"

namespace pb200 {

__global__ void k_fr_to_mont(const Fr* in, Fr* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fp_to_mont(in[i]);
}
__global__ void k_fr_from_mont(const Fr* in, Fr* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fp_from_mont(in[i]);
}

void fr_to_mont(Context* ctx, const Fr* in, Fr* out, uint64_t n) {
  if (!n) return;
  k_fr_to_mont<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, n);
  ctx->launches++;
  PB_CUDA(cudaGetLastError());
}
void fr_from_mont(Context* ctx, const Fr* in, Fr* out, uint64_t n) {
  if (!n) return;
  k_fr_from_mont<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in, out, n);
  ctx->launches++;
  PB_CUDA(cudaGetLastError());
}

// ---- element-wise ring operations on canonical vectors (poly.py:23-109) -------------------------------------------
// Polynomial.__add__ / __sub__ / __mul__ / __truediv__ with a Polynomial or a Scalar operand, and Polynomial.shift,
// for device-resident operands.  Canonical in, canonical out: a * b = mont_mul(mont_mul(a, b), R^2); a scalar operand
// arrives in Montgomery form so one product suffices; x / y uses py_ecc's convention inv(0) == 0 (FQ.__truediv__),
// with one safegcd inversion shared by 8 divisions (Montgomery's trick over the strided set {t, t+T, ...}).
enum VecOp { VOP_ADD = 0, VOP_SUB, VOP_MUL, VOP_DIV, VOP_ADD_S, VOP_SUB_S, VOP_MUL_S, VOP_ADD_S0, VOP_SUB_S0, VOP_SHIFT };

__global__ void k_vec_op(int op, const Fr* a, const Fr* b, Fr s_canon, Fr s_mont, Fr* out, uint64_t n, uint64_t shift) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr x = a[op == VOP_SHIFT ? (i + shift) % n : i], r;
  switch (op) {
    case VOP_ADD: r = fp_add(x, b[i]); break;
    case VOP_SUB: r = fp_sub(x, b[i]); break;
    case VOP_MUL: r = fp_mul(fp_mul(x, b[i]), Fr::r2()); break;
    case VOP_ADD_S: r = fp_add(x, s_canon); break;
    case VOP_SUB_S: r = fp_sub(x, s_canon); break;
    case VOP_MUL_S: r = fp_mul(x, s_mont); break;
    case VOP_ADD_S0: r = i == 0 ? fp_add(x, s_canon) : x; break;
    case VOP_SUB_S0: r = i == 0 ? fp_sub(x, s_canon) : x; break;
    default: r = x; break;  // VOP_SHIFT
  }
  out[i] = r;
}

__global__ void __launch_bounds__(128) k_vec_div(const Fr* a, const Fr* b, Fr* out, uint64_t n, uint64_t T) {
  const int CH = 8;
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  Fr pref[CH], den[CH];
  Fr run = Fr::one();
  int cnt = 0;
  for (int k = 0; k < CH; k++) {
    uint64_t i = t + (uint64_t)k * T;
    if (i >= n) break;
    Fr d = fp_to_mont(b[i]);
    if (d.is_zero()) d = Fr::one();
    den[k] = d;
    pref[k] = run;
    run = fp_mul(run, d);
    cnt++;
  }
  Fr inv = fp_inv_gcd(run);
  for (int k = cnt - 1; k >= 0; k--) {
    uint64_t i = t + (uint64_t)k * T;
    Fr ik = fp_mul(inv, pref[k]);     // 1 / b_i, Montgomery form
    inv = fp_mul(inv, den[k]);
    out[i] = b[i].is_zero() ? Fr::zero() : fp_mul(a[i], ik);  // canonical * Montgomery -> canonical
  }
}

void fr_vec_op(Context* ctx, int op, const Fr* a, const Fr* b, const Fr& scalar_canonical, Fr* out, uint64_t n,
               uint64_t shift) {
  if (!n) return;
  PB_CHECK(op >= VOP_ADD && op <= VOP_SHIFT, "bad vector operation");
  if (op == VOP_DIV) {
    uint64_t T = (n + 7) / 8;
    k_vec_div<<<(unsigned)((T + 127) / 128), 128, 0, ctx->stream>>>(a, b, out, n, T);
  } else {
    PB_CHECK(op != VOP_SHIFT || a != out, "shift cannot run in place");
    k_vec_op<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(op, a, b, scalar_canonical, fp_to_mont(scalar_canonical),
                                                                  out, n, shift);
  }
  ctx->launches++;
  PB_CUDA(cudaGetLastError());
}

// ---- block-wide sum of Fr (any form; addition is form-agnostic) ------------------------------
template <int NT>
__device__ __forceinline__ Fr block_sum(Fr v, Fr* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int d = NT >> 1; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] = fp_add(sh[threadIdx.x], sh[threadIdx.x + d]);
    __syncthreads();
  }
  Fr r = sh[0];
  __syncthreads();
  return r;
}

// ---- barycentric evaluation: (x^n - 1)/n * sum_i v_i w^i / (x - w^i), inv(0) = 0 ---------------
// vals canonical or Montgomery (the sum keeps the form of vals); x, w in Montgomery form.
__global__ void __launch_bounds__(128) k_bary_partial(const Fr* vals, uint64_t n, Fr x, Fr w, Fr w_inv, Fr* partial) {
  const int CH = 8;
  __shared__ Fr sh[128];
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = t * CH;
  Fr acc = Fr::zero();
  if (i0 < n) {
    int cnt = (int)min((uint64_t)CH, n - i0);
    Fr pref[CH];
    Fr wk = fp_pow_u64(w, i0);
    Fr run = Fr::one();
    for (int k = 0; k < cnt; k++) {
      Fr d = fp_sub(x, wk);
      if (d.is_zero()) d = Fr::one();
      pref[k] = run;
      run = fp_mul(run, d);
      if (k + 1 < cnt) wk = fp_mul(wk, w);
    }
    Fr inv = fp_inv_gcd(run);
    for (int k = cnt - 1; k >= 0; k--) {
      Fr d = fp_sub(x, wk);
      bool z = d.is_zero();
      if (z) d = Fr::one();
      Fr ik = fp_mul(inv, pref[k]);
      inv = fp_mul(inv, d);
      if (!z) {
        Fr term = fp_mul(vals[i0 + k], fp_mul(wk, ik));
        acc = fp_add(acc, term);
      }
      wk = fp_mul(wk, w_inv);
    }
  }
  Fr s = block_sum<128>(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out = factor * sum(partial[0..m))   (factor Montgomery)
__global__ void __launch_bounds__(128) k_sum_scale(const Fr* partial, uint32_t m, Fr factor, Fr* out) {
  __shared__ Fr sh[128];
  Fr acc = Fr::zero();
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) acc = fp_add(acc, partial[i]);
  Fr s = block_sum<128>(acc, sh);
  if (threadIdx.x == 0) out[0] = fp_mul(s, factor);
}

Fr fr_from_u64(uint64_t x);
Fr fr_root_of_unity(int log_n);

// result has the form (canonical / Montgomery) of vals
void barycentric_eval(Context* ctx, const Fr* d_vals, int log_n, const Fr& x_mont, Fr* h_out) {
  uint64_t n = (uint64_t)1 << log_n;
  Fr w = fr_root_of_unity(log_n);
  Fr w_inv = fp_inv(w);
  uint64_t threads = (n + 7) / 8;
  uint32_t blocks = (uint32_t)((threads + 127) / 128);
  ctx->scratch[0].ensure((size_t)(blocks + 1) * 32);
  Fr* partial = ctx->scratch[0].as<Fr>();
  k_bary_partial<<<blocks, 128, 0, ctx->stream>>>(d_vals, n, x_mont, w, w_inv, partial);
  // (x^n - 1) / n
  Fr xn = fp_pow_u64(x_mont, n);
  Fr factor = fp_mul(fp_sub(xn, Fr::one()), fp_inv(fr_from_u64(n)));
  k_sum_scale<<<1, 128, 0, ctx->stream>>>(partial, blocks, factor, partial + blocks);
  ctx->launches += 2;
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaMemcpyAsync(h_out, partial + blocks, 32, cudaMemcpyDeviceToHost, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// ---- modmul throughput micro-benchmark --------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) k_bench_modmul(F* sink, uint32_t iters, F seed) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  F a = seed, b = seed;
  a.v[0] ^= t;
  b.v[1] ^= t * 2654435761u;
  a.v[7] &= 0x0fffffffu;
  b.v[7] &= 0x0fffffffu;
  // two independent dependency chains per thread
  for (uint32_t i = 0; i < iters; i += 2) {
    a = fp_mul(a, a);
    b = fp_mul(b, b);
  }
  F r = fp_add(a, b);
  if (r.v[0] == 0x12345678u && r.v[3] == 42u) sink[t & 1023] = r;  // practically never; defeats DCE
}

float bench_modmul(Context* ctx, int field, uint64_t threads, uint32_t iters) {
  ctx->scratch[0].ensure(1024 * 32);
  cudaEvent_t e0, e1;
  PB_CUDA(cudaEventCreate(&e0));
  PB_CUDA(cudaEventCreate(&e1));
  unsigned blocks = (unsigned)((threads + 255) / 256);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {  // first run warms up
    PB_CUDA(cudaEventRecord(e0, ctx->stream));
    if (field == 0) k_bench_modmul<Fr><<<blocks, 256, 0, ctx->stream>>>(ctx->scratch[0].as<Fr>(), iters, Fr::r2());
    else k_bench_modmul<Fq><<<blocks, 256, 0, ctx->stream>>>(ctx->scratch[0].as<Fq>(), iters, Fq::r2());
    ctx->launches++;
    PB_CUDA(cudaEventRecord(e1, ctx->stream));
    PB_CUDA(cudaEventSynchronize(e1));
    PB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms;
}

}  // namespace pb200
