// DFMA-pipe Montgomery product (fieldd.cuh) vs the IMAD product (field.cuh): throughput and a device-side
// cross-check  fp_mul(a,b) == 16 * fpd_mul(a,b)  (the radices differ by 2^4)
#include <cstdio>
#include "fieldd.cuh"
using namespace pb200;
template <int MODE>  // 0: IMAD, 1: DFMA, 2: even warps IMAD / odd warps DFMA
__global__ void __launch_bounds__(256) k(Fq* sink, uint32_t iters, Fq seed) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq a = seed, b = seed;
  a.v[0] ^= t; b.v[1] ^= t * 2654435761u; a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu;
  bool use_d = MODE == 1 || (MODE == 2 && ((threadIdx.x >> 5) & 1));
  Fq r;
  if (use_d) {
    FpD<FqParams> x = fpd_from_u32(a), y = fpd_from_u32(b);
    for (uint32_t i = 0; i < iters; i += 2) { x = fpd_mul(x, y); y = fpd_mul(y, x); }
    r = fp_add(fpd_to_u32(x), fpd_to_u32(y));
  } else {
    for (uint32_t i = 0; i < iters; i += 2) { a = fp_mul(a, b); b = fp_mul(b, a); }
    r = fp_add(a, b);
  }
  if (r.v[0] == 0x12345678u && r.v[3] == 42u) sink[t & 1023] = r;
}
__global__ void check(uint32_t* bad, Fq seed) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq a = seed, b = seed;
  a.v[0] ^= t * 747796405u; b.v[1] ^= t * 2654435761u; a.v[5] ^= t; a.v[7] &= 0x0fffffffu; b.v[7] &= 0x0fffffffu;
  for (int it = 0; it < 8; it++) {
    Fq m = fp_mul(a, b);
    Fq d = fpd_to_u32(fpd_mul(fpd_from_u32(a), fpd_from_u32(b)));
    for (int k = 0; k < 4; k++) d = fp_dbl(d);
    if (m != d) atomicAdd(bad, 1u);
    a = m; b = fp_add(b, m);
  }
}
template <int MODE> void run(const char* name, Fq* sink) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned blocks = 148 * 8; uint32_t iters = 2048;
  k<MODE><<<blocks, 256>>>(sink, iters, Fq::r2());
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(sink, iters, Fq::r2());
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-22s %.3f ms  %.2f Gmul/s\n", name, ms, (double)blocks * 256 * iters / ms / 1e6);
}
int main() {
  Fq* sink; cudaMalloc(&sink, 1024 * 32);
  uint32_t* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  check<<<1024, 256>>>(bad, Fq::r2());
  uint32_t h; cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
  printf("cross-check mismatches: %u (of %d)  [%s]\n", h, 1024 * 256 * 8, cudaGetErrorString(cudaGetLastError()));
  run<0>("IMAD (field.cuh)", sink);
  run<1>("DFMA (fieldd.cuh)", sink);
  run<2>("half IMAD / half DFMA", sink);
  return 0;
}
