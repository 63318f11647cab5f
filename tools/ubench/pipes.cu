// pipe throughput micro-benchmarks on B200: DFMA (fp64), IMAD.WIDE (fmaheavy), IADD3 (alu), and mixes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, uint64_t* outi, double seed, uint32_t iseed) {
  double d[8];
  uint64_t w[8];
  uint32_t a[8];
  for (int i = 0; i < 8; i++) { d[i] = seed + i + threadIdx.x; w[i] = iseed + i * 977 + threadIdx.x; a[i] = iseed * (i + 3) + threadIdx.x; }
  double m = seed * 1.0000001, c = 3.0;
  uint32_t mi = iseed | 1;
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 3 || MODE == 4) d[i] = __fma_rz(d[i], m, c);
      if (MODE == 1 || MODE == 3 || MODE == 5) w[i] = (uint64_t)(uint32_t)w[i] * mi + w[i];  // IMAD.WIDE
      if (MODE == 2 || MODE == 4 || MODE == 5) { a[i] = a[i] + (a[(i + 1) & 7] ^ mi); }       // alu
    }
  }
  double s = 0; uint64_t t = 0;
  for (int i = 0; i < 8; i++) { s += d[i]; t += w[i] + a[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> void run(const char* name, int ops_per_iter) {
  int blocks = 148 * 8, threads = 256;
  double* o; uint64_t* oi;
  cudaMalloc(&o, blocks * threads * 8); cudaMalloc(&oi, blocks * threads * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(o, oi, 1.5, 12345);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(o, oi, 1.5, 12345);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * ITER * 8 * ops_per_iter;
  // per SM per clock assuming 1.965 GHz
  printf("%-28s %8.3f ms  %7.2f Gops/s  %6.1f ops/clk/SM (at 1965 MHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(o); cudaFree(oi);
}
int main() {
  run<0>("DFMA", 1);
  run<1>("IMAD.WIDE", 1);
  run<2>("IADD/LOP (alu)", 2);
  run<3>("DFMA + IMAD.WIDE", 2);
  run<4>("DFMA + alu", 3);
  run<5>("IMAD.WIDE + alu", 3);
  return 0;
}
