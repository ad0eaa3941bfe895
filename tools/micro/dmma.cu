// Is FP64 mma.sync (DMMA m8n8k4) on sm_100a a separate pipe from DFMA, and what is its rate?
//   mode 0: all warps DFMA;  mode 1: all warps DMMA;  mode 2: even warps DFMA, odd warps DMMA
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double *out, double a, double b, int iters, int mode) {
  const int warp = threadIdx.x >> 5;
  const bool use_mma = mode == 1 || (mode == 2 && (warp & 1));
  double acc[2 * NACC];
#pragma unroll
  for (int i = 0; i < 2 * NACC; ++i) acc[i] = threadIdx.x + i;
  if (use_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) dmma(acc[2 * i], acc[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 2 * NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  double s = 0;
  for (int i = 0; i < 2 * NACC; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  void *buf; cudaMalloc(&buf, 1024);
  const int iters = 4096, sms = 148 * 2;
  constexpr int NACC = 16;
  for (int warps : {8, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      float ms = timeit([&] { k<NACC><<<sms, warps * 32>>>((double *) buf, 1.0000001, 1e-9, iters, mode); });
      // FMA count: DFMA warp-iteration = 32 lanes * 2*NACC ; DMMA warp-iteration = NACC * 256
      double fma_dfma = 0, fma_dmma = 0;
      const double wtot = (double) sms * warps * iters;
      if (mode == 0) fma_dfma = wtot * 32 * 2 * NACC;
      if (mode == 1) fma_dmma = wtot * NACC * 256;
      if (mode == 2) { fma_dfma = wtot / 2 * 32 * 2 * NACC; fma_dmma = wtot / 2 * NACC * 256; }
      printf("warps/CTA=%2d mode=%d: %.3f ms  DFMA %.2f T/s  DMMA %.2f Tfma/s  total %.2f Tfma/s\n", warps, mode, ms,
             fma_dfma / ms * 1e-9, fma_dmma / ms * 1e-9, (fma_dfma + fma_dmma) / ms * 1e-9);
    }
  }
  return 0;
}
