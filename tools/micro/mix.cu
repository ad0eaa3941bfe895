// Does a half-rate DFMA block the issue port for other pipes?  mix of NF DFMAs + NA ALU ops / + NX F2F per iteration.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NA, int NX>
__global__ void k_mix(double *out, double a, double b, int iters, int seed) {
  double acc[NF > 0 ? NF : 1];
  int ia[NA > 0 ? NA : 1];
  float fx[NX > 0 ? NX : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x + i;
#pragma unroll
  for (int i = 0; i < NA; ++i) ia[i] = seed + i;
#pragma unroll
  for (int i = 0; i < NX; ++i) fx[i] = seed + i + 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NF > NA ? NF : NA); ++i) {
      if (i < NF) acc[i] = fma(acc[i], a, b);
      if (i < NA) ia[i] = (ia[i] ^ (ia[i] >> 3)) + seed;   // 2 ALU ops (LOP3/SHF + IADD)
      if (i < NX) { double d = (double) fx[i]; fx[i] = __int_as_float(__double2hiint(d)); }
    }
  }
  double s = 0;
  for (int i = 0; i < NF; ++i) s += acc[i];
  for (int i = 0; i < NA; ++i) s += ia[i];
  for (int i = 0; i < NX; ++i) s += fx[i];
  if (s == 123.456) out[0] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
template <int NF, int NA, int NX> void run(int warps) {
  void *buf; cudaMalloc(&buf, 1024);
  const int iters = 2048, sms = 148;
  float ms = timeit([&] { k_mix<NF, NA, NX><<<sms, warps * 32>>>((double *) buf, 1.0000001, 1e-9, iters, 3); });
  double cyc = ms * 1e-3 * 1.965e9;                         // SM cycles
  double per_iter_per_smsp = cyc / iters / (warps / 4.0);   // cycles per (warp-iteration) per SMSP
  printf("NF=%2d NA(x2 ops)=%2d NX=%2d warps/SM=%2d : %.3f ms, %.1f cycles per warp-iteration per SMSP\n", NF, NA, NX, warps, ms, per_iter_per_smsp);
  cudaFree(buf);
}
int main() {
  for (int w : {4, 16}) {
    run<32, 0, 0>(w); run<0, 16, 0>(w); run<32, 8, 0>(w); run<32, 16, 0>(w); run<32, 32, 0>(w);
    run<32, 0, 4>(w); run<32, 16, 4>(w); run<16, 16, 0>(w);
  }
  return 0;
}
