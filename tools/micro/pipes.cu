// Pipe-throughput probes for B200 (sm_100a): DFMA, F2F.F64.F32, SHFL, FFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k_dfma(double *out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0; for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}
template <int ILP> __global__ void k_ffma(float *out, float a, float b, int iters) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0; for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}
template <int ILP> __global__ void k_cvt(double *out, float seed, int iters) {
  float f[ILP]; double acc = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) f[i] = seed + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) { double d = (double) f[i]; f[i] = __int_as_float(__float_as_int(f[i]) + __double2hiint(d)); }
  }
  for (int i = 0; i < ILP; ++i) acc += f[i];
  if (acc == 123.456) out[0] = acc;
}
template <int ILP> __global__ void k_shfl(float *out, float seed, int iters) {
  float f[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) f[i] = seed + threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) f[i] = __shfl_sync(0xffffffffu, f[i], (threadIdx.x | 3) & 31);
  }
  float s = 0; for (int i = 0; i < ILP; ++i) s += f[i];
  if (s == 123.456f) out[0] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  void *buf; cudaMalloc(&buf, 1024);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; printf("SMs %d clock %d kHz\n", sms, p.clockRate);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32; int blocks = sms * 1;
    if (threads > 1024) continue;
    float ms = timeit([&] { k_dfma<16><<<blocks, threads>>>((double *) buf, 1.0000001, 1e-9, iters); });
    double ops = (double) blocks * threads * 16.0 * iters;
    printf("DFMA  warps/SM %2d: %.3f ms  %.2f TFMA/s  (%.1f lanes/clk/SM @1.965GHz)\n", warps, ms, ops / ms / 1e9, ops / ms / 1e6 / sms / 1.965e3 / 1e3*1e3);
  }
  for (int warps : {8, 32}) {
    int threads = warps * 32; int blocks = sms;
    float ms = timeit([&] { k_ffma<16><<<blocks, threads>>>((float *) buf, 1.0000001f, 1e-9f, iters); });
    double ops = (double) blocks * threads * 16.0 * iters;
    printf("FFMA  warps/SM %2d: %.3f ms  %.2f TFMA/s\n", warps, ms, ops / ms / 1e9);
    ms = timeit([&] { k_cvt<8><<<blocks, threads>>>((double *) buf, 1.5f, iters); });
    ops = (double) blocks * threads * 8.0 * iters;
    printf("F2F.F64.F32 (+int ops) warps/SM %2d: %.3f ms  %.2f Tcvt/s (%.1f lanes/clk/SM)\n", warps, ms, ops / ms / 1e9, ops / ms / 1e6 / sms / 1.965e3);
    ms = timeit([&] { k_shfl<8><<<blocks, threads>>>((float *) buf, 1.5f, iters); });
    printf("SHFL  warps/SM %2d: %.3f ms  %.2f Tshfl/s (%.1f lanes/clk/SM)\n", warps, ms, ops / ms / 1e9, ops / ms / 1e6 / sms / 1.965e3);
  }
  return 0;
}
