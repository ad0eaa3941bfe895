// pointwise.cu -- per-sample epilogues.
//
// unsharp combine: the point pass of UnsharpMaskImage (MagickCore/effect.c:4310-4384):
//   d = p - blur;  out = (|2d| < QuantumRange*threshold) ? p : p + gain*d     (double -> float)
// applied to every channel (all carry the Update trait on the accelerated path).
#include "mb200_internal.h"

#include <cuda_runtime.h>

namespace mb200 {
namespace {

__global__ void __launch_bounds__(256) unsharp_kernel(const float4 *__restrict__ src, float4 *__restrict__ blur,
                                                      size_t n4, double gain, double qthreshold) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 p = __ldg(src + i);
  float4 b = blur[i];
  auto one = [&](float pv, float bv) -> float {
    double pixel = static_cast<double>(pv) - static_cast<double>(bv);
    if (fabs(2.0 * pixel) < qthreshold) pixel = static_cast<double>(pv);
    else pixel = static_cast<double>(pv) + gain * pixel;
    return static_cast<float>(pixel);
  };
  b.x = one(p.x, b.x); b.y = one(p.y, b.y); b.z = one(p.z, b.z); b.w = one(p.w, b.w);
  blur[i] = b;
}

__global__ void __launch_bounds__(256) unsharp_tail_kernel(const float *__restrict__ src, float *__restrict__ blur,
                                                           size_t begin, size_t n, double gain, double qthreshold) {
  const size_t i = begin + static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const double pv = static_cast<double>(src[i]);
  double pixel = pv - static_cast<double>(blur[i]);
  if (fabs(2.0 * pixel) < qthreshold) pixel = pv;
  else pixel = pv + gain * pixel;
  blur[i] = static_cast<float>(pixel);
}

}  // namespace

int launch_unsharp_combine(const float *src, float *blur_inout, size_t n, double gain, double quantum_threshold,
                           void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(blur_inout)) & 15) == 0;
  const size_t n4 = aligned ? n / 4 : 0;
  if (n4) {
    unsharp_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<const float4 *>(src), reinterpret_cast<float4 *>(blur_inout), n4, gain, quantum_threshold);
    count_launch();
  }
  if (n4 * 4 < n) {
    const size_t rest = n - n4 * 4;
    unsharp_tail_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, s>>>(src, blur_inout, n4 * 4, n, gain,
                                                                                 quantum_threshold);
    count_launch();
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "unsharp launch");
  return MB200_OK;
}

}  // namespace mb200
