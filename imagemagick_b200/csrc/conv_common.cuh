// conv_common.cuh -- device helpers shared by the 1-D convolution kernels (conv1d.cu: DFMA streaming kernels,
// conv_mma.cu: FP64 mma.sync kernels): the reference's PerceptibleReciprocal clamp, the reciprocal, UnsharpMask's point pass.
#pragma once

#include <cuda_runtime.h>

namespace mb200 {
namespace {

constexpr double kQuantumScale = 1.0 / 65535.0;
constexpr double kEpsilon = 1.0e-12;

// A zero-padded tap multiplies a sample OUTSIDE the reference's window; 0 * (+-inf | NaN) = NaN would poison outputs
// the reference computes from finite samples only (HDRI pixels may be non-finite).  Padded launches therefore test
// every sample (exponent field all ones) and take a predicated slow path for the few steps that carry one.
__device__ __forceinline__ bool nonfinite_bits(float v) { return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u; }
__device__ __forceinline__ bool nonfinite_bits(double v) {
  return (static_cast<unsigned>(__double2hiint(v)) & 0x7ff00000u) == 0x7ff00000u;
}

// PerceptibleReciprocal's clamp (pixel-accessor.h:242-254: 1/x if |x| >= MagickEpsilon else sign/MagickEpsilon) applied
// to gamma = QS * den, i.e. |den| is raised to MagickEpsilon / QuantumScale, exactly.
__device__ __forceinline__ double clamp_denominator(double den) {
  // |den| < eps/QS decided exactly as ONE 64-bit unsigned comparison of the magnitude bits (positive doubles order like
  // their bit patterns): ISETP + ISETP.EX, then two selects -- one instruction more than r01's high-word-only test.
  constexpr double kTiny = kEpsilon / kQuantumScale;           // 6.5535e-8
  const unsigned hi = static_cast<unsigned>(__double2hiint(den));
  const unsigned long long mag = (static_cast<unsigned long long>(hi & 0x7fffffffu) << 32) | static_cast<unsigned>(__double2loint(den));
  const unsigned long long tiny = static_cast<unsigned long long>(__double_as_longlong(kTiny));
  if (mag < tiny)
    den = __hiloint2double(static_cast<int>((hi & 0x80000000u) | static_cast<unsigned>(tiny >> 32)), static_cast<int>(tiny & 0xffffffffu));
  return den;
}

// UnsharpMaskImage's point pass (effect.c:4358-4364) on the float-rounded blur value, in the reference's operation
// order with unfused double arithmetic: bit-identical to running it as a separate pass.
__device__ __forceinline__ float unsharp_point(float p, float blurred, double gain, double qthreshold) {
  const double d = __dsub_rn(static_cast<double>(p), static_cast<double>(blurred));
  if (fabs(__dmul_rn(2.0, d)) < qthreshold) return p;
  return static_cast<float>(__dadd_rn(static_cast<double>(p), __dmul_rn(gain, d)));
}

// 1/g to ~1 ulp: MUFU.RCP64H seed (one XU op, ~20 bits) + two FP64 Newton steps.
__device__ __forceinline__ double fast_reciprocal(double g) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(g));
  const double e0 = fma(-g, r0, 1.0);
  const double r1 = fma(r0, e0, r0);              // ~2^-40
  const double e1 = fma(-g, r1, 1.0);
  return fma(r1, e1, r1);                         // ~1 ulp of double: keeps the first pass of a two-pass operator
}                                                 // bit-identical to the reference's quotient in all but ~1e-8 of the samples

__device__ __forceinline__ double shfl_double(double v, int lane) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, lane);
  hi = __shfl_sync(0xffffffffu, hi, lane);
  return __hiloint2double(hi, lo);
}

}  // namespace
}  // namespace mb200
